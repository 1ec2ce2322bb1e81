# round-2 measurement set (one B200): run under gpurun, outputs in gpurun_out/, summaries copied to profiles/
set -x
python tools/gemm_dw_ab.py > gpurun_out/r02_gemm_dw_ab.txt 2>&1
python tools/gemm_skinny_ab.py > gpurun_out/r02_gemm_skinny_ab.txt 2>&1
python tools/spmm_sweep.py 32 50 64 128 200 256 512 > gpurun_out/r02_spmm_sweep.txt 2>&1
# the launch list of the bench command itself (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_launches_bench.log 2>&1
# ncu --set full of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:spmm_csr_kernel -s 1 -c 1 -o gpurun_out/r02_prof_spmm_f50 python tools/spmm_one.py 50 > gpurun_out/r02_ncu_f50.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_csr_narrow -s 1 -c 1 -o gpurun_out/r02_prof_spmm_f64 python tools/spmm_one.py 64 > gpurun_out/r02_ncu_f64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 2 -c 1 -o gpurun_out/r02_prof_gemm_tma_200x178 python tools/gemm_tma_one.py 2927963 200 178 1 > gpurun_out/r02_ncu_tma.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 2 -c 1 -o gpurun_out/r02_prof_gemm_tma_50x200 python tools/gemm_tma_one.py 2927963 50 200 1 > gpurun_out/r02_ncu_tma2.log 2>&1
