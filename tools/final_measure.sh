# round-2 final measurement set (one B200): run under gpurun, outputs in gpurun_out/, summaries copied to profiles/
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_final_test_gpu.log 2>&1
python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err
# the launch list of the bench command itself (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_final_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_final_launches_bench.log 2>&1
# ncu --set full of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:spmm_csr_staged -s 1 -c 1 -o gpurun_out/r02_final_prof_spmm_f50_window python tools/spmm_one.py 50 > gpurun_out/r02_final_ncu_f50.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_csr_staged -s 1 -c 1 -o gpurun_out/r02_final_prof_spmm_f50_pitch64 python tools/spmm_one.py 50 64 > gpurun_out/r02_final_ncu_f50p64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_tn_kernel -s 2 -c 1 -o gpurun_out/r02_final_prof_gemm_tma_tn python tools/gemm_tn_decomp.py > gpurun_out/r02_final_ncu_tn.log 2>&1
# multi-GPU lines (separate gpurun calls, --gpus N):
#   python -m pytest tests/test_gpu_multi.py -q                                                   (N = 2)
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
#       bench.py --gpus N --steps 20 --warmup 3 > gpurun_out/r02_final_bench_nN.json
