set -x
python bench.py > gpurun_out/f_ddi.json 2> gpurun_out/f_ddi.err
python bench.py --workload collab > gpurun_out/f_collab.json 2> gpurun_out/f_collab.err
python bench.py --workload citation2 --steps 5 --warmup 3 > gpurun_out/f_cit.json 2> gpurun_out/f_cit.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_f_ddi.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/launches_f_ddi.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_csr -s 1 -c 1 -o gpurun_out/prof_spmm_f200_final python tools/spmm_one.py 200 > gpurun_out/ncu_f200.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_csr -s 1 -c 1 -o gpurun_out/prof_spmm_f50_final python tools/spmm_one.py 50 > gpurun_out/ncu_f50.log 2>&1
python tools/spmm_sweep.py 50 64 100 128 200 256 512 > gpurun_out/f_spmm_sweep.txt 2>&1
timeout 600 python tools/microbench.py sweep > gpurun_out/f_config5_sweep.txt 2>&1
tail -3 gpurun_out/f_config5_sweep.txt
for f in f_ddi f_collab f_cit; do python -c "
import json
d=json.load(open('gpurun_out/$f.json'))
print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'] if d.get('cpu_baseline') else None, d['roofline'].get('frac'))
"; done
