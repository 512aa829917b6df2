#!/bin/bash
# round-2 final 1-GPU pass: full GPU test-suite, headline bench, launch list, ncu --set full of the three kernels that
# carry the step (rank-256 update GEMM, m x 64 x m product GEMM, bulge chasing)
mkdir -p gpurun_out
T=r02c
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/${T}_bench_N20000.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_N20000.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print(json.dumps({k:round(v,5) for k,v in s.items()}))
print('value', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'pageable', round(d['e2e_pageable']['value'],4), 'roof', round(d['roofline']['achieved'],2), round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
print('parity', json.dumps(d.get('parity_vs_oracle_fixture'))[:600])
print('cpu', json.dumps(d.get('cpu_baseline'))[:400]); print('pair', json.dumps(d.get('same_config_pair')))
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/fit_probe.py 20000 10 0.001 once > gpurun_out/${T}_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches.csv full > gpurun_out/${T}_launch_summary_N20000.txt 2>&1
head -14 gpurun_out/${T}_launch_summary_N20000.txt; tail -1 gpurun_out/${T}_launch_summary_N20000.txt
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:dgemm_kernelILi64ELi64ELi2ELi4ELb0ELb1ELb1E -s 60 -c 1 -o gpurun_out/${T}_update256 python tools/fit_probe.py 20000 10 0.001 once > gpurun_out/${T}_update256.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:dgemm_kernelILi128ELi64ELi4ELi2ELb0ELb0ELb1E -s 400 -c 1 -o gpurun_out/${T}_zprod python tools/fit_probe.py 20000 10 0.001 once > gpurun_out/${T}_zprod.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chase_ws_kernel -c 1 -o gpurun_out/${T}_chase python tools/fit_probe.py 20000 10 0.001 once > gpurun_out/${T}_chase.log 2>&1
for k in update256 zprod chase; do ncu -i gpurun_out/${T}_$k.ncu-rep --page raw --csv > gpurun_out/${T}_${k}_raw.csv 2>/dev/null; python tools/ncu_key_metrics.py gpurun_out/${T}_${k}_raw.csv > gpurun_out/${T}_${k}_ncu_key.csv 2>/dev/null; rm -f gpurun_out/${T}_$k.ncu-rep; done
wc -l gpurun_out/${T}_*_ncu_key.csv
