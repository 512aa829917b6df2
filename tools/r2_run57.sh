#!/bin/bash
# round-2 final 1-GPU pass: full GPU test-suite, headline bench, launch list
mkdir -p gpurun_out
T=r02e
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/${T}_bench_N20000.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_N20000.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print(json.dumps({k:round(v,5) for k,v in s.items()}))
print('value', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'pageable', round(d['e2e_pageable']['value'],4), 'roof', round(d['roofline']['achieved'],2), round(d['roofline']['frac'],3), 'share', round(d['roofline']['share_of_step'],3), 'launches', d['gpu_launches'])
print('parity', d.get('parity_vs_oracle_fixture',{}).get('within_tolerance'))
print({k:round(v['frac'],3) for k,v in d['stage_roofline'].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/fit_probe.py 20000 10 0.001 once > gpurun_out/${T}_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches.csv full > gpurun_out/${T}_launch_summary_N20000.txt 2>&1
head -24 gpurun_out/${T}_launch_summary_N20000.txt; tail -1 gpurun_out/${T}_launch_summary_N20000.txt
