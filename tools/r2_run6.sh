#!/bin/bash
mkdir -p gpurun_out
for mode in la nola la; do
  if [ $mode = nola ]; then export BK_SY2SB_NOLA=1; else unset BK_SY2SB_NOLA; fi
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run6_bench_$mode.json 2> gpurun_out/r2_run6_bench_$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_run6_bench_$mode.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print('$mode', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'steps', [round(x,4) for x in d['per_step_seconds']], 'sy2sb', round(s['t_sy2sb'],4), 'eigen', round(s['t_eigen'],4), 'vcov', round(s['t_vcov'],4), 'lam', round(s['t_lambda'],5), 'kernel', round(s['t_kernel'],5), 'roof', round(d['roofline']['achieved'],2), 'launches', d['gpu_launches'])
PY
done
unset BK_SY2SB_NOLA
timeout 300 python tools/e2e_pageable_probe.py > gpurun_out/r2_run6_e2e_probe.json 2> gpurun_out/r2_run6_e2e_probe.err; cat gpurun_out/r2_run6_e2e_probe.json
