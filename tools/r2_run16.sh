#!/bin/bash
# re-entry check: in-kernel cycle profiles of the latency-bound kernels + the bench line of the committed state
mkdir -p gpurun_out
BK_CHASE_PROF=1 BK_QR_PROF=1 BK_Q2_PROF=1 timeout 300 python tools/fit_probe.py 20000 10 0.001 > gpurun_out/r2_run16_prof.log 2>&1
grep -E "prof|t_total" gpurun_out/r2_run16_prof.log | cut -c1-900
timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run16_bench.json 2> gpurun_out/r2_run16_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_run16_bench.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print(json.dumps({k:round(v,5) for k,v in s.items()}))
print(round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'roof', round(d['roofline']['achieved'],2), round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
PY
