#!/bin/bash
mkdir -p gpurun_out
export BK_ALLOC_VERBOSE=1
for mode in nola la; do
  if [ $mode = nola ]; then export BK_SY2SB_NOLA=1; else unset BK_SY2SB_NOLA; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run8_bench_$mode.json 2> gpurun_out/r2_run8_bench_$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_run8_bench_$mode.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print('$mode', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'steps', [round(x,4) for x in d['per_step_seconds']], 'sy2sb', round(s['t_sy2sb'],4), 'eigen', round(s['t_eigen'],4), 'vcov', round(s['t_vcov'],4))
PY
  grep -c alloc gpurun_out/r2_run8_bench_$mode.err; grep alloc gpurun_out/r2_run8_bench_$mode.err | sort | uniq -c | sort -rn | head -20
done
unset BK_SY2SB_NOLA; unset BK_ALLOC_VERBOSE
timeout 300 python tools/e2e_pageable_probe.py > gpurun_out/r2_run8_e2e_probe.json 2> gpurun_out/r2_run8_e2e_probe.err; cat gpurun_out/r2_run8_e2e_probe.json
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fit.py -m gpu -x -q 2>&1 | tail -3
