#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_twostage.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
  echo "== BK_GEMM_NOPREFETCH=$v"
  if [ $v = 1 ]; then export BK_GEMM_NOPREFETCH=1; else unset BK_GEMM_NOPREFETCH; fi
  timeout 120 python tools/gemm_rmw_bench.py 2>&1 | grep -E "lower=2|k=2048"
  timeout 120 python tools/gemm_shape_bench.py 2>&1 | grep -E "n=64 k=16384|n=286|k=8192 lower"
done > gpurun_out/r2_run14_prefetch.txt 2>&1
cat gpurun_out/r2_run14_prefetch.txt
unset BK_GEMM_NOPREFETCH
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run14_bench.json 2> gpurun_out/r2_run14_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_run14_bench.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print('bench', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), round(d['e2e_pageable']['value'],4), 'steps', [round(x,4) for x in d['per_step_seconds']], 'sy2sb', round(s['t_sy2sb'],4), 'eigen', round(s['t_eigen'],4), 'dc', round(s['t_dc'],4), 'lam', round(s['t_lambda'],5), 'kernel', round(s['t_kernel'],5), 'roof', round(d['roofline']['achieved'],2), d['roofline']['frac'])
PY
