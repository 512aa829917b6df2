#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_twostage.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python tools/gemm_rmw_bench.py 2>&1 | grep -E "lower=2|k=2048"
timeout 120 python tools/gemm_shape_bench.py 2>&1 | grep -E "n=64 k=16384|k=8192 lower"
for v in "1 128" "2 64" "2 96"; do
  set -- $v
  BK_QR_PER_SM=$1 BK_QR_ROWS=$2 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run15_bench.json 2> gpurun_out/r2_run15_bench.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_run15_bench.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print('QR per_sm/rows $v:', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), round(d['e2e_pageable']['value'],4), 'sy2sb', round(s['t_sy2sb'],4), 'eigen', round(s['t_eigen'],4), 'dc', round(s['t_dc'],4), 'kernel', round(s['t_kernel'],5), 'roof', round(d['roofline']['achieved'],2), round(d['roofline']['frac'],3))
PY
done
