#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_ops.py tests/test_gpu_fit.py -m gpu -x -q > gpurun_out/r2_run32_pytest.log 2>&1; tail -3 gpurun_out/r2_run32_pytest.log
timeout 200 python tools/fit_probe.py 20000 10 0.001 > gpurun_out/r2_run32.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/r2_run32.log'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('t_total','t_eigen','t_sy2sb','t_sb2st','t_dc','t_backtransform','t_q2','t_q1','lambda','lastkeeper','gpu_launches','band_gemm_seconds')})
PY
