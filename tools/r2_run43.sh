#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_twostage.py -m gpu -x -q > gpurun_out/r2_run43_pytest.log 2>&1; tail -3 gpurun_out/r2_run43_pytest.log
BK_CHASE_TRACE=3000 BK_CHASE_TRACE_T0=200 timeout 200 python tools/fit_probe.py 20000 10 0.001 once > gpurun_out/r2_run43_trace.log 2>&1
grep "chase trace\] 3000 " gpurun_out/r2_run43_trace.log | head -4 | cut -c1-220
grep "chase lag ns\] hop 8:" gpurun_out/r2_run43_trace.log | cut -c1-330
timeout 200 python tools/fit_probe.py 20000 10 0.001 > gpurun_out/r2_run43.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/r2_run43.log'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('t_total','t_eigen','t_sy2sb','t_sb2st','t_dc','t_backtransform','lambda','lastkeeper')})
PY
