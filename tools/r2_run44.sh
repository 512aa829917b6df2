#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_twostage.py -m gpu -x -q > gpurun_out/r2_run44_pytest.log 2>&1; tail -2 gpurun_out/r2_run44_pytest.log
for rep in 1 2; do
timeout 200 python tools/fit_probe.py 20000 10 0.001 > gpurun_out/r2_run44_$rep.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/r2_run44_$rep.log'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('t_total','t_eigen','t_sy2sb','t_sb2st','t_dc','t_backtransform')})
PY
done
