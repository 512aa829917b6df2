#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r2_run55_pytest.log 2>&1; tail -2 gpurun_out/r2_run55_pytest.log
timeout 200 python tools/fit_probe.py 20000 10 0.001 > gpurun_out/r2_run55.log 2>&1
python - <<PY
import json
ls=[json.loads(l) for l in open('gpurun_out/r2_run55.log') if l.startswith('{')]
d=ls[-1]; print({k:d[k] for k in ('t_total','t_eigen','t_lambda','t_coef','t_vcov','t_deriv','t_kernel','lambda','lastkeeper','n_probes','n_passes')})
PY
