#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_ops.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r2_run54_pytest.log 2>&1; tail -3 gpurun_out/r2_run54_pytest.log
timeout 200 python tools/fit_probe.py 20000 10 0.001 > gpurun_out/r2_run54.log 2>&1
python - <<PY
import json
ls=[json.loads(l) for l in open('gpurun_out/r2_run54.log') if l.startswith('{')]
d=ls[-1]; print({k:d[k] for k in ('t_total','t_eigen','t_lambda','t_coef','t_vcov','t_deriv','t_kernel','lambda','lastkeeper','n_probes','n_passes')})
PY
