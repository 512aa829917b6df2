#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_twostage.py -m gpu -x -q > gpurun_out/r2_run33_pytest.log 2>&1; tail -2 gpurun_out/r2_run33_pytest.log
for v in new seq; do
if [ $v = seq ]; then export BK_LARFT_SEQ=1; fi
timeout 200 python tools/fit_probe.py 20000 10 0.001 > gpurun_out/r2_run33_$v.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/r2_run33_$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', {k:d[k] for k in ('t_total','t_eigen','t_sy2sb','t_sb2st','t_dc','t_q2','t_q1','gpu_launches','band_gemm_seconds')})
PY
done
