#!/bin/bash
mkdir -p gpurun_out
python - <<PY
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
from bigkrls_b200 import _lib
lib=_lib.load(); ctx=_lib.default_context(0); r=C.c_double()
for kind,name in ((0,'dfma'),(1,'dmma'),(6,'dmma+dfma mixed')):
    _lib.check(lib.bk_microbench(ctx.handle, kind, 0, 0, C.byref(r))); print(name, round(r.value,2), 'TF/s')
PY
timeout 300 python -m pytest tests/test_gpu_twostage.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/fit_probe.py 20000 10 0.001 > gpurun_out/r2_run34.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/r2_run34.log'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('t_total','t_eigen','t_sy2sb','t_sb2st','t_dc','t_q2','t_q1','gpu_launches','band_gemm_seconds')})
PY
