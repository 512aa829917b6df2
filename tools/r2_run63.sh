#!/bin/bash
for lz in 1 0; do
BK_DC_LAZY=$lz FIT_REPS=4 timeout 200 python tools/fit_probe.py 3000 10 0.001 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lazy=$lz', {k:round(d[k],5) for k in ('wall','t_total','t_eigen','t_tridiag','t_dc','t_backtransform','t_lambda','t_kernel')})"
done
