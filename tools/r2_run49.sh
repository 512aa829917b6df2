#!/bin/bash
mkdir -p gpurun_out
BK_SY2SB_PIPE_TRACE=20 timeout 200 python tools/fit_probe.py 20000 10 0.001 > gpurun_out/r2_run49.log 2>&1
grep "sy2sb pipe" gpurun_out/r2_run49.log | tail -8
