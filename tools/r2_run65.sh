#!/bin/bash
# ncu --set full of the Q2 back-transformation kernel (n = 8000, 286 columns: first chunk of 143)
mkdir -p gpurun_out
T=${1:-r02g_q2}
timeout 500 ncu --set full --clock-control none --import-source on -k regex:q2_ -c 2 -f -o gpurun_out/$T \
  python tools/q2_probe.py 8000 286 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log; ls -la gpurun_out/$T.ncu-rep
