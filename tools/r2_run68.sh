#!/bin/bash
# ncu --set full (source-level samples) of the bulge-chasing kernel, n = 8000
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:chase_ws -c 1 -f -o gpurun_out/r02j_chase \
  python tools/q2_probe.py 8000 36 > gpurun_out/r02j_chase_ncu.log 2>&1
tail -3 gpurun_out/r02j_chase_ncu.log; ls -la gpurun_out/r02j_chase.ncu-rep
