#!/bin/bash
# Q2 back-transformation (pass window in registers, 4 sweeps as one block reflector): parity tests, seconds at the headline size
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_twostage.py -x -q -m gpu -k "test_twostage or matches_onestage" 2>&1 | tail -3
BK_Q2_PROF=1 timeout 300 python tools/q2_probe.py 20000 286 36 2>&1 | awk '!seen[$0]++' | tail -8
