#!/bin/bash
# Q2 back-transformation on the DMMA path (8 sweeps per block reflector): parity tests, then seconds at the headline size
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_twostage.py -x -q -m gpu -k "test_twostage or matches_onestage" 2>&1 | tail -5
timeout 200 python tools/q2_probe.py 20000 286 36 80 2>&1 | tail -4
