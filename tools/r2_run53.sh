#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_twostage.py -m gpu -x -q > gpurun_out/r2_run53_pytest.log 2>&1; tail -3 gpurun_out/r2_run53_pytest.log
timeout 300 python tools/q2_probe.py 20000 36 72 143 2>&1 | tail -4
