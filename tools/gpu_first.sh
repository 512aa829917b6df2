#!/bin/bash
# First GPU contact: microbenchmarks (FP64 roofline denominators), stage tests, a mid-size fit.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python tools/microbench.py > gpurun_out/microbench.json 2> gpurun_out/microbench.err
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/test_ops.log 2>&1
echo "ops exit $?" >> gpurun_out/test_ops.log
timeout 900 python -m pytest tests/test_gpu_fit.py -x -q -m gpu > gpurun_out/test_fit.log 2>&1
echo "fit exit $?" >> gpurun_out/test_fit.log
timeout 600 python tools/fit_probe.py 5000 10 > gpurun_out/probe5000.log 2>&1
tail -5 gpurun_out/test_ops.log gpurun_out/test_fit.log gpurun_out/probe5000.log
cat gpurun_out/microbench.json
