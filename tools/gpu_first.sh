#!/bin/bash
# First GPU contact: microbenchmarks (FP64 roofline denominators), stage tests, a mid-size fit.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in gauss temp_kernel dgemm crossprod loo eigen deriv neffective error; do
  timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "$grp" > gpurun_out/ops_$grp.log 2>&1
  echo "exit $?" >> gpurun_out/ops_$grp.log
  echo "== $grp: $(tail -2 gpurun_out/ops_$grp.log | tr '\n' ' ')"
done
for grp in mtcars synthetic config1 predict crossvalidate validation; do
  timeout 900 python -m pytest tests/test_gpu_fit.py -q -m gpu -k "$grp" > gpurun_out/fit_$grp.log 2>&1
  echo "exit $?" >> gpurun_out/fit_$grp.log
  echo "== fit $grp: $(tail -2 gpurun_out/fit_$grp.log | tr '\n' ' ')"
done
timeout 600 python tools/microbench.py > gpurun_out/microbench.json 2> gpurun_out/microbench.err
timeout 600 python tools/fit_probe.py 5000 10 > gpurun_out/probe5000.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/fit_probe.py 300 4 > gpurun_out/sanitize.log 2>&1
tail -4 gpurun_out/probe5000.log
tail -5 gpurun_out/sanitize.log
head -c 1500 gpurun_out/microbench.json
