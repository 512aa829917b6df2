#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  echo "== BK_GEMM_2STAGE=$v"
  if [ $v = 1 ]; then export BK_GEMM_2STAGE=1; else unset BK_GEMM_2STAGE; fi
  timeout 120 python tools/gemm_rmw_bench.py 2>&1 | grep -E "lower=2"
done > gpurun_out/r2_run13_2stage.txt 2>&1
cat gpurun_out/r2_run13_2stage.txt
