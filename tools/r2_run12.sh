#!/bin/bash
mkdir -p gpurun_out
for s in 0 4000 8000 16000 32000 64000; do
  echo "== BK_GEMM_STAGGER=$s"
  BK_GEMM_STAGGER=$s timeout 120 python tools/gemm_rmw_bench.py 2>&1 | grep -E "k=256|k=128 lower=2|k=2048"
  BK_GEMM_STAGGER=$s timeout 120 python tools/gemm_shape_bench.py 2>&1 | grep -E "n=64 k=16384|k=8192"
done > gpurun_out/r2_run12_stagger.txt 2>&1
cat gpurun_out/r2_run12_stagger.txt
