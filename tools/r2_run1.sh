#!/bin/bash
# round-2 first GPU pass: full -m gpu suite (incl. the new full-size fixtures), pageable e2e probe, short bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_run1_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run1_pytest.log
timeout 300 python tools/e2e_pageable_probe.py > gpurun_out/r2_run1_e2e_probe.json 2> gpurun_out/r2_run1_e2e_probe.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_run1_bench.json 2> gpurun_out/r2_run1_bench.err
tail -5 gpurun_out/r2_run1_pytest.log; cat gpurun_out/r2_run1_e2e_probe.json; tail -c 1500 gpurun_out/r2_run1_bench.json
