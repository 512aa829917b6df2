#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_golden.py -m gpu -x -q -k "synthetic or config4 or config3 or config2" > gpurun_out/r2_run11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run11_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_worker.py nccl > gpurun_out/r2_run11_dist_worker.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2_run11_dist_worker.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --config 4 --steps 1 --warmup 1 > gpurun_out/r2_run11_config4_2gpu.json 2> gpurun_out/r2_run11_config4_2gpu.err; echo "c4 rc=$?" >> gpurun_out/r2_run11_config4_2gpu.err
tail -3 gpurun_out/r2_run11_pytest.log; grep -E "OK|rror|assert|rc=" gpurun_out/r2_run11_dist_worker.log | cut -c1-250 | tail -12; tail -c 700 gpurun_out/r2_run11_config4_2gpu.json; tail -3 gpurun_out/r2_run11_config4_2gpu.err
