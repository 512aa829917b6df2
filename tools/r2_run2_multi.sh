#!/bin/bash
# round-2 multi-GPU pass (gpurun --gpus 2): IPC probe, peer collectives + partitioned-fit parity, short bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_run9_gpus.txt
nvidia-smi topo -m >> gpurun_out/r2_run9_gpus.txt 2>&1

timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_worker.py nccl > gpurun_out/r2_run9_dist_worker.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2_run9_dist_worker.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/r2_run9_bench2.json 2> gpurun_out/r2_run9_bench2.err; echo "bench rc=$?" >> gpurun_out/r2_run9_bench2.err
grep -E "OK|Error|error|assert|rc=" gpurun_out/r2_run9_dist_worker.log | tail -20; tail -c 600 gpurun_out/r2_run9_bench2.json; tail -3 gpurun_out/r2_run9_bench2.err
