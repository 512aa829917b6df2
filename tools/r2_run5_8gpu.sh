#!/bin/bash
# 8-GPU pass: peer collectives + partitioned-fit parity on 8 ranks, headline bench at N=8, configs 4 and 5
mkdir -p gpurun_out
NG=${1:-8}
nvidia-smi -L > gpurun_out/r2_run5_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tests/dist_worker.py nccl > gpurun_out/r2_run5_dist_worker_${NG}gpu.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2_run5_dist_worker_${NG}gpu.log
timeout 400 $TR --master-port 29542 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/r2_run5_bench_${NG}gpu.json 2> gpurun_out/r2_run5_bench_${NG}gpu.err; echo "bench rc=$?" >> gpurun_out/r2_run5_bench_${NG}gpu.err
timeout 400 $TR --master-port 29543 bench.py --gpus $NG --config 4 --steps 2 --warmup 1 > gpurun_out/r2_run5_config4_${NG}gpu.json 2> gpurun_out/r2_run5_config4_${NG}gpu.err; echo "c4 rc=$?" >> gpurun_out/r2_run5_config4_${NG}gpu.err
timeout 400 $TR --master-port 29544 bench.py --gpus $NG --config 5 --steps 1 --warmup 1 > gpurun_out/r2_run5_config5_${NG}gpu.json 2> gpurun_out/r2_run5_config5_${NG}gpu.err; echo "c5 rc=$?" >> gpurun_out/r2_run5_config5_${NG}gpu.err
grep -E "OK|rror|assert|rc=" gpurun_out/r2_run5_dist_worker_${NG}gpu.log | tail -14
for f in bench config4 config5; do echo "== $f"; tail -c 900 gpurun_out/r2_run5_${f}_${NG}gpu.json; tail -2 gpurun_out/r2_run5_${f}_${NG}gpu.err; done
