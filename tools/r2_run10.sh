#!/bin/bash
mkdir -p gpurun_out
NG=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29551 tools/dist_prof.py 20000 > gpurun_out/r2_run10_prof_${NG}gpu.log 2>&1
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_run10_prof_${NG}gpu.log | tail -40
