#!/bin/bash
# 8-GPU pass: partitioned-fit parity on 8 ranks, headline bench at N=8, config 5
mkdir -p gpurun_out
NG=${1:-8}
T=r02e
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29541 tests/dist_worker.py nccl > gpurun_out/${T}_dist_worker_${NG}gpu.log 2>&1; echo "worker rc=$?" >> gpurun_out/${T}_dist_worker_${NG}gpu.log
timeout 300 $TR --master-port 29542 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/${T}_bench_${NG}gpu.json 2> gpurun_out/${T}_bench_${NG}gpu.err; echo "bench rc=$?" >> gpurun_out/${T}_bench_${NG}gpu.err
timeout 300 $TR --master-port 29544 bench.py --gpus $NG --config 5 --steps 1 --warmup 1 > gpurun_out/${T}_config5_${NG}gpu.json 2> gpurun_out/${T}_config5_${NG}gpu.err; echo "c5 rc=$?" >> gpurun_out/${T}_config5_${NG}gpu.err
grep -E "OK|rror|assert|rc=" gpurun_out/${T}_dist_worker_${NG}gpu.log | tail -12 | cut -c1-260
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_${NG}gpu.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print(json.dumps({k:round(v,5) for k,v in s.items()}))
print('value', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'pageable', round(d['e2e_pageable']['value'],4))
p=d.get('parity_vs_oracle_fixture'); print('parity ok:', all(x.get('within_tolerance') for x in p) if isinstance(p,list) else p)
d=json.loads(open('gpurun_out/${T}_config5_${NG}gpu.json').read().strip().splitlines()[-1]); print('config5', round(d['value'],4))
PY
tail -n 2 gpurun_out/${T}_bench_${NG}gpu.err
