#!/bin/bash
# 2-GPU pass: partitioned-fit parity with every rank chasing / D&C and the back-transformation split by columns
mkdir -p gpurun_out
NG=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tests/dist_worker.py nccl > gpurun_out/r2_run31_dist_worker_${NG}gpu.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2_run31_dist_worker_${NG}gpu.log
timeout 400 $TR --master-port 29542 bench.py --gpus $NG --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run31_bench_${NG}gpu.json 2> gpurun_out/r2_run31_bench_${NG}gpu.err; echo "bench rc=$?" >> gpurun_out/r2_run31_bench_${NG}gpu.err
grep -E "OK|rror|assert|rc=" gpurun_out/r2_run31_dist_worker_${NG}gpu.log | tail -14
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_run31_bench_${NG}gpu.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print(json.dumps({k:round(v,5) for k,v in s.items()}))
print('value', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'parity', d.get('parity_vs_fixture', d.get('fixture', None)))
PY
tail -3 gpurun_out/r2_run31_bench_${NG}gpu.err
