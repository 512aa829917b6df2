#!/bin/bash
# 2-GPU bench line (fixture comparison on every rank included) on the state with the DMMA back-transformation
mkdir -p gpurun_out
NG=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29542 bench.py --gpus $NG --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_${NG}gpu.json 2> gpurun_out/r02i_bench_${NG}gpu.err; echo "bench rc=$?" >> gpurun_out/r02i_bench_${NG}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02i_bench_${NG}gpu.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print(json.dumps({k:round(v,5) for k,v in s.items()}))
print('value', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'parity', d.get('parity_vs_oracle_fixture'))
PY
tail -3 gpurun_out/r02i_bench_${NG}gpu.err
