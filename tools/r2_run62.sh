#!/bin/bash
# round-2 closing 1-GPU pass: smoke(), full GPU test-suite, headline bench, reference arm (bounded)
mkdir -p gpurun_out
T=${1:-r02f}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/${T}_bench_N20000.json 2> gpurun_out/${T}_bench.err; tail -1 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_N20000.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print(json.dumps({k:round(v,5) for k,v in s.items()}))
print('value', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'pageable', round(d['e2e_pageable']['value'],4), 'roof', round(d['roofline']['achieved'],2), round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('parity', d.get('parity_vs_oracle_fixture',{}).get('within_tolerance'), 'pair', d.get('same_config_pair',{}).get('ratio'))
print({k:round(v['frac'],3) for k,v in d['stage_roofline'].items()})
PY
