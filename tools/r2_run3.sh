#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run3_pytest.log
BK_SY2SB_LOOKAHEAD=1 timeout 600 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_golden.py -m gpu -x -q -k "twostage or config3 or config2" > gpurun_out/r2_run3_pytest_la.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run3_pytest_la.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run3_bench_pair.json 2> gpurun_out/r2_run3_bench_pair.err
BK_SY2SB_LOOKAHEAD=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run3_bench_la.json 2> gpurun_out/r2_run3_bench_la.err
BK_SY2SB_LOOKAHEAD=1 BK_QR_ROWS=256 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run3_bench_la256.json 2> gpurun_out/r2_run3_bench_la256.err
BK_SY2SB_LOOKAHEAD=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_run3_launches_la.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_run3_ncu.log 2>&1
tail -3 gpurun_out/r2_run3_pytest.log; tail -3 gpurun_out/r2_run3_pytest_la.log
for f in pair la la256; do python - <<PY
import json
d=json.loads(open('gpurun_out/r2_run3_bench_$f.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print('$f', d['value'], 'e2e', d['e2e']['value'], 'sy2sb', s['t_sy2sb'], 'eigen', s['t_eigen'], 'lam', s['t_lambda'], 'roof', d['roofline']['achieved'], d['roofline']['share_of_step'], 'launches', d['gpu_launches'])
PY
done
