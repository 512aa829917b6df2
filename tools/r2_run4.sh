#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_golden.py tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r2_run4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run4_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_run4_bench.json 2> gpurun_out/r2_run4_bench.err
BK_SY2SB_NOLA=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_run4_bench_nola.json 2> gpurun_out/r2_run4_bench_nola.err
T=r02b
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/fit_probe.py 20000 10 0.001 once > gpurun_out/${T}_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches.csv full > gpurun_out/${T}_launch_summary_N20000.txt 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:dgemm_kernelILi64ELi64ELi2ELi4ELb0ELb1ELb1E -s 60 -c 1 -o gpurun_out/${T}_update256 python tools/fit_probe.py 20000 10 0.001 once > gpurun_out/${T}_update256.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gauss_tile_kernel -c 1 -o gpurun_out/${T}_gauss python tools/fit_probe.py 20000 10 0.001 once > gpurun_out/${T}_gauss.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:dgemm_kernelILi128ELi32 -c 1 -o gpurun_out/${T}_kpass python tools/fit_probe.py 20000 10 0.001 once > gpurun_out/${T}_kpass.log 2>&1
for k in update256 gauss kpass; do ncu -i gpurun_out/${T}_$k.ncu-rep --page raw --csv > gpurun_out/${T}_${k}_raw.csv 2>/dev/null; done
tail -3 gpurun_out/r2_run4_pytest.log
for f in bench bench_nola; do python - <<PY
import json
d=json.loads(open('gpurun_out/r2_run4_$f.json').read().strip().splitlines()[-1])
s=d['stage_seconds']; print('$f', d['value'], 'e2e', d['e2e']['value'], 'sy2sb', s['t_sy2sb'], 'eigen', s['t_eigen'], 'lam', s['t_lambda'], 'kernel', s['t_kernel'], 'roof', d['roofline']['achieved'], d['roofline']['share_of_step'], 'launches', d['gpu_launches'])
PY
done
head -25 gpurun_out/${T}_launch_summary_N20000.txt
