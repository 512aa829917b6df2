#!/bin/bash
# closing pass on the final library: smoke(), GPU test-suite, ncu launch list of one headline fit
mkdir -p gpurun_out
T=r02k
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/fit_probe.py 20000 10 0.001 once > gpurun_out/${T}_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches.csv full > gpurun_out/${T}_launch_summary_N20000.txt 2>&1
rm -f gpurun_out/${T}_launches.csv
head -12 gpurun_out/${T}_launch_summary_N20000.txt; tail -1 gpurun_out/${T}_launch_summary_N20000.txt
