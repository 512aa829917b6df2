#!/bin/bash
# ncu evidence for the round (B200_PROFILING.md recipe).  Usage: tools/ncu_capture.sh <tag> [N]
TAG=${1:-r01}
N=${2:-20000}
mkdir -p gpurun_out
# (1) every launch of one fit with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/fit_probe.py $N 10 0.001 once > gpurun_out/${TAG}_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_summary.txt 2>&1
# (2) the dominant kernel, full set, one launch from the middle of the reduction
ncu --set full --clock-control none --import-source on -k regex:sytrd_panel -s 40 -c 1 \
    -o gpurun_out/${TAG}_sytrd python tools/fit_probe.py $N 10 0.001 once > gpurun_out/${TAG}_sytrd.log 2>&1
ncu -i gpurun_out/${TAG}_sytrd.ncu-rep --page raw --csv > gpurun_out/${TAG}_sytrd_raw.csv 2>/dev/null
tail -30 gpurun_out/${TAG}_launch_summary.txt
