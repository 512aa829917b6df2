#!/bin/bash
# ncu evidence for the round (B200_PROFILING.md recipe).  Usage: tools/ncu_capture.sh <tag> [N]
TAG=${1:-r01}
N=${2:-20000}
mkdir -p gpurun_out
# (1) every launch of one fit with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/fit_probe.py $N 10 0.001 once > gpurun_out/${TAG}_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv full > gpurun_out/${TAG}_launch_summary.txt 2>&1
# (2) the dominant kernel (stage-1 DGEMM, rank-128 symmetric update), full set, one launch from the middle
ncu --set full --clock-control none --import-source on -k regex:dgemm_kernel -s 1200 -c 6 \
    -o gpurun_out/${TAG}_dgemm python tools/fit_probe.py $N 10 0.001 once > gpurun_out/${TAG}_dgemm.log 2>&1
ncu -i gpurun_out/${TAG}_dgemm.ncu-rep --page raw --csv > gpurun_out/${TAG}_dgemm_raw.csv 2>/dev/null
head -30 gpurun_out/${TAG}_launch_summary.txt
