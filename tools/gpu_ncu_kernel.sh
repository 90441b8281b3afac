#!/bin/bash
# ncu --set full of one kernel (regex) from a prof_sweep.py run, summaries exported on the box
# usage: tools/gpu_ncu_kernel.sh <tag> <kernel regex> <skip> [prof_sweep args]
TAG=$1; RE=$2; SKIP=$3; shift; shift; shift
mkdir -p gpurun_out
REP=/tmp/${TAG}_k
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c 1 -f -o $REP python tools/prof_sweep.py "$@" > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${TAG}_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/${TAG}_src.csv 45 > gpurun_out/${TAG}_hotspots.txt
head -48 gpurun_out/${TAG}_hotspots.txt | cut -c1-170
