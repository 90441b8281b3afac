#!/bin/bash
# ncu --set full of one kernel (regex) from a 148-job K=4096 wave: tools/gpu_prof_kernel.sh <tag> <regex> [prof_sweep args]
TAG=$1; RE=$2; shift; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s 1 -c 1 -f -o gpurun_out/${TAG} \
    python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 "$@" > gpurun_out/${TAG}.log 2>&1
tail -3 gpurun_out/${TAG}.log
