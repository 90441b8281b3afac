#!/bin/bash
# ncu --set full of k_hap_fb (148 full-panel passes), summaries exported on the box
TAG=${1:-hap}
mkdir -p gpurun_out
REP=/tmp/${TAG}_hap
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hap_fb -c 1 -f -o $REP python tools/bench_haploid.py 148 > gpurun_out/${TAG}_hap_ncu.log 2>&1
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/${TAG}_hap_raw.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${TAG}_hap_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/${TAG}_hap_src.csv 50 > gpurun_out/${TAG}_hap_hotspots.txt
head -40 gpurun_out/${TAG}_hap_hotspots.txt | cut -c1-160
