#!/bin/bash
# ncu --set full of one k_sweep launch of a common-SNP wave and of an all-SNP wave (K = 4096, 148 jobs); the reports are
# turned into CSV / per-line summaries ON THE BOX (two .ncu-rep files exceed what gpurun copies back).
# usage: tools/gpu_ncu_sweep.sh <tag>
TAG=${1:-ncu}
mkdir -p gpurun_out
for KIND in common allsnp; do
  EXTRA=""; [ $KIND = allsnp ] && EXTRA="--all-snps"
  REP=/tmp/${TAG}_${KIND}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 9 -c 1 -f -o $REP \
      python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 $EXTRA > gpurun_out/${TAG}_${KIND}_ncu.log 2>&1
  ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/${TAG}_${KIND}_raw.csv 2>/dev/null
  ncu -i $REP.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${TAG}_${KIND}_src.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/${TAG}_${KIND}_src.csv 60 > gpurun_out/${TAG}_${KIND}_hotspots.txt
  tail -2 gpurun_out/${TAG}_${KIND}_ncu.log
done
ls -la gpurun_out
