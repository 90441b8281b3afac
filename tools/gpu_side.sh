#!/bin/bash
# ncu --set full of the side kernels (one launch each), exported to CSV on the box (the reports themselves are too large to bring back)
TAG=${1:-k1}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_haploid_pass.py -m gpu -x -q -s 2>&1 | tail -12 ) | tee gpurun_out/${TAG}_hap.log
cap() { # kernel-regex extra-args
  KN=$1; shift
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KN} -c 1 -f -o /tmp/${TAG}_${KN} python tools/prof_sweep.py --K 4096 --jobs 148 --its 3 "$@" > gpurun_out/${TAG}_${KN}.log 2>&1
  ncu -i /tmp/${TAG}_${KN}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${KN}_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_${KN}.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${TAG}_${KN}_source.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/${TAG}_${KN}_source.csv 25 > gpurun_out/${TAG}_${KN}_hotspots.txt 2>&1
  rm -f /tmp/${TAG}_${KN}.ncu-rep
}
cap k_make_eG
cap k_build_tables
cap k_happrobs
cap k_assemble_all --all-snps
ls -la gpurun_out | grep ${TAG}
