#!/bin/bash
mkdir -p gpurun_out
for W in chr20_2Mb_1x_K512 nipt_2Mb_0.5x_K2048; do
  ( timeout 900 python bench.py --steps 2 --warmup 3 --workload $W ) > gpurun_out/wl_${W}.json 2> gpurun_out/wl_${W}.err
  echo "$W exit $?"; tail -4 gpurun_out/wl_${W}.err
done
