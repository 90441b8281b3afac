#!/bin/bash
# per-phase clock breakdown only (instrumented variant) + plain timing, common and all-SNP waves
TAG=${1:-c}
mkdir -p gpurun_out
QUILT_B200_LIB=$PWD/quilt_b200/libquiltgpu_clk.so timeout 600 python tools/prof_sweep.py --K 4096 --jobs 148 --its 8 > gpurun_out/${TAG}_clk_common.log 2>&1
timeout 600 python tools/prof_sweep.py --K 4096 --jobs 148 --its 8 > gpurun_out/${TAG}_plain_common.log 2>&1
timeout 600 python tools/prof_sweep.py --K 4096 --jobs 148 --its 8 --all-snps > gpurun_out/${TAG}_plain_all.log 2>&1
grep -h QBCLK gpurun_out/${TAG}_clk_common.log | head -2
tail -n 1 gpurun_out/${TAG}_clk_common.log gpurun_out/${TAG}_plain_common.log gpurun_out/${TAG}_plain_all.log
