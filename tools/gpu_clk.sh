#!/bin/bash
# per-phase clock breakdown of the sweep kernel (instrumented variant, tools/build_variant.py clk -DQB_CLK=1) + GPU tests
TAG=${1:-c1}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
QUILT_B200_LIB=$PWD/quilt_b200/libquiltgpu_clk.so timeout 600 python tools/prof_sweep.py --K 4096 --jobs 148 --its 8 > gpurun_out/${TAG}_clk_common.log 2>&1
QUILT_B200_LIB=$PWD/quilt_b200/libquiltgpu_clk.so timeout 600 python tools/prof_sweep.py --K 4096 --jobs 148 --its 8 --all-snps > gpurun_out/${TAG}_clk_all.log 2>&1
timeout 600 python tools/prof_sweep.py --K 4096 --jobs 148 --its 8 > gpurun_out/${TAG}_plain_common.log 2>&1
grep -h QBCLK gpurun_out/${TAG}_clk_common.log gpurun_out/${TAG}_clk_all.log | head -4
tail -n 2 gpurun_out/${TAG}_clk_common.log gpurun_out/${TAG}_plain_common.log
