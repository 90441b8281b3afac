#!/bin/bash
# tests of the class path + launch list of a common wave + per-phase clocks (instrumented variant)
TAG=${1:-c}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "gibbs or golden or full_size" ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -n 3 gpurun_out/${TAG}_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_launch_summary_common.txt
QUILT_B200_LIB=$PWD/quilt_b200/libquiltgpu_clk.so timeout 600 python tools/prof_sweep.py --K 4096 --jobs 148 --its 8 > gpurun_out/${TAG}_clk_common.log 2>&1
grep -h QBCLK gpurun_out/${TAG}_clk_common.log | head -1
tail -n 1 gpurun_out/${TAG}_clk_common.log
