#!/bin/bash
# GPU parity tests + one ncu --set full capture of the sweep kernel (148-job K=4096 wave)
TAG=${1:-p}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 9 -c 1 -f -o gpurun_out/${TAG}_sweep \
    python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
