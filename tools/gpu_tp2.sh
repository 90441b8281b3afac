#!/bin/bash
# GPU parity tests + timing of a 148-job K=4096 wave under several class-path thresholds
TAG=${1:-tp}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -n 3 gpurun_out/${TAG}_pytest.log
python tools/exp_sweep.py "CLS_MIN=1" "CLS_MIN=6" "CLS_MIN=10" "CLS_MIN=100" "ALL,CLS_MIN=1" "ALL,CLS_MIN=6" "ALL,CLS_MIN=10" "ALL,CLS_MIN=100" | tee gpurun_out/${TAG}_exp.log
