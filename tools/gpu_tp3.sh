#!/bin/bash
# GPU parity tests + timing of a 148-job K=4096 wave: class-path thresholds on common-SNP calls, classes forced on / off on all-SNP calls
TAG=${1:-tp3}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -n 3 gpurun_out/${TAG}_pytest.log
python tools/exp_sweep.py "CLS_MIN=8" "CLS_MIN=3" "CLS_MIN=1" "CLASSES=0" "ALL" "ALL,CLASSES=1,CLS_MIN=3" "ALL,CLASSES=1,CLS_MIN=1" | tee gpurun_out/${TAG}_exp.log
