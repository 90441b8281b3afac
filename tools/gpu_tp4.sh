#!/bin/bash
# quick: sweep-related GPU parity tests + timing of a 148-job K=4096 wave (common-SNP with / without classes, all-SNP with / without)
TAG=${1:-tp4}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_full_size.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -n 3 gpurun_out/${TAG}_pytest.log
python tools/exp_sweep.py "$@" | tee gpurun_out/${TAG}_exp.log
