#!/bin/bash
# GPU parity tests + plain timing of a 148-job K=4096 wave (common and all-SNP)
TAG=${1:-tp}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -n 3 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/prof_sweep.py --K 4096 --jobs 148 --its 8 > gpurun_out/${TAG}_plain_common.log 2>&1
timeout 600 python tools/prof_sweep.py --K 4096 --jobs 148 --its 8 --all-snps > gpurun_out/${TAG}_plain_all.log 2>&1
tail -n 1 gpurun_out/${TAG}_plain_common.log gpurun_out/${TAG}_plain_all.log
