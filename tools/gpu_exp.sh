#!/bin/bash
TAG=${1:-e}; shift
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python tools/exp_sweep.py "$@" | tee gpurun_out/${TAG}_exp.jsonl
