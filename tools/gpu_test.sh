#!/bin/bash
# GPU parity tests only; extra pytest args after the tag
TAG=${1:-t}; shift
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q "$@" ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -40 gpurun_out/${TAG}_pytest.log
