#!/bin/bash
# GPU parity tests + haploid-pass throughput
TAG=${1:-r2b}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/bench_haploid.py 296 > gpurun_out/${TAG}_haploid.json 2> gpurun_out/${TAG}_haploid.err; tail -2 gpurun_out/${TAG}_haploid.err; cat gpurun_out/${TAG}_haploid.json
