#!/bin/bash
# GPU parity suite + bench lines (tiny workload first as a smoke of the chained path, then the headline workload)
TAG=${1:-t1}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
( timeout 600 python bench.py --workload tiny --steps 2 --warmup 3 ) > gpurun_out/${TAG}_bench_tiny.json 2> gpurun_out/${TAG}_bench_tiny.err
echo "tiny bench exit $?"; tail -3 gpurun_out/${TAG}_bench_tiny.err
( time timeout 1500 python bench.py --steps 2 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
tail -6 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | cut -c1-600
