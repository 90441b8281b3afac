#!/bin/bash
# GPU parity suite + one short bench line (with the parity gate) of the headline workload
TAG=${1:-t1}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
( time timeout 1500 python bench.py --steps 2 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
tail -4 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | cut -c1-1500
