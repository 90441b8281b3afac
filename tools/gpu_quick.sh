#!/bin/bash
# quick iteration: GPU parity tests, one bench line (no CPU arm), ncu launch list of a 148-job K=4096 wave
TAG=${1:-q}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
( timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
tail -4 gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 > gpurun_out/${TAG}_launches.log 2>&1
