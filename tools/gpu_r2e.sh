#!/bin/bash
# GPU parity tests, full-panel pass throughput, ncu launch list of one NIPT wave (K = 2048, ff = 0.1, 0.5x)
TAG=${1:-r2e}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/bench_haploid.py 296 > gpurun_out/${TAG}_haploid.json 2> gpurun_out/${TAG}_haploid.err; tail -2 gpurun_out/${TAG}_haploid.err; cat gpurun_out/${TAG}_haploid.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_nipt.csv \
    python tools/prof_sweep.py --K 2048 --jobs 148 --its 6 --ff 0.1 --coverage 0.5 > gpurun_out/${TAG}_launches_nipt.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_nipt.csv | tee gpurun_out/${TAG}_launch_summary_nipt.txt
