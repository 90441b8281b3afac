#!/bin/bash
# GPU parity tests, one bench line without the CPU arm, ncu launch lists of a common-SNP and an all-SNP wave
TAG=${1:-tb}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -n 3 gpurun_out/${TAG}_pytest.log
( timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
tail -n 3 gpurun_out/${TAG}_bench.err | cut -c1-300
cut -c1-900 gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_all.csv python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 --all-snps > gpurun_out/${TAG}_launches_all.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_launch_summary_common.txt
python tools/launch_summary.py gpurun_out/${TAG}_launches_all.csv | tee gpurun_out/${TAG}_launch_summary_allsnp.txt
