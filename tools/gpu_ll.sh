#!/bin/bash
# tests + launch lists (common / all-SNP wave)
TAG=${1:-ll}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -n 3 gpurun_out/${TAG}_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_all.csv python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 --all-snps > gpurun_out/${TAG}_launches_all.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_launch_summary_common.txt
python tools/launch_summary.py gpurun_out/${TAG}_launches_all.csv | tee gpurun_out/${TAG}_launch_summary_allsnp.txt
