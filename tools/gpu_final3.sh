#!/bin/bash
# final build of the round: all GPU tests, smoke, bench line + reference arm, NIPT / K512 workloads, NIPT launch list, full-panel pass
TAG=${1:-fin3}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
( time timeout 1500 python bench.py --steps 2 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err | cut -c1-300
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 3 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_nipt.csv \
    python tools/prof_sweep.py --K 2048 --jobs 148 --its 6 --ff 0.1 --coverage 0.5 > gpurun_out/${TAG}_launches_nipt.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_nipt.csv > gpurun_out/${TAG}_launches_nipt_summary.txt
timeout 600 python tools/bench_haploid.py 296 > gpurun_out/${TAG}_haploid.json 2> gpurun_out/${TAG}_haploid.err
for W in nipt_2Mb_0.5x_K2048 chr20_2Mb_1x_K512; do
  ( timeout 1200 python bench.py --steps 2 --warmup 3 --workload $W ) > gpurun_out/${TAG}_wl_${W}.json 2> gpurun_out/${TAG}_wl_${W}.err
  echo "$W exit $?"; tail -2 gpurun_out/${TAG}_wl_${W}.err | cut -c1-300
done
