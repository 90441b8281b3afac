#!/bin/bash
# Round-end evidence in one gpurun call: GPU parity tests, smoke, bench line + reference arm, ncu launch lists (common / all-SNP / NIPT waves),
# the other BASELINE.json shapes.  usage: tools/gpu_final.sh <tag>
TAG=${1:-fin}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
( time timeout 1500 python bench.py --steps 2 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err | cut -c1-300
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 3 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_all.csv \
    python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 --all-snps > gpurun_out/${TAG}_launches_all.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_nipt.csv \
    python tools/prof_sweep.py --K 2048 --jobs 148 --its 6 --ff 0.1 --coverage 0.5 > gpurun_out/${TAG}_launches_nipt.log 2>&1
for k in launches launches_all launches_nipt; do python tools/launch_summary.py gpurun_out/${TAG}_$k.csv > gpurun_out/${TAG}_${k}_summary.txt; done
timeout 600 python tools/bench_haploid.py 296 > gpurun_out/${TAG}_haploid.json 2> gpurun_out/${TAG}_haploid.err
for W in chr20_2Mb_1x_K512 nipt_2Mb_0.5x_K2048 chr20_2Mb_1x_K8192_20k; do
  ( timeout 1200 python bench.py --steps 2 --warmup 3 --workload $W ) > gpurun_out/${TAG}_wl_${W}.json 2> gpurun_out/${TAG}_wl_${W}.err
  echo "$W exit $?"; tail -3 gpurun_out/${TAG}_wl_${W}.err | cut -c1-300
done
ls -la gpurun_out
