#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench line (+ reference arm), ncu launch list, ncu --set full captures of the sweep kernel.
# usage: tools/gpu_round.sh [tag]   (outputs under gpurun_out/<tag>_*)
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
( time timeout 1200 python bench.py --steps 2 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 3 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_all.csv \
    python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 --all-snps > gpurun_out/${TAG}_launches_all.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 9 -c 1 -f -o gpurun_out/${TAG}_sweep \
    python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 9 -c 1 -f -o gpurun_out/${TAG}_sweep_all \
    python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 --all-snps > gpurun_out/${TAG}_ncu_full_all.log 2>&1
ls -la gpurun_out | grep ${TAG}
