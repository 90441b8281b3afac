#!/bin/bash
# final build: all GPU tests, smoke, bench line + reference arm, section-timing report of one wave, ncu --set full of both sweep instances
TAG=${1:-fin2}
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
QUILT_B200_SECTIONS=1 timeout 600 python tools/prof_sweep.py --K 4096 --jobs 148 --its 21 > gpurun_out/${TAG}_sections_common.txt 2>&1
QUILT_B200_SECTIONS=1 timeout 600 python tools/prof_sweep.py --K 4096 --jobs 148 --its 21 --all-snps > gpurun_out/${TAG}_sections_allsnp.txt 2>&1
bash tools/gpu_ncu_sweep.sh ${TAG}
ls -la gpurun_out
