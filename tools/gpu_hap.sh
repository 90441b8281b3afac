#!/bin/bash
TAG=${1:-h1}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_haploid_pass.py -m gpu -x -q -s 2>&1 | tail -25 ) | tee gpurun_out/${TAG}_hap.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_haploid_pass.py -m gpu -x -q -k "special or equals_reference" 2>&1 | tail -8 | tee gpurun_out/${TAG}_hap_memcheck.log
# launch list of one common-SNP wave and one all-SNP wave (side kernels after the make_eG / assemble_all rewrite)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_all.csv python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 --all-snps > gpurun_out/${TAG}_launches_all.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_launch_summary_common.txt; python tools/launch_summary.py gpurun_out/${TAG}_launches_all.csv | tee gpurun_out/${TAG}_launch_summary_allsnp.txt
