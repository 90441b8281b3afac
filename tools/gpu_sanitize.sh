#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on the shared-memory heavy kernels) over small GPU parity tests
TAG=${1:-san}
mkdir -p gpurun_out
SEL='test_gibbs_nipt_block_short or test_gibbs_both_sweep_instances or test_gibbs_shard or test_gpu_pass_equals_reference or test_gpu_pass_special or test_two_or_three_grids or test_section_timing or test_gpu_select_equals_oracle or test_gpu_ingest_equals_oracle or test_gpu_vcf_column'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/${TAG}_memcheck.log
tail -6 gpurun_out/${TAG}_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "test_gibbs_nipt_block_short or test_gibbs_both_sweep_instances or test_gpu_pass_special_haplotypes" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/${TAG}_racecheck.log
tail -6 gpurun_out/${TAG}_racecheck.log
