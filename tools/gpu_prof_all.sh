#!/bin/bash
TAG=${1:-pa}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 9 -c 1 -f -o gpurun_out/${TAG}_sweep_all \
    python tools/prof_sweep.py --K 4096 --jobs 148 --its 6 --all-snps > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
python - <<'PY'
import sys
sys.path.insert(0,'.')
from quilt_b200 import synth
import numpy as np
w = synth.make_world(20260118, K_full=5008, nSNPs=32000, region_bp=3_000_000, all_snps_factor=3)
sr = synth.make_sample_reads(w, 50, coverage=1.0, region_bp=3_000_000)
for name, r in (("common", sr.common), ("all", sr.all)):
    cnt = np.diff(r.offsets)
    print(name, "reads", r.nReads, "SNPs/read mean", cnt.mean(), "hist", np.bincount(np.minimum(cnt, 12)))
PY
