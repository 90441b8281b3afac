"""Small driver for ncu: a handful of production-shaped calls at the benchmark's K so one sweep launch can be captured."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quilt_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--K", type=int, default=4096)
ap.add_argument("--jobs", type=int, default=8)
ap.add_argument("--its", type=int, default=6)
ap.add_argument("--all-snps", action="store_true")
ap.add_argument("--iterative", action="store_true")
ap.add_argument("--nsnps", type=int, default=32000)
ap.add_argument("--ff", type=float, default=0.0, help="> 0: three-haplotype (NIPT) calls")
ap.add_argument("--coverage", type=float, default=1.0)
a = ap.parse_args()
w = synth.make_world(20260118, K_full=max(a.K + 100, 5008), nSNPs=a.nsnps, region_bp=int(3_000_000 * a.nsnps / 32000), all_snps_factor=3 if a.all_snps else 0)
calls = []
for j in range(a.jobs):
    sr = synth.make_sample_reads(w, 50 + j, coverage=a.coverage, region_bp=int(3_000_000 * a.nsnps / 32000), n_true_haps=3 if a.ff > 0 else 2,
                                 hap_probs=(0.5, 0.5 - a.ff / 2, a.ff / 2) if a.ff > 0 else None)
    calls.append(synth.make_call(w, sr.all if a.all_snps else sr.common, 70 + j, K=a.K, all_snps=a.all_snps, first_iteration=a.iterative,
                                 n_burn_in=a.its - 1, n_sample=1, block_its=(3,), ff=a.ff))
lib = api.GpuLib()
b = api.Batch(lib, calls)
sections = os.environ.get("QUILT_B200_SECTIONS") == "1"  # per-kernel device time of the second run (quilt_gpu_section_timing)
for i in range(2):
    if sections and i == 1:
        lib.section_timing(True)
    b.run()
    b.sync()
    print(b.timing(), file=sys.stderr)
if sections:
    print(lib.section_report())
    lib.section_timing(False)
b.free()
