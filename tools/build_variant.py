"""Build an EXPERIMENT variant of the library: python tools/build_variant.py <tag> -DQB_CLK=1 ...
-> quilt_b200/libquiltgpu_<tag>.so (git-ignored; load it with QUILT_B200_LIB=<path>).  The product build is quilt_b200/build.py."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quilt_b200 import build as qb  # noqa: E402

tag, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(os.path.dirname(qb.SO), f"libquiltgpu_{tag}.so")
cmd = [qb.nvcc()] + qb.NVCC_FLAGS + extra + ["-o", out] + [os.path.join(qb.CSRC, s) for s in qb.SOURCES]
r = subprocess.run(cmd, capture_output=True, text=True)
open(out + ".log", "w").write(r.stdout + r.stderr)
print(out if r.returncode == 0 else (r.stdout + r.stderr)[-3000:])
sys.exit(r.returncode)
