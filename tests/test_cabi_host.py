"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/quilt_b200.h declares, argument
validation fails loudly without a device, the sample sharding rule, the schedule and the gloo broadcast."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from quilt_b200 import cabi, dist, schedule, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(prefix, header=("include", "quilt_b200.h")):
    hdr = open(os.path.join(ROOT, *header)).read()
    return sorted(set(re.findall(r"\b(" + prefix + r"_[A-Za-z0-9_]+)\s*\(", hdr)))


def test_gpu_library_exports_every_declared_symbol():
    from quilt_b200 import api, build

    so = build.build()  # nvcc cross-compiles sm_100a without a GPU
    lib = C.CDLL(so)
    names = _declared("quilt_gpu")
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/quilt_b200.h but not exported by libquiltgpu.so"
    api.GpuLib(so)  # the ctypes binding attaches argtypes to all of them


def test_oracle_library_exports_every_declared_symbol(oracle):
    names = _declared("quilt_oracle", ("oracle", "quilt_oracle.h"))
    assert len(names) >= 6
    for n in names:
        assert hasattr(oracle.lib, n), n


def test_product_header_declares_no_test_infrastructure():
    """the oracle's / compiled reference's entry points live under oracle/, not in the product's public header"""
    assert _declared("quilt_oracle") == [] and _declared("quilt_ref") == []


def test_no_cpu_fallback_without_device(small_world, small_reads):
    """without a CUDA device the product path errors out (status 4 / 1), it never computes on the host"""
    from quilt_b200 import api

    lib = api.GpuLib()
    if lib.device_count() > 0:
        pytest.skip("a GPU is visible: the no-device path cannot be exercised here")
    call = synth.make_call(small_world, small_reads.common, 1, K=32)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        lib.gibbs(call)


def test_product_sources_never_touch_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu arm may use oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "quilt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle_py|dlopen\([^)]*oracle|CDLL\([^)]*oracle|quilt_oracle_\w+\s*\(", txt, re.M), f


def test_struct_layout_matches_header():
    """ctypes mirrors must have the C layout: check sizes against a tiny C program compiled from the header"""
    src = r'''
#include "quilt_b200.h"
#include <stdio.h>
#include <stddef.h>
int main(){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(QuiltPanel), sizeof(QuiltReads), sizeof(QuiltGibbsArgs), sizeof(QuiltGibbsOut),
  offsetof(QuiltGibbsArgs, flags), offsetof(QuiltGibbsOut, read_category)); return 0; }
'''
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    want = [C.sizeof(cabi.QuiltPanel), C.sizeof(cabi.QuiltReads), C.sizeof(cabi.QuiltGibbsArgs), C.sizeof(cabi.QuiltGibbsOut),
            cabi.QuiltGibbsArgs.flags.offset, cabi.QuiltGibbsOut.read_category.offset]
    assert [int(x) for x in out] == want


def test_sample_range_is_a_partition():
    for n in (0, 1, 7, 256, 4096):
        for w in (1, 2, 3, 8):
            r = [dist.sample_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_schedule_matches_quilt2_defaults(small_world, small_reads):
    calls = schedule.sample_calls(small_world, small_reads, 3, K=64)
    assert len(calls) == 32  # 8 chains x (3 + 1)
    it = [bool(c.flags & cabi.F_GIBBS_INITIALIZE_ITERATIVELY) for c in calls]
    rc = [bool(c.flags & cabi.F_MAKE_EMATREAD_RARE_COMMON) for c in calls]
    assert sum(it) == 8 and sum(rc) == 8
    for c in calls:
        assert c.n_gibbs_burn_in_its == 20 and c.n_gibbs_sample_its == 1 and tuple(c.block_gibbs_iterations) == (3, 6, 9)
        assert (c.nGrids == small_world.nGrids_all) == bool(c.flags & cabi.F_MAKE_EMATREAD_RARE_COMMON)


_GLOO = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as td
from quilt_b200 import dist, synth
rank, local, world, dev = dist.init("gloo")
w = synth.make_world(42, K_full=200, nSNPs=640, region_bp=60_000, all_snps_factor=3) if rank == 0 else None
w = dist.broadcast_world(w, src=0)
ref = synth.make_world(42, K_full=200, nSNPs=640, region_bp=60_000, all_snps_factor=3)
for name in ("hapMatcherR", "distinctHapsB", "distinctHapsIE", "special_matrix", "special_helper", "snp_is_common", "common_snp_index", "rare_hap_offsets", "rare_hap_snps"):
    assert np.array_equal(getattr(w.panel, name), getattr(ref.panel, name)), name
assert np.array_equal(w.bits_all, ref.bits_all) and np.array_equal(w.transMatRate_all, ref.transMatRate_all)
lo, hi = dist.sample_range(7, rank, world)
t = dist.max_over_ranks(float(rank + 1))
s = dist.sum_over_ranks(float(hi - lo))
assert t == world and s == 7.0, (t, s)
# INFO counters: every rank holds the sums over its own samples; the writer needs the totals (quilt.R:957-961, writers.R:38-47)
rng = np.random.default_rng(100 + rank)
mine = {"infoCount": rng.random((640, 2)), "afCount": rng.random(640), "hweCount": rng.integers(0, 5, (640, 3)).astype(float), "alleleCount": rng.random((640, 2))}
tot = dist.allreduce_info_counts(mine)
want = {k: sum(np.asarray({"infoCount": np.random.default_rng(100 + r).random((640, 2))}["infoCount"]) for r in range(world)) for k in ("infoCount",)}
assert np.allclose(tot["infoCount"], want["infoCount"]) and tot["hweCount"].shape == (640, 3)
sc = dist.info_scores(tot, N=7)
assert sc["info"].shape == (640,) and np.all(sc["info"] >= 0)
dist.barrier()
print("ok", rank)
'''


def test_gloo_world_size_2_broadcast_and_reductions(tmp_path):
    """the N > 1 plumbing (broadcast of the prepared reference, max/sum over ranks) on CPU with gloo"""
    script = tmp_path / "g.py"
    script.write_text(_GLOO % {"root": ROOT})
    port = 29500 + (os.getpid() % 500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the B200 arm) on the tiny workload"""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["value"] > 0
    from oracle import ref_py

    assert d["cpu_baseline"]["kind"] == ("reference" if ref_py.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_rcpp_shim_compiles_with_the_gpu_back_end():
    """the shim in its PRODUCT configuration (back end quilt_gpu_gibbs) compiles against the stand-in headers; its executed
    CPU configuration (oracle back end) is tests/test_shim_executes.py"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I" + os.path.join(root, "oracle", "refshim"), "-I" + os.path.join(root, "include"),
                        os.path.join(root, "shim", "quilt_gpu_shim.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_roofline_traffic_is_keyed_by_workload():
    """bench.py reports the ncu DRAM traffic only for a workload that has a capture under profiles/ (null otherwise)"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t = bench.ncu_traffic("chr20_2Mb_1x_K4096")
    assert t is not None and t["dram_bytes_per_launch"] > 1e10
    # the kernel moves fewer bytes than the dense-emission algorithmic figure of SURVEY.md section 8(d)
    for kind in ("common", "allsnp"):
        assert t[kind]["dram_bytes_per_launch"] < t[kind]["algorithmic_bytes_of_captured_launch"]
    assert bench.ncu_traffic("chr20_2Mb_1x_K512") is None
    assert bench.ncu_traffic("no such workload") is None
