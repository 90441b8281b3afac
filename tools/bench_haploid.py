"""Throughput of the full-panel haploid pass (row a10): n passes over the 5008-haplotype x 1000-grid panel in one batch call, device time
from CUDA events, achieved GB/s against the algorithmic bytes (K_full * nGrids * (2 symbol bytes + 8 alphaHat written + 8 read) per pass),
and the compiled reference (oracle/_ref) timed on the host cores beside it.   python tools/bench_haploid.py [n_passes]"""
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from quilt_b200 import api, cabi, synth  # noqa: E402
from test_haploid_pass import make_gl, thinned_cols  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
w = synth.make_world(20260118, K_full=5008, nSNPs=32000, region_bp=3_000_000)
K, T, nS = w.panel.K_full, w.panel.nGrids, w.panel.nSNPs
cols = thinned_cols(T)
n_thin = int(cols.max()) + 1
gls = []
for s in range(8):
    sr = synth.make_sample_reads(w, 4000 + s, coverage=1.0, region_bp=3_000_000)
    H = np.random.default_rng(s).integers(1, 3, sr.common.nReads)
    gls += [make_gl(sr.common, nS, H, 1), make_gl(sr.common, nS, H, 2)]
lib = api.GpuLib()
fn = lib.lib.quilt_gpu_haploid_dosage_versus_refs_batch
fn.argtypes = [C.c_int32, C.POINTER(cabi.QuiltHaploidArgs), C.POINTER(cabi.QuiltHaploidOut)]
fn.restype = C.c_int
lib.lib.quilt_gpu_haploid_last_timing.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
args = (cabi.QuiltHaploidArgs * n)()
outs = (cabi.QuiltHaploidOut * n)()
ps = w.panel.c_struct()
tm = cabi.f64(w.transMatRate)
keep = []
flags = cabi.HF_RETURN_DOSAGE | cabi.HF_GET_BEST_HAPS  # production: dosage + best haplotypes (functions.R:1990-2000, return_gamma_t FALSE)
for i in range(n):
    gl = gls[i % len(gls)]
    a, o = args[i], outs[i]
    a.panel = C.pointer(ps)
    a.gl, a.transMatRate_t, a.gammaSmall_cols_to_get = cabi._ptr(gl, cabi._pd), cabi._ptr(tm, cabi._pd), cabi._ptr(cols, cabi._pi)
    a.n_thinned, a.K_top_matches, a.best_cap, a.min_emission_prob_normalization_threshold, a.flags = n_thin, 5, 32, 1e-100, flags
    d, c = np.zeros(nS), np.zeros(T)
    bh, bv, bc = np.zeros((n_thin, 32), np.int32), np.zeros((n_thin, 32)), np.zeros(n_thin, np.int32)
    keep += [d, c, bh, bv, bc]
    o.dosage, o.c, o.best_haps, o.best_haps_values, o.best_haps_count = cabi._ptr(d, cabi._pd), cabi._ptr(c, cabi._pd), cabi._ptr(bh, cabi._pi), cabi._ptr(bv, cabi._pd), cabi._ptr(bc, cabi._pi)
res = {}
for rep in range(3):
    t0 = time.perf_counter()
    rc = fn(n, args, outs)
    wall = time.perf_counter() - t0
    assert rc == 0, lib.last_error()
    ms, by = C.c_double(), C.c_double()
    lib.lib.quilt_gpu_haploid_last_timing(C.byref(ms), C.byref(by))
    res = {"passes": n, "kernel_ms": ms.value, "wall_ms": 1e3 * wall, "passes_per_s_device": n / (ms.value / 1e3), "passes_per_s_e2e": n / wall,
           "algorithmic_GB": by.value / 1e9, "achieved_GBps": by.value / 1e9 / (ms.value / 1e3)}
    print(res, file=sys.stderr)
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
res["hbm_peak_GBps"] = peak
res["roofline_frac"] = res["achieved_GBps"] / peak
# the compiled reference on the host cores
from oracle.ref_py import Ref  # noqa: E402

ref = Ref()
cores = len(os.sched_getaffinity(0))
fl = cabi.HF_RETURN_DOSAGE | cabi.HF_GET_BEST_HAPS


def one(i):
    t0 = time.perf_counter()
    r = ref.haploid_dosage(w.panel, gls[i % len(gls)], w.transMatRate, cols, flags=fl)
    return time.perf_counter() - t0, r


t0 = time.perf_counter()
with ThreadPoolExecutor(max_workers=cores) as ex:
    rr = list(ex.map(one, range(2 * cores)))
wall = time.perf_counter() - t0
res["cpu_reference"] = {"cores": cores, "passes": 2 * cores, "passes_per_s": 2 * cores / wall, "seconds_per_pass": float(np.mean([x[0] for x in rr]))}
res["speedup_device"] = res["passes_per_s_device"] / res["cpu_reference"]["passes_per_s"]
res["speedup_e2e"] = res["passes_per_s_e2e"] / res["cpu_reference"]["passes_per_s"]
# parity of pass 0 against the reference
d0 = np.ctypeslib.as_array(outs[0].dosage, shape=(nS,))
res["max_d_dosage_vs_reference"] = float(np.max(np.abs(d0 - rr[0][1]["dosage"])))
print(json.dumps(res))
