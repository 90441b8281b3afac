"""Regenerates tests/golden/*.npz: small seeded calls and the ORACLE's outputs for them.

These are regression pins, not reference-generated vectors: the reference cannot be built or run in this image
(it needs R / Rcpp / RcppArmadillo) and ships no golden vectors for this path, so parity stays "unpinned" at that level
(DESIGN.md).  What the fixtures pin: (a) the oracle itself across refactors, (b) the CUDA path against a frozen copy of
the oracle's answers, including inputs stored verbatim so the check does not depend on the synthetic generator.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle_py import Oracle  # noqa: E402
from quilt_b200 import cabi, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (world kwargs, reads kwargs, call kwargs)
    "diploid_common_iterative": (dict(seed=11, K_full=150, nSNPs=640, region_bp=60_000, all_snps_factor=3), dict(seed=12, coverage=1.5, region_bp=60_000),
                                 dict(seed=13, K=64, first_iteration=True)),
    "diploid_common_replayed": (dict(seed=11, K_full=150, nSNPs=640, region_bp=60_000, all_snps_factor=3), dict(seed=14, coverage=1.5, region_bp=60_000),
                                dict(seed=15, K=96, first_iteration=False, sort_haps=False)),
    "diploid_all_snps": (dict(seed=11, K_full=150, nSNPs=640, region_bp=60_000, all_snps_factor=3), dict(seed=16, coverage=1.0, region_bp=60_000),
                         dict(seed=17, K=64, all_snps=True)),
    "diploid_special_haps": (dict(seed=21, K_full=150, nSNPs=640, region_bp=60_000, nMaxDH=5, n_founders=30), dict(seed=22, coverage=1.5, region_bp=60_000),
                             dict(seed=23, K=64, first_iteration=False)),
}

PANEL_FIELDS = ["hapMatcherR", "distinctHapsB", "distinctHapsIE", "special_matrix", "special_helper", "snp_is_common", "common_snp_index",
                "rare_hap_offsets", "rare_hap_snps"]
CALL_FIELDS = ["which_haps_to_use", "transMatRate_tc_H", "L_grid", "smooth_cm", "H0", "runif_reads", "runif_block", "runif_shard"]
SCALARS = ["nGrids", "nSNPs", "ff", "n_gibbs_burn_in_its", "n_gibbs_sample_its", "first_read_for_gibbs_initialization", "maxDifferenceBetweenReads",
           "Jmax", "class_sum_cutoff", "shuffle_bin_radius", "block_gibbs_quantile_prob", "flags"]


def build_case(wk, rk, ck):
    w = synth.make_world(**wk)
    sr = synth.make_sample_reads(w, **rk)
    all_snps = ck.get("all_snps", False)
    call = synth.make_call(w, sr.all if all_snps else sr.common, **ck)
    return w, call


def save(name, w, call, res):
    d = {}
    for f in PANEL_FIELDS:
        v = getattr(w.panel, f)
        if v is not None:
            d["panel_" + f] = v
    d["panel_ref_error"] = np.float64(w.panel.ref_error)
    d["panel_nSNPs"] = np.int64(w.panel.nSNPs)
    for f in ("offsets", "u", "bq", "wif0"):
        d["reads_" + f] = getattr(call.reads, f)
    for f in CALL_FIELDS:
        d["call_" + f] = np.asarray(getattr(call, f))
    for f in SCALARS:
        d["call_" + f] = np.asarray(getattr(call, f))
    d["call_block_gibbs_iterations"] = np.asarray(call.block_gibbs_iterations, dtype=np.int32)
    for f in ("hapProbs_t", "genProbsM_t", "genProbsF_t", "H", "H_class", "per_it_likelihoods", "read_category"):
        d["out_" + f] = getattr(res, f)
    d["out_underflow_problem"] = np.asarray(res.underflow_problem)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def load(path):
    """-> (GibbsCall, dict of expected outputs)"""
    z = np.load(path)
    p = {f: (z["panel_" + f] if "panel_" + f in z.files else None) for f in PANEL_FIELDS}
    panel = cabi.Panel(ref_error=float(z["panel_ref_error"]), nSNPs=int(z["panel_nSNPs"]), **p)
    reads = cabi.Reads(offsets=z["reads_offsets"], u=z["reads_u"], bq=z["reads_bq"], wif0=z["reads_wif0"])
    kw = {f: z["call_" + f] for f in CALL_FIELDS}
    for f in SCALARS:
        v = z["call_" + f]
        kw[f] = float(v) if v.dtype.kind == "f" else int(v)
    call = cabi.GibbsCall(panel=panel, reads=reads, block_gibbs_iterations=tuple(int(x) for x in z["call_block_gibbs_iterations"]), **kw)
    exp = {f[4:]: z[f] for f in z.files if f.startswith("out_")}
    return call, exp


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    orc = Oracle()
    for name, (wk, rk, ck) in CASES.items():
        w, call = build_case(wk, rk, ck)
        res = orc.gibbs(call)
        save(name, w, call, res)
        print(name, "reads", call.reads.nReads, "K", call.K, "T", call.nGrids, os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")
