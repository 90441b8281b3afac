"""Regenerates tests/golden/*.npz: small seeded calls and the REFERENCE's outputs for them.

The outputs come from oracle/_ref/libquiltref.so — the unmodified reference sources
(/root/reference/QUILT/src/{copied-from-stitch, gibbs-small, gibbs-nipt, gibbs-nipt-block}.cpp) compiled against the
header-only RcppArmadillo stand-in in oracle/refshim/ — i.e. they are reference-generated vectors.  The fixtures pin
(a) the oracle's restatement (tests/test_golden.py::test_oracle_reproduces_golden, bit for bit), (b) the CUDA path
(::test_gpu_matches_golden) and (c) the stand-in build itself across refactors; inputs are stored verbatim so the
checks do not depend on the synthetic generator, and /root/reference is not needed to run them.

    python tools/make_golden.py          # needs /root/reference (or a prebuilt oracle/_ref/libquiltref.so)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_py import Ref  # noqa: E402
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
    "nipt_common_iterative": (dict(seed=11, K_full=150, nSNPs=640, region_bp=60_000, all_snps_factor=3), dict(seed=18, coverage=1.5, region_bp=60_000),
                              dict(seed=19, K=64, first_iteration=True, ff=0.1)),
    "nipt_all_snps": (dict(seed=11, K_full=150, nSNPs=640, region_bp=60_000, all_snps_factor=3), dict(seed=24, coverage=1.0, region_bp=60_000),
                      dict(seed=25, K=96, all_snps=True, ff=0.2)),
}

PANEL_FIELDS = ["hapMatcherR", "distinctHapsB", "distinctHapsIE", "special_matrix", "special_helper", "snp_is_common", "common_snp_index",
                "rare_hap_offsets", "rare_hap_snps"]
CALL_FIELDS = ["which_haps_to_use", "transMatRate_tc_H", "L_grid", "smooth_cm", "H0", "runif_reads", "runif_block", "runif_shard"]
OPTIONAL_CALL_FIELDS = ["runif_H_class"]
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
    for f in OPTIONAL_CALL_FIELDS:
        if getattr(call, f) is not None:
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
    for f in OPTIONAL_CALL_FIELDS:
        if "call_" + f in z.files:
            kw[f] = z["call_" + f]
    for f in SCALARS:
        v = z["call_" + f]
        kw[f] = float(v) if v.dtype.kind == "f" else int(v)
    call = cabi.GibbsCall(panel=panel, reads=reads, block_gibbs_iterations=tuple(int(x) for x in z["call_block_gibbs_iterations"]), **kw)
    exp = {f[4:]: z[f] for f in z.files if f.startswith("out_")}
    return call, exp


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref = Ref()
    for name, (wk, rk, ck) in CASES.items():
        w, call = build_case(wk, rk, ck)
        res = ref.gibbs(call)
        save(name, w, call, res)
        print(name, "reads", call.reads.nReads, "K", call.K, "T", call.nGrids, os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")
