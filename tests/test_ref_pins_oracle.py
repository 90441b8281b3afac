"""The oracle is pinned to the reference ITSELF: oracle/_ref/libquiltref.so is the unmodified
/root/reference/QUILT/src/{copied-from-stitch, gibbs-small, gibbs-nipt, gibbs-nipt-block}.cpp compiled against the
header-only RcppArmadillo stand-in (oracle/refshim/).  These CPU tests run the reference's own
rcpp_forwardBackwardGibbsNIPT (and component functions) next to the oracle's restatement on the same inputs and demand
identical read labels / H_class / read categories and bit-identical state and probabilities (both libraries are
built without FMA contraction, so "equal" means equal).
"""
import glob
import os
import sys

import numpy as np
import pytest

from quilt_b200 import cabi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden  # noqa: E402

FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


@pytest.fixture(scope="session")
def ref():
    from oracle import ref_py

    if not ref_py.available() and not ref_py.can_build():
        pytest.skip("oracle/_ref/libquiltref.so not built and /root/reference not present")
    return ref_py.Ref()


def _lk(x):
    return np.nan_to_num(x, nan=-7e300, posinf=1e300, neginf=-1e300)


def _same(r, o, nh=None, state=False):
    assert bool(r.underflow_problem) == bool(o.underflow_problem)
    assert np.array_equal(r.H, o.H)
    assert np.array_equal(r.H_class, o.H_class)
    assert np.array_equal(r.read_category, o.read_category)
    assert np.array_equal(r.H_sample_its, o.H_sample_its)
    for f in ("hapProbs_t", "genProbsM_t", "genProbsF_t"):
        assert np.array_equal(getattr(r, f), getattr(o, f)), f
    assert np.array_equal(_lk(r.per_it_likelihoods), _lk(o.per_it_likelihoods))
    if state:
        for h in range(nh):
            for f in ("alphaHat_t", "betaHat_t", "eMatGrid_t", "c"):
                assert np.array_equal(getattr(r, f)[h], getattr(o, f)[h]), (f, h)
        assert np.array_equal(r.eMatRead_t, o.eMatRead_t)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_reference_reproduces_golden(ref, path):
    """the committed fixtures are the REFERENCE's answers (tools/make_golden.py runs oracle/_ref)"""
    call, exp = make_golden.load(path)
    r = ref.gibbs(call)
    assert np.array_equal(r.H, exp["H"]) and np.array_equal(r.H_class, exp["H_class"]) and np.array_equal(r.read_category, exp["read_category"])
    for f in ("hapProbs_t", "genProbsM_t", "genProbsF_t"):
        assert np.array_equal(getattr(r, f), exp[f]), f
    assert np.array_equal(_lk(r.per_it_likelihoods), _lk(exp["per_it_likelihoods"]))


CASES = [
    # name, reads, make_call kwargs
    ("diploid_iterative_K200", "common", dict(seed=1, K=200, first_iteration=True)),
    ("diploid_replayed_unsorted_K600", "common", dict(seed=2, K=600, first_iteration=False, sort_haps=False)),
    ("diploid_all_snps_K300", "all", dict(seed=3, K=300, all_snps=True)),
    ("diploid_three_sampling_sweeps", "common", dict(seed=5, K=150, first_iteration=False, n_sample=3)),
    ("diploid_no_block_gibbs", "common", dict(seed=6, K=100, first_iteration=True, block_its=())),
    ("nipt_production_iterative_K200_ff10", "common", dict(seed=33, K=200, first_iteration=True, ff=0.1)),
    ("nipt_production_replayed_K400_ff20", "common", dict(seed=34, K=400, first_iteration=False, ff=0.2)),
    ("nipt_all_snps_K300_ff10", "all", dict(seed=35, K=300, all_snps=True, ff=0.1)),
    ("nipt_sweeps_only_ff25", "common", dict(seed=36, K=128, first_iteration=False, ff=0.25, n_burn_in=6, n_sample=2, block_its=())),
]


@pytest.mark.parametrize("name,which,kw", CASES, ids=[c[0] for c in CASES])
def test_reference_equals_oracle_whole_call(ref, oracle, small_world, small_reads, name, which, kw):
    reads = small_reads.all if which == "all" else small_reads.common
    call = synth.make_call(small_world, reads, **kw)
    call.flags |= cabi.F_RETURN_ALPHA | cabi.F_RETURN_EXTRA
    r, o = ref.gibbs(call), oracle.gibbs(call)
    assert int(np.sum(r.H != call.H0)) > 0, "the call must do something"
    _same(r, o, nh=2 if call.ff == 0 else 3, state=True)


def test_reference_equals_oracle_special_haplotypes(ref, oracle):
    """nMaxDH = 5 forces most (hap, grid) words through the special-matrix binary search"""
    w = synth.make_world(21, K_full=150, nSNPs=640, region_bp=60_000, nMaxDH=5, n_founders=30)
    sr = synth.make_sample_reads(w, 22, coverage=1.5, region_bp=60_000)
    call = synth.make_call(w, sr.common, 23, K=64, first_iteration=False)
    _same(ref.gibbs(call), oracle.gibbs(call))


@pytest.mark.parametrize("all_snps", [False, True])
def test_reference_equals_oracle_emissions(ref, oracle, small_world, small_reads, all_snps):
    reads = small_reads.all if all_snps else small_reads.common
    call = synth.make_call(small_world, reads, 9, K=257, all_snps=all_snps, first_iteration=False)
    call.flags &= ~cabi.F_DISABLE_READ_CATEGORY_USAGE
    e1, c1 = ref.make_eMatRead_t(call)
    e2, c2 = oracle.make_eMatRead_t(call)
    assert np.array_equal(e1, e2) and np.array_equal(c1, c2)
    assert set(np.unique(c1)) <= {0, 1, 2, 3} and len(np.unique(c1)) >= 2


def test_reference_equals_oracle_forward_backward(ref, oracle):
    rng = np.random.default_rng(4)
    K, T = 97, 60
    e = np.asfortranarray(rng.uniform(0.01, 1.0, size=(K, T)))
    sig = rng.uniform(0.9, 0.999, size=T - 1)
    tm = np.asfortranarray(np.stack([sig, 1 - sig]))
    for x, y in zip(ref.forward_backward(e, tm), oracle.forward_backward(e, tm)):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("all_snps", [False, True])
def test_reference_equals_oracle_panel_words(ref, oracle, small_world, all_snps):
    """rcpp_int_expand + rcpp_simple_binary_matrix_search (reference code) vs the oracle's word lookup"""
    rng = np.random.default_rng(8)
    which = rng.choice(small_world.panel.K_full, size=77, replace=False).astype(np.int32) + 1
    assert np.array_equal(ref.unpack_panel(small_world.panel, which, all_snps), oracle.unpack_panel(small_world.panel, which, all_snps))


def test_scripted_random_stream_detects_misalignment(ref, small_world, small_reads):
    """a shard pass that is not 'every pair' draws runif(n_blocks - 1), a data-dependent length the flat ABI does not
    script: the driver must refuse instead of replaying a misaligned stream"""
    call = synth.make_call(small_world, small_reads.common, 1, K=64, first_iteration=False)
    call.flags &= ~cabi.F_SHARD_CHECK_EVERY_PAIR
    with pytest.raises(RuntimeError):
        ref.gibbs(call)
