"""The Rcpp shim (shim/quilt_gpu_shim.cpp) is compiled against the stand-in headers, LINKED and EXECUTED on the CPU:
mock R objects in (the 63 arguments, built the way functions.R:2566-2678 builds them) -> the shim -> C ABI (the oracle as
back end) -> named R list out, compared field by field with the compiled reference run on the same objects with the same
random generator — including the position the generator is left at (NIPT's data-dependent draws, underflow early return).
"""
import ctypes as C
import os

import numpy as np
import pytest

from quilt_b200 import cabi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libshimharness.so")


@pytest.fixture(scope="module")
def harness():
    from oracle import ref_py

    if ref_py.can_build():
        ref_py.build()
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libshimharness.so not built and /root/reference not present")
    lib = C.CDLL(SO)
    lib.shim_harness_run.argtypes = [C.POINTER(cabi.QuiltGibbsArgs), C.c_uint64, C.c_int, C.POINTER(cabi.QuiltGibbsOut), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_int32)]
    lib.shim_harness_run.restype = C.c_int
    lib.shim_harness_last_error.restype = C.c_char_p

    def run(call, seed, via):
        a, o = cabi.QuiltGibbsArgs(), cabi.QuiltGibbsOut()
        call.fill(a)
        res = cabi.alloc_out(call, o)
        pos, nel = C.c_int64(), C.c_int32()
        rc = lib.shim_harness_run(C.byref(a), seed, via, C.byref(o), C.byref(pos), C.byref(nel))
        if rc != 0:
            raise RuntimeError(lib.shim_harness_last_error().decode())
        res.underflow_problem = bool(o.underflow_problem)
        return res, pos.value, nel.value

    return run


CASES = [
    ("diploid_iterative", "common", dict(seed=1, K=120, first_iteration=True)),
    ("diploid_replayed", "common", dict(seed=2, K=200, first_iteration=False, sort_haps=False)),
    ("diploid_all_snps", "all", dict(seed=3, K=100, all_snps=True)),
    ("nipt_production", "common", dict(seed=33, K=120, first_iteration=True, ff=0.1)),
    ("nipt_all_snps", "all", dict(seed=35, K=100, all_snps=True, ff=0.2)),
    ("nipt_no_block_gibbs", "common", dict(seed=36, K=64, first_iteration=False, ff=0.25, n_burn_in=4, n_sample=2, block_its=())),
]


@pytest.mark.parametrize("name,which,kw", CASES, ids=[c[0] for c in CASES])
def test_shim_equals_reference(harness, small_world, small_reads, name, which, kw):
    reads = small_reads.all if which == "all" else small_reads.common
    call = synth.make_call(small_world, reads, **kw)
    r, pos_r, nel_r = harness(call, 12345, 0)   # the reference's rcpp_forwardBackwardGibbsNIPT
    s, pos_s, nel_s = harness(call, 12345, 1)   # the shim, oracle back end
    assert not r.underflow_problem and not s.underflow_problem
    assert nel_r == nel_s, "the returned list has a different number of elements"
    assert np.array_equal(r.H, s.H) and np.array_equal(r.H_class, s.H_class)
    assert np.array_equal(r.H_sample_its, s.H_sample_its), "double_list_of_ending_read_labels differs"
    for f in ("hapProbs_t", "genProbsM_t", "genProbsF_t"):
        assert np.array_equal(getattr(r, f), getattr(s, f)), f
    lk = lambda x: np.nan_to_num(x, nan=-7e300, posinf=1e300, neginf=-1e300)  # noqa: E731
    assert np.array_equal(lk(r.per_it_likelihoods), lk(s.per_it_likelihoods))
    assert pos_r == pos_s, f"R's random stream is left at {pos_s} by the shim, at {pos_r} by the reference"
    assert int(np.sum(r.H != call.H0)) > 0


def test_shim_stream_position_after_underflow(harness, small_world):
    """underflow early return (gibbs-nipt.cpp:2959-2969): the reference stops drawing; the shim must rewind to the same position"""
    sr = synth.make_sample_reads(small_world, 9, coverage=60.0, region_bp=300_000)
    call = synth.make_call(small_world, sr.common, 26, K=100, first_iteration=False, maxDifferenceBetweenReads=1e300)
    r, pos_r, nel_r = harness(call, 777, 0)
    s, pos_s, nel_s = harness(call, 777, 1)
    assert r.underflow_problem and s.underflow_problem
    assert nel_r == nel_s == 1   # list(underflow_problem = TRUE)
    assert pos_r == pos_s


def test_shim_rejects_unsupported_arguments(harness, small_world, small_reads):
    """a shard pass that is not 'every pair' is outside the accelerated space: Rcpp::stop, not a silent different computation"""
    call = synth.make_call(small_world, small_reads.common, 1, K=64, first_iteration=False)
    call.flags &= ~cabi.F_SHARD_CHECK_EVERY_PAIR
    with pytest.raises(RuntimeError, match="shard_check_every_pair"):
        harness(call, 1, 1)
