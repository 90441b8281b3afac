"""GPU parity on the edge cases of the path: tiny / ragged shapes, quirks the reference's code paths depend on.

Every case runs the full production-shaped call (or a short one) through the C ABI and the oracle on identical inputs."""
import numpy as np
import pytest

from quilt_b200 import cabi, synth
from test_gpu_parity import _compare, _run_both

pytestmark = pytest.mark.gpu


def _tiny_world(seed, nSNPs, K_full=120, **kw):
    return synth.make_world(seed, K_full=K_full, nSNPs=nSNPs, region_bp=max(3000, nSNPs * 94), **kw)


@pytest.mark.parametrize("nSNPs", [33, 64, 70])
@pytest.mark.parametrize("K", [33, 100])
def test_two_or_three_grids(gpu, oracle, nSNPs, K):
    """T = 2 / 3 grids (partial last grid), K not a multiple of 32"""
    w = _tiny_world(500 + nSNPs, nSNPs)
    sr = synth.make_sample_reads(w, 1, coverage=4.0, region_bp=max(3000, nSNPs * 94))
    for first in (False, True):
        call = synth.make_call(w, sr.common, 41, K=K, first_iteration=first, n_burn_in=4, n_sample=1, block_its=(1,))
        g, o = _run_both(gpu, oracle, call)
        _compare(f"T={w.nGrids} K={K} iterative={first}", g, o)


def test_sparse_reads_many_empty_grids(gpu, oracle, small_world):
    """0.05x coverage: most grids have no read (forward skips the emission, backward takes the no-read branch)"""
    sr = synth.make_sample_reads(small_world, 11, coverage=0.05, region_bp=300_000)
    assert sr.common.nReads < small_world.nGrids
    for first in (False, True):
        call = synth.make_call(small_world, sr.common, 42, K=300, first_iteration=first)
        g, o = _run_both(gpu, oracle, call)
        _compare(f"sparse reads iterative={first}", g, o)


def test_single_read(gpu, oracle, small_world):
    sr = synth.make_sample_reads(small_world, 12, n_reads=1, region_bp=300_000)
    assert sr.common.nReads == 1
    call = synth.make_call(small_world, sr.common, 43, K=64, first_iteration=False, n_burn_in=2, n_sample=1, block_its=(0,))
    g, o = _run_both(gpu, oracle, call)
    _compare("single read", g, o)


def test_zero_base_quality_and_Jmax_truncation(gpu, oracle, small_world, small_reads):
    """bq == 0 leaves (pR, pA) at the previous SNP's / read's values (gibbs-small.cpp:172-181); Jmax cuts long reads"""
    r = small_reads.common
    bq = r.bq.copy()
    bq[::7] = 0
    reads = cabi.Reads(offsets=r.offsets, u=r.u, bq=bq, wif0=r.wif0)
    call = synth.make_call(small_world, reads, 44, K=200, first_iteration=False, n_burn_in=3, n_sample=1, block_its=(1,), Jmax=1)
    g, o = _run_both(gpu, oracle, call)
    _compare("bq = 0 / Jmax = 1", g, o)
    eg, cg = gpu.make_eMatRead_t(call)
    eo, co = oracle.make_eMatRead_t(call)
    assert np.array_equal(cg, co)
    assert np.max(np.abs(eg - eo) / eo) < 1e-14


def test_emission_floor(gpu, oracle, small_world):
    """maxDifferenceBetweenReads = 10: the floor 1 / maxDifferenceBetweenReads is hit by most mismatching haplotypes"""
    sr = synth.make_sample_reads(small_world, 13, coverage=2.0, region_bp=300_000)
    call = synth.make_call(small_world, sr.common, 45, K=256, first_iteration=False, maxDifferenceBetweenReads=10.0)
    g, o = _run_both(gpu, oracle, call)
    _compare("emission floor", g, o)


def test_high_coverage_grids_exceed_one_staging_buffer(gpu, oracle, small_world):
    """30x: > 48 reads per grid and > 704 table entries, i.e. the chunked staging path of the sweep kernel"""
    sr = synth.make_sample_reads(small_world, 14, coverage=30.0, region_bp=300_000)
    call = synth.make_call(small_world, sr.all, 46, K=128, all_snps=True, n_burn_in=3, n_sample=1, block_its=(1,))
    g, o = _run_both(gpu, oracle, call)
    _compare("30x all-SNP", g, o)


def test_no_rescale_and_no_record(gpu, oracle, small_world, small_reads):
    """rescale_eMatRead_t = FALSE and record_read_set = FALSE (non-production flag values of param_list).  Without
    record_read_set the reference's H_class is an empty vector, which its block-Gibbs code would index out of bounds:
    the combination is only defined without block episodes."""
    flags = cabi.FLAGS_QUILT2_DIPLOID & ~cabi.F_RESCALE_EMATREAD & ~cabi.F_RECORD_READ_SET
    call = synth.make_call(small_world, small_reads.common, 47, K=200, first_iteration=False, flags=flags, n_burn_in=4, n_sample=1, block_its=())
    g, o = _run_both(gpu, oracle, call)
    _compare("no rescale / no record", g, o)


def test_long_region_many_grids(gpu, oracle):
    """T = 2500 common / 7500 all-SNP grids and ~27k reads (a 7.5 Mb window at 1x): larger offsets than the benchmark shape"""
    w = synth.make_world(777, K_full=700, nSNPs=80_000, region_bp=7_500_000, all_snps_factor=3)
    sr = synth.make_sample_reads(w, 778, coverage=0.55, region_bp=7_500_000)
    call = synth.make_call(w, sr.common, 48, K=512, first_iteration=True)
    g, o = _run_both(gpu, oracle, call)
    _compare(f"long region common T={w.nGrids} R={sr.common.nReads}", g, o, state=False)
    call = synth.make_call(w, sr.all, 49, K=256, all_snps=True, sort_haps=False, n_burn_in=5, n_sample=1, block_its=(2,))
    g, o = _run_both(gpu, oracle, call)
    _compare(f"long region all-SNP T={w.nGrids_all} R={sr.all.nReads}", g, o, state=False)


def test_section_timing_switch(gpu, small_world, small_reads):
    """the per-section timing switch (reference: print_extra_timing_information, copied-from-stitch.cpp:31-45): every kernel of a
    Gibbs call shows up with its launch count and a positive device time; switched off, nothing is recorded"""
    call = synth.make_call(small_world, small_reads.common, 11, K=200, n_burn_in=4, n_sample=1, block_its=(2,))
    gpu.section_timing(True)
    n0 = gpu.kernel_launches()
    gpu.gibbs_batch([call])
    n = gpu.kernel_launches() - n0
    rep = gpu.section_report()
    rows = {ln.split()[0]: ln.split() for ln in rep.splitlines()[1:] if ln.strip()}
    assert int(rows["k_sweep"][1]) == 5 and float(rows["k_sweep"][2]) > 0
    assert "k_shard" in rows and "k_happrobs" in rows and "k_build_tables" in rows
    assert int(rows["TOTAL"][1]) == n
    gpu.section_timing(False)
    gpu.gibbs_batch([call])
    assert gpu.section_report().splitlines()[-1].split()[1] == "0"
