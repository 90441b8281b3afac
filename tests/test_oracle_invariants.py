"""CPU tests of the oracle (test infrastructure): the reference's own differential invariants, re-expressed on our
seeded fixtures.  The reference holds no golden vectors for this path — its unit tests assert
that two formulations of the same quantity agree — so these are the pins the oracle can be held to:

  * fast backward (1/K shortcut) == generic backward           (test-unit-gibbs-nipt-parts.R:177-264)
  * incremental forward step == column of the full forward      (test-unit-gibbs-nipt-parts.R:1-174)
  * emissions from the compressed panel == emissions from the inflated alleles (test-unit-gibbs-small.R:18-335)
  * read categories 1/2/3 (sparse updates) == dense updates     (test-unit-gibbs-diploid.R:56-129)
  * hapProbs/genProbs from bit-packed words == dense computation (test-unit-gibbs-small.R:339-490)
  * outputs are probabilities: GP columns sum to 1, DS in [0, 2] (check_quilt_output, test-drivers.R:1-89)
"""
import numpy as np
import pytest

from quilt_b200 import cabi, synth


def test_backward_fast_equals_generic(oracle):
    rng = np.random.default_rng(1)
    K, T = 61, 40
    e = np.asfortranarray(rng.uniform(0.05, 1.0, size=(K, T)))
    ghr = (rng.random(T) < 0.7).astype(np.uint8)
    e[:, ghr == 0] = 1.0  # grids without reads carry emission 1
    sig = rng.uniform(0.9, 0.999, size=T - 1)
    tm = np.asfortranarray(np.stack([sig, 1 - sig]))
    _, _, c = oracle.forward_backward(e, tm)
    b1, b2 = oracle.backward_pair(e, tm, c, ghr)
    np.testing.assert_allclose(b1, b2, rtol=1e-12)


@pytest.mark.parametrize("faster", [False, True])
def test_incremental_forward_equals_full(oracle, faster):
    rng = np.random.default_rng(2)
    K, T = 50, 30
    e = np.asfortranarray(rng.uniform(0.05, 1.0, size=(K, T)))
    sig = rng.uniform(0.9, 0.999, size=T - 1)
    tm = np.asfortranarray(np.stack([sig, 1 - sig]))
    a, _, c = oracle.forward_backward(e, tm)
    ghr = np.ones(T, dtype=np.uint8)
    for g in (1, 7, T - 1):
        a2 = a.copy(order="F")
        a2[:, g] = 0
        c2 = c.copy()
        c2[g] = 1.0  # the incremental step multiplies the stored c by its own normaliser
        a3, c3 = oracle.forward_one(e, tm, ghr, g, a2, c2, faster)
        np.testing.assert_allclose(a3[:, g], a[:, g], rtol=1e-12)
        np.testing.assert_allclose(c3[g], c[g], rtol=1e-12)


def _dense_eMatRead(world, reads, which, all_snps, maxdiff=1e10, eps=1e-3):
    """straight from the truth alleles, no compressed structures"""
    bits = (world.bits_all if all_snps else world.bits_common)[which - 1]  # [K, nSNPs]
    K = bits.shape[0]
    R = reads.nReads
    out = np.ones((K, R))
    has_carrier = bits.any(axis=0)
    is_common = world.panel.snp_is_common if all_snps else None
    pR, pA = 1.0, 1.0
    for r in range(R):
        for j in range(reads.offsets[r], reads.offsets[r + 1]):
            bq, u = int(reads.bq[j]), int(reads.u[j])
            if bq < 0:
                e = 10 ** (bq / 10)
                pR, pA = 1 - e, e / 3
            if bq > 0:
                e = 10 ** (-bq / 10)
                pR, pA = e / 3, 1 - e
            if all_snps and not is_common[u] and not has_carrier[u]:
                continue  # skipped when rescaling (gibbs-small.cpp:404-411)
            ek = np.where(bits[:, u] == 1, 1 - eps, eps)
            out[:, r] *= ek * pA + (1 - ek) * pR
        out[:, r] = np.maximum(out[:, r] / out[:, r].max(), 1 / maxdiff)
    return out


@pytest.mark.parametrize("all_snps", [False, True])
def test_emissions_compressed_equals_inflated(oracle, small_world, small_reads, all_snps):
    reads = small_reads.all if all_snps else small_reads.common
    call = synth.make_call(small_world, reads, 31, K=80, all_snps=all_snps, first_iteration=False)
    e, cat = oracle.make_eMatRead_t(call)
    ref = _dense_eMatRead(small_world, reads, call.which_haps_to_use, all_snps)
    np.testing.assert_allclose(e, ref, rtol=1e-12)
    assert e.max() <= 1.0 and e.min() >= 1e-10


def test_emissions_with_special_haplotypes(oracle):
    """nMaxDH small -> most haplotypes are 'special' and go through the binary search (gibbs-small.cpp:69-105)"""
    w = synth.make_world(5, K_full=300, nSNPs=640, region_bp=60_000, nMaxDH=6, n_founders=40)
    assert w.panel.special_matrix.shape[0] > 100
    sr = synth.make_sample_reads(w, 6, coverage=2.0, region_bp=60_000)
    call = synth.make_call(w, sr.common, 32, K=120, first_iteration=False)
    e, _ = oracle.make_eMatRead_t(call)
    words = oracle.unpack_panel(w.panel, call.which_haps_to_use)
    truth = synth.pack_bits(w.bits_common[call.which_haps_to_use - 1])
    # the reference's search returns 0 for a grid whose special group has exactly one row (sic, gibbs-small.cpp:76-78):
    # everywhere else the looked-up words are the true ones
    helper = w.panel.special_helper
    single = (helper[:, 1] - helper[:, 0] == 0) & (helper[:, 0] > 0)
    ok = ~single
    assert np.array_equal(words[:, ok], truth[:, ok])


def test_categories_do_not_change_results(oracle, small_world, small_reads):
    """sparse updates for read categories 2/3 vs every read handled densely (disable_read_category_usage)"""
    a = synth.make_call(small_world, small_reads.common, 33, K=90, first_iteration=False)
    b = synth.make_call(small_world, small_reads.common, 33, K=90, first_iteration=False)
    b.flags |= cabi.F_FORCE_RESET_READ_CATEGORY_0  # categories 2/3 -> 0, category 1 kept (still skipped in diploid mode)
    ra, rb = oracle.gibbs(a), oracle.gibbs(b)
    assert np.array_equal(ra.H, rb.H)
    np.testing.assert_allclose(ra.hapProbs_t, rb.hapProbs_t, atol=1e-9)


def test_outputs_are_probabilities(oracle, small_world, small_reads):
    for all_snps in (False, True):
        reads = small_reads.all if all_snps else small_reads.common
        r = oracle.gibbs(synth.make_call(small_world, reads, 34, K=100, all_snps=all_snps))
        assert not r.underflow_problem
        np.testing.assert_allclose(r.genProbsM_t.sum(axis=0), 1.0, atol=2e-3)  # check_quilt_output: GP sums to 1 +- 0.002
        assert r.hapProbs_t.min() >= 0 and r.hapProbs_t.max() <= 1
        assert np.all((r.dosage >= 0) & (r.dosage <= 2))
        assert set(np.unique(r.H)) <= {1, 2}


def test_hapProbs_match_dense_gamma(oracle, small_world, small_reads):
    """gamma -> hapProbs through the bit-packed words == gamma @ alleles"""
    call = synth.make_call(small_world, small_reads.common, 35, K=70, first_iteration=False)
    call.flags |= cabi.F_RETURN_ALPHA
    r = oracle.gibbs(call)
    bits = small_world.bits_common[call.which_haps_to_use - 1].astype(np.float64)
    eps = small_world.panel.ref_error
    for h in range(2):
        gamma = r.alphaHat_t[h] * r.betaHat_t[h] / r.c[h][None, :]  # [K, T]
        np.testing.assert_allclose(gamma.sum(axis=0), 1.0, rtol=1e-9)
        g_snp = np.repeat(gamma, 32, axis=1)[:, : small_world.nSNPs]
        hp = (g_snp * (bits * (1 - eps) + (1 - bits) * eps)).sum(axis=0)
        np.testing.assert_allclose(r.hapProbs_t[h], hp, atol=1e-10)


def test_imputation_is_informative(oracle):
    """end-to-end sanity in the style of the reference's acceptance tests: with all panel haplotypes selected and 8x reads
    the dosage tracks the truth (check_quilt_output tolerates |DS - truth| < 0.1 on its own simulations)"""
    w = synth.make_world(9, K_full=120, nSNPs=960, region_bp=90_000, n_founders=8, switch_rate=5e-4)
    sr = synth.make_sample_reads(w, 10, coverage=8.0, region_bp=90_000, sample_switch_rate=0.0)
    call = synth.make_call(w, sr.common, 36, K=120)
    r = oracle.gibbs(call)
    truth = sr.truth_haps.sum(axis=0)
    r2 = np.corrcoef(r.dosage, truth)[0, 1] ** 2
    assert r2 > 0.9, r2
