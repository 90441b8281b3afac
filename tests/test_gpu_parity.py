"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): integer outputs (labels H, H_class, read categories, allele words) bit-exact;
floating point within the tolerance written next to each assert.  K-long sums are reduced in a different order on
the GPU (tree vs Armadillo's two accumulators), so state arrays agree to ~1e-12 relative, not bitwise.
"""
import numpy as np
import pytest

from quilt_b200 import cabi, synth

pytestmark = pytest.mark.gpu

RTOL_STATE = 1e-9   # alpha / beta / c / eMatGrid after many sweeps (sum-order noise accumulates ~1e-13)
ATOL_DS = 1e-4      # north_star: DS / GP within 1e-4


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def _compare(tag, g, o, state=True):
    """print + assert the whole result of one call"""
    msgs = []
    assert g.underflow_problem == o.underflow_problem, f"{tag}: underflow {g.underflow_problem} vs {o.underflow_problem}"
    nH = int(np.sum(g.H != o.H))
    nC = int(np.sum(g.H_class != o.H_class))
    nK = int(np.sum(g.read_category != o.read_category))
    dhp = float(np.max(np.abs(g.hapProbs_t - o.hapProbs_t)))
    dgm = float(np.max(np.abs(g.genProbsM_t - o.genProbsM_t)))
    dgf = float(np.max(np.abs(g.genProbsF_t - o.genProbsF_t)))
    msgs.append(f"H mismatches {nH}/{g.H.size}, H_class {nC}, category {nK}, |dhap| {dhp:.3e} |dgenM| {dgm:.3e} |dgenF| {dgf:.3e}")
    if state and g.alphaHat_t is not None:
        for h in range(len(g.alphaHat_t)):
            msgs.append(
                f"  hap{h}: alpha rel {_rel(g.alphaHat_t[h], o.alphaHat_t[h]):.3e} beta rel {_rel(g.betaHat_t[h], o.betaHat_t[h]):.3e} "
                f"eG rel {_rel(g.eMatGrid_t[h], o.eMatGrid_t[h]):.3e} c rel {_rel(g.c[h], o.c[h]):.3e}"
            )
    fin = np.isfinite(o.per_it_likelihoods)
    same_inf = np.array_equal(fin, np.isfinite(g.per_it_likelihoods))
    dl = _rel(g.per_it_likelihoods[fin], o.per_it_likelihoods[fin]) if same_inf else np.inf
    msgs.append(f"  per_it_likelihoods rel {dl:.3e} (inf pattern equal: {same_inf})")
    print(f"[{tag}] " + "\n".join(msgs))
    assert nK == 0, f"{tag}: read categories differ"
    assert nH == 0, f"{tag}: read labels differ"
    assert nC == 0, f"{tag}: H_class differs"
    assert dhp <= ATOL_DS and dgm <= ATOL_DS and dgf <= ATOL_DS, f"{tag}: dosage-level outputs differ"
    # GT = argmax of the genotype probabilities must be identical
    assert np.array_equal(np.argmax(g.genProbsM_t, axis=0), np.argmax(o.genProbsM_t, axis=0)), f"{tag}: GT differs"
    if state and g.alphaHat_t is not None:
        for h in range(len(g.alphaHat_t)):
            assert _rel(g.c[h], o.c[h]) < RTOL_STATE, f"{tag}: c[{h}]"
            assert _rel(g.eMatGrid_t[h], o.eMatGrid_t[h]) < RTOL_STATE, f"{tag}: eMatGrid[{h}]"
            assert _rel(g.alphaHat_t[h], o.alphaHat_t[h]) < RTOL_STATE, f"{tag}: alpha[{h}]"
            assert _rel(g.betaHat_t[h], o.betaHat_t[h]) < RTOL_STATE, f"{tag}: beta[{h}]"
    assert same_inf and dl < 1e-8, f"{tag}: per_it_likelihoods"


# ---------------------------------------------------------------------------------------------- components
@pytest.mark.parametrize("K", [64, 200, 600])
def test_unpack_panel_common(gpu, oracle, small_world, K):
    rng = np.random.default_rng(K)
    which = rng.choice(small_world.panel.K_full, size=K, replace=False) + 1  # unsorted on purpose (Appendix D.10)
    wg = gpu.unpack_panel(small_world.panel, which, all_snps=False)
    wo = oracle.unpack_panel(small_world.panel, which, all_snps=False)
    assert np.array_equal(wg, wo)


def test_unpack_panel_all_snps(gpu, oracle, small_world):
    which = np.sort(np.random.default_rng(5).choice(small_world.panel.K_full, size=300, replace=False)) + 1
    wg = gpu.unpack_panel(small_world.panel, which, all_snps=True)
    wo = oracle.unpack_panel(small_world.panel, which, all_snps=True)
    assert np.array_equal(wg, wo)
    # and they are the truth alleles of the synthetic panel
    bits = synth.unpack_words(wg, small_world.nSNPs_all)
    assert np.array_equal(bits, small_world.bits_all[which - 1])


@pytest.mark.parametrize("K", [100, 600])
@pytest.mark.parametrize("all_snps", [False, True])
def test_make_eMatRead_t(gpu, oracle, small_world, small_reads, K, all_snps):
    reads = small_reads.all if all_snps else small_reads.common
    call = synth.make_call(small_world, reads, 11, K=K, all_snps=all_snps, first_iteration=False)
    if all_snps:
        call.flags &= ~cabi.F_DISABLE_READ_CATEGORY_USAGE  # look at the real categories as well
    eg, cg = gpu.make_eMatRead_t(call)
    eo, co = oracle.make_eMatRead_t(call)
    print(f"eMatRead K={K} all={all_snps}: max abs diff {np.max(np.abs(eg - eo)):.3e}, categories {np.bincount(co, minlength=4)}")
    assert np.array_equal(cg, co)
    assert np.array_equal(eg, eo), "emission table arithmetic is element-wise identical to the reference order -> bit-exact"


def test_make_eMatRead_t_dense_and_gather(gpu, oracle, small_world):
    """reads with > NBMAX SNPs (dense columns), with gaps (gather mode) and spanning > 3 grids"""
    rng = np.random.default_rng(3)
    nS = small_world.nSNPs
    offs, u, bq, wif = [0], [], [], []
    for g in range(0, small_world.nGrids, 3):
        base = 32 * g
        kinds = [np.arange(base + 3, base + 3 + 14), np.array([base + 1, base + 4, base + 30, base + 33]), np.array([base + 31, base + 32]),
                 np.arange(base + 20, min(base + 20 + 80, nS), 9)]
        for idx in kinds:
            idx = idx[idx < nS]
            if idx.size == 0:
                continue
            q = rng.integers(20, 41, size=idx.size) * rng.choice([-1, 1], size=idx.size)
            u += list(idx)
            bq += list(q)
            offs.append(len(u))
            wif.append(int(idx[idx.size // 2] // 32))
    order = np.argsort(wif, kind="stable")
    offs = np.asarray(offs)
    gather = np.concatenate([np.arange(offs[i], offs[i + 1]) for i in order])
    new_offs = np.concatenate([[0], np.cumsum(np.diff(offs)[order])])
    reads = cabi.Reads(offsets=new_offs, u=np.asarray(u)[gather], bq=np.asarray(bq)[gather], wif0=np.asarray(wif)[order])
    call = synth.make_call(small_world, reads, 12, K=300, first_iteration=False)
    eg, cg = gpu.make_eMatRead_t(call)
    eo, co = oracle.make_eMatRead_t(call)
    assert np.array_equal(cg, co)
    assert np.array_equal(eg, eo)
    # Jmax truncation
    call.Jmax = 2
    eg, cg = gpu.make_eMatRead_t(call)
    eo, co = oracle.make_eMatRead_t(call)
    assert np.array_equal(cg, co) and np.array_equal(eg, eo)


@pytest.mark.parametrize("K,T", [(37, 5), (512, 64), (600, 33), (1500, 20), (3000, 12)])
def test_forward_backward(gpu, oracle, K, T):
    rng = np.random.default_rng(K + T)
    e = np.asfortranarray(rng.uniform(0.05, 1.0, size=(K, T)))
    sig = rng.uniform(0.9, 0.9999, size=T - 1)
    tm = np.asfortranarray(np.stack([sig, 1 - sig]))
    ag, bg, cg = gpu.forward_backward(e, tm)
    ao, bo, co = oracle.forward_backward(e, tm)
    print(f"fb K={K} T={T}: alpha {_rel(ag, ao):.2e} beta {_rel(bg, bo):.2e} c {_rel(cg, co):.2e}")
    assert _rel(cg, co) < 1e-12 and _rel(ag, ao) < 1e-11 and _rel(bg, bo) < 1e-11


# ---------------------------------------------------------------------------------------------- whole calls
def _run_both(gpu, oracle, call):
    call.flags |= cabi.F_RETURN_ALPHA
    return gpu.gibbs(call), oracle.gibbs(call)


@pytest.mark.parametrize("K", [200, 600])
@pytest.mark.parametrize("iterative", [False, True])
@pytest.mark.parametrize("its", [(0, 0), (1, 0), (2, 0), (3, 1)])
def test_gibbs_short(gpu, oracle, small_world, small_reads, K, iterative, its):
    """init only / one sweep / two sweeps / 3 + 1 sampling sweep, no block Gibbs: bisects the sweep kernel"""
    call = synth.make_call(small_world, small_reads.common, 21, K=K, first_iteration=iterative, n_burn_in=its[0], n_sample=its[1], block_its=())
    g, o = _run_both(gpu, oracle, call)
    _compare(f"short K={K} iterative={iterative} its={its}", g, o)


@pytest.mark.parametrize("K", [200, 600])
def test_gibbs_shard(gpu, oracle, small_world, small_reads, K):
    call = synth.make_call(small_world, small_reads.common, 22, K=K, first_iteration=False, n_burn_in=3, n_sample=1, block_its=(1,))
    g, o = _run_both(gpu, oracle, call)
    _compare(f"shard K={K}", g, o)


@pytest.mark.parametrize("K", [200, 600, 1500, 3000])
@pytest.mark.parametrize("iterative", [False, True])
def test_gibbs_production_common(gpu, oracle, K, iterative):
    """the production common-SNP call: 20 + 1 sweeps, block Gibbs at 3/6/9 (functions.R:620-706)"""
    w = synth.make_world(99 + K, K_full=max(K + 100, 700), nSNPs=2240, region_bp=210_000)
    sr = synth.make_sample_reads(w, K, coverage=1.0, region_bp=210_000)
    call = synth.make_call(w, sr.common, 23, K=K, first_iteration=iterative)
    g, o = _run_both(gpu, oracle, call)
    _compare(f"production K={K} iterative={iterative}", g, o)


@pytest.mark.parametrize("classes", ["0", "1"])
@pytest.mark.parametrize("kind", ["common", "iterative", "all_snps", "diverse_panel"])
def test_gibbs_both_sweep_instances(gpu, oracle, small_world, small_reads, monkeypatch, classes, kind):
    """the sweep kernel has two instances (sweep.cuh): reads decided on haplotype-class totals (classes.cuh; chosen for calls with many reads per
    grid) and the K-long walk.  Both are forced here on the same calls, including grids that fall back inside the class instance: reads kept as
    dense columns (all-SNP call), initialisation / pass-through reads (iterative), and a panel with more distinct haplotypes per grid than
    CLS_MAX (every haplotype carries private flips)."""
    monkeypatch.setenv("QUILT_B200_CLASSES", classes)
    if kind == "diverse_panel":
        w = synth.make_world(4242, K_full=900, nSNPs=1600, region_bp=150_000, n_founders=400, flip_rate=0.03, nMaxDH=255)
        sr = synth.make_sample_reads(w, 43, coverage=2.0, region_bp=150_000)
        call = synth.make_call(w, sr.common, 44, K=800, first_iteration=False)
    elif kind == "all_snps":
        call = synth.make_call(small_world, small_reads.all, 24, K=600, all_snps=True)
    else:
        call = synth.make_call(small_world, small_reads.common, 25, K=600, first_iteration=(kind == "iterative"))
    g, o = _run_both(gpu, oracle, call)
    _compare(f"sweep instance classes={classes} {kind}", g, o)


@pytest.mark.parametrize("K", [200, 600])
def test_gibbs_production_all_snps(gpu, oracle, small_world, small_reads, K):
    """the final all-SNP (rare/common) call (rare_common.R:325-391)"""
    call = synth.make_call(small_world, small_reads.all, 24, K=K, all_snps=True)
    g, o = _run_both(gpu, oracle, call)
    _compare(f"all-SNP K={K}", g, o)


@pytest.mark.parametrize("K", [200, 600, 1500])
@pytest.mark.parametrize("iterative", [False, True])
@pytest.mark.parametrize("ff", [0.1, 0.25])
def test_gibbs_nipt_sweeps(gpu, oracle, small_world, small_reads, K, iterative, ff):
    """three haplotypes (NIPT, ff > 0): 6 + 2 sweeps of the general read-label resampler, no block Gibbs"""
    w = small_world if K <= 500 else synth.make_world(77 + K, K_full=K + 100, nSNPs=2240, region_bp=210_000)
    sr = small_reads if K <= 500 else synth.make_sample_reads(w, K, coverage=0.5, region_bp=210_000)
    call = synth.make_call(w, sr.common, 31, K=K, first_iteration=iterative, ff=ff, n_burn_in=6, n_sample=2, block_its=())
    g, o = _run_both(gpu, oracle, call)
    _compare(f"NIPT sweeps K={K} iterative={iterative} ff={ff}", g, o)


@pytest.mark.parametrize("K", [200, 600])
def test_gibbs_nipt_block_short(gpu, oracle, small_world, small_reads, K):
    """NIPT with one block-Gibbs episode after the second sweep: bisects define-blocks / resampler / re-forward"""
    call = synth.make_call(small_world, small_reads.common, 32, K=K, first_iteration=False, ff=0.2, n_burn_in=3, n_sample=1, block_its=(1,))
    g, o = _run_both(gpu, oracle, call)
    _compare(f"NIPT block short K={K}", g, o)


@pytest.mark.parametrize("K", [200, 600, 2048])
@pytest.mark.parametrize("iterative", [False, True])
def test_gibbs_nipt_production(gpu, oracle, K, iterative):
    """the production NIPT call: 20 + 1 sweeps, block Gibbs (six label permutations, H_class resampling) at 3/6/9"""
    w = synth.make_world(55 + K, K_full=max(K + 100, 700), nSNPs=2240, region_bp=210_000)
    sr = synth.make_sample_reads(w, K + 1, coverage=0.5, region_bp=210_000)
    call = synth.make_call(w, sr.common, 33, K=K, first_iteration=iterative, ff=0.1)
    g, o = _run_both(gpu, oracle, call)
    _compare(f"NIPT production K={K} iterative={iterative}", g, o)


@pytest.mark.parametrize("K", [4100, 5000, 8192])
@pytest.mark.parametrize("iterative", [False, True])
def test_gibbs_production_large_K(gpu, oracle, K, iterative):
    """Ksubset in (4096, 8192]: the sweep / shard kernels run as two-CTA clusters (4096 states per CTA, block sums
    exchanged through distributed shared memory)"""
    w = synth.make_world(300 + K, K_full=K + 60, nSNPs=1600, region_bp=150_000)
    sr = synth.make_sample_reads(w, K, coverage=1.0, region_bp=150_000)
    call = synth.make_call(w, sr.common, 51, K=K, first_iteration=iterative)
    g, o = _run_both(gpu, oracle, call)
    _compare(f"large K={K} iterative={iterative}", g, o)


def test_forward_backward_large_K(gpu, oracle):
    rng = np.random.default_rng(8)
    K, T = 6000, 9
    e = np.asfortranarray(rng.random((K, T)) * 0.9 + 0.05)
    sigma = rng.random(T - 1) * 0.5 + 0.5
    tm = np.asfortranarray(np.stack([sigma, 1 - sigma]))
    ag, bg, cg = gpu.forward_backward(e, tm)
    ao, bo, co = oracle.forward_backward(e, tm)
    assert _rel(cg, co) < 1e-12 and _rel(ag, ao) < 1e-11 and _rel(bg, bo) < 1e-11


def test_gibbs_unsorted_haps_and_sampling_its(gpu, oracle, small_world, small_reads):
    """which_haps_to_use unsorted (after mspbwt selection, Appendix D.10), 3 sampling sweeps averaged"""
    call = synth.make_call(small_world, small_reads.common, 25, K=300, sort_haps=False, first_iteration=False, n_burn_in=5, n_sample=3, block_its=(2,))
    g, o = _run_both(gpu, oracle, call)
    _compare("unsorted / 3 sampling its", g, o)


def test_gibbs_batch_mixed(gpu, oracle, small_world):
    """a batch mixing shapes and flags: every job must match its own oracle run"""
    calls = []
    for s in range(6):
        sr = synth.make_sample_reads(small_world, 100 + s, coverage=0.5 + 0.25 * s, region_bp=300_000)
        calls.append(synth.make_call(small_world, sr.common, 200 + s, K=200 if s % 2 else 333, first_iteration=(s % 3 == 0)))
        calls.append(synth.make_call(small_world, sr.all, 300 + s, K=200, all_snps=True))
    res = gpu.gibbs_batch(calls)
    for i, (c, g) in enumerate(zip(calls, res)):
        _compare(f"batch job {i}", g, oracle.gibbs(c), state=False)


def _with_stream(call, seed=99, slack=7):
    """the same call with its episode uniforms supplied as one flat unif_rand() stream (QuiltGibbsArgs.unif_stream)"""
    R, T = call.reads.nReads, call.nGrids
    shard = bool(call.flags & cabi.F_DO_SHARD_BLOCK_GIBBS)
    per = 8 * R + (R if call.ff > 0 else 0) + (T - 1 if shard else 0)
    call.unif_stream = np.random.default_rng(seed).random(len(call.block_gibbs_iterations) * per + slack)
    return call


@pytest.mark.parametrize("kw", [dict(K=200, first_iteration=True), dict(K=300, first_iteration=False, ff=0.1), dict(K=150, first_iteration=True, ff=0.25),
                                dict(K=200, first_iteration=False, n_sample=3)],
                         ids=["diploid", "nipt_ff10", "nipt_ff25_iterative", "diploid_three_sampling_sweeps"])
def test_episode_stream_mode(gpu, oracle, small_world, small_reads, kw):
    """episode uniforms as one flat stream consumed in the reference's order: for NIPT the number of H_class draws per episode is
    data-dependent, so the position of every later episode is; the library must report how far the reference would have read"""
    call = _with_stream(synth.make_call(small_world, small_reads.common, 41, **kw))
    g, o = gpu.gibbs(call), oracle.gibbs(call)
    _compare(f"stream mode {kw}", g, o, state=False)
    assert np.array_equal(g.H_sample_its, o.H_sample_its), "labels after each sampling sweep (double_list_of_ending_read_labels)"
    assert g.n_unif_consumed == o.n_unif_consumed > 0
    assert g.underflow_iteration == o.underflow_iteration == -1


def test_underflow_iteration_and_stream_position(gpu, oracle, small_world):
    sr = synth.make_sample_reads(small_world, 9, coverage=60.0, region_bp=300_000)
    call = _with_stream(synth.make_call(small_world, sr.common, 26, K=100, first_iteration=False, maxDifferenceBetweenReads=1e300))
    g, o = gpu.gibbs(call), oracle.gibbs(call)
    assert g.underflow_problem and o.underflow_problem
    assert g.underflow_iteration == o.underflow_iteration >= 0
    assert g.n_unif_consumed == o.n_unif_consumed


def test_panel_cache_is_keyed_by_content(gpu, oracle, small_world, small_reads):
    """two different panels living at the SAME host addresses (the first one overwritten in place) must not share a device copy"""
    w2 = synth.make_world(777, K_full=small_world.panel.K_full, nSNPs=small_world.panel.nSNPs, region_bp=300_000, all_snps_factor=3)
    call = synth.make_call(small_world, small_reads.common, 43, K=150, first_iteration=False)
    g1 = gpu.gibbs(call)
    p1, p2 = small_world.panel, w2.panel
    saved = {}
    for f in ("hapMatcherR", "distinctHapsB", "distinctHapsIE", "special_helper"):
        a, b = getattr(p1, f), getattr(p2, f)
        if a.shape == b.shape:
            saved[f] = a.copy()
            a[...] = b
    if len(saved) < 3:
        pytest.skip("the two synthetic panels differ in shape")
    try:
        o2 = oracle.gibbs(call)
        g2 = gpu.gibbs(call)
        assert np.array_equal(g2.H, o2.H), "a stale device panel was used"
        assert not np.array_equal(g1.hapProbs_t, g2.hapProbs_t)
    finally:
        for f, v in saved.items():
            getattr(p1, f)[...] = v


def test_underflow_reported(gpu, oracle, small_world, small_reads):
    """maxDifferenceBetweenReads huge + many reads per grid -> the reference reports underflow instead of failing"""
    sr = synth.make_sample_reads(small_world, 9, coverage=60.0, region_bp=300_000)
    call = synth.make_call(small_world, sr.common, 26, K=100, first_iteration=False, maxDifferenceBetweenReads=1e300)
    g, o = gpu.gibbs(call), oracle.gibbs(call)
    print("underflow:", g.underflow_problem, o.underflow_problem)
    assert g.underflow_problem == o.underflow_problem


def test_bad_arguments(gpu, small_world, small_reads):
    QuiltGpuError = RuntimeError  # api.QuiltGpuError derives from it

    call = synth.make_call(small_world, small_reads.common, 27, K=50)
    call.which_haps_to_use = call.which_haps_to_use.copy()
    call.which_haps_to_use[3] = small_world.panel.K_full + 5
    with pytest.raises(QuiltGpuError):
        gpu.gibbs(call)
    call = synth.make_call(small_world, small_reads.common, 27, K=50)
    call.reads = cabi.Reads(offsets=call.reads.offsets, u=call.reads.u, bq=call.reads.bq, wif0=call.reads.wif0[::-1].copy())
    with pytest.raises(QuiltGpuError):
        gpu.gibbs(call)
