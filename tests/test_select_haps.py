"""Haplotype re-selection between Gibbs calls (select_new_haps_mspbwt_v3, QUILT/R/mspbwt.R:230-474).

CPU: properties of the oracle section (oracle/select_oracle.cpp: in-tree R restated, mspbwt matching by the documented contract —
"parity unpinned" for that un-vendored step).  GPU: the device kernels return the oracle's list element for element, standalone
and chained from one staged batch into the next without a host round trip."""
import numpy as np
import pytest

from quilt_b200 import api, cabi, synth


def _panel_words(panel):
    """32-SNP word of every (haplotype, grid) whose symbol is in the table (0 = special -> -1)"""
    hm = panel.hapMatcherR.astype(np.int64)
    db = panel.distinctHapsB.astype(np.int64) & 0xFFFFFFFF
    w = np.full(hm.shape, -1, dtype=np.int64)
    for g in range(hm.shape[1]):
        ok = hm[:, g] > 0
        w[ok, g] = db[hm[ok, g] - 1, g]
    return w


def _hap_as_probs(world, k):
    """hapProbs_t whose rounding reproduces panel haplotype k exactly (0.9 / 0.1)"""
    bits = world.bits_common[k].astype(np.float64)
    return 0.1 + 0.8 * bits


def test_oracle_finds_the_haplotype_itself(oracle, small_world):
    """feeding panel haplotypes as the sample's haplotypes: each is its own longest match in every subset, so it is selected"""
    w = small_world
    hp = np.zeros((3, w.panel.nSNPs), order="F")
    hp[0], hp[1] = _hap_as_probs(w, 17), _hap_as_probs(w, 401)
    which, n_unique = oracle.select_haps(w.panel, hp, Knew=50)
    assert 18 in which and 402 in which
    assert len(set(which.tolist())) == len(which) == min(50, n_unique)
    assert which.min() >= 1 and which.max() <= w.panel.K_full


def test_oracle_cut_and_order(oracle, small_world):
    """more haplotypes found than Knew -> the coverage-weighted ranking, interleaved over the two haplotypes, cut at Knew"""
    w = small_world
    hp = np.zeros((3, w.panel.nSNPs), order="F")
    hp[0], hp[1] = _hap_as_probs(w, 3), _hap_as_probs(w, 250)
    full, n_unique = oracle.select_haps(w.panel, hp, Knew=w.panel.K_full)
    assert len(full) == n_unique > 12
    cut, n_unique2 = oracle.select_haps(w.panel, hp, Knew=12)
    assert n_unique2 == n_unique and len(cut) == 12 and len(set(cut.tolist())) == 12
    assert set(cut.tolist()) <= set(full.tolist())
    # the two source haplotypes match over the whole region: they carry the largest weights of their lists, hence lead the interleave
    assert set(cut[:2].tolist()) == {4, 251}


def test_oracle_nothing_matches(oracle, small_world):
    """a haplotype whose words are not in the panel's tables finds nothing: the caller pads with sample() (mspbwt.R:357-366)"""
    w = small_world
    hp = np.full((3, w.panel.nSNPs), 0.5 + 1e-9, order="F")  # all-alt words
    hp[:, ::2] = 0.0
    which, n_unique = oracle.select_haps(w.panel, hp, Knew=20)
    words = _panel_words(w.panel)
    z = 0
    for b in range(1, 32, 2):
        z |= 1 << b
    expected_hits = int(np.sum(words == z))
    assert (n_unique == 0) == (expected_hits == 0)


def test_oracle_padding_rule(oracle, small_world):
    w = small_world
    hp = np.zeros((3, w.panel.nSNPs), order="F")
    hp[0], hp[1] = _hap_as_probs(w, 17), _hap_as_probs(w, 401)
    which, n_unique = oracle.select_haps(w.panel, hp, Knew=w.panel.K_full)
    K = min(w.panel.K_full, n_unique + 40)
    pu = np.random.default_rng(3).random(K)
    padded = oracle.select_haps_padded(w.panel, hp, K, pu)
    assert np.array_equal(padded[:n_unique], which[:n_unique])
    assert len(set(padded.tolist())) == K and padded.min() >= 1 and padded.max() <= w.panel.K_full


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("Knew", [12, 60, 600])
@pytest.mark.parametrize("source", ["panel_haps", "gibbs_call", "nipt_call"])
def test_gpu_select_equals_oracle(gpu, oracle, small_world, small_reads, Knew, source):
    w = small_world
    nHap = 2
    if source == "panel_haps":
        hp = np.zeros((3, w.panel.nSNPs), order="F")
        hp[0], hp[1] = _hap_as_probs(w, 3), _hap_as_probs(w, 250)
    elif source == "gibbs_call":
        hp = oracle.gibbs(synth.make_call(w, small_reads.common, 1, K=200, first_iteration=True)).hapProbs_t
    else:
        nHap = 3
        hp = oracle.gibbs(synth.make_call(w, small_reads.common, 33, K=200, first_iteration=True, ff=0.2)).hapProbs_t
    wo, no = oracle.select_haps(w.panel, hp, Knew, nHap=nHap)
    wg, ng = gpu.select_haps(w.panel, hp, Knew, nHap=nHap)
    print(f"{source} Knew={Knew}: found {len(wo)} of {no} unique")
    assert ng == no
    assert np.array_equal(wg, wo)


@pytest.mark.gpu
@pytest.mark.parametrize("nind,L,M", [(1, 3, 1), (4, 1, 1), (4, 3, 3), (7, 8, 2)])
def test_gpu_select_parameter_space(gpu, oracle, small_world, nind, L, M):
    w = small_world
    hp = np.zeros((3, w.panel.nSNPs), order="F")
    hp[0], hp[1] = _hap_as_probs(w, 99), _hap_as_probs(w, 100)
    # a few flipped SNPs break the runs so that shorter matches of other haplotypes are reported too
    rng = np.random.default_rng(5)
    flip = rng.choice(w.panel.nSNPs, 40, replace=False)
    hp[0, flip] = 1.0 - hp[0, flip]
    for Knew in (10, 300):
        wo, no = oracle.select_haps(w.panel, hp, Knew, mspbwt_nindices=nind, mspbwtL=L, mspbwtM=M)
        wg, ng = gpu.select_haps(w.panel, hp, Knew, mspbwt_nindices=nind, mspbwtL=L, mspbwtM=M)
        assert ng == no and np.array_equal(wg, wo)


@pytest.mark.gpu
def test_gpu_chained_selection_without_host_round_trip(gpu, oracle, small_world, small_reads):
    """call -> select -> call on the device: the second batch's jobs receive their haplotype lists from the first batch's hapProbs_t"""
    w = small_world
    K = 300
    first = [synth.make_call(w, small_reads.common, 60 + j, K=K, first_iteration=True) for j in range(3)]
    second = [synth.make_call(w, small_reads.common, 70 + j, K=K, first_iteration=False, sort_haps=False) for j in range(3)]
    pu = np.random.default_rng(9).random(3 * K)
    b1, b2 = api.Batch(gpu, first), api.Batch(gpu, second)
    b1.run()
    b1.sync()
    r1 = b1.fetch()
    b1.chain_select_into(b2, pu)
    lists = [b2.which_haps(j) for j in range(3)]
    b2.run()
    b2.sync()
    r2 = b2.fetch()
    b1.free()
    b2.free()
    for j in range(3):
        expect = oracle.select_haps_padded(w.panel, r1[j].hapProbs_t, K, pu[j * K:(j + 1) * K])
        assert np.array_equal(lists[j], expect), f"job {j}: device list differs from the oracle's"
        assert len(set(lists[j].tolist())) == K
        # the second call really ran on the selected haplotypes: identical to the oracle given that list
        c = second[j]
        c.which_haps_to_use = expect
        o = oracle.gibbs(c)
        assert np.array_equal(r2[j].H, o.H)
        assert np.max(np.abs(r2[j].hapProbs_t - o.hapProbs_t)) <= 1e-4


@pytest.mark.gpu
def test_gpu_chained_host_buffer_calls(gpu, oracle, small_world, small_reads):
    """the same chain through quilt_gpu_gibbs_batch_chained (host buffers in / out, waves pipelined): the intermediate call keeps its
    probabilities on the device (QUILT_F_OUTPUT_NO_PROBS), the second call's lists are selected there"""
    w = small_world
    K = 250
    first = [synth.make_call(w, small_reads.common, 80 + j, K=K, first_iteration=True) for j in range(4)]
    second = [synth.make_call(w, small_reads.common, 90 + j, K=K, first_iteration=False, sort_haps=False) for j in range(4)]
    for c in first:
        c.flags |= cabi.F_OUTPUT_NO_PROBS
    pu = np.random.default_rng(11).random(4 * K)
    r1, kept = api.run_prepared_chained(gpu, gpu.prepare(first), None, None, keep=True)
    r2, none = api.run_prepared_chained(gpu, gpu.prepare(second), kept, pu, keep=False)
    kept.free()
    assert none is None
    for j in range(4):
        o1 = oracle.gibbs(first[j])
        assert np.array_equal(r1[j].H, o1.H)   # labels are shipped even when the probabilities stay on the device
        expect = oracle.select_haps_padded(w.panel, o1.hapProbs_t, K, pu[j * K:(j + 1) * K])
        c = second[j]
        c.which_haps_to_use = expect
        o2 = oracle.gibbs(c)
        assert np.array_equal(r2[j].H, o2.H)
        assert np.max(np.abs(r2[j].hapProbs_t - o2.hapProbs_t)) <= 1e-4
        assert np.max(np.abs(r2[j].genProbsM_t - o2.genProbsM_t)) <= 1e-4   # formed on the host from hapProbs_t (one sampling sweep)
        assert np.max(np.abs(r2[j].genProbsF_t - o2.genProbsF_t)) <= 1e-4


@pytest.mark.gpu
def test_gpu_whole_chain_in_one_call(gpu, oracle, small_world, small_reads):
    """quilt_gpu_gibbs_chain: three stages (two common-SNP calls, then the all-SNP call) with host buffers, one pipeline; every stage
    equals the CPU chain (oracle Gibbs call -> oracle selection -> next call)"""
    w = small_world
    K, n = 200, 3
    stages = [[synth.make_call(w, small_reads.common, 100 + j, K=K, first_iteration=True) for j in range(n)],
              [synth.make_call(w, small_reads.common, 110 + j, K=K, first_iteration=False, sort_haps=False) for j in range(n)],
              [synth.make_call(w, small_reads.all, 120 + j, K=K, all_snps=True, sort_haps=False) for j in range(n)]]
    for c in stages[0]:
        c.flags |= cabi.F_OUTPUT_NO_PROBS
    rng = np.random.default_rng(13)
    pads = [rng.random(n * K), rng.random(n * K)]
    res = api.run_chain_prepared(gpu, [gpu.prepare(st) for st in stages], pads)
    for j in range(n):
        prev = oracle.gibbs(stages[0][j])
        assert np.array_equal(res[0][j].H, prev.H)
        for s in (1, 2):
            c = stages[s][j]
            c.which_haps_to_use = oracle.select_haps_padded(w.panel, prev.hapProbs_t, K, pads[s - 1][j * K:(j + 1) * K])
            cur = oracle.gibbs(c)
            assert np.array_equal(res[s][j].H, cur.H), (s, j)
            assert np.max(np.abs(res[s][j].hapProbs_t - cur.hapProbs_t)) <= 1e-4
            assert np.max(np.abs(res[s][j].genProbsM_t - cur.genProbsM_t)) <= 1e-4
            prev = cur
