"""GPU tests at BASELINE.json's full size (chr20 2 Mb @1x, K = 4096 of 5008 haplotypes, T = 1000 / 3000 grids):
one call checked against the oracle, the rest through size-independent properties (probabilities, determinism,
independence of the batch order, imputation quality against the simulated truth)."""
import numpy as np
import pytest

from quilt_b200 import postprocess, synth

pytestmark = pytest.mark.gpu

K = 4096


@pytest.fixture(scope="module")
def world():
    return synth.make_world(20260118, K_full=5008, nSNPs=32000, region_bp=3_000_000, all_snps_factor=3)


@pytest.fixture(scope="module")
def calls(world):
    out = []
    for s in range(2):
        sr = synth.make_sample_reads(world, 4000 + s, coverage=1.0, region_bp=3_000_000)
        out.append((sr, synth.make_call(world, sr.common, 100 + s, K=K, first_iteration=True)))
        out.append((sr, synth.make_call(world, sr.common, 200 + s, K=K, first_iteration=False, sort_haps=False)))
        out.append((sr, synth.make_call(world, sr.all, 300 + s, K=K, all_snps=True, sort_haps=False)))
    return out


@pytest.fixture(scope="module")
def results(gpu, calls):
    return gpu.gibbs_batch([c for _, c in calls])


def _same_as_cpu(g, o, tag):
    """north_star's bar: phased labels and GT identical, DS / GP within 1e-4 (tolerance written here)"""
    assert g.underflow_problem == o.underflow_problem, tag
    assert np.array_equal(g.H, o.H), f"{tag}: {int(np.sum(g.H != o.H))} read labels differ"
    assert np.array_equal(g.H_class, o.H_class), tag
    assert np.array_equal(g.read_category, o.read_category), tag
    for f in ("hapProbs_t", "genProbsM_t", "genProbsF_t"):
        d = float(np.max(np.abs(getattr(g, f) - getattr(o, f))))
        assert d <= 1e-4, f"{tag}: max |d {f}| = {d:.3e}"
    assert np.array_equal(np.argmax(g.genProbsM_t, axis=0), np.argmax(o.genProbsM_t, axis=0)), f"{tag}: GT differs"
    print(f"[{tag}] labels / H_class / GT identical; max |d hapProbs| = {np.max(np.abs(g.hapProbs_t - o.hapProbs_t)):.2e}")


def _cpu_many(oracle, calls):
    """the CPU checker on several calls at once (ctypes releases the GIL; one call takes 5 - 25 s at this size)"""
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=min(len(calls), 8)) as ex:
        return list(ex.map(oracle.gibbs, calls))


def test_full_size_all_call_kinds_match_oracle(gpu, oracle, calls, results):
    """all three call kinds of the benchmark at its full size — iterative-initialisation call and replayed call on the
    common SNPs (K = 4096, T = 1000) and the all-SNP rare/common call (T = 3000): labels, H_class, read categories and GT
    identical to the CPU oracle (itself pinned bit for bit to the compiled reference), DS / GP within 1e-4"""
    idx = [0, 1, 2]
    cpu = _cpu_many(oracle, [calls[i][1] for i in idx])
    for i, o in zip(idx, cpu):
        _same_as_cpu(results[i], o, f"K=4096 call kind {['iterative', 'replayed', 'all-SNP'][i]}")


def test_full_size_nipt_call_matches_oracle(gpu, oracle, world):
    """BASELINE.json config #4 at its benchmark shape: NIPT, K = 2048, T = 1000, 0.5x, fetal fraction 10 %, 20 + 1 sweeps with
    the three-haplotype block Gibbs at 3/6/9"""
    sr = synth.make_sample_reads(world, 4100, coverage=0.5, region_bp=3_000_000, n_true_haps=3, hap_probs=(0.5, 0.45, 0.05))
    cs = [synth.make_call(world, sr.common, 400, K=2048, first_iteration=True, ff=0.1),
          synth.make_call(world, sr.common, 401, K=2048, first_iteration=False, sort_haps=False, ff=0.1)]
    for tag, g, o in zip(("iterative", "replayed"), gpu.gibbs_batch(cs), _cpu_many(oracle, cs)):
        _same_as_cpu(g, o, f"NIPT K=2048 T=1000 {tag}")


def test_full_size_K8192_from_large_panel_matches_oracle(gpu, oracle):
    """config #5's kernel shape: K = 8192 (two-CTA cluster kernels) selected from a 20 000-haplotype panel, T = 1000"""
    w = synth.make_world(20260119, K_full=20000, nSNPs=32000, region_bp=3_000_000)
    sr = synth.make_sample_reads(w, 4200, coverage=1.0, region_bp=3_000_000)
    call = synth.make_call(w, sr.common, 500, K=8192, first_iteration=False, sort_haps=False)
    _same_as_cpu(gpu.gibbs(call), oracle.gibbs(call), "K=8192 of 20000, T=1000")


def test_outputs_are_probabilities(calls, results):
    for (sr, call), r in zip(calls, results):
        assert not r.underflow_problem
        hp = r.hapProbs_t[:2]
        assert np.all(np.isfinite(hp)) and hp.min() >= 0.0 and hp.max() <= 1.0 + 1e-12
        assert np.allclose(r.genProbsM_t.sum(axis=0), 1.0, atol=1e-9)
        assert set(np.unique(r.H)) <= {1, 2}
        lik = r.per_it_likelihoods
        assert lik.shape == (21, 13) and np.all(lik[:, 2] == np.arange(1, 22))


def test_deterministic_and_batch_order_independent(gpu, calls, results):
    """same inputs -> bit-identical outputs, whatever the position of a call in the batch (wave / slot assignment)"""
    order = [4, 2, 0, 5, 3, 1]
    again = gpu.gibbs_batch([calls[i][1] for i in order])
    for pos, i in enumerate(order):
        a, b = results[i], again[pos]
        assert np.array_equal(a.H, b.H)
        assert np.array_equal(a.hapProbs_t, b.hapProbs_t)
        assert np.array_equal(a.genProbsM_t, b.genProbsM_t)
        assert np.array_equal(a.per_it_likelihoods, b.per_it_likelihoods)


def test_imputation_tracks_the_truth(calls, results):
    """dosage from the final all-SNP calls correlates with the simulated truth at 1x (sanity of the whole path;
    the reference's acceptance tests use accuracy thresholds of the same kind, test-drivers.R:1-89)"""
    for i in (2, 5):
        sr, _ = calls[i]
        acc = postprocess.SampleAccumulator(results[i].hapProbs_t.shape[1])
        acc.add(results[i].hapProbs_t)
        ds = acc.finalize()["dosage"]
        truth = sr.truth_haps[:2].sum(axis=0)
        poly = truth.std() > 0
        r2 = np.corrcoef(ds, truth)[0, 1] ** 2 if poly else 1.0
        print(f"all-SNP call {i}: r2(dosage, truth) = {r2:.3f}")
        assert r2 > 0.5
