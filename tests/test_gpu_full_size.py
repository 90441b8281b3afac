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


def test_full_size_call_matches_oracle(gpu, oracle, calls, results):
    """one production common-SNP call at the benchmark size: labels and GT identical, DS / GP within 1e-4"""
    _, call = calls[1]
    g, o = results[1], oracle.gibbs(call)
    assert g.underflow_problem == o.underflow_problem
    assert np.array_equal(g.H, o.H) and np.array_equal(g.H_class, o.H_class)
    assert np.max(np.abs(g.hapProbs_t - o.hapProbs_t)) <= 1e-4
    assert np.max(np.abs(g.genProbsM_t - o.genProbsM_t)) <= 1e-4
    assert np.array_equal(np.argmax(g.genProbsM_t, axis=0), np.argmax(o.genProbsM_t, axis=0))


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
