"""R-side reductions after the Gibbs calls: host mirror checked on hand-computed cases and invariants."""
import numpy as np
import pytest

from quilt_b200 import postprocess
from quilt_b200 import postprocess as pp


def test_accumulator_matches_direct_formula():
    rng = np.random.default_rng(3)
    calls = [rng.random((3, 50)) for _ in range(5)]
    acc = pp.SampleAccumulator(50)
    for c in calls:
        acc.add(c)
    out = acc.finalize()
    dosage = sum(c[0] + c[1] for c in calls) / 5
    assert np.allclose(out["dosage"], dosage, atol=1e-15)
    assert np.allclose(out["gp_t"].sum(axis=0), 1.0, atol=1e-12)  # genotype probabilities of every call sum to 1
    # expected dosage from the averaged genotype probabilities equals the averaged dosage
    assert np.allclose(out["gp_t"][1] + 2 * out["gp_t"][2], dosage, atol=1e-12)


def test_accumulator_nipt_has_fetal_track():
    rng = np.random.default_rng(4)
    acc = pp.SampleAccumulator(20, method="nipt")
    c = rng.random((3, 20))
    acc.add(c)
    out = acc.finalize()
    assert np.allclose(out["fet_dosage"], c[0] + c[2])
    assert np.allclose(out["fet_gp_t"][2], c[0] * c[2])


def test_recast_haps_cases():
    # columns: hom ref forced, hom alt forced, het with a1 > a2, het with a1 <= a2, consistent (untouched)
    hd1 = np.array([0.6, 0.4, 0.7, 0.2, 0.9])
    hd2 = np.array([0.1, 0.3, 0.6, 0.2, 0.1])
    gp = np.array([[0.8, 0.1, 0.1], [0.1, 0.2, 0.7], [0.1, 0.8, 0.1], [0.2, 0.7, 0.1], [0.1, 0.8, 0.1]])
    r1, r2 = pp.recast_haps(hd1, hd2, gp)
    assert r1.tolist() == [0, 1, 1, 0, 0.9]
    assert r2.tolist() == [0, 1, 0, 1, 0.1]
    gt = pp.phased_genotypes(hd1, hd2, gp)
    assert gt.tolist() == [[0, 0], [1, 1], [1, 0], [0, 1], [1, 0]]


def test_recast_haps_ties_take_first_maximum():
    gp = np.array([[0.5, 0.5, 0.0]])
    r1, r2 = pp.recast_haps(np.array([0.9]), np.array([0.2]), gp)  # rounded sum 1, arg-max (first) 0 -> forced hom ref
    assert (r1[0], r2[0]) == (0.0, 0.0)


@pytest.mark.gpu
def test_device_sample_summary_equals_the_host_mirror(gpu, small_world, small_reads):
    """row a12 on the device (quilt_gpu_samples_summary): dosage / gp_t sums in R's order, final division, recast_haps, phased GT and the
    per-rank INFO counters from the device-resident hapProbs_t of a run batch == the numpy mirror of the R code, bit for bit"""
    from quilt_b200 import api, synth

    w = small_world
    calls = [synth.make_call(w, small_reads.common, 200 + j, K=150, first_iteration=(j % 2 == 0)) for j in range(8)]
    b = api.Batch(gpu, calls)
    b.run()
    b.sync()
    res = b.fetch()
    # two "samples": stored calls 0..2 / 4..6, phasing calls 3 / 7
    samples = [([(b, 0), (b, 1), (b, 2)], (b, 3)), ([(b, 4), (b, 5), (b, 6)], (b, 7))]
    outs, cnt = api.samples_summary(gpu, samples, w.panel.nSNPs)
    b.free()
    e_tot, f_tot, a_tot = np.zeros(w.panel.nSNPs), np.zeros(w.panel.nSNPs), np.zeros(w.panel.nSNPs)
    hwe = np.zeros((w.panel.nSNPs, 3))
    for (stored, ph), o in zip([([0, 1, 2], 3), ([4, 5, 6], 7)], outs):
        acc = postprocess.SampleAccumulator(w.panel.nSNPs)
        for j in stored:
            acc.add(res[j].hapProbs_t)
        hd1, hd2 = postprocess.recast_haps(res[ph].hapProbs_t[0], res[ph].hapProbs_t[1], acc.gp_t.T)
        fin = acc.finalize()
        assert np.array_equal(o["dosage"], fin["dosage"]) and np.array_equal(o["gp_t"], fin["gp_t"])
        assert np.array_equal(o["hd"][:, 0], hd1) and np.array_equal(o["hd"][:, 1], hd2)
        assert np.array_equal(o["gt"], np.stack([np.round(hd1), np.round(hd2)], axis=1).astype(np.int8))
        eij = np.round(1000 * (fin["gp_t"][1] + 2 * fin["gp_t"][2])) / 1000
        fij = np.round(1000 * (fin["gp_t"][1] + 4 * fin["gp_t"][2])) / 1000
        e_tot = e_tot + eij
        f_tot = f_tot + (fij - eij * eij)
        a_tot = a_tot + eij / 2
        mg = np.zeros(w.panel.nSNPs, dtype=int)
        mv = fin["gp_t"][0].copy()
        for i in (1, 2):
            wch = fin["gp_t"][i] > mv
            mg[wch] = i
            mv[wch] = fin["gp_t"][i][wch]
        hwe[np.arange(w.panel.nSNPs), mg] += 1
    assert np.array_equal(cnt["infoCount"][:, 0], e_tot) and np.array_equal(cnt["infoCount"][:, 1], f_tot)
    assert np.array_equal(cnt["afCount"], a_tot) and np.array_equal(cnt["hweCount"], hwe)
