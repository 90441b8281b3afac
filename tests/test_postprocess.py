"""R-side reductions after the Gibbs calls: host mirror checked on hand-computed cases and invariants."""
import numpy as np

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
