"""Committed fixtures (tests/golden/*.npz, made by tools/make_golden.py): inputs stored verbatim + the oracle's frozen answers.
CPU: the oracle still reproduces them bit for bit.  GPU: the CUDA path matches them (labels exact, DS/GP 1e-4)."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden  # noqa: E402

FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def test_fixtures_exist():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(oracle, path):
    call, exp = make_golden.load(path)
    r = oracle.gibbs(call)
    assert np.array_equal(r.H, exp["H"]) and np.array_equal(r.H_class, exp["H_class"]) and np.array_equal(r.read_category, exp["read_category"])
    for f in ("hapProbs_t", "genProbsM_t", "genProbsF_t"):
        assert np.array_equal(getattr(r, f), exp[f]), f
    assert np.array_equal(np.nan_to_num(r.per_it_likelihoods, posinf=1e300), np.nan_to_num(exp["per_it_likelihoods"], posinf=1e300))


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_gpu_matches_golden(gpu, path):
    call, exp = make_golden.load(path)
    r = gpu.gibbs(call)
    assert bool(r.underflow_problem) == bool(exp["underflow_problem"])
    assert np.array_equal(r.read_category, exp["read_category"])
    assert np.array_equal(r.H, exp["H"]), "phased read labels must be identical"
    assert np.array_equal(r.H_class, exp["H_class"])
    for f in ("hapProbs_t", "genProbsM_t", "genProbsF_t"):
        assert np.max(np.abs(getattr(r, f) - exp[f])) <= 1e-4, f  # north_star: DS / GP within 1e-4
    assert np.array_equal(np.argmax(r.genProbsM_t, axis=0), np.argmax(exp["genProbsM_t"], axis=0)), "GT must be identical"
