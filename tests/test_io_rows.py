"""Rows on either side of the Gibbs path: pileup ingestion and the per-sample VCF column, CUDA vs the NumPy
restatement of the R code (oracle/io_rows_oracle.py) — integer outputs and text identical, allele counts bit for bit."""
import numpy as np
import pytest

from oracle import io_rows_oracle as orc
from quilt_b200 import synth


def _pileup(seed, nSNPs=3200, region_bp=300_000, coverage=2.0, shuffle=True):
    """reads in BAM (start) order with their central SNP, the way loadBamAndConvert hands them over — NOT ordered by grid"""
    rng = np.random.default_rng(seed)
    pos = np.sort(rng.choice(np.arange(1, region_bp), size=nSNPs, replace=False))
    n_reads = int(coverage * region_bp / 150)
    starts = np.sort(rng.integers(1, region_bp - 150, size=n_reads))
    lo, hi = np.searchsorted(pos, starts), np.searchsorted(pos, starts + 150)
    keep = hi > lo
    lo, hi = lo[keep], hi[keep]
    if shuffle:  # mate pairs / long inserts: the central SNP is not monotone in the read order
        p = np.arange(lo.shape[0])
        sw = rng.random(p.shape[0]) < 0.3
        p[sw] = np.clip(p[sw] + rng.integers(-5, 6, size=int(sw.sum())), 0, p.shape[0] - 1)
        lo, hi = lo[p], hi[p]
    offsets = np.concatenate([[0], np.cumsum(hi - lo)]).astype(np.int32)
    u = np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)]).astype(np.int32)
    q = rng.integers(5, 41, size=u.shape[0])
    bq = np.where(rng.random(u.shape[0]) < 0.5, q, -q).astype(np.int32)
    bq[rng.random(u.shape[0]) < 0.01] = 0
    central = np.array([u[offsets[r] + (offsets[r + 1] - offsets[r]) // 2] for r in range(lo.shape[0])], dtype=np.int32)
    grid = (np.arange(nSNPs) // 32).astype(np.int32)
    return offsets, u, bq, central, grid, (nSNPs + 31) // 32


def test_oracle_allele_count_and_order_invariants():
    offsets, u, bq, central, grid, T = _pileup(5)
    o = orc.ingest_pileup(offsets, u, bq, central, grid, T)
    assert np.all(np.diff(o["wif0"]) >= 0) and o["first_read_of_grid"][-1] == len(central)
    # the same multiset of (SNP, bq) entries, read by read
    for q in (0, 17, len(central) - 1):
        r = o["order"][q]
        assert np.array_equal(o["u"][o["offsets"][q]:o["offsets"][q + 1]], u[offsets[r]:offsets[r + 1]])
    depth = np.bincount(u[bq != 0], minlength=len(grid))
    assert np.allclose(o["alleleCount"][:, 1], depth - (2 / 3) * np.bincount(u, weights=10.0 ** (-np.abs(bq) / 10) * (bq != 0), minlength=len(grid)), atol=1e-9)
    assert orc.make_vcf_column(np.array([[0.0005], [0.9990], [0.0005]]), np.array([[0.4995, 0.5]]))[0] == "0|0:0.001,0.999,0.001:1.000:0.499,0.500"


@pytest.mark.gpu
@pytest.mark.parametrize("seed,cov", [(1, 1.0), (2, 8.0), (3, 0.05)])
def test_gpu_ingest_equals_oracle(gpu, seed, cov):
    from quilt_b200 import api

    offsets, u, bq, central, grid, T = _pileup(seed, coverage=cov)
    g = api.ingest_pileup(gpu, offsets, u, bq, central, grid, T)
    o = orc.ingest_pileup(offsets, u, bq, central, grid, T)
    for k in ("order", "offsets", "u", "bq", "wif0", "first_read_of_grid", "grid_has_read"):
        assert np.array_equal(g[k], o[k]), k
    assert np.array_equal(g["alleleCount"], o["alleleCount"]), "allele counts must be bit-identical (same order of additions)"


@pytest.mark.gpu
def test_gpu_ingest_feeds_the_gibbs_call(gpu, oracle, small_world):
    """ingested reads == the reads synth hands to the Gibbs path (it orders by the central SNP's grid the same way)"""
    from quilt_b200 import api, cabi

    sr = synth.make_sample_reads(small_world, 11, coverage=1.0, region_bp=300_000)
    rd = sr.common
    rng = np.random.default_rng(3)
    perm = rng.permutation(rd.nReads)  # undo the order
    off = np.concatenate([[0], np.cumsum(np.diff(rd.offsets)[perm])]).astype(np.int32)
    gather = np.concatenate([np.arange(rd.offsets[r], rd.offsets[r + 1]) for r in perm])
    u, bq = rd.u[gather].astype(np.int32), rd.bq[gather].astype(np.int32)
    central = np.array([u[off[r] + (off[r + 1] - off[r]) // 2] for r in range(rd.nReads)], dtype=np.int32)
    grid = (np.arange(small_world.nSNPs) // 32).astype(np.int32)
    g = api.ingest_pileup(gpu, off, u, bq, central, grid, small_world.nGrids)
    assert np.array_equal(g["wif0"], rd.wif0) and np.array_equal(np.sort(perm[g["order"]]), np.arange(rd.nReads))
    reads = cabi.Reads(offsets=g["offsets"], u=g["u"], bq=g["bq"], wif0=g["wif0"])
    call = synth.make_call(small_world, reads, 5, K=200, first_iteration=False, n_burn_in=2, n_sample=1, block_its=())
    a, b = gpu.gibbs(call), oracle.gibbs(call)
    assert np.array_equal(a.H, b.H)


@pytest.mark.gpu
def test_gpu_vcf_column_equals_oracle(gpu):
    from quilt_b200 import api

    rng = np.random.default_rng(9)
    n = 5000
    h = rng.random((n, 2))
    h[:200] = np.round(h[:200] * 1000) / 1000 + 0.0005  # values at the rounding boundary of the third decimal
    h[200:300, 0] = 0.5
    h = np.clip(h, 0, 1)
    gp = np.stack([(1 - h[:, 0]) * (1 - h[:, 1]), h[:, 0] * (1 - h[:, 1]) + h[:, 1] * (1 - h[:, 0]), h[:, 0] * h[:, 1]], axis=0)
    gp[:, 300:400] = np.round(gp[:, 300:400] * 2000) / 2000
    g = api.make_vcf_column(gpu, gp, h)
    o = orc.make_vcf_column(gp, h)
    bad = [i for i in range(n) if g[i] != o[i]]
    assert not bad, (bad[:5], [g[i] for i in bad[:5]], [o[i] for i in bad[:5]])


def test_prepared_reference_pack_round_trip(tmp_path):
    """the binary, memory-mappable prepared reference: every array comes back identical, mapped read-only, and two 'ranks' share the file"""
    from quilt_b200 import refpack

    w = synth.make_world(77, K_full=300, nSNPs=960, region_bp=90_000, all_snps_factor=3)
    path = str(tmp_path / "ref.qb2")
    n = refpack.save_panel(path, w.panel)
    assert n > w.panel.hapMatcherR.nbytes
    for mm in (True, False):
        p = refpack.load_panel(path, mmap=mm)
        for name in refpack._NAMES:
            assert np.array_equal(getattr(p, name), getattr(w.panel, name)), name
        assert p.ref_error == w.panel.ref_error and p.nSNPs == w.panel.nSNPs and p.K_full == 300
        assert p.c_struct().nMaxDH == w.panel.c_struct().nMaxDH
    with pytest.raises(ValueError):
        open(path, "r+b").write(b"XXXXXXXX")
        refpack.load_panel(path)
