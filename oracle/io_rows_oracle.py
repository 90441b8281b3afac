"""TEST INFRASTRUCTURE (CPU oracle) for the rows on either side of the Gibbs path: a NumPy / pure-Python
restatement of the R code.  Only tests may import this.

PARITY UNPINNED for the three STITCH functions the R code calls and this tree does not vendor (STITCH is a separate package):
  * convertScaledBQtoProbs(bq):  bq < 0 -> (1 - eps, eps / 3), bq > 0 -> (eps / 3, 1 - eps), eps = 10^(-|bq| / 10) — the same expression the
    vendored emission code uses for its (pR, pA) pairs (QUILT/src/gibbs-small.cpp:172-181);
  * snap_sampleReads_to_grid(sampleReads, grid): the read's second field (0-based central SNP) becomes grid[central SNP]; reads are handed on
    ordered by it (QUILT/src/gibbs-nipt.cpp:811 needs non-decreasing wif; QUILT/R/functions.R:295-298, :314-316);
  * rcpp_make_column_of_vcf(gp_t, use_state_probabilities = TRUE, q_t): "GT:GP:DS:HD" with three decimals (header QUILT/R/writers.R:10-36).
Pinned to the reference tree: increment2N (QUILT/src/copied-from-stitch.cpp:573-579), get_alleleCount (QUILT/R/functions.R:2779-2800), the GT
replacement by the phased genotype (:1436-1442)."""
import numpy as np


def convertScaledBQtoProbs(bq):
    bq = np.asarray(bq, dtype=np.float64)
    out = np.zeros((bq.shape[0], 2))
    for i, b in enumerate(bq):
        if b < 0:
            eps = 10.0 ** (b / 10)
            out[i] = (1 - eps, eps * (1.0 / 3.0))
        elif b > 0:
            eps = 10.0 ** (-b / 10)
            out[i] = (eps * (1.0 / 3.0), 1 - eps)
    return out


def increment2N(yT, xT, y, z):
    """QUILT/src/copied-from-stitch.cpp:573-579"""
    x = np.zeros(xT + 1)
    for t in range(yT):
        x[int(z[t])] = x[int(z[t])] + y[t]
    return x


def get_alleleCount(u, bq, nSNPs):
    """QUILT/R/functions.R:2779-2800 on the flattened sampleReads -> [nSNPs, 2] (the third column of the R array stays 0 here)"""
    p = convertScaledBQtoProbs(bq)
    c1 = increment2N(p.shape[0], nSNPs - 1, p[:, 0], u)
    c2 = increment2N(p.shape[0], nSNPs - 1, p[:, 1], u)
    return np.stack([c2, c1 + c2], axis=1)


def ingest_pileup(offsets, u, bq, central_snp, grid, nGrids):
    R = len(offsets) - 1
    wif = np.asarray(grid)[np.asarray(central_snp)]
    order = np.argsort(wif, kind="stable")
    cnt = np.diff(offsets)[order]
    new_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    gather = np.concatenate([np.arange(offsets[r], offsets[r + 1]) for r in order]) if R else np.zeros(0, np.int64)
    first = np.searchsorted(wif[order], np.arange(nGrids + 1), side="left").astype(np.int32)
    has = np.zeros(nGrids, np.uint8)
    has[wif] = 1
    return {"order": order.astype(np.int32), "offsets": new_off, "u": np.asarray(u)[gather].astype(np.int32), "bq": np.asarray(bq)[gather].astype(np.int32),
            "wif0": wif[order].astype(np.int32), "first_read_of_grid": first, "grid_has_read": has, "alleleCount": get_alleleCount(u, bq, len(grid))}


def make_vcf_column(gp_t, hd):
    """QUILT/R/functions.R:1421-1442 (diploid, output_gt_phased_genotypes = TRUE)"""
    out = []
    for s in range(gp_t.shape[1]):
        g0, g1, g2 = (float(x) for x in gp_t[:, s])
        h1, h2 = float(hd[s, 0]), float(hd[s, 1])
        out.append("%d|%d:%.3f,%.3f,%.3f:%.3f:%.3f,%.3f" % (int(np.rint(h1)), int(np.rint(h2)), g0, g1, g2, g1 + 2 * g2, h1, h2))
    return out
