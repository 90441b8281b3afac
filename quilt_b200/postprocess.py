"""R-side reductions that follow the Gibbs calls of one sample, as a host mirror in numpy.

Reference (QUILT/R/functions.R): the running sums over the non-phasing chains' calls past the seek burn-in
(`:999-1020`, all-SNP variant `:1100-1122`), the final normalisation (`:1304-1325`) and `recast_haps`
(`:3180-3209`), which reconciles the rounded phasing haplotypes with the arg-max genotype.  These are O(nSNPs)
vector operations on the hapProbs_t rows returned by `rcpp_forwardBackwardGibbsNIPT`; they stay on the host.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np


def genotype_probabilities(h1: np.ndarray, h2: np.ndarray) -> np.ndarray:
    """rbind((1 - hap1) * (1 - hap2), (1 - hap1) * hap2 + hap1 * (1 - hap2), hap1 * hap2)  (functions.R:1004-1005)"""
    return np.stack([(1 - h1) * (1 - h2), (1 - h1) * h2 + h1 * (1 - h2), h1 * h2])


@dataclass
class SampleAccumulator:
    """dosage / gp_t running sums of one sample (diploid) or mat_* / fet_* (NIPT)"""

    nSNPs: int
    method: str = "diploid"
    nDosage: int = 0
    dosage: np.ndarray = field(default=None)
    gp_t: np.ndarray = field(default=None)
    fet_dosage: Optional[np.ndarray] = None
    fet_gp_t: Optional[np.ndarray] = None

    def __post_init__(self):
        self.dosage = np.zeros(self.nSNPs)
        self.gp_t = np.zeros((3, self.nSNPs))
        if self.method != "diploid":
            self.fet_dosage = np.zeros(self.nSNPs)
            self.fet_gp_t = np.zeros((3, self.nSNPs))

    def add(self, hapProbs_t: np.ndarray):
        """one stored call: hapProbs_t is the [3 x nSNPs] matrix of the call (row 3 unused when diploid)"""
        h1, h2 = hapProbs_t[0], hapProbs_t[1]
        self.dosage = self.dosage + h1 + h2
        self.gp_t = self.gp_t + genotype_probabilities(h1, h2)
        if self.method != "diploid":
            h3 = hapProbs_t[2]
            self.fet_dosage = self.fet_dosage + h1 + h3
            self.fet_gp_t = self.fet_gp_t + genotype_probabilities(h1, h3)
        self.nDosage += 1

    def finalize(self):
        """division by the number of stored calls (functions.R:1304-1325)"""
        if self.nDosage == 0:
            raise ValueError("no call was accumulated")
        out = {"dosage": self.dosage / self.nDosage, "gp_t": self.gp_t / self.nDosage}
        if self.method != "diploid":
            out["fet_dosage"] = self.fet_dosage / self.nDosage
            out["fet_gp_t"] = self.fet_gp_t / self.nDosage
        return out


def recast_haps(hd1: np.ndarray, hd2: np.ndarray, gp: np.ndarray, force_round: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """functions.R:3180-3209.  gp is [nSNPs x 3]; returns the haplotype dosages made consistent with the arg-max genotype
    (first maximum wins, like the reference's strict `>` scan).  R's round() and numpy's both round half to even."""
    hd1 = np.array(hd1, dtype=np.float64, copy=True)
    hd2 = np.array(hd2, dtype=np.float64, copy=True)
    gp = np.asarray(gp, dtype=np.float64)
    if force_round:
        hd1 = np.round(hd1)
        hd2 = np.round(hd2)
    gt1 = np.round(hd1) + np.round(hd2)
    max_val = gp[:, 0].copy()
    gt3 = np.zeros(gp.shape[0])
    for i in (1, 2):
        w = gp[:, i] > max_val
        gt3[w] = i
        max_val[w] = gp[w, i]
    to_change = gt3 != gt1
    z = to_change & (gt3 == 0)
    hd1[z] = 0
    hd2[z] = 0
    t = to_change & (gt3 == 2)
    hd1[t] = 1
    hd2[t] = 1
    o = to_change & (gt3 == 1)
    a1, a2 = hd1[o].copy(), hd2[o].copy()
    hd1[o] = np.where(a1 > a2, 1.0, 0.0)
    hd2[o] = np.where(a1 > a2, 0.0, 1.0)
    return hd1, hd2


def phased_genotypes(hd1: np.ndarray, hd2: np.ndarray, gp: np.ndarray) -> np.ndarray:
    """phased GT per SNP as an int array [nSNPs x 2] = round(recast haplotypes); the VCF column formatter (STITCH's
    rcpp_make_column_of_vcf, out of scope) prints it as a|b"""
    r1, r2 = recast_haps(hd1, hd2, gp)
    return np.stack([np.round(r1), np.round(r2)], axis=1).astype(np.int8)
