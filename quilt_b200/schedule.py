"""The per-sample call schedule of QUILT2, restated over the C ABI.

For each of nGibbsSamples (+1 phasing) chains: n_seek_its Gibbs calls on the common SNPs (the first with
iterative initialisation, QUILT/R/functions.R:575-706), then one call on ALL SNPs
(impute_final_gibbs_with_rare_common, QUILT/R/rare_common.R:109-411).  The haplotype re-selection between
calls (mspbwt, functions.R:856) is host-side and out of scope here: the benchmark and the
parity tests replay `which_haps_to_use` lists, exactly as §8(c) prescribes.
"""
from __future__ import annotations

from typing import List

import numpy as np

from . import cabi, synth


def sample_calls(
    world: synth.World,
    reads: synth.SampleReads,
    seed: int,
    K: int,
    nGibbsSamples: int = 7,
    n_seek_its: int = 3,
    impute_rare_common: bool = True,
    ff: float = 0.0,
) -> List[cabi.GibbsCall]:
    """all Gibbs calls of one sample at QUILT2 defaults: (nGibbsSamples + 1) x (n_seek_its common + 1 all-SNP)"""
    calls: List[cabi.GibbsCall] = []
    rng = np.random.default_rng(seed)
    for chain in range(nGibbsSamples + 1):
        for i_it in range(n_seek_its):
            calls.append(
                synth.make_call(world, reads.common, int(rng.integers(1 << 31)), K=K, first_iteration=(i_it == 0), sort_haps=(i_it == 0), ff=ff)
            )
        if impute_rare_common and reads.all is not None:
            calls.append(synth.make_call(world, reads.all, int(rng.integers(1 << 31)), K=K, all_snps=True, sort_haps=False, ff=ff))
    return calls


def average_dosage(results: List[cabi.GibbsResult]) -> np.ndarray:
    """dosage = mean over the given calls of hap1 + hap2 (functions.R:999-1006, :1304-1325)"""
    acc = None
    for r in results:
        d = r.hapProbs_t[0] + r.hapProbs_t[1]
        acc = d if acc is None else acc + d
    return acc / len(results)
