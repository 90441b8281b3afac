"""Seeded synthetic inputs in the shape of the reference's objects.

Builds, with NumPy only:
  * a reference panel of mosaic haplotypes, bit-packed 32 SNPs/word (LSB first, as rhb_t,
    QUILT/src/copied-from-stitch.cpp:95-104) and compressed the way QUILT2 keeps it after
    QUILT_prepare_reference (hapMatcherR / distinctHapsB / distinctHapsIE / eMatDH_special_matrix(+helper);
    format pinned by QUILT/tests/testthat/test-unit-reference-single.R:273-303);
  * rare/common extras for the all-SNP stage (snp_is_common, common_snp_index, rare_per_hap_info;
    QUILT/R/rare_common.R:202-322);
  * low-coverage reads as sampleReads = list(J, wif, bq, u) flattened to CSR
    (recipe after QUILT/R/test-drivers.R:127-319: bq sign encodes the allele, |bq| the phred);
  * transition rates sigma_g = exp(-nGen * d_cM / 100) (QUILT/R/prepare_reference_functions.R:89-108)
    stored as transMatRate_tc_H rows (sigma, 1 - sigma) (test-drivers.R:352).

This is data plumbing for tests and bench.py; nothing here is on the compute path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import cabi


def pack_bits(bits: np.ndarray) -> np.ndarray:
    """bits [K, nSNPs] {0,1} -> rhb_t [K, T] uint32, bit b of word g = SNP 32 g + b."""
    K, n = bits.shape
    T = (n + 31) // 32
    pad = np.zeros((K, T * 32), dtype=np.uint32)
    pad[:, :n] = bits
    w = pad.reshape(K, T, 32) << np.arange(32, dtype=np.uint32)[None, None, :]
    return np.bitwise_or.reduce(w, axis=2).astype(np.uint32)


def unpack_words(words: np.ndarray, nSNPs: int) -> np.ndarray:
    K, T = words.shape
    b = (words[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1
    return b.reshape(K, T * 32)[:, :nSNPs].astype(np.uint8)


def compress_rhb_t(rhb_t: np.ndarray, nSNPs: int, nMaxDH: int, ref_error: float):
    """Our own restatement of what STITCH::make_rhb_t_equality emits (un-vendored).

    Per grid: the <= nMaxDH most frequent distinct words go to distinctHapsB[, g] (most frequent
    first); hapMatcherR[k, g] is the 1-based row, or 0 for a "special" haplotype whose word is kept
    in eMatDH_special_matrix (rows grouped by grid, sorted by 0-based hap within a grid) with
    eMatDH_special_matrix_helper[g, ] = 1-based (first, last) row of the grid's group.
    """
    K, T = rhb_t.shape
    assert nMaxDH <= 255
    hapMatcherR = np.zeros((K, T), dtype=np.uint8, order="F")
    distinctHapsB = np.zeros((nMaxDH, T), dtype=np.int32, order="F")
    distinctHapsIE = np.full((nMaxDH, nSNPs), ref_error, dtype=np.float64, order="F")
    special_rows = []
    helper = np.zeros((T, 2), dtype=np.int32, order="F")
    n_sp = 0
    for g in range(T):
        col = rhb_t[:, g]
        vals, inv, counts = np.unique(col, return_inverse=True, return_counts=True)
        order = np.argsort(-counts, kind="stable")
        keep = order[:nMaxDH]
        rank = np.zeros(vals.shape[0], dtype=np.int64)  # 0 = special
        rank[keep] = np.arange(1, keep.shape[0] + 1)
        hapMatcherR[:, g] = rank[inv]
        distinctHapsB[: keep.shape[0], g] = vals[keep].astype(np.uint32).view(np.int32)
        s, e = 32 * g, min(32 * (g + 1), nSNPs)
        bits = (vals[keep].astype(np.uint32)[:, None] >> np.arange(e - s, dtype=np.uint32)[None, :]) & 1
        distinctHapsIE[: keep.shape[0], s:e] = np.where(bits == 1, 1 - ref_error, ref_error)
        sp = np.nonzero(rank[inv] == 0)[0]
        if sp.size:
            helper[g, 0] = n_sp + 1
            helper[g, 1] = n_sp + sp.size
            special_rows.append(np.stack([sp.astype(np.int32), col[sp].astype(np.uint32).view(np.int32)], axis=1))
            n_sp += sp.size
    special = np.concatenate(special_rows, axis=0) if special_rows else np.zeros((0, 2), dtype=np.int32)
    return hapMatcherR, distinctHapsB, distinctHapsIE, np.asfortranarray(special), helper


@dataclass
class World:
    """Everything one synthetic region needs (panel + coordinates + transition rates)."""

    panel: cabi.Panel
    bits_common: np.ndarray  # [K_full, nSNPs] uint8 truth alleles at the common SNPs
    pos: np.ndarray  # [nSNPs] physical positions (common SNPs)
    L_grid: np.ndarray  # [T]
    transMatRate: np.ndarray  # [2, T-1]
    smooth_cm: np.ndarray  # [T-1]
    # all-SNP stage (None when not generated)
    bits_all: Optional[np.ndarray] = None  # [K_full, nSNPs_all]
    pos_all: Optional[np.ndarray] = None
    L_grid_all: Optional[np.ndarray] = None
    transMatRate_all: Optional[np.ndarray] = None
    smooth_cm_all: Optional[np.ndarray] = None

    @property
    def nSNPs(self):
        return self.panel.nSNPs

    @property
    def nGrids(self):
        return self.panel.nGrids

    @property
    def nSNPs_all(self):
        return self.panel.nSNPs_all

    @property
    def nGrids_all(self):
        return (self.panel.nSNPs_all + 31) // 32


def _grid_stuff(pos: np.ndarray, nGen: float, rate_cM_per_Mb: float = 1.0):
    n = pos.shape[0]
    T = (n + 31) // 32
    L_grid = np.array([int(np.mean(pos[32 * g : min(32 * (g + 1), n)])) for g in range(T)], dtype=np.int32)
    d_cM = np.diff(L_grid.astype(np.float64)) * rate_cM_per_Mb / 1e6
    sigma = np.exp(-nGen * d_cM / 100.0)
    tm = np.asfortranarray(np.stack([sigma, 1 - sigma], axis=0))
    smooth_cm = np.maximum(d_cM, 1e-8)
    return L_grid, tm, smooth_cm


def make_world(
    seed: int,
    K_full: int = 5008,
    nSNPs: int = 32000,
    region_bp: int = 3_000_000,
    n_founders: int = 64,
    switch_rate: float = 1e-4,
    flip_rate: float = 1e-3,
    nMaxDH: int = 255,
    ref_error: float = 1e-3,
    nGen: float = 100.0,
    all_snps_factor: int = 0,
    rare_max_carriers: int = 5,
) -> World:
    """all_snps_factor = 3 adds an all-SNP axis with nSNPs*3 sites, 1/3 of them the common ones."""
    rng = np.random.default_rng(seed)
    af = rng.beta(0.2, 0.2, size=nSNPs)
    founders = (rng.random((n_founders, nSNPs)) < af[None, :]).astype(np.uint8)
    # mosaic copying
    bits = np.empty((K_full, nSNPs), dtype=np.uint8)
    for k in range(K_full):
        sw = np.nonzero(rng.random(nSNPs) < switch_rate)[0]
        bounds = np.concatenate([[0], sw, [nSNPs]])
        src = rng.integers(0, n_founders, size=bounds.shape[0] - 1)
        seg = np.repeat(src, np.diff(bounds))
        bits[k] = founders[seg, np.arange(nSNPs)]
    flips = rng.random((K_full, nSNPs)) < flip_rate
    bits ^= flips.astype(np.uint8)
    rhb_t = pack_bits(bits)
    hm, dB, dIE, sp, helper = compress_rhb_t(rhb_t, nSNPs, nMaxDH, ref_error)

    if all_snps_factor and all_snps_factor > 1:
        nAll = nSNPs * all_snps_factor
        pos_all = np.sort(rng.choice(np.arange(1, region_bp), size=nAll, replace=False)).astype(np.int64)
        common_idx = np.sort(rng.choice(nAll, size=nSNPs, replace=False))
        snp_is_common = np.zeros(nAll, dtype=np.uint8)
        snp_is_common[common_idx] = 1
        common_snp_index = np.zeros(nAll, dtype=np.int32)
        common_snp_index[common_idx] = np.arange(1, nSNPs + 1)
        pos = pos_all[common_idx]
        rare_idx = np.nonzero(snp_is_common == 0)[0]
        bits_all = np.zeros((K_full, nAll), dtype=np.uint8)
        bits_all[:, common_idx] = bits
        per_hap = [[] for _ in range(K_full)]
        ncar = rng.integers(1, rare_max_carriers + 1, size=rare_idx.shape[0])
        for s, n in zip(rare_idx, ncar):
            ks = rng.choice(K_full, size=n, replace=False)
            bits_all[ks, s] = 1
            for k in ks:
                per_hap[k].append(s + 1)  # 1-based all-SNP index
        offs = np.zeros(K_full + 1, dtype=np.int64)
        offs[1:] = np.cumsum([len(x) for x in per_hap])
        snps = np.array([s for x in per_hap for s in sorted(x)], dtype=np.int32)
    else:
        pos = np.sort(rng.choice(np.arange(1, region_bp), size=nSNPs, replace=False)).astype(np.int64)
        snp_is_common = common_snp_index = offs = snps = None
        bits_all = pos_all = None

    panel = cabi.Panel(
        hapMatcherR=hm,
        distinctHapsB=dB,
        distinctHapsIE=dIE,
        special_matrix=sp,
        special_helper=helper,
        ref_error=ref_error,
        nSNPs=nSNPs,
        snp_is_common=snp_is_common,
        common_snp_index=common_snp_index,
        rare_hap_offsets=offs,
        rare_hap_snps=snps,
    )
    L_grid, tm, scm = _grid_stuff(pos, nGen)
    w = World(panel=panel, bits_common=bits, pos=pos, L_grid=L_grid, transMatRate=tm, smooth_cm=scm)
    if bits_all is not None:
        w.bits_all, w.pos_all = bits_all, pos_all
        w.L_grid_all, w.transMatRate_all, w.smooth_cm_all = _grid_stuff(pos_all, nGen)
    return w


@dataclass
class SampleReads:
    common: cabi.Reads
    all: Optional[cabi.Reads]
    truth_haps: np.ndarray  # [2 or 3, nSNPs(_all)] true alleles of the sample's haplotypes on the widest axis
    read_true_hap: np.ndarray


def make_sample_reads(
    world: World,
    seed: int,
    coverage: float = 1.0,
    read_len: int = 150,
    region_bp: int = 3_000_000,
    n_true_haps: int = 2,
    hap_probs: Optional[Tuple[float, ...]] = None,
    sample_switch_rate: float = 2e-5,
    n_reads: Optional[int] = None,
) -> SampleReads:
    """Reads copied from the sample's true (mosaic-of-panel) haplotypes with phred-scaled base error.

    When the world has an all-SNP axis the same physical reads are returned twice: restricted to the
    common SNPs (indices on the common axis) and over all SNPs (allSNP_sampleReads, functions.R:1051).
    """
    rng = np.random.default_rng(seed)
    use_all = world.bits_all is not None
    bits = world.bits_all if use_all else world.bits_common
    pos = world.pos_all if use_all else world.pos
    nS = bits.shape[1]
    K_full = bits.shape[0]
    truth = np.empty((n_true_haps, nS), dtype=np.uint8)
    for h in range(n_true_haps):
        sw = np.nonzero(rng.random(nS) < sample_switch_rate)[0]
        bounds = np.concatenate([[0], sw, [nS]])
        src = rng.integers(0, K_full, size=bounds.shape[0] - 1)
        seg = np.repeat(src, np.diff(bounds))
        truth[h] = bits[seg, np.arange(nS)]
    if n_reads is None:
        n_reads = int(round(coverage * region_bp / read_len))
    starts = np.sort(rng.integers(1, region_bp - read_len, size=n_reads))
    lo = np.searchsorted(pos, starts, side="left")
    hi = np.searchsorted(pos, starts + read_len, side="left")
    if hap_probs is None:
        hap_probs = tuple([1.0 / n_true_haps] * n_true_haps)
    src_hap = rng.choice(n_true_haps, size=n_reads, p=np.asarray(hap_probs))
    # flat (read, SNP) incidence on the widest axis; qualities / errors drawn once so both views agree
    lens = hi - lo
    read_of = np.repeat(np.arange(n_reads), lens)
    snp_of = np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)]) if lens.sum() else np.zeros(0, np.int64)
    q = rng.integers(20, 41, size=snp_of.shape[0])
    err = rng.random(snp_of.shape[0]) < 10.0 ** (-q / 10.0)
    obs = truth[src_hap[read_of], snp_of] ^ err.astype(np.uint8)
    bq = np.where(obs == 1, q, -q).astype(np.int32)

    def view(keep_mask, u_vals):
        r_of, u, b = read_of[keep_mask], u_vals[keep_mask], bq[keep_mask]
        cnt = np.bincount(r_of, minlength=n_reads)
        has = np.nonzero(cnt > 0)[0]
        offs = np.concatenate([[0], np.cumsum(cnt[has])])
        centre = offs[:-1] + cnt[has] // 2
        wif = (u[centre] // 32).astype(np.int32)
        order = np.argsort(wif, kind="stable")  # reads ordered by wif0 (gibbs-nipt.cpp:811 assumes non-decreasing)
        gather = np.concatenate([np.arange(offs[i], offs[i + 1]) for i in order]) if order.size else np.zeros(0, np.int64)
        new_offs = np.concatenate([[0], np.cumsum(cnt[has][order])]).astype(np.int32)
        return cabi.Reads(offsets=new_offs, u=u[gather], bq=b[gather], wif0=wif[order]), has[order]

    if use_all:
        is_c = world.panel.snp_is_common[snp_of] == 1
        rc, keep_c = view(is_c, world.panel.common_snp_index[snp_of].astype(np.int64) - 1)
        ra, _ = view(np.ones(snp_of.shape[0], dtype=bool), snp_of)
    else:
        rc, keep_c = view(np.ones(snp_of.shape[0], dtype=bool), snp_of)
        ra = None
    return SampleReads(common=rc, all=ra, truth_haps=truth, read_true_hap=src_hap[keep_c])


def make_call(
    world: World,
    reads: cabi.Reads,
    seed: int,
    K: int,
    all_snps: bool = False,
    which_haps_to_use: Optional[np.ndarray] = None,
    sort_haps: bool = True,
    first_iteration: bool = True,
    H0: Optional[np.ndarray] = None,
    n_burn_in: int = 20,
    n_sample: int = 1,
    block_its=(3, 6, 9),
    ff: float = 0.0,
    flags: Optional[int] = None,
    **kw,
) -> cabi.GibbsCall:
    """One Gibbs call with the production argument values (QUILT/R/functions.R:620-706, rare_common.R:325-391)
    and every uniform the reference would draw inside the call supplied from a seeded NumPy stream."""
    rng = np.random.default_rng(seed)
    R = reads.nReads
    T = world.nGrids_all if all_snps else world.nGrids
    nS = world.nSNPs_all if all_snps else world.nSNPs
    if which_haps_to_use is None:
        which_haps_to_use = rng.choice(world.panel.K_full, size=K, replace=False) + 1
        if sort_haps:
            which_haps_to_use = np.sort(which_haps_to_use)
    nh = 2 if ff == 0 else 3
    if H0 is None:
        if nh == 2:
            H0 = rng.integers(1, 3, size=R)
        else:
            H0 = rng.choice([1, 2, 3], size=R, p=[0.5, 0.5 - ff / 2, ff / 2])
    if flags is None:
        if ff == 0:
            flags = cabi.FLAGS_QUILT2_DIPLOID
        else:
            flags = (
                cabi.F_PERFORM_BLOCK_GIBBS
                | cabi.F_SHARD_CHECK_EVERY_PAIR
                | cabi.F_RESCALE_EMATREAD
                | cabi.F_RECORD_READ_SET
                | cabi.F_USE_SMOOTH_CM_IN_BLOCK_GIBBS
            )
        if all_snps:
            # rare_common.R:325-391 leaves disable_read_category_usage at its default TRUE (functions.R:2409)
            flags |= cabi.F_MAKE_EMATREAD_RARE_COMMON | cabi.F_DISABLE_READ_CATEGORY_USAGE
        elif first_iteration:
            flags |= cabi.F_GIBBS_INITIALIZE_ITERATIVELY  # functions.R:591
    n_full = n_burn_in + n_sample
    n_ep = len(block_its)
    return cabi.GibbsCall(
        panel=world.panel,
        reads=reads,
        which_haps_to_use=np.asarray(which_haps_to_use, dtype=np.int32),
        nGrids=T,
        nSNPs=nS,
        transMatRate_tc_H=world.transMatRate_all if all_snps else world.transMatRate,
        L_grid=world.L_grid_all if all_snps else world.L_grid,
        smooth_cm=world.smooth_cm_all if all_snps else world.smooth_cm,
        H0=np.asarray(H0, dtype=np.int32),
        runif_reads=rng.random(R * n_full),
        runif_block=rng.random(max(n_ep, 1) * R),
        runif_shard=rng.random(max(n_ep, 1) * max(T - 1, 1)),
        runif_H_class=rng.random(max(n_ep, 1) * R) if ff > 0 else None,
        ff=ff,
        n_gibbs_burn_in_its=n_burn_in,
        n_gibbs_sample_its=n_sample,
        block_gibbs_iterations=tuple(block_its),
        first_read_for_gibbs_initialization=int(rng.integers(0, max(R, 1))),
        flags=flags,
        **kw,
    )
