"""ctypes mirror of include/quilt_b200.h (the C ABI) plus numpy marshalling helpers.

The structures here are field-for-field copies of the header; both the GPU library
(quilt_b200/csrc -> libquiltgpu.so) and the CPU oracle (oracle/libquiltoracle.so, tests only)
take them.  Matrices are column-major (R / Armadillo) with K fastest — build numpy arrays with
order="F".
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

# flags (names follow param_list, QUILT/R/functions.R:2566-2599)
F_SAMPLE_IS_DIPLOID = 1 << 0
F_GIBBS_INITIALIZE_ITERATIVELY = 1 << 1
F_PERFORM_BLOCK_GIBBS = 1 << 2
F_DO_SHARD_BLOCK_GIBBS = 1 << 3
F_SHARD_CHECK_EVERY_PAIR = 1 << 4
F_DISABLE_READ_CATEGORY_USAGE = 1 << 5
F_FORCE_RESET_READ_CATEGORY_0 = 1 << 6
F_MAKE_EMATREAD_RARE_COMMON = 1 << 7
F_RESCALE_EMATREAD = 1 << 8
F_RECORD_READ_SET = 1 << 9
F_USE_SMOOTH_CM_IN_BLOCK_GIBBS = 1 << 10
F_RETURN_ALPHA = 1 << 11
F_RETURN_EXTRA = 1 << 12
F_GIBBS_INITIALIZE_AT_FIRST_READ = 1 << 13
F_OUTPUT_NO_PROBS = 1 << 14

FLAGS_QUILT2_DIPLOID = (
    F_SAMPLE_IS_DIPLOID
    | F_PERFORM_BLOCK_GIBBS
    | F_DO_SHARD_BLOCK_GIBBS
    | F_SHARD_CHECK_EVERY_PAIR
    | F_RESCALE_EMATREAD
    | F_RECORD_READ_SET
    | F_USE_SMOOTH_CM_IN_BLOCK_GIBBS
)

OK, ERR_CUDA, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_NO_DEVICE = 0, 1, 2, 3, 4

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)
_pu8 = C.POINTER(C.c_uint8)
_pi64 = C.POINTER(C.c_int64)
_pu32 = C.POINTER(C.c_uint32)


class QuiltPanel(C.Structure):
    _fields_ = [
        ("K_full", C.c_int32),
        ("nGrids", C.c_int32),
        ("nSNPs", C.c_int32),
        ("nMaxDH", C.c_int32),
        ("hapMatcherR", _pu8),
        ("distinctHapsB", _pi),
        ("distinctHapsIE", _pd),
        ("eMatDH_special_matrix", _pi),
        ("n_special", C.c_int32),
        ("eMatDH_special_matrix_helper", _pi),
        ("ref_error", C.c_double),
        ("nSNPs_all", C.c_int32),
        ("snp_is_common", _pu8),
        ("common_snp_index", _pi),
        ("rare_hap_offsets", _pi64),
        ("rare_hap_snps", _pi),
    ]


class QuiltReads(C.Structure):
    _fields_ = [
        ("nReads", C.c_int32),
        ("offsets", _pi),
        ("u", _pi),
        ("bq", _pi),
        ("wif0", _pi),
    ]


class QuiltGibbsArgs(C.Structure):
    _fields_ = [
        ("panel", C.POINTER(QuiltPanel)),
        ("reads", QuiltReads),
        ("K", C.c_int32),
        ("which_haps_to_use", _pi),
        ("nGrids", C.c_int32),
        ("nSNPs", C.c_int32),
        ("transMatRate_tc_H", _pd),
        ("L_grid", _pi),
        ("smooth_cm", _pd),
        ("ff", C.c_double),
        ("n_gibbs_burn_in_its", C.c_int32),
        ("n_gibbs_sample_its", C.c_int32),
        ("block_gibbs_iterations", _pi),
        ("n_block_gibbs_iterations", C.c_int32),
        ("H0", _pi),
        ("first_read_for_gibbs_initialization", C.c_int32),
        ("runif_reads", _pd),
        ("runif_block", _pd),
        ("runif_shard", _pd),
        ("runif_H_class", _pd),
        ("maxDifferenceBetweenReads", C.c_double),
        ("Jmax", C.c_int32),
        ("class_sum_cutoff", C.c_double),
        ("shuffle_bin_radius", C.c_int32),
        ("block_gibbs_quantile_prob", C.c_double),
        ("flags", C.c_uint32),
        ("unif_stream", _pd),
        ("n_unif_stream", C.c_int64),
    ]


class QuiltGibbsOut(C.Structure):
    _fields_ = [
        ("underflow_problem", C.c_int32),
        ("hapProbs_t", _pd),
        ("genProbsM_t", _pd),
        ("genProbsF_t", _pd),
        ("H", _pi),
        ("H_class", _pi),
        ("per_it_likelihoods", _pd),
        ("alphaHat_t", _pd * 3),
        ("betaHat_t", _pd * 3),
        ("eMatGrid_t", _pd * 3),
        ("c", _pd * 3),
        ("eMatRead_t", _pd),
        ("read_category", _pi),
        ("H_sample_its", _pi),
        ("n_unif_consumed", C.c_int64),
        ("underflow_iteration", C.c_int32),
    ]


class QuiltSelectArgs(C.Structure):
    _fields_ = [
        ("panel", C.POINTER(QuiltPanel)),
        ("nHap", C.c_int32),
        ("hapProbs_t", _pd),
        ("Knew", C.c_int32),
        ("mspbwt_nindices", C.c_int32),
        ("mspbwtL", C.c_int32),
        ("mspbwtM", C.c_int32),
    ]


class QuiltSummaryCall(C.Structure):
    _fields_ = [("batch", C.c_void_p), ("job", C.c_int32)]


class QuiltSampleSummary(C.Structure):
    _fields_ = [
        ("n_calls", C.c_int32),
        ("calls", C.POINTER(QuiltSummaryCall)),
        ("phasing", QuiltSummaryCall),
        ("dosage", _pd),
        ("gp_t", _pd),
        ("hd", _pd),
        ("gt", C.POINTER(C.c_int8)),
    ]


class QuiltPileup(C.Structure):
    _fields_ = [
        ("nReads", C.c_int32),
        ("nSNPs", C.c_int32),
        ("nGrids", C.c_int32),
        ("offsets", _pi),
        ("u", _pi),
        ("bq", _pi),
        ("central_snp", _pi),
        ("grid", _pi),
    ]


class QuiltIngestOut(C.Structure):
    _fields_ = [
        ("order", _pi),
        ("offsets", _pi),
        ("u", _pi),
        ("bq", _pi),
        ("wif0", _pi),
        ("first_read_of_grid", _pi),
        ("grid_has_read", C.POINTER(C.c_uint8)),
        ("alleleCount", _pd),
    ]


VCF_RECORD = 39

HF_RETURN_DOSAGE, HF_RETURN_BETAHAT, HF_RETURN_GAMMA, HF_GET_BEST_HAPS, HF_RETURN_ALPHAHAT = 1, 2, 4, 8, 16


class QuiltHaploidArgs(C.Structure):
    _fields_ = [
        ("panel", C.POINTER(QuiltPanel)),
        ("gl", _pd),
        ("transMatRate_t", _pd),
        ("gammaSmall_cols_to_get", _pi),
        ("n_thinned", C.c_int32),
        ("K_top_matches", C.c_int32),
        ("best_cap", C.c_int32),
        ("min_emission_prob_normalization_threshold", C.c_double),
        ("flags", C.c_uint32),
    ]


class QuiltHaploidOut(C.Structure):
    _fields_ = [
        ("dosage", _pd),
        ("c", _pd),
        ("alphaHat_t", _pd),
        ("betaHat_t", _pd),
        ("gamma_t", _pd),
        ("best_haps", _pi),
        ("best_haps_values", _pd),
        ("best_haps_count", _pi),
    ]


def _ptr(a: Optional[np.ndarray], typ):
    if a is None:
        return C.cast(None, typ)
    return a.ctypes.data_as(typ)


def f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "W"] + (["F"] if order == "F" else ["C"]))


def i32(a, order="F"):
    return np.require(np.asarray(a, dtype=np.int32), requirements=["A", "W"] + (["F"] if order == "F" else ["C"]))


@dataclass
class Panel:
    """Prepared reference in QUILT2's compressed form."""

    hapMatcherR: np.ndarray  # [K_full, T] uint8, F order
    distinctHapsB: np.ndarray  # [nMaxDH, T] int32
    distinctHapsIE: np.ndarray  # [nMaxDH, nSNPs] float64
    special_matrix: np.ndarray  # [n_special, 2] int32
    special_helper: np.ndarray  # [T, 2] int32
    ref_error: float
    nSNPs: int
    # rare/common extras
    snp_is_common: Optional[np.ndarray] = None  # [nSNPs_all] uint8
    common_snp_index: Optional[np.ndarray] = None  # [nSNPs_all] int32 (1-based)
    rare_hap_offsets: Optional[np.ndarray] = None  # [K_full+1] int64
    rare_hap_snps: Optional[np.ndarray] = None  # int32 (1-based all-SNP index)
    _c: Optional[QuiltPanel] = field(default=None, repr=False)

    def __post_init__(self):
        self.hapMatcherR = np.require(self.hapMatcherR, dtype=np.uint8, requirements=["F", "A"])
        self.distinctHapsB = i32(self.distinctHapsB)
        self.distinctHapsIE = f64(self.distinctHapsIE)
        self.special_matrix = i32(self.special_matrix.reshape(-1, 2) if self.special_matrix.size else np.zeros((1, 2)))
        self.special_helper = i32(self.special_helper)
        if self.snp_is_common is not None:
            self.snp_is_common = np.ascontiguousarray(self.snp_is_common, dtype=np.uint8)
            self.common_snp_index = np.ascontiguousarray(self.common_snp_index, dtype=np.int32)
            self.rare_hap_offsets = np.ascontiguousarray(self.rare_hap_offsets, dtype=np.int64)
            self.rare_hap_snps = np.ascontiguousarray(
                self.rare_hap_snps if self.rare_hap_snps.size else np.zeros(1), dtype=np.int32
            )

    @property
    def K_full(self):
        return self.hapMatcherR.shape[0]

    @property
    def nGrids(self):
        return self.hapMatcherR.shape[1]

    @property
    def nSNPs_all(self):
        return 0 if self.snp_is_common is None else int(self.snp_is_common.shape[0])

    def c_struct(self) -> QuiltPanel:
        if self._c is None:
            p = QuiltPanel()
            p.K_full = self.K_full
            p.nGrids = self.nGrids
            p.nSNPs = self.nSNPs
            p.nMaxDH = self.distinctHapsB.shape[0]
            p.hapMatcherR = _ptr(self.hapMatcherR, _pu8)
            p.distinctHapsB = _ptr(self.distinctHapsB, _pi)
            p.distinctHapsIE = _ptr(self.distinctHapsIE, _pd)
            p.eMatDH_special_matrix = _ptr(self.special_matrix, _pi)
            p.n_special = self.special_matrix.shape[0]
            p.eMatDH_special_matrix_helper = _ptr(self.special_helper, _pi)
            p.ref_error = self.ref_error
            p.nSNPs_all = self.nSNPs_all
            p.snp_is_common = _ptr(self.snp_is_common, _pu8)
            p.common_snp_index = _ptr(self.common_snp_index, _pi)
            p.rare_hap_offsets = _ptr(self.rare_hap_offsets, _pi64)
            p.rare_hap_snps = _ptr(self.rare_hap_snps, _pi)
            self._c = p
        return self._c


@dataclass
class Reads:
    """sampleReads flattened to CSR."""

    offsets: np.ndarray
    u: np.ndarray
    bq: np.ndarray
    wif0: np.ndarray

    def __post_init__(self):
        self.offsets = np.ascontiguousarray(self.offsets, dtype=np.int32)
        self.u = np.ascontiguousarray(self.u, dtype=np.int32)
        self.bq = np.ascontiguousarray(self.bq, dtype=np.int32)
        self.wif0 = np.ascontiguousarray(self.wif0, dtype=np.int32)

    @property
    def nReads(self):
        return int(self.wif0.shape[0])

    def c_struct(self) -> QuiltReads:
        r = QuiltReads()
        r.nReads = self.nReads
        r.offsets = _ptr(self.offsets, _pi)
        r.u = _ptr(self.u, _pi)
        r.bq = _ptr(self.bq, _pi)
        r.wif0 = _ptr(self.wif0, _pi)
        return r


@dataclass
class GibbsCall:
    """One rcpp_forwardBackwardGibbsNIPT call (argument meaning as QUILT/R/functions.R:2614-2678)."""

    panel: Panel
    reads: Reads
    which_haps_to_use: np.ndarray  # 1-based
    nGrids: int
    nSNPs: int
    transMatRate_tc_H: np.ndarray  # [2, T-1]
    L_grid: np.ndarray
    smooth_cm: np.ndarray
    H0: np.ndarray
    runif_reads: np.ndarray
    runif_block: np.ndarray
    runif_shard: np.ndarray
    runif_H_class: Optional[np.ndarray] = None
    unif_stream: Optional[np.ndarray] = None  # episode stream: replaces runif_block / runif_shard / runif_H_class (quilt_b200.h)
    ff: float = 0.0
    n_gibbs_burn_in_its: int = 20
    n_gibbs_sample_its: int = 1
    block_gibbs_iterations: Sequence[int] = (3, 6, 9)
    first_read_for_gibbs_initialization: int = 0
    maxDifferenceBetweenReads: float = 1e10
    Jmax: int = 10000
    class_sum_cutoff: float = 0.06
    shuffle_bin_radius: int = 5000
    block_gibbs_quantile_prob: float = 0.95
    flags: int = FLAGS_QUILT2_DIPLOID
    _keep: list = field(default_factory=list, repr=False)

    @property
    def K(self):
        return int(np.asarray(self.which_haps_to_use).shape[0])

    @property
    def n_full_its(self):
        return self.n_gibbs_burn_in_its + self.n_gibbs_sample_its

    def fill(self, a: QuiltGibbsArgs):
        self.which_haps_to_use = np.ascontiguousarray(self.which_haps_to_use, dtype=np.int32)
        self.transMatRate_tc_H = f64(self.transMatRate_tc_H)
        self.L_grid = np.ascontiguousarray(self.L_grid, dtype=np.int32)
        self.smooth_cm = np.ascontiguousarray(self.smooth_cm, dtype=np.float64)
        self.H0 = np.ascontiguousarray(self.H0, dtype=np.int32)
        self.runif_reads = np.ascontiguousarray(self.runif_reads, dtype=np.float64)
        self.runif_block = np.ascontiguousarray(self.runif_block, dtype=np.float64)
        self.runif_shard = np.ascontiguousarray(self.runif_shard, dtype=np.float64)
        if self.runif_H_class is not None:
            self.runif_H_class = np.ascontiguousarray(self.runif_H_class, dtype=np.float64)
        if self.unif_stream is not None:
            self.unif_stream = np.ascontiguousarray(self.unif_stream, dtype=np.float64)
        bgi = np.ascontiguousarray(np.asarray(self.block_gibbs_iterations, dtype=np.int32))
        self._keep = [bgi]
        a.panel = C.pointer(self.panel.c_struct())
        a.reads = self.reads.c_struct()
        a.K = self.K
        a.which_haps_to_use = _ptr(self.which_haps_to_use, _pi)
        a.nGrids = self.nGrids
        a.nSNPs = self.nSNPs
        a.transMatRate_tc_H = _ptr(self.transMatRate_tc_H, _pd)
        a.L_grid = _ptr(self.L_grid, _pi)
        a.smooth_cm = _ptr(self.smooth_cm, _pd)
        a.ff = self.ff
        a.n_gibbs_burn_in_its = self.n_gibbs_burn_in_its
        a.n_gibbs_sample_its = self.n_gibbs_sample_its
        a.block_gibbs_iterations = _ptr(bgi, _pi)
        a.n_block_gibbs_iterations = int(bgi.shape[0])
        a.H0 = _ptr(self.H0, _pi)
        a.first_read_for_gibbs_initialization = self.first_read_for_gibbs_initialization
        a.runif_reads = _ptr(self.runif_reads, _pd)
        a.runif_block = _ptr(self.runif_block, _pd)
        a.runif_shard = _ptr(self.runif_shard, _pd)
        a.runif_H_class = _ptr(self.runif_H_class, _pd)
        a.unif_stream = _ptr(self.unif_stream, _pd)
        a.n_unif_stream = 0 if self.unif_stream is None else int(self.unif_stream.shape[0])
        a.maxDifferenceBetweenReads = self.maxDifferenceBetweenReads
        a.Jmax = self.Jmax
        a.class_sum_cutoff = self.class_sum_cutoff
        a.shuffle_bin_radius = self.shuffle_bin_radius
        a.block_gibbs_quantile_prob = self.block_gibbs_quantile_prob
        a.flags = self.flags


@dataclass
class GibbsResult:
    """Named like the reference's return list (gibbs-nipt.cpp:3217-3306)."""

    underflow_problem: bool
    hapProbs_t: np.ndarray
    genProbsM_t: np.ndarray
    genProbsF_t: np.ndarray
    H: np.ndarray
    H_class: np.ndarray
    per_it_likelihoods: np.ndarray
    alphaHat_t: Optional[List[np.ndarray]] = None
    betaHat_t: Optional[List[np.ndarray]] = None
    eMatGrid_t: Optional[List[np.ndarray]] = None
    c: Optional[List[np.ndarray]] = None
    eMatRead_t: Optional[np.ndarray] = None
    read_category: Optional[np.ndarray] = None
    H_sample_its: Optional[np.ndarray] = None  # [nReads, n_gibbs_sample_its] labels after each sampling sweep
    n_unif_consumed: int = 0
    underflow_iteration: int = -1

    @property
    def double_list_of_ending_read_labels(self):
        if self.H_sample_its is not None:
            return [[self.H_sample_its[:, i] for i in range(self.H_sample_its.shape[1])]]
        return [[self.H]]

    @property
    def dosage(self):
        # functions.R:2719
        return self.genProbsM_t[1, :] + 2 * self.genProbsM_t[2, :]


def alloc_out(call: GibbsCall, o: QuiltGibbsOut) -> GibbsResult:
    nS, R, K, T = call.nSNPs, call.reads.nReads, call.K, call.nGrids
    nh = 2 if (call.flags & F_SAMPLE_IS_DIPLOID) else 3
    res = GibbsResult(
        underflow_problem=False,
        hapProbs_t=np.empty((3, nS), order="F"),
        genProbsM_t=np.empty((3, nS), order="F"),
        genProbsF_t=np.empty((3, nS), order="F"),
        H=np.zeros(R, dtype=np.int32),
        H_class=np.zeros(R, dtype=np.int32),
        per_it_likelihoods=np.zeros((1 if call.n_gibbs_sample_its == 0 else call.n_full_its, 13), order="F"),
        read_category=np.zeros(R, dtype=np.int32),
    )
    o.hapProbs_t = _ptr(res.hapProbs_t, _pd)
    o.genProbsM_t = _ptr(res.genProbsM_t, _pd)
    o.genProbsF_t = _ptr(res.genProbsF_t, _pd)
    o.H = _ptr(res.H, _pi)
    o.H_class = _ptr(res.H_class, _pi)
    o.per_it_likelihoods = _ptr(res.per_it_likelihoods, _pd)
    o.read_category = _ptr(res.read_category, _pi)
    if call.n_gibbs_sample_its > 0:
        res.H_sample_its = np.zeros((R, call.n_gibbs_sample_its), dtype=np.int32, order="F")
        o.H_sample_its = _ptr(res.H_sample_its, _pi)
    if call.flags & F_RETURN_ALPHA:
        res.alphaHat_t = [np.zeros((K, T), order="F") for _ in range(nh)]
        res.betaHat_t = [np.zeros((K, T), order="F") for _ in range(nh)]
        res.eMatGrid_t = [np.zeros((K, T), order="F") for _ in range(nh)]
        res.c = [np.zeros(T) for _ in range(nh)]
        for h in range(nh):
            o.alphaHat_t[h] = _ptr(res.alphaHat_t[h], _pd)
            o.betaHat_t[h] = _ptr(res.betaHat_t[h], _pd)
            o.eMatGrid_t[h] = _ptr(res.eMatGrid_t[h], _pd)
            o.c[h] = _ptr(res.c[h], _pd)
    if call.flags & F_RETURN_EXTRA:
        res.eMatRead_t = np.zeros((K, R), order="F")
        o.eMatRead_t = _ptr(res.eMatRead_t, _pd)
    return res


def declare(lib: C.CDLL, prefix: str):
    """Attach argtypes/restype for the entry points shared by both libraries."""
    pa, po = C.POINTER(QuiltGibbsArgs), C.POINTER(QuiltGibbsOut)
    getattr(lib, f"{prefix}_gibbs").argtypes = [pa, po]
    getattr(lib, f"{prefix}_gibbs").restype = C.c_int
    getattr(lib, f"{prefix}_make_eMatRead_t").argtypes = [pa, _pd, _pi]
    getattr(lib, f"{prefix}_make_eMatRead_t").restype = C.c_int
    getattr(lib, f"{prefix}_unpack_panel").argtypes = [C.POINTER(QuiltPanel), C.c_int32, _pi, C.c_int32, _pu32]
    getattr(lib, f"{prefix}_unpack_panel").restype = C.c_int
    getattr(lib, f"{prefix}_forward_backward").argtypes = [C.c_int32, C.c_int32, _pd, _pd, _pd, _pd, _pd]
    getattr(lib, f"{prefix}_forward_backward").restype = C.c_int
    if hasattr(lib, f"{prefix}_select_haps"):
        getattr(lib, f"{prefix}_select_haps").argtypes = [C.POINTER(QuiltSelectArgs), _pi, _pi, _pi]
        getattr(lib, f"{prefix}_select_haps").restype = C.c_int


class _LibAPI:
    """Calls shared by the GPU library and the oracle (same ABI, different prefix)."""

    prefix = ""
    lib: C.CDLL

    def gibbs(self, call: GibbsCall) -> GibbsResult:
        a, o = QuiltGibbsArgs(), QuiltGibbsOut()
        call.fill(a)
        res = alloc_out(call, o)
        rc = getattr(self.lib, f"{self.prefix}_gibbs")(C.byref(a), C.byref(o))
        if rc != OK:
            raise RuntimeError(f"{self.prefix}_gibbs failed with status {rc}: {self.last_error()}")
        res.underflow_problem = bool(o.underflow_problem)
        res.n_unif_consumed = int(o.n_unif_consumed)
        res.underflow_iteration = int(o.underflow_iteration)
        return res

    def make_eMatRead_t(self, call: GibbsCall):
        a = QuiltGibbsArgs()
        call.fill(a)
        e = np.zeros((call.K, call.reads.nReads), order="F")
        cat = np.zeros(call.reads.nReads, dtype=np.int32)
        rc = getattr(self.lib, f"{self.prefix}_make_eMatRead_t")(C.byref(a), _ptr(e, _pd), _ptr(cat, _pi))
        if rc != OK:
            raise RuntimeError(f"{self.prefix}_make_eMatRead_t failed with status {rc}: {self.last_error()}")
        return e, cat

    def unpack_panel(self, panel: Panel, which_haps_to_use, all_snps=False) -> np.ndarray:
        w = np.ascontiguousarray(which_haps_to_use, dtype=np.int32)
        nG = (panel.nSNPs_all + 31) // 32 if all_snps else panel.nGrids
        words = np.zeros((w.shape[0], nG), dtype=np.uint32, order="F")
        ps = panel.c_struct()
        rc = getattr(self.lib, f"{self.prefix}_unpack_panel")(
            C.byref(ps), w.shape[0], _ptr(w, _pi), int(all_snps), _ptr(words, _pu32)
        )
        if rc != OK:
            raise RuntimeError(f"{self.prefix}_unpack_panel failed with status {rc}: {self.last_error()}")
        return words

    def forward_backward(self, eMatGrid_t: np.ndarray, transMatRate_tc_H: np.ndarray):
        e = f64(eMatGrid_t)
        t = f64(transMatRate_tc_H)
        K, T = e.shape
        a, b, c = np.zeros((K, T), order="F"), np.zeros((K, T), order="F"), np.zeros(T)
        rc = getattr(self.lib, f"{self.prefix}_forward_backward")(
            K, T, _ptr(e, _pd), _ptr(t, _pd), _ptr(a, _pd), _ptr(b, _pd), _ptr(c, _pd)
        )
        if rc != OK:
            raise RuntimeError(f"{self.prefix}_forward_backward failed with status {rc}: {self.last_error()}")
        return a, b, c

    def select_haps(self, panel: Panel, hapProbs_t: np.ndarray, Knew: int, nHap: int = 2, mspbwt_nindices: int = 4, mspbwtL: int = 3, mspbwtM: int = 1):
        """select_new_haps_mspbwt_v3 (QUILT/R/mspbwt.R:230-474) -> (which_haps_to_use[:n_found] 1-based, n_unique); the caller pads a
        short list with sample() as mspbwt.R:381-399 does"""
        hp = f64(hapProbs_t)
        assert hp.shape == (3, panel.nSNPs)
        a = QuiltSelectArgs()
        ps = panel.c_struct()
        a.panel = C.pointer(ps)
        a.nHap, a.hapProbs_t, a.Knew = nHap, _ptr(hp, _pd), Knew
        a.mspbwt_nindices, a.mspbwtL, a.mspbwtM = mspbwt_nindices, mspbwtL, mspbwtM
        which = np.zeros(Knew, dtype=np.int32)
        nf, nu = C.c_int32(), C.c_int32()
        rc = getattr(self.lib, f"{self.prefix}_select_haps")(C.byref(a), _ptr(which, _pi), C.byref(nf), C.byref(nu))
        if rc != OK:
            raise RuntimeError(f"{self.prefix}_select_haps failed with status {rc}: {self.last_error()}")
        return which[: nf.value].copy(), nu.value

    def haploid_dosage(self, panel: Panel, gl, transMatRate_t, cols_to_get=None, K_top_matches=5, best_cap=32, threshold=1e-100,
                       flags=HF_RETURN_DOSAGE | HF_RETURN_BETAHAT | HF_RETURN_GAMMA | HF_RETURN_ALPHAHAT):
        """Rcpp_haploid_dosage_versus_refs (reference-single.cpp:2189-2413) for one haplotype; -> dict of outputs"""
        fn = getattr(self.lib, f"{self.prefix}_haploid_dosage_versus_refs")
        fn.argtypes = [C.POINTER(QuiltHaploidArgs), C.POINTER(QuiltHaploidOut)]
        fn.restype = C.c_int
        glf, tm = f64(gl), f64(transMatRate_t)
        K, T, nS = panel.K_full, panel.nGrids, panel.nSNPs
        assert glf.shape == (2, nS) and tm.shape == (2, T - 1)
        cols = np.full(T, -1, dtype=np.int32) if cols_to_get is None else np.ascontiguousarray(cols_to_get, dtype=np.int32)
        n_thin = int(cols.max()) + 1 if cols.size and cols.max() >= 0 else 0
        if n_thin > 0:
            flags |= HF_GET_BEST_HAPS
        a, o = QuiltHaploidArgs(), QuiltHaploidOut()
        ps = panel.c_struct()
        a.panel = C.pointer(ps)
        a.gl, a.transMatRate_t, a.gammaSmall_cols_to_get = _ptr(glf, _pd), _ptr(tm, _pd), _ptr(cols, _pi)
        a.n_thinned, a.K_top_matches, a.best_cap = n_thin, K_top_matches, best_cap
        a.min_emission_prob_normalization_threshold, a.flags = threshold, flags
        res = {"dosage": np.zeros(nS), "c": np.zeros(T)}
        o.dosage, o.c = _ptr(res["dosage"], _pd), _ptr(res["c"], _pd)
        for name, bit in (("alphaHat_t", HF_RETURN_ALPHAHAT), ("betaHat_t", HF_RETURN_BETAHAT), ("gamma_t", HF_RETURN_GAMMA)):
            if flags & bit:
                res[name] = np.zeros((K, T), order="F")
                setattr(o, name, _ptr(res[name], _pd))
        if n_thin > 0:
            res["best_haps"] = np.full((n_thin, best_cap), -1, dtype=np.int32)
            res["best_haps_values"] = np.zeros((n_thin, best_cap))
            res["best_haps_count"] = np.zeros(n_thin, dtype=np.int32)
            o.best_haps, o.best_haps_values, o.best_haps_count = _ptr(res["best_haps"], _pi), _ptr(res["best_haps_values"], _pd), _ptr(res["best_haps_count"], _pi)
        rc = fn(C.byref(a), C.byref(o))
        if rc != OK:
            raise RuntimeError(f"{self.prefix}_haploid_dosage_versus_refs failed with status {rc}: {self.last_error()}")
        return res

    def last_error(self) -> str:
        return ""
