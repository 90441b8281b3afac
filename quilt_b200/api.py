"""Host-side mirror of the reference's interface for the Gibbs hot path, over the C ABI of libquiltgpu.so.

`rcpp_forwardBackwardGibbsNIPT(...)` keeps the reference's argument names and meaning
(QUILT/R/RcppExports.R:215-217, production call site QUILT/R/functions.R:2614-2678) for the arguments
that influence results on this path, and returns an object whose fields are named like the reference's
return list (QUILT/src/gibbs-nipt.cpp:3217-3306).  The random numbers the reference draws from R's RNG
inside the .Call are inputs here — the library never generates randomness.

There is no CPU fallback: if libquiltgpu.so is missing or no B200 is visible the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import cabi
from .cabi import GibbsCall, GibbsResult, Panel, Reads

_HERE = os.path.dirname(os.path.abspath(__file__))
# QUILT_B200_LIB: experiment builds only (tools/build_variant.py); the product library is libquiltgpu.so next to this file
SO_PATH = os.environ.get("QUILT_B200_LIB") or os.path.join(_HERE, "libquiltgpu.so")


class QuiltGpuError(RuntimeError):
    pass


class GpuLib(cabi._LibAPI):
    """ctypes binding of libquiltgpu.so (include/quilt_b200.h)."""

    prefix = "quilt_gpu"

    def __init__(self, path: str = SO_PATH):
        if not os.path.exists(path):
            raise QuiltGpuError(
                f"{path} is missing: build it with `python -m quilt_b200.build` (nvcc, sm_100a). There is no CPU fallback."
            )
        self.lib = C.CDLL(path)
        cabi.declare(self.lib, self.prefix)
        L = self.lib
        pa, po = C.POINTER(cabi.QuiltGibbsArgs), C.POINTER(cabi.QuiltGibbsOut)
        L.quilt_gpu_gibbs_batch.argtypes = [C.c_int32, pa, po]
        L.quilt_gpu_gibbs_batch.restype = C.c_int
        L.quilt_gpu_batch_stage.argtypes = [C.c_int32, pa, C.POINTER(C.c_void_p)]
        L.quilt_gpu_batch_stage.restype = C.c_int
        for f in ("run", "sync", "free"):
            getattr(L, f"quilt_gpu_batch_{f}").argtypes = [C.c_void_p]
            getattr(L, f"quilt_gpu_batch_{f}").restype = C.c_int
        L.quilt_gpu_batch_fetch.argtypes = [C.c_void_p, po]
        L.quilt_gpu_batch_fetch.restype = C.c_int
        L.quilt_gpu_batch_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.quilt_gpu_batch_timing.restype = C.c_int
        L.quilt_gpu_batch_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        L.quilt_gpu_batch_bytes.restype = C.c_int
        L.quilt_gpu_device_count.restype = C.c_int
        L.quilt_gpu_set_device.argtypes = [C.c_int32]
        L.quilt_gpu_set_device.restype = C.c_int
        L.quilt_gpu_last_error.restype = C.c_char_p
        L.quilt_gpu_kernel_launches.restype = C.c_int64
        L.quilt_gpu_release_panel_cache.restype = None
        L.quilt_gpu_section_timing.argtypes = [C.c_int32]
        L.quilt_gpu_section_timing.restype = C.c_int
        L.quilt_gpu_section_report.argtypes = [C.c_char_p, C.c_int64]
        L.quilt_gpu_section_report.restype = C.c_int64
        L.quilt_gpu_batch_chain_select.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
        L.quilt_gpu_batch_chain_select.restype = C.c_int
        L.quilt_gpu_gibbs_batch_chained.argtypes = [C.c_int32, pa, po, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double),
                                                   C.POINTER(C.c_void_p)]
        L.quilt_gpu_gibbs_batch_chained.restype = C.c_int
        L.quilt_gpu_gibbs_chain.argtypes = [C.c_int32, C.c_int32, C.POINTER(pa), C.POINTER(po), C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.POINTER(C.c_double))]
        L.quilt_gpu_gibbs_chain.restype = C.c_int
        L.quilt_gpu_batch_chain_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.quilt_gpu_batch_chain_timing.restype = C.c_int
        L.quilt_gpu_batch_which_haps.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        L.quilt_gpu_batch_which_haps.restype = C.c_int

    # ---- housekeeping
    def last_error(self) -> str:
        return (self.lib.quilt_gpu_last_error() or b"").decode()

    def device_count(self) -> int:
        return int(self.lib.quilt_gpu_device_count())

    def set_device(self, device: int):
        self._check(self.lib.quilt_gpu_set_device(int(device)), "quilt_gpu_set_device")

    def kernel_launches(self) -> int:
        return int(self.lib.quilt_gpu_kernel_launches())

    def release_panel_cache(self):
        self.lib.quilt_gpu_release_panel_cache()

    def section_timing(self, enable: bool = True):
        """the reference's print_extra_timing_information switch (copied-from-stitch.cpp:31-45): per-kernel device time"""
        self._check(self.lib.quilt_gpu_section_timing(int(bool(enable))), "quilt_gpu_section_timing")

    def section_report(self) -> str:
        n = int(self.lib.quilt_gpu_section_report(None, 0))
        buf = C.create_string_buffer(max(n, 1))
        self.lib.quilt_gpu_section_report(buf, n)
        return buf.value.decode()

    def _check(self, rc: int, what: str):
        if rc != cabi.OK:
            raise QuiltGpuError(f"{what} failed with status {rc}: {self.last_error()}")

    # ---- batches
    def gibbs_batch(self, calls: Sequence[GibbsCall]) -> List[GibbsResult]:
        """n independent calls through quilt_gpu_gibbs_batch (host buffers in, host buffers out)."""
        return self.run_prepared(self.prepare(calls))

    def prepare(self, calls: Sequence[GibbsCall], touch: bool = False):
        """argument / result structs and host result buffers for `run_prepared` (a caller that processes batch after
        batch keeps them, like R keeps its per-worker scratch matrices, QUILT/R/quilt.R:731-762)"""
        calls = list(calls)
        n = len(calls)
        args = (cabi.QuiltGibbsArgs * n)()
        outs = (cabi.QuiltGibbsOut * n)()
        for i, c in enumerate(calls):
            c.fill(args[i])
        res = [cabi.alloc_out(c, outs[i]) for i, c in enumerate(calls)]
        if touch:
            for r in res:
                r.hapProbs_t.fill(0)
                r.genProbsM_t.fill(0)
                r.genProbsF_t.fill(0)
        return calls, args, outs, res

    def run_prepared(self, prep) -> List[GibbsResult]:
        calls, args, outs, res = prep
        self._check(self.lib.quilt_gpu_gibbs_batch(len(calls), args, outs), "quilt_gpu_gibbs_batch")
        for i, r in enumerate(res):
            r.underflow_problem = bool(outs[i].underflow_problem)
            r.n_unif_consumed = int(outs[i].n_unif_consumed)
            r.underflow_iteration = int(outs[i].underflow_iteration)
        return res


class KeptBatch:
    """Handle of a batch whose results stay resident on the device after a host-buffer call (the `prev` of the next chained call)."""

    def __init__(self, lib: GpuLib, handle):
        self.lib, self._h = lib, handle

    def free(self):
        if self._h:
            self.lib.lib.quilt_gpu_batch_free(self._h)
            self._h = None


def run_prepared_chained(lib: GpuLib, prep, prev: Optional[KeptBatch], pad_unif, keep: bool, mspbwt_nindices=4, mspbwtL=3, mspbwtM=1):
    """quilt_gpu_gibbs_batch_chained: host buffers in / out with per-wave pipelining; the calls' haplotype lists come from the device-resident
    results of `prev` (None: from the arguments).  -> (results, KeptBatch or None)"""
    calls, args, outs, res = prep
    kept = C.c_void_p()
    pu = None if pad_unif is None else np.ascontiguousarray(pad_unif, dtype=np.float64)
    lib._check(lib.lib.quilt_gpu_gibbs_batch_chained(len(calls), args, outs, prev._h if prev is not None else None, mspbwt_nindices, mspbwtL, mspbwtM,
                                                     cabi._ptr(pu, cabi._pd), C.byref(kept) if keep else None), "quilt_gpu_gibbs_batch_chained")
    for i, r in enumerate(res):
        r.underflow_problem = bool(outs[i].underflow_problem)
        r.n_unif_consumed = int(outs[i].n_unif_consumed)
        r.underflow_iteration = int(outs[i].underflow_iteration)
    return res, (KeptBatch(lib, kept) if keep and kept.value else None)


def samples_summary(lib: GpuLib, samples, nSNPs: int):
    """quilt_gpu_samples_summary: per-sample dosage / gp_t / recast phasing haplotypes / phased GT and this rank's INFO counters, computed on
    the device from the device-resident hapProbs_t of batches that have run.
    samples: list of (stored_calls, phasing_call) with stored_calls = [(Batch-or-KeptBatch, job), ...] in the reference's accumulation order.
    -> (list of dicts, {"infoCount": [nSNPs, 2], "afCount": [nSNPs], "hweCount": [nSNPs, 3]})"""
    n = len(samples)
    arr = (cabi.QuiltSampleSummary * n)()
    keep, outs = [], []
    for i, (stored, phasing) in enumerate(samples):
        calls = (cabi.QuiltSummaryCall * len(stored))()
        for c, (b, j) in enumerate(stored):
            calls[c].batch, calls[c].job = b._h, j
        keep.append(calls)
        o = {"dosage": np.zeros(nSNPs), "gp_t": np.zeros((3, nSNPs), order="F"), "hd": np.zeros((nSNPs, 2), order="F"), "gt": np.zeros((nSNPs, 2), dtype=np.int8, order="F")}
        outs.append(o)
        a = arr[i]
        a.n_calls, a.calls = len(stored), calls
        a.phasing.batch, a.phasing.job = phasing[0]._h, phasing[1]
        a.dosage, a.gp_t, a.hd = cabi._ptr(o["dosage"], cabi._pd), cabi._ptr(o["gp_t"], cabi._pd), cabi._ptr(o["hd"], cabi._pd)
        a.gt = o["gt"].ctypes.data_as(C.POINTER(C.c_int8))
    cnt = {"infoCount": np.zeros((nSNPs, 2), order="F"), "afCount": np.zeros(nSNPs), "hweCount": np.zeros((nSNPs, 3), order="F")}
    fn = lib.lib.quilt_gpu_samples_summary
    fn.argtypes = [C.c_int32, C.c_int32, C.POINTER(cabi.QuiltSampleSummary), cabi._pd, cabi._pd, cabi._pd]
    fn.restype = C.c_int
    lib._check(fn(n, nSNPs, arr, cabi._ptr(cnt["infoCount"], cabi._pd), cabi._ptr(cnt["afCount"], cabi._pd), cabi._ptr(cnt["hweCount"], cabi._pd)),
               "quilt_gpu_samples_summary")
    return outs, cnt


def ingest_pileup(lib: GpuLib, offsets, u, bq, central_snp, grid, nGrids: int):
    """quilt_gpu_ingest_pileup: one sample's flat pileup -> the reads in the order the Gibbs path needs them + allele counts
    (QUILT/R/functions.R:243-316, :2779-2800).  -> dict(order, offsets, u, bq, wif0, first_read_of_grid, grid_has_read, alleleCount)"""
    offsets, u, bq, central_snp, grid = (np.ascontiguousarray(x, dtype=np.int32) for x in (offsets, u, bq, central_snp, grid))
    R, nS, nU = offsets.shape[0] - 1, grid.shape[0], int(offsets[-1])
    a = cabi.QuiltPileup(R, nS, nGrids, cabi._ptr(offsets, cabi._pi), cabi._ptr(u, cabi._pi), cabi._ptr(bq, cabi._pi), cabi._ptr(central_snp, cabi._pi), cabi._ptr(grid, cabi._pi))
    res = {"order": np.zeros(R, np.int32), "offsets": np.zeros(R + 1, np.int32), "u": np.zeros(nU, np.int32), "bq": np.zeros(nU, np.int32), "wif0": np.zeros(R, np.int32),
           "first_read_of_grid": np.zeros(nGrids + 1, np.int32), "grid_has_read": np.zeros(nGrids, np.uint8), "alleleCount": np.zeros((nS, 2), order="F")}
    o = cabi.QuiltIngestOut(*(cabi._ptr(res[k], cabi._pi) for k in ("order", "offsets", "u", "bq", "wif0", "first_read_of_grid")),
                            res["grid_has_read"].ctypes.data_as(C.POINTER(C.c_uint8)), cabi._ptr(res["alleleCount"], cabi._pd))
    fn = lib.lib.quilt_gpu_ingest_pileup
    fn.argtypes = [C.POINTER(cabi.QuiltPileup), C.POINTER(cabi.QuiltIngestOut)]
    fn.restype = C.c_int
    lib._check(fn(C.byref(a), C.byref(o)), "quilt_gpu_ingest_pileup")
    return res


def make_vcf_column(lib: GpuLib, gp_t: np.ndarray, hd: np.ndarray):
    """quilt_gpu_make_vcf_column: GT:GP:DS:HD text of one diploid sample (QUILT/R/functions.R:1408-1463) -> list of str, one per SNP"""
    gp_t = np.asfortranarray(gp_t, dtype=np.float64)
    hd = np.asfortranarray(hd, dtype=np.float64)
    nS = gp_t.shape[1]
    out = np.zeros((nS, cabi.VCF_RECORD), dtype=np.uint8)
    fn = lib.lib.quilt_gpu_make_vcf_column
    fn.argtypes = [C.c_int32, cabi._pd, cabi._pd, C.c_char_p]
    fn.restype = C.c_int
    lib._check(fn(nS, cabi._ptr(gp_t, cabi._pd), cabi._ptr(hd, cabi._pd), out.ctypes.data_as(C.c_char_p)), "quilt_gpu_make_vcf_column")
    return [bytes(r).decode("ascii") for r in out]


def run_chain_prepared(lib: GpuLib, preps, pads, mspbwt_nindices=4, mspbwtL=3, mspbwtM=1):
    """quilt_gpu_gibbs_chain: all stages of the call chain in one call (host buffers in / out).  preps: one lib.prepare(...) per stage;
    pads: [n x Ksubset] uniforms per link.  -> list of result lists"""
    ns = len(preps)
    n = len(preps[0][0])
    pa, po = C.POINTER(cabi.QuiltGibbsArgs), C.POINTER(cabi.QuiltGibbsOut)
    a_arr = (pa * ns)(*[C.cast(p[1], pa) for p in preps])
    o_arr = (po * ns)(*[C.cast(p[2], po) for p in preps])
    pus = [np.ascontiguousarray(x, dtype=np.float64) for x in pads]
    pd = C.POINTER(C.c_double)
    p_arr = (pd * max(ns - 1, 1))(*[cabi._ptr(x, cabi._pd) for x in pus])
    lib._check(lib.lib.quilt_gpu_gibbs_chain(ns, n, a_arr, o_arr, mspbwt_nindices, mspbwtL, mspbwtM, p_arr), "quilt_gpu_gibbs_chain")
    out = []
    for calls, args, outs, res in preps:
        for i, r in enumerate(res):
            r.underflow_problem = bool(outs[i].underflow_problem)
            r.n_unif_consumed = int(outs[i].n_unif_consumed)
            r.underflow_iteration = int(outs[i].underflow_iteration)
        out.append(res)
    return out


class Batch:
    """Staged form: inputs resident in HBM after __init__, `run()` only launches kernels."""

    def __init__(self, lib: GpuLib, calls: Sequence[GibbsCall]):
        self.lib = lib
        self.calls = list(calls)
        n = len(self.calls)
        self._args = (cabi.QuiltGibbsArgs * n)()
        for i, c in enumerate(self.calls):
            c.fill(self._args[i])
        self._h = C.c_void_p()
        lib._check(lib.lib.quilt_gpu_batch_stage(n, self._args, C.byref(self._h)), "quilt_gpu_batch_stage")

    def run(self):
        self.lib._check(self.lib.lib.quilt_gpu_batch_run(self._h), "quilt_gpu_batch_run")

    def sync(self):
        self.lib._check(self.lib.lib.quilt_gpu_batch_sync(self._h), "quilt_gpu_batch_sync")

    def timing(self):
        t, s, n = C.c_double(), C.c_double(), C.c_int32()
        self.lib._check(self.lib.lib.quilt_gpu_batch_timing(self._h, C.byref(t), C.byref(s), C.byref(n)), "quilt_gpu_batch_timing")
        return {"total_ms": t.value, "sweep_ms": s.value, "n_sweep_launches": n.value}

    def bytes(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_double()
        self.lib._check(self.lib.lib.quilt_gpu_batch_bytes(self._h, C.byref(a), C.byref(b), C.byref(c)), "quilt_gpu_batch_bytes")
        return {"h2d_bytes": a.value, "d2h_bytes": b.value, "sweep_algorithmic_bytes": c.value}

    def fetch(self) -> List[GibbsResult]:
        n = len(self.calls)
        outs = (cabi.QuiltGibbsOut * n)()
        res = [cabi.alloc_out(c, outs[i]) for i, c in enumerate(self.calls)]
        self.lib._check(self.lib.lib.quilt_gpu_batch_fetch(self._h, outs), "quilt_gpu_batch_fetch")
        for i, r in enumerate(res):
            r.underflow_problem = bool(outs[i].underflow_problem)
            r.n_unif_consumed = int(outs[i].n_unif_consumed)
            r.underflow_iteration = int(outs[i].underflow_iteration)
        return res

    def chain_select_into(self, nxt: "Batch", pad_unif: np.ndarray, mspbwt_nindices: int = 4, mspbwtL: int = 3, mspbwtM: int = 1):
        """select_new_haps_mspbwt_v3 on the device for every job of this (run) batch; the lists become the
        which_haps_to_use of the corresponding jobs of `nxt` (staged, not yet run): no host round trip (QUILT/R/functions.R:856-868)"""
        pu = np.ascontiguousarray(pad_unif, dtype=np.float64)
        assert pu.size >= len(self.calls) * nxt.calls[0].K
        self.lib._check(self.lib.lib.quilt_gpu_batch_chain_select(self._h, nxt._h, mspbwt_nindices, mspbwtL, mspbwtM, cabi._ptr(pu, cabi._pd)),
                        "quilt_gpu_batch_chain_select")

    def chain_ms(self) -> float:
        v = C.c_double()
        self.lib._check(self.lib.lib.quilt_gpu_batch_chain_timing(self._h, C.byref(v)), "quilt_gpu_batch_chain_timing")
        return v.value

    def which_haps(self, job: int) -> np.ndarray:
        out = np.zeros(self.calls[job].K, dtype=np.int32)
        self.lib._check(self.lib.lib.quilt_gpu_batch_which_haps(self._h, job, cabi._ptr(out, cabi._pi)), "quilt_gpu_batch_which_haps")
        return out

    def free(self):
        if self._h:
            self.lib.lib.quilt_gpu_batch_free(self._h)
            self._h = C.c_void_p()


_LIB: Optional[GpuLib] = None


def lib() -> GpuLib:
    global _LIB
    if _LIB is None:
        _LIB = GpuLib()
    return _LIB


def rcpp_forwardBackwardGibbsNIPT(
    sampleReads: Reads,
    *,
    panel: Panel,
    which_haps_to_use,
    transMatRate_tc_H,
    H,
    runif_reads,
    runif_block,
    runif_shard,
    L_grid,
    smooth_cm,
    nGrids: int,
    nSNPs: int,
    ff: float = 0.0,
    n_gibbs_burn_in_its: int = 20,
    n_gibbs_sample_its: int = 1,
    block_gibbs_iterations=(3, 6, 9),
    first_read_for_gibbs_initialization: int = 0,
    maxDifferenceBetweenReads: float = 1e10,
    Jmax: int = 10000,
    class_sum_cutoff: float = 0.06,
    shuffle_bin_radius: int = 5000,
    block_gibbs_quantile_prob: float = 0.95,
    sample_is_diploid: bool = True,
    gibbs_initialize_iteratively: bool = False,
    perform_block_gibbs: bool = True,
    do_shard_block_gibbs: bool = True,
    shard_check_every_pair: bool = True,
    disable_read_category_usage: bool = False,
    force_reset_read_category_zero: bool = False,
    make_eMatRead_t_rare_common: bool = False,
    rescale_eMatRead_t: bool = True,
    record_read_set: bool = True,
    use_smooth_cm_in_block_gibbs: bool = True,
    runif_H_class=None,
) -> GibbsResult:
    """One Gibbs call on the GPU; argument names follow the reference (functions.R:2614-2678)."""
    flags = 0
    for on, bit in (
        (sample_is_diploid, cabi.F_SAMPLE_IS_DIPLOID),
        (gibbs_initialize_iteratively, cabi.F_GIBBS_INITIALIZE_ITERATIVELY),
        (perform_block_gibbs, cabi.F_PERFORM_BLOCK_GIBBS),
        (do_shard_block_gibbs, cabi.F_DO_SHARD_BLOCK_GIBBS),
        (shard_check_every_pair, cabi.F_SHARD_CHECK_EVERY_PAIR),
        (disable_read_category_usage, cabi.F_DISABLE_READ_CATEGORY_USAGE),
        (force_reset_read_category_zero, cabi.F_FORCE_RESET_READ_CATEGORY_0),
        (make_eMatRead_t_rare_common, cabi.F_MAKE_EMATREAD_RARE_COMMON),
        (rescale_eMatRead_t, cabi.F_RESCALE_EMATREAD),
        (record_read_set, cabi.F_RECORD_READ_SET),
        (use_smooth_cm_in_block_gibbs, cabi.F_USE_SMOOTH_CM_IN_BLOCK_GIBBS),
    ):
        if on:
            flags |= bit
    call = GibbsCall(
        panel=panel,
        reads=sampleReads,
        which_haps_to_use=np.asarray(which_haps_to_use, dtype=np.int32),
        nGrids=nGrids,
        nSNPs=nSNPs,
        transMatRate_tc_H=transMatRate_tc_H,
        L_grid=L_grid,
        smooth_cm=smooth_cm,
        H0=np.asarray(H, dtype=np.int32),
        runif_reads=runif_reads,
        runif_block=runif_block,
        runif_shard=runif_shard,
        runif_H_class=runif_H_class,
        ff=ff,
        n_gibbs_burn_in_its=n_gibbs_burn_in_its,
        n_gibbs_sample_its=n_gibbs_sample_its,
        block_gibbs_iterations=tuple(block_gibbs_iterations),
        first_read_for_gibbs_initialization=first_read_for_gibbs_initialization,
        maxDifferenceBetweenReads=maxDifferenceBetweenReads,
        Jmax=Jmax,
        class_sum_cutoff=class_sum_cutoff,
        shuffle_bin_radius=shuffle_bin_radius,
        block_gibbs_quantile_prob=block_gibbs_quantile_prob,
        flags=flags,
    )
    return lib().gibbs(call)
