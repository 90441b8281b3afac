"""ctypes loader for oracle/libquiltoracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from quilt_b200 import cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libquiltoracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "quilt_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "libquiltoracle.so"])
    return _SO


class Oracle(cabi._LibAPI):
    prefix = "quilt_oracle"

    def __init__(self):
        if not os.path.exists(_SO):
            build()
        self.lib = C.CDLL(_SO)
        cabi.declare(self.lib, self.prefix)
        pd, pu8 = C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        self.lib.quilt_oracle_select_haps_padded.argtypes = [C.POINTER(cabi.QuiltSelectArgs), pd, C.POINTER(C.c_int32)]
        self.lib.quilt_oracle_select_haps_padded.restype = C.c_int
        self.lib.quilt_oracle_backward_pair.argtypes = [C.c_int32, C.c_int32, pd, pd, pd, pu8, pd, pd]
        self.lib.quilt_oracle_forward_one_pair.argtypes = [C.c_int32, C.c_int32, pd, pd, pu8, C.c_int32, pd, pd, C.c_int32]

    def select_haps_padded(self, panel, hapProbs_t, Knew, pad_unif, nHap=2, mspbwt_nindices=4, mspbwtL=3, mspbwtM=1):
        """selection + the chain-mode completion of a short list (quilt_gpu_batch_chain_select's rule)"""
        hp = cabi.f64(hapProbs_t)
        a = cabi.QuiltSelectArgs()
        ps = panel.c_struct()
        a.panel = C.pointer(ps)
        a.nHap, a.hapProbs_t, a.Knew = nHap, cabi._ptr(hp, cabi._pd), Knew
        a.mspbwt_nindices, a.mspbwtL, a.mspbwtM = mspbwt_nindices, mspbwtL, mspbwtM
        pu = np.ascontiguousarray(pad_unif, dtype=np.float64)
        which = np.zeros(Knew, dtype=np.int32)
        rc = self.lib.quilt_oracle_select_haps_padded(C.byref(a), cabi._ptr(pu, cabi._pd), cabi._ptr(which, cabi._pi))
        if rc != cabi.OK:
            raise RuntimeError(f"quilt_oracle_select_haps_padded failed with status {rc}")
        return which

    def backward_pair(self, eMatGrid_t, tm, c, grid_has_read):
        e, t = cabi.f64(eMatGrid_t), cabi.f64(tm)
        K, T = e.shape
        cc = np.ascontiguousarray(c, dtype=np.float64)
        g = np.ascontiguousarray(grid_has_read, dtype=np.uint8)
        b1, b2 = np.zeros((K, T), order="F"), np.zeros((K, T), order="F")
        p = cabi._ptr
        self.lib.quilt_oracle_backward_pair(K, T, p(e, cabi._pd), p(t, cabi._pd), p(cc, cabi._pd), p(g, cabi._pu8), p(b1, cabi._pd), p(b2, cabi._pd))
        return b1, b2

    def forward_one(self, eMatGrid_t, tm, grid_has_read, iGrid, alphaHat_t, c, faster: bool):
        e, t = cabi.f64(eMatGrid_t), cabi.f64(tm)
        K, T = e.shape
        a = cabi.f64(alphaHat_t).copy(order="F")
        cc = np.array(c, dtype=np.float64)
        g = np.ascontiguousarray(grid_has_read, dtype=np.uint8)
        p = cabi._ptr
        self.lib.quilt_oracle_forward_one_pair(K, T, p(e, cabi._pd), p(t, cabi._pd), p(g, cabi._pu8), iGrid, p(a, cabi._pd), p(cc, cabi._pd), int(faster))
        return a, cc
