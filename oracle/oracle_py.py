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
        self.lib.quilt_oracle_backward_pair.argtypes = [C.c_int32, C.c_int32, pd, pd, pd, pu8, pd, pd]
        self.lib.quilt_oracle_forward_one_pair.argtypes = [C.c_int32, C.c_int32, pd, pd, pu8, C.c_int32, pd, pd, C.c_int32]

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
