"""ctypes loader for oracle/_ref/libquiltref.so — TEST INFRASTRUCTURE ONLY.

The library is the reference's own C++ (the unmodified files under /root/reference/QUILT/src) compiled against the
header-only RcppArmadillo stand-in in oracle/refshim/ (recipe: oracle/refshim/Makefile).  It can only be (re)built
where /root/reference exists (the build container); the GPU box uses the prebuilt .so that travels with the snapshot.

Only tests/, tools/make_golden.py, __graft_entry__ and bench.py (cpu_baseline / --impl reference) may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from quilt_b200 import cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libquiltref.so")
REFERENCE_SRC = "/root/reference/QUILT/src"


def available() -> bool:
    return os.path.exists(_SO)


def can_build() -> bool:
    return os.path.isdir(REFERENCE_SRC)


def build(force: bool = False) -> str:
    """make -C oracle/refshim; a no-op (the prebuilt library is kept) when /root/reference is absent."""
    if not can_build():
        if not available():
            raise RuntimeError("oracle/_ref/libquiltref.so is missing and /root/reference is not present to build it")
        return _SO
    cmd = ["make", "-s", "-C", os.path.join(_HERE, "refshim"), "-j8"] + (["-B"] if force else [])
    subprocess.check_call(cmd)
    return _SO


class Ref(cabi._LibAPI):
    prefix = "quilt_ref"

    def __init__(self):
        if not available():
            build()
        self.lib = C.CDLL(_SO)
        cabi.declare(self.lib, self.prefix)
        self.lib.quilt_ref_last_error.restype = C.c_char_p

    def last_error(self) -> str:
        return (self.lib.quilt_ref_last_error() or b"").decode()
