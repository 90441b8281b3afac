"""Binary, memory-mappable form of the prepared reference (second half).

The reference keeps the output of QUILT_prepare_reference as an RData blob that every worker load()s and deserialises
(QUILT/R/quilt-prepare-reference.R:484-525, loaded again in QUILT/R/quilt.R).  Here the arrays the path needs (cabi.Panel: hapMatcherR,
distinctHapsB, distinctHapsIE, eMatDH_special_matrix(+helper), rare/common extras) are written once, raw and 256-byte
aligned behind a small header, so that the ranks of a box map the same file pages read-only (np.memmap: no parse, no copy, one page-cache
copy for all ranks) and the arrays can be handed to the C ABI / uploaded as they lie.

Layout: magic "QB2REF01", uint32 header length, UTF-8 JSON header {"ref_error", "nSNPs", "arrays": [{name, dtype, shape, order, offset,
nbytes}]}, padding to 256 bytes, then the arrays in Fortran (column-major: R's) order at their offsets.  Little endian.
"""
from __future__ import annotations

import json
import struct

import numpy as np

from . import cabi

MAGIC = b"QB2REF01"
_ALIGN = 256
_NAMES = ("hapMatcherR", "distinctHapsB", "distinctHapsIE", "special_matrix", "special_helper", "snp_is_common", "common_snp_index", "rare_hap_offsets", "rare_hap_snps")


def save_panel(path: str, panel: cabi.Panel) -> int:
    """-> bytes written"""
    arrs = [(n, getattr(panel, n)) for n in _NAMES if getattr(panel, n) is not None]
    metas, off = [], 0
    for n, a in arrs:
        off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        metas.append({"name": n, "dtype": a.dtype.str, "shape": list(a.shape), "order": "F", "offset": off, "nbytes": int(a.nbytes)})
        off += a.nbytes
    hdr = json.dumps({"ref_error": panel.ref_error, "nSNPs": int(panel.nSNPs), "arrays": metas}).encode()
    base = (len(MAGIC) + 4 + len(hdr) + _ALIGN - 1) // _ALIGN * _ALIGN
    with open(path, "wb") as fh:
        fh.write(MAGIC + struct.pack("<I", len(hdr)) + hdr)
        fh.write(b"\0" * (base - fh.tell()))
        for (n, a), m in zip(arrs, metas):
            fh.write(b"\0" * (base + m["offset"] - fh.tell()))
            fh.write(np.asfortranarray(a).tobytes(order="F"))
        return fh.tell()


def load_panel(path: str, mmap: bool = True) -> cabi.Panel:
    with open(path, "rb") as fh:
        if fh.read(len(MAGIC)) != MAGIC:
            raise ValueError(f"{path}: not a prepared-reference pack")
        (n,) = struct.unpack("<I", fh.read(4))
        hdr = json.loads(fh.read(n).decode())
    base = (len(MAGIC) + 4 + n + _ALIGN - 1) // _ALIGN * _ALIGN
    kw = {}
    for m in hdr["arrays"]:
        shape = tuple(m["shape"])
        if mmap:
            a = np.memmap(path, dtype=np.dtype(m["dtype"]), mode="r", offset=base + m["offset"], shape=shape, order="F")
        else:
            a = np.fromfile(path, dtype=np.dtype(m["dtype"]), count=int(np.prod(shape)), offset=base + m["offset"]).reshape(shape, order="F")
        kw[m["name"]] = a
    return cabi.Panel(ref_error=hdr["ref_error"], nSNPs=hdr["nSNPs"], **kw)
