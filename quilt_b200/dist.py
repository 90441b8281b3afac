"""Multi-GPU plumbing: samples shard across ranks with no data-path collective.

The only exchange is the one-time broadcast of the prepared reference (and, for synthetic benchmarks, the
truth alleles reads are simulated from) from rank 0, over NCCL on a GPU box or gloo in the CPU tests, and a
max-reduction of the timed region.  One process per GPU; rendezvous comes from the torchrun environment.
"""
from __future__ import annotations

import os
from dataclasses import fields
from typing import List, Optional, Tuple

import numpy as np


def env_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: Optional[str] = None):
    """init_process_group from the torchrun environment; returns (rank, local_rank, world_size, device)."""
    import torch
    import torch.distributed as td

    rank, local_rank, world = env_world()
    if world > 1 and not td.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        td.init_process_group(backend=backend, rank=rank, world_size=world)
    dev = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")
    return rank, local_rank, world, dev


def sample_range(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous sample ranges per worker, the rule the reference uses for its forked workers
    (QUILT/R/quilt.R:690-692 via STITCH getSampleRange): the first (n mod W) workers take one extra sample"""
    base, extra = divmod(n_samples, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world_arrays(w) -> List[Tuple[str, str, Optional[np.ndarray]]]:
    out = []
    p = w.panel
    for name in ("hapMatcherR", "distinctHapsB", "distinctHapsIE", "special_matrix", "special_helper", "snp_is_common", "common_snp_index",
                 "rare_hap_offsets", "rare_hap_snps"):
        out.append(("panel", name, getattr(p, name)))
    for f in fields(w):
        if f.name == "panel":
            continue
        out.append(("world", f.name, getattr(w, f.name)))
    return out


def broadcast_world(w, src: int = 0, device=None):
    """Broadcast a synth.World (prepared reference + coordinates + truth alleles) from rank `src`.

    Rank `src` passes its world, the others pass None.  All array payloads travel as ONE uint8 tensor
    (NCCL broadcast over NVLink on a GPU box); shapes/dtypes travel as a small object list."""
    import torch
    import torch.distributed as td

    from . import cabi, synth

    if not td.is_initialized() or td.get_world_size() == 1:
        return w
    rank = td.get_rank()
    meta = [None]
    if rank == src:
        arrs = _world_arrays(w)
        meta[0] = {
            "arrays": [(k, n, None if a is None else (a.shape, a.dtype.str, bool(a.flags.f_contiguous and a.ndim > 1))) for k, n, a in arrs],
            "ref_error": w.panel.ref_error,
            "nSNPs": w.panel.nSNPs,
        }
    td.broadcast_object_list(meta, src=src)
    m = meta[0]
    sizes = []
    for _, _, d in m["arrays"]:
        sizes.append(0 if d is None else int(np.prod(d[0])) * np.dtype(d[1]).itemsize)
    offs = np.concatenate([[0], np.cumsum([(s + 15) // 16 * 16 for s in sizes])]).astype(np.int64)
    total = int(offs[-1])
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")
    if rank == src:
        host = np.zeros(total, dtype=np.uint8)
        for (k, n, a), o, s in zip(_world_arrays(w), offs[:-1], sizes):
            if a is not None and s:
                host[o : o + s] = np.frombuffer(a.tobytes(order="A"), dtype=np.uint8)
        blob = torch.from_numpy(host).to(device)
    else:
        blob = torch.empty(total, dtype=torch.uint8, device=device)
    td.broadcast(blob, src=src)
    if rank == src:
        return w
    host = blob.cpu().numpy()
    got = {"panel": {}, "world": {}}
    for (k, n, d), o, s in zip(m["arrays"], offs[:-1], sizes):
        if d is None:
            got[k][n] = None
            continue
        shape, dt, forder = d
        a = np.frombuffer(host[o : o + s].tobytes(), dtype=np.dtype(dt)).reshape(shape, order="F" if forder else "C")
        got[k][n] = a.copy(order="F" if forder else "C")
    pk = got["panel"]
    panel = cabi.Panel(
        hapMatcherR=pk["hapMatcherR"],
        distinctHapsB=pk["distinctHapsB"],
        distinctHapsIE=pk["distinctHapsIE"],
        special_matrix=pk["special_matrix"],
        special_helper=pk["special_helper"],
        ref_error=m["ref_error"],
        nSNPs=m["nSNPs"],
        snp_is_common=pk["snp_is_common"],
        common_snp_index=pk["common_snp_index"],
        rare_hap_offsets=pk["rare_hap_offsets"],
        rare_hap_snps=pk["rare_hap_snps"],
    )
    return synth.World(panel=panel, **got["world"])


def max_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as td

    if not td.is_initialized() or td.get_world_size() == 1:
        return float(x)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([x], dtype=torch.float64, device=device)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as td

    if not td.is_initialized() or td.get_world_size() == 1:
        return float(x)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([x], dtype=torch.float64, device=device)
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return float(t.item())


def allreduce_info_counts(counts: dict, device=None) -> dict:
    """The one data reduction of the path: every rank holds the INFO-score / allele-frequency / HWE counters of
    its own samples (infoCount [nSNPs x 2], afCount [nSNPs], hweCount [nSNPs x 3], alleleCount [nSNPs x 2]; QUILT/R/quilt.R:957-961) and
    the writer needs their sums over all samples (QUILT/R/writers.R:38-47).  One all-reduce (NCCL on a GPU box, gloo in the CPU tests) of
    the concatenated counters, < 10 MB."""
    import torch
    import torch.distributed as td

    if not td.is_initialized() or td.get_world_size() == 1:
        return {k: np.array(v, dtype=np.float64, copy=True) for k, v in counts.items()}
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")
    keys = sorted(counts)
    flat = np.concatenate([np.asarray(counts[k], dtype=np.float64).ravel(order="F") for k in keys])
    t = torch.from_numpy(flat).to(device)
    td.all_reduce(t, op=td.ReduceOp.SUM)
    flat = t.cpu().numpy()
    out, o = {}, 0
    for k in keys:
        a = np.asarray(counts[k])
        out[k] = flat[o:o + a.size].reshape(a.shape, order="F").copy()
        o += a.size
    return out


def info_scores(total: dict, N: int) -> dict:
    """the writer's finalisation of the summed counters (QUILT/R/writers.R:48-60): INFO score and estimated allele frequency"""
    theta = total["infoCount"][:, 0] / 2 / N
    denom = 2 * N * theta * (1 - theta)
    with np.errstate(divide="ignore", invalid="ignore"):
        info = 1 - total["infoCount"][:, 1] / denom
    info[(np.round(theta, 2) == 0) | (np.round(theta, 2) == 1)] = 1
    info[info < 0] = 0
    return {"info": info, "estimatedAlleleFrequency": total["afCount"] / N}


def barrier():
    import torch.distributed as td

    if td.is_initialized() and td.get_world_size() > 1:
        td.barrier()


def shutdown():
    import torch.distributed as td

    if td.is_initialized():
        td.destroy_process_group()
