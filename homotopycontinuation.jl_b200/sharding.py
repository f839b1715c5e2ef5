"""Multi-GPU: paths are independent, so a batch is sharded by path index, one process per GPU, with
no collective on the hot path and one final gather of the PathResult SoA on rank 0.

This is the device-level counterpart of the reference's work distribution (file:line):
  threaded_solve   src/solve.jl:628-709   atomic next-index counter, results stored by path index
  many_solve       src/solve.jl:815-881   parameter sweeps: one solve per parameter point
Within a device the atomic counter lives in the kernel (device-side path queue); across devices the
index range is split into contiguous shards, which keeps the result order deterministic
(src/solve.jl:637, 670).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .capi import BatchResults


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous shard; shard sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def concat_results(parts: list[BatchResults]) -> BatchResults:
    """Concatenate per-shard results in rank order == path order."""
    n = parts[0].n
    out = {}
    for f in dataclasses.fields(BatchResults):
        if f.name in ("n", "N"):
            continue
        out[f.name] = np.concatenate([getattr(p, f.name) for p in parts], axis=0)
    return BatchResults(n=n, N=sum(p.N for p in parts), **out)


def gather_results(res: BatchResults, dist=None, dst: int = 0) -> BatchResults | None:
    """Final gather on `dst` (returns None on the other ranks).  `dist` is torch.distributed with an
    initialised process group (nccl on the GPU box, gloo in the CPU tests) or None for one process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return res
    world, rank = dist.get_world_size(), dist.get_rank()
    buf = [None] * world if rank == dst else None
    dist.gather_object(res, buf, dst=dst)
    return concat_results(buf) if rank == dst else None


def class_counts(res: BatchResults, mode: int = 0) -> dict:
    """Per-path class counts (before multiplicity clustering; `result.statistics` gives the reference's
    ResultStatistics).  mode 0 / 2: return codes are EndgameTrackerCode values (src/endgame_tracker.jl:100-117);
    mode 1 (plain Tracker batches): TrackerCode values (src/tracker.jl:166-176), where nothing is "at infinity" and
    every code but success is a failure."""
    ok = res.return_code == 1
    real = ok & (np.abs(res.solution.imag).max(axis=1) < 1e-6)       # src/path_result.jl:280-298
    at_inf = np.isin(res.return_code, (2, 3)) if mode != 1 else np.zeros(res.N, dtype=bool)
    return {"paths": int(res.N), "success": int(ok.sum()), "nonsingular": int((ok & (res.singular == 0)).sum()),
            "singular": int((ok & (res.singular == 1)).sum()), "real": int(real.sum()),
            "at_infinity": int(at_inf.sum()), "failed": int((~ok & ~at_inf).sum())}
