"""Batches of independent planning queries, sharded across the GPUs of one node.

SURVEY.md 8e: goal sweeps / many costmaps are independent units -> query q belongs to exactly
one rank (contiguous block partition), every rank runs the same single-GPU path on its shard,
and there is NO collective on the data path.  ``torch.distributed`` is only used, optionally,
to gather the (small) waypoint lists on rank 0 afterwards (NCCL or gloo, whatever the
process group was initialised with).  The reference has no counterpart: it plans one query
per interpreter (Coupled_motion_planner.py:1092).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition: ranks < n % world get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist_info():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), dist
    except Exception:
        pass
    return 0, 1, None


def solve_chunk_gpu(cost, goals, starts, tau: float = 0.5):
    """Default per-chunk worker: full-field solve + one path per query on the current CUDA
    device.  ``cost`` is one shared (rows, cols) map or (n, rows, cols) maps (numpy or torch)."""
    import torch
    from . import engine
    dev = torch.device("cuda", torch.cuda.current_device())
    c = cost if isinstance(cost, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(cost))
    c = c.to(dev, non_blocking=True)
    n = len(goals)
    T = engine.solve2d(c, np.asarray(goals, dtype=np.int32), nq=n, sync=False)
    out, cnt, st = engine.trace2d(T, np.asarray(starts, dtype=np.float64), np.asarray(goals, dtype=np.float64), tau)
    engine.finish(dev)
    # only the rows that were written travel to the host: the path slab is (n, 30002, 2) fp64 = 480 KB per query
    keep = torch.arange(out.shape[1], device=dev)[None, :] < cnt[:, None]
    flat = out[keep].cpu().numpy()
    cnt, st = cnt.cpu().numpy(), st.cpu().numpy()
    ends = np.cumsum(cnt)
    return [(flat[ends[i] - cnt[i]:ends[i]].copy(), int(st[i])) for i in range(n)]


def solve_queries(cost, goals: Sequence, starts: Sequence, tau: float = 0.5, chunk: int = 64,
                  gather: bool = False, solve_fn: Optional[Callable] = None):
    """Plan ``len(goals)`` independent queries (path from starts[q] to goals[q]).

    cost: (rows, cols) shared by all queries, or (Q, rows, cols) one map per query.
    Each rank processes its contiguous shard in chunks of ``chunk`` queries (bounds device
    memory: chunk * rows * cols * 8 B of fields).  Returns ``(lo, results)`` with
    ``results[i] = (path ndarray (K,2), status)`` for global query ``lo + i``; with
    ``gather=True`` rank 0 instead gets ``(0, all results in query order)`` and the other
    ranks ``(lo, their own)``.
    """
    rank, world, dist = _dist_info()
    Q = len(goals)
    if len(starts) != Q:
        raise ValueError("goals and starts must have the same length")
    per_query = hasattr(cost, "ndim") and cost.ndim == 3 or (hasattr(cost, "dim") and cost.dim() == 3)
    if per_query and cost.shape[0] != Q:
        raise ValueError("per-query costmaps must have one map per query")
    lo, hi = shard_bounds(Q, rank, world)
    fn = solve_fn or solve_chunk_gpu
    results: List = []
    for a in range(lo, hi, max(1, chunk)):
        b = min(hi, a + max(1, chunk))
        cc = cost[a:b] if per_query else cost
        results.extend(fn(cc, [list(g) for g in goals[a:b]], [list(s) for s in starts[a:b]], tau))
    if gather and dist is not None and world > 1:
        bucket = [None] * world if rank == 0 else None
        dist.gather_object((lo, results), bucket, dst=0)
        if rank == 0:
            merged: List = []
            for part_lo, part in sorted(bucket, key=lambda t: t[0]):
                assert part_lo == len(merged)
                merged.extend(part)
            return 0, merged
    return lo, results
