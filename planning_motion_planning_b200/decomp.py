"""Row-slab domain decomposition of ONE very large 2D map (optional).

Each slab owns a contiguous block of rows plus a one-row halo on every interior side.  A halo
row is a copy of the neighbour slab's boundary row: locally it has cost +inf (so it is never
relaxed here) but its T values feed the slab's first/last interior row.  The solve alternates

    resolve every slab whose halo improved (``fmb_resolve2d_f64``: continue from the current T)
    exchange boundary rows with the neighbours (64 KiB per 8192-wide fp64 row)

until no halo row improves anywhere.  The exchange is the only communication: NCCL point-to-point
over NVLink between ranks (``solve2d_slabs_dist``), or plain device copies when one process holds
all slabs (``solve2d_slabs_local``).  A single-source front sweeps through the slabs one after
another, so this buys capacity, not speed; one B200 holds 8192^2 outright.
The reference has no counterpart (single process, FastMarching.py:92-112).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np

INF = float("inf")


def slab_bounds(rows: int, nslabs: int, s: int, align: int = 32) -> Tuple[int, int]:
    """Rows [lo, hi) of slab s; interior cuts are multiples of `align` (the tile height)."""
    if nslabs < 1 or not (0 <= s < nslabs):
        raise ValueError("bad slab index")
    blocks = (rows + align - 1) // align
    base, extra = divmod(blocks, nslabs)
    b_lo = s * base + min(s, extra)
    b_hi = b_lo + base + (1 if s < extra else 0)
    return min(rows, b_lo * align), min(rows, b_hi * align)


class Slab:
    """One slab's local arrays (any array namespace with numpy-like indexing: numpy or torch)."""

    def __init__(self, cost, lo: int, hi: int, goal, xp):
        rows, cols = cost.shape
        self.lo, self.hi, self.cols = lo, hi, cols
        self.th = 1 if lo > 0 else 0            # halo row above / below
        self.bh = 1 if hi < rows else 0
        self.xp = xp
        self.cost = cost[lo - self.th:hi + self.bh].clone() if hasattr(cost, "clone") else cost[lo - self.th:hi + self.bh].copy()
        if self.th:
            self.cost[0] = INF
        if self.bh:
            self.cost[-1] = INF
        self.T = xp.full_like(self.cost, INF)
        gx, gy = int(goal[0]), int(goal[1])
        self.seed = [gx, gy - lo + self.th] if lo <= gy < hi else [-1, -1]
        self.first = True
        self.top_dirty = self.bot_dirty = False

    @property
    def halo_rows(self) -> int:
        return (1 if self.th else 0) | (2 if self.bh else 0)

    def needs_work(self) -> bool:
        return (self.first and self.seed[0] >= 0) or self.top_dirty or self.bot_dirty

    def run(self, resolve_fn: Callable):
        activate = (1 if self.top_dirty else 0) | (2 if self.bot_dirty else 0)
        resolve_fn(self.cost, self.T, self.seed if self.first else [-1, -1], activate, self.halo_rows)
        self.first = False
        self.top_dirty = self.bot_dirty = False

    def boundary_rows(self):
        """(first interior row, last interior row): what the neighbours need."""
        n = self.T.shape[0]
        return self.T[self.th], self.T[n - 1 - self.bh]

    def apply_halo(self, from_above=None, from_below=None):
        """Install received rows; remember which side improved."""
        if self.th and from_above is not None:
            if bool((from_above < self.T[0]).any()):
                self.T[0] = self.xp.minimum(self.T[0], from_above)
                self.top_dirty = True
        if self.bh and from_below is not None:
            if bool((from_below < self.T[-1]).any()):
                self.T[-1] = self.xp.minimum(self.T[-1], from_below)
                self.bot_dirty = True

    def interior(self):
        n = self.T.shape[0]
        return self.T[self.th:n - self.bh]


def gpu_resolve(cost, T, seed, activate: int, halo_rows: int):
    """resolve_fn on the current CUDA device (torch tensors), through the C ABI."""
    import torch
    from . import _capi, engine
    rows, cols = T.shape
    L = _capi.lib()
    ws = engine._workspace(L.fmb_workspace_bytes_2d(rows, cols, 1), T.device)
    sd = torch.tensor(seed, dtype=torch.int32, device=T.device)
    _capi.check(L.fmb_resolve2d_f64(cost.data_ptr(), cost.stride(0), T.data_ptr(), T.stride(0), rows, cols,
                                    sd.data_ptr(), int(activate), int(halo_rows), ws.data_ptr(), ws.numel(),
                                    torch.cuda.current_stream().cuda_stream))
    engine.finish(T.device)


def solve2d_slabs_local(cost, goal, nslabs: int, resolve_fn: Optional[Callable] = None, max_rounds: int = 10000):
    """All slabs in one process (one GPU, or the CPU emulator in tests).  Returns (T, rounds)."""
    try:
        import torch
        is_torch = isinstance(cost, torch.Tensor)
    except Exception:
        is_torch = False
    if is_torch:
        import torch as xp
    else:
        xp = np
    resolve_fn = resolve_fn or gpu_resolve
    rows = cost.shape[0]
    slabs = [Slab(cost, *slab_bounds(rows, nslabs, s), goal, xp) for s in range(nslabs)]
    slabs = [s for s in slabs if s.hi > s.lo]
    rounds = 0
    while any(s.needs_work() for s in slabs):
        rounds += 1
        if rounds > max_rounds:
            raise RuntimeError("domain decomposition did not converge")
        for s in slabs:
            if s.needs_work():
                s.run(resolve_fn)
        for a, b in zip(slabs[:-1], slabs[1:]):          # a above b
            a_last = a.boundary_rows()[1]
            b_first = b.boundary_rows()[0]
            a.apply_halo(from_below=b_first)
            b.apply_halo(from_above=a_last)
    T = xp.cat([s.interior() for s in slabs], 0) if is_torch else np.concatenate([s.interior() for s in slabs], 0)
    return T, rounds


def solve2d_slabs_dist(cost, goal, resolve_fn: Optional[Callable] = None, max_rounds: int = 10000):
    """One slab per rank of the initialised torch.distributed group.  `cost` is the full map on every
    rank (torch tensor on the rank's device, or CPU for gloo tests); returns (lo, hi, T rows, rounds).
    Communication per round: one row to/from each neighbour (P2P) + a 1-int all-reduce."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    resolve_fn = resolve_fn or gpu_resolve
    rows, cols = cost.shape
    lo, hi = slab_bounds(rows, world, rank)
    slab = Slab(cost, lo, hi, goal, torch) if hi > lo else None
    # neighbours that actually own rows
    owners = [r for r in range(world) if slab_bounds(rows, world, r)[1] > slab_bounds(rows, world, r)[0]]
    up = owners[owners.index(rank) - 1] if slab is not None and owners.index(rank) > 0 else None
    down = owners[owners.index(rank) + 1] if slab is not None and owners.index(rank) + 1 < len(owners) else None
    dev = cost.device
    rounds = 0
    while True:
        work = torch.tensor([1 if (slab is not None and slab.needs_work()) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(work, op=dist.ReduceOp.MAX)
        if int(work[0]) == 0:
            break
        rounds += 1
        if rounds > max_rounds:
            raise RuntimeError("domain decomposition did not converge")
        if slab is not None and slab.needs_work():
            slab.run(resolve_fn)
        if slab is not None:
            first, last = slab.boundary_rows()
            first, last = first.contiguous(), last.contiguous()
            r_up = torch.empty(cols, dtype=cost.dtype, device=dev)
            r_dn = torch.empty(cols, dtype=cost.dtype, device=dev)
            ops = []
            if up is not None:
                ops += [dist.P2POp(dist.isend, first, up), dist.P2POp(dist.irecv, r_up, up)]
            if down is not None:
                ops += [dist.P2POp(dist.isend, last, down), dist.P2POp(dist.irecv, r_dn, down)]
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            slab.apply_halo(from_above=r_up if up is not None else None, from_below=r_dn if down is not None else None)
    if slab is None:
        return lo, hi, torch.empty((0, cols), dtype=cost.dtype, device=dev), rounds
    return lo, hi, slab.interior(), rounds
