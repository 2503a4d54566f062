"""Helpers shared by the two drop-in modules: argument normalisation and the
emulation of the reference's early-exit ("partial field") semantics."""
from __future__ import annotations

import numpy as np
import torch

from planning_motion_planning_b200 import _capi, engine

_EXC = {
    _capi.TRACE_VALUEERROR: (ValueError, "cannot convert float NaN to integer"),
    _capi.TRACE_INDEXERROR: (IndexError, "index out of bounds for the field"),
    _capi.TRACE_OVERFLOW: (OverflowError, "cannot convert float infinity to integer"),
}


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("FastMarching (B200 build) needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_c_field(a):
    """Return (C-contiguous float64 view or copy, transposed?) for an array of any order.

    The planner passes an F-ordered view (``cMap.T``, Coupled_motion_planner.py:1226).  The
    4-neighbour update is symmetric in its two axes (FastMarching.py:17-29), so an F-ordered
    map is solved as its C-ordered transpose with x and y swapped -- no copy.
    """
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2 and a.flags.f_contiguous and not a.flags.c_contiguous:
        return a.T, True
    return np.ascontiguousarray(a), False


def node2(p, swap):
    x, y = int(p[0]), int(p[1])
    return (y, x) if swap else (x, y)


def check_node2(p, rows, cols):
    x, y = p
    if not (0 <= x < cols and 0 <= y < rows):
        # the reference wraps negative indices / raises IndexError at the array edge
        raise IndexError(f"node {list(p)} is outside the {rows}x{cols} map")


def raise_trace(status: int):
    if status in _EXC:
        exc, msg = _EXC[status]
        raise exc(msg)


# --------------------------------------------------------------------------
# Early-exit emulation.  The reference stops popping as soon as
# the start node is accepted (FastMarching3D.py:141-142) or the two fronts meet
# (FastMarching.py:150-155) and returns a PARTIAL field: accepted cells hold final
# values, narrow-band cells hold their last tentative value, the rest is +inf.
# The solver always produces the full field; pop ranks are recovered by a stable
# sort of T (ties in row-major order) and the partial field is rebuilt from them.
def pop_ranks(T: torch.Tensor) -> torch.Tensor:
    """rank[c] = number of nodes popped before-or-with c (source = 0, unreached = huge)."""
    flat = T.reshape(-1)
    order = torch.sort(flat, stable=True).indices
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=T.device)
    rank[~torch.isfinite(flat)] = torch.iinfo(torch.int64).max
    return rank.reshape(T.shape)


def accepted_neighbour(acc: torch.Tensor) -> torch.Tensor:
    """True where at least one face neighbour is accepted (any number of dims)."""
    nb = torch.zeros_like(acc)
    for d in range(acc.dim()):
        n = acc.shape[d]
        if n < 2:
            continue
        lo = [slice(None)] * acc.dim()
        hi = [slice(None)] * acc.dim()
        lo[d] = slice(0, n - 1)
        hi[d] = slice(1, n)
        nb[tuple(hi)] |= acc[tuple(lo)]
        nb[tuple(lo)] |= acc[tuple(hi)]
    return nb
