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
    """int32 rank[c] = number of nodes popped before-or-with c (source = 0, unreached = INT32_MAX)."""
    flat = T.reshape(-1)
    order = torch.sort(flat, stable=True).indices
    rank = torch.empty(flat.numel(), dtype=torch.int32, device=T.device)
    rank[order] = torch.arange(flat.numel(), dtype=torch.int32, device=T.device)
    rank[~torch.isfinite(flat)] = torch.iinfo(torch.int32).max
    return rank.reshape(T.shape)


def truncate(T: torch.Tensor, cost: torch.Tensor, rank: torch.Tensor, k: int) -> torch.Tensor:
    """Partial field after k pops (accepted final, narrow band tentative, rest +inf), rebuilt on the
    device by libfm_b200's fmb_truncate{2d,3d}_f64 (csrc/truncate.cuh)."""
    T = T.contiguous()
    cost = cost.contiguous()
    rank = rank.contiguous()
    out = torch.empty_like(T)
    scratch = torch.empty(T.numel(), dtype=torch.int32, device=T.device)      # compacted narrow-band cells
    counters = torch.zeros(2, dtype=torch.int32, device=T.device)
    L = _capi.lib()
    stream = torch.cuda.current_stream().cuda_stream
    if T.dim() == 2:
        rc = L.fmb_truncate2d_f64(T.data_ptr(), cost.data_ptr(), rank.data_ptr(), T.shape[0], T.shape[1], int(k),
                                  out.data_ptr(), scratch.data_ptr(), counters.data_ptr(), stream)
    else:
        rc = L.fmb_truncate3d_f64(T.data_ptr(), cost.data_ptr(), rank.data_ptr(), T.shape[0], T.shape[1], T.shape[2],
                                  int(k), out.data_ptr(), scratch.data_ptr(), counters.data_ptr(), stream)
    _capi.check(rc)
    return out
