"""Device-resident API of the B200 Fast Marching path (torch tensors in / out).

PyTorch is plumbing here: device memory, streams, host<->device copies.  All
numerics run in libfm_b200.so (hand-written sm_100a kernels) through the C ABI
of include/fm_b200.h.  There is no CPU fallback: without a CUDA device or
without the library every call raises.

Conventions are the reference's (src/FastMarching): fields are indexed
``[y, x]`` / ``[y, x, z]``, nodes are ``[x, y]`` / ``[x, y, z]``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _capi

_WS = {}            # (device index, stream) -> workspace tensor
_LAST_STATS = {}


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (this path has no CPU implementation)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    key = (device.index, _stream())
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def _as_seeds(seeds, nq: int, dim: int, device) -> torch.Tensor:
    if isinstance(seeds, torch.Tensor):
        s = seeds.to(device=device, dtype=torch.int32).reshape(-1, dim).contiguous()
    else:
        s = torch.tensor(seeds, dtype=torch.int32).reshape(-1, dim).to(device)
    if s.shape[0] != nq:
        raise ValueError(f"expected {nq} seeds, got {s.shape[0]}")
    return s


def _as_points(p, dim: int, device) -> torch.Tensor:
    if isinstance(p, torch.Tensor):
        return p.to(device=device, dtype=torch.float64).reshape(-1, dim).contiguous()
    import numpy as np
    return torch.from_numpy(np.ascontiguousarray(np.asarray(p, dtype=np.float64).reshape(-1, dim))).to(device)


def last_stats() -> dict:
    """Counters of the most recent finished solve on the current device."""
    return dict(_LAST_STATS.get(torch.cuda.current_device(), {}))


def finish(device=None) -> dict:
    """Synchronise, raise on device-side failure, return the solver counters."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    ws = _WS.get((dev.index, _stream()))
    if ws is None:
        torch.cuda.current_stream().synchronize()
        return {}
    st = _capi.FmbStats()
    _capi.check(_capi.lib().fmb_finish(ws.data_ptr(), ws.numel(), _stream(), C.byref(st)))
    _LAST_STATS[dev.index] = st.as_dict()
    return st.as_dict()


# ------------------------------------------------------------------ 2D ------
def solve2d(cost: torch.Tensor, seeds, out: Optional[torch.Tensor] = None, nq: Optional[int] = None,
            sync: bool = True) -> torch.Tensor:
    """Full total-cost field(s) of the Eikonal equation |grad T| = cost, T(seed) = 0.

    cost  (rows, cols): one map shared by every query, or (nq, rows, cols); float64/float32;
          +inf = obstacle.
    seeds (nq, 2) int [x, y].
    Returns T of shape (nq, rows, cols) (same dtype).  Replaces the sequential loop of
    FastMarching.py:92-112 / the two fronts of :114-162.
    """
    _require_cuda(cost, "cost")
    if cost.dtype not in (torch.float64, torch.float32):
        raise TypeError("cost must be float64 or float32")
    if cost.dim() == 2:
        shared = True
        rows, cols = cost.shape
        if nq is None:
            nq = len(seeds) if not isinstance(seeds, torch.Tensor) else int(seeds.reshape(-1, 2).shape[0])
    elif cost.dim() == 3:
        shared = False
        nq, rows, cols = cost.shape
    else:
        raise ValueError("cost must be (rows, cols) or (nq, rows, cols)")
    if cost.stride(-1) != 1:
        cost = cost.contiguous()
    dev = cost.device
    s = _as_seeds(seeds, nq, 2, dev)
    if out is None:
        out = torch.empty((nq, rows, cols), dtype=cost.dtype, device=dev)
    else:
        if out.shape != (nq, rows, cols) or out.dtype != cost.dtype or not out.is_contiguous():
            raise ValueError("out must be a contiguous (nq, rows, cols) tensor of the cost dtype")
    L = _capi.lib()
    nbytes = L.fmb_workspace_bytes_2d(rows, cols, nq)
    ws = _workspace(nbytes, dev)
    fn = L.fmb_solve2d_f64 if cost.dtype == torch.float64 else L.fmb_solve2d_f32
    cost_pitch = cost.stride(-2)
    cost_q = 0 if shared else cost.stride(0)
    with torch.cuda.device(dev):
        _capi.check(fn(cost.data_ptr(), cost_pitch, cost_q, out.data_ptr(), cols, rows * cols, rows, cols, nq,
                       s.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        if sync:
            finish(dev)
    return out


def trace2d(T: torch.Tensor, init, end, tau: float = 0.5, field_of_path=None,
            max_steps: Optional[int] = None, log_blocks: int = 0):
    """Gradient-descent paths over 2D field(s) (FastMarching.py:164-236).

    T (rows, cols) or (nf, rows, cols) float64; init/end (np, 2) [x, y] in cell units.
    Returns (paths (np, cap, 2) float64, count (np,) int32, status (np,) int32).  log_blocks > 0: also the log of the
    field cells each path read (fmb_trace2d_logged_f64): (..., blocks (np, log_blocks, 2) int32, nblocks (np,) int32).
    """
    _require_cuda(T, "T")
    if T.dtype != torch.float64:
        raise TypeError("trace2d needs a float64 field")
    if T.dim() == 2:
        T = T.unsqueeze(0)
    T = T.contiguous()
    nf, rows, cols = T.shape
    dev = T.device
    i = _as_points(init, 2, dev)
    e = _as_points(end, 2, dev)
    npaths = i.shape[0]
    if e.shape[0] != npaths:
        raise ValueError("init and end must have the same number of rows")
    if field_of_path is None:
        if nf == 1 and npaths > 1:
            fop = torch.zeros(npaths, dtype=torch.int32, device=dev)
        elif nf == npaths:
            fop = None
        else:
            raise ValueError("field_of_path required when #fields != #paths")
    else:
        fop = torch.as_tensor(field_of_path, dtype=torch.int32).to(dev).contiguous()
    if max_steps is None:
        max_steps = int(round(15000 / tau))
    cap = max_steps + 2
    out = torch.empty((npaths, cap, 2), dtype=torch.float64, device=dev)
    count = torch.empty(npaths, dtype=torch.int32, device=dev)
    status = torch.empty(npaths, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        if log_blocks > 0:
            blocks = torch.empty((npaths, log_blocks, 2), dtype=torch.int32, device=dev)
            nblocks = torch.empty(npaths, dtype=torch.int32, device=dev)
            _capi.check(_capi.lib().fmb_trace2d_logged_f64(T.data_ptr(), cols, rows * cols, rows, cols, npaths,
                                                           fop.data_ptr() if fop is not None else None,
                                                           i.data_ptr(), e.data_ptr(), float(tau), max_steps,
                                                           out.data_ptr(), cap, count.data_ptr(), status.data_ptr(),
                                                           blocks.data_ptr(), nblocks.data_ptr(), int(log_blocks), _stream()))
            return out, count, status, blocks, nblocks
        _capi.check(_capi.lib().fmb_trace2d_f64(T.data_ptr(), cols, rows * cols, rows, cols, npaths,
                                                fop.data_ptr() if fop is not None else None,
                                                i.data_ptr(), e.data_ptr(), float(tau), max_steps,
                                                out.data_ptr(), cap, count.data_ptr(), status.data_ptr(), _stream()))
    return out, count, status


# ------------------------------------------------------------------ 3D ------
def solve3d(cost: torch.Tensor, seeds, out: Optional[torch.Tensor] = None, nq: Optional[int] = None,
            sync: bool = True, exact: bool = False) -> torch.Tensor:
    """3D analogue of :func:`solve2d` (FastMarching3D.py:126-145, full field).

    cost (ny, nx, nz) shared or (nq, ny, nx, nz); seeds (nq, 3) [x, y, z].
    exact: solve with the reference's own rounding of ``**2`` on NumPy scalars (libm pow, csrc/pow2_glibc.cuh): the
    field is then the exact fixed point of the reference's arithmetic; needed where exact ties of the pop order matter
    (the early exit of ``FastMarching3D.computeTmap``).  True = the whole solve in that arithmetic
    (``fmb_solve3d_exact_f64``); "polish" = the fast solve followed by ``fmb_polish3d_f64``.
    """
    _require_cuda(cost, "cost")
    if cost.dtype not in (torch.float64, torch.float32):
        raise TypeError("cost must be float64 or float32")
    if cost.dim() == 3:
        shared = True
        ny, nx, nz = cost.shape
        if nq is None:
            nq = len(seeds) if not isinstance(seeds, torch.Tensor) else int(seeds.reshape(-1, 3).shape[0])
    elif cost.dim() == 4:
        shared = False
        nq, ny, nx, nz = cost.shape
    else:
        raise ValueError("cost must be (ny, nx, nz) or (nq, ny, nx, nz)")
    cost = cost.contiguous()
    dev = cost.device
    s = _as_seeds(seeds, nq, 3, dev)
    if out is None:
        out = torch.empty((nq, ny, nx, nz), dtype=cost.dtype, device=dev)
    L = _capi.lib()
    ws = _workspace(L.fmb_workspace_bytes_3d(ny, nx, nz, nq), dev)
    fn = L.fmb_solve3d_f64 if cost.dtype == torch.float64 else L.fmb_solve3d_f32
    with torch.cuda.device(dev):
        if exact == "polish":
            if cost.dtype != torch.float64:
                raise TypeError("the exact arithmetic exists for float64 only")
            _capi.check(fn(cost.data_ptr(), 0 if shared else ny * nx * nz, out.data_ptr(), ny * nx * nz, ny, nx, nz, nq,
                           s.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
            finish(dev)            # the polish pass reuses the workspace: the solve's own failures are reported first
            _capi.check(L.fmb_polish3d_f64(cost.data_ptr(), 0 if shared else ny * nx * nz, out.data_ptr(), ny * nx * nz,
                                           ny, nx, nz, nq, s.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        else:
            if exact:
                if cost.dtype != torch.float64:
                    raise TypeError("the exact arithmetic exists for float64 only")
                fn = L.fmb_solve3d_exact_f64
            _capi.check(fn(cost.data_ptr(), 0 if shared else ny * nx * nz, out.data_ptr(), ny * nx * nz, ny, nx, nz, nq,
                           s.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        if sync:
            finish(dev)
    return out


def trace3d(T: torch.Tensor, init, end, tau: float = 0.5, field_of_path=None,
            max_steps: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Gradient-descent paths over 3D field(s) (FastMarching3D.py:198-271)."""
    _require_cuda(T, "T")
    if T.dtype != torch.float64:
        raise TypeError("trace3d needs a float64 field")
    if T.dim() == 3:
        T = T.unsqueeze(0)
    T = T.contiguous()
    nf, ny, nx, nz = T.shape
    dev = T.device
    i = _as_points(init, 3, dev)
    e = _as_points(end, 3, dev)
    npaths = i.shape[0]
    if field_of_path is None:
        if nf == 1 and npaths > 1:
            fop = torch.zeros(npaths, dtype=torch.int32, device=dev)
        elif nf == npaths:
            fop = None
        else:
            raise ValueError("field_of_path required when #fields != #paths")
    else:
        fop = torch.as_tensor(field_of_path, dtype=torch.int32).to(dev).contiguous()
    if max_steps is None:
        max_steps = int(round(15000 / tau))
    cap = max_steps + 2
    out = torch.empty((npaths, cap, 3), dtype=torch.float64, device=dev)
    count = torch.empty(npaths, dtype=torch.int32, device=dev)
    status = torch.empty(npaths, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().fmb_trace3d_f64(T.data_ptr(), ny * nx * nz, ny, nx, nz, npaths,
                                                fop.data_ptr() if fop is not None else None,
                                                i.data_ptr(), e.data_ptr(), float(tau), max_steps,
                                                out.data_ptr(), cap, count.data_ptr(), status.data_ptr(), _stream()))
    return out, count, status
