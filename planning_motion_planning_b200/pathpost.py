"""Path post-processing on the device: what the planner does to the traced paths right after the tracers.

  stitch_rover_paths[_device]   Coupled_motion_planner.py:1232-1234
      roverPath = resolution * (vstack(flipud(pathS), pathG[1:]) + 1)
  smooth_resample_arm[_device]  Coupled_motion_planner.py:1641-1671
      per-axis scaling, scipy.signal.savgol_filter(., 11, 3) (mode='interp'), shift back to the global frame,
      last row := the sample pose, interp1d(range(n), .)(linspace(0, n - 1, m))

The ``_device`` forms take the tracer's own output tensors (``engine.trace2d`` / ``trace3d``: paths (np, cap, D),
count (np,)) and stay on the GPU: a batch of queries hands m x 3 / (nS + nG - 1) x 2 waypoints to the host instead of
30 002-row slabs.  Kernels: csrc/pathpost.cuh through fmb_path_stitch2d_f64 / fmb_path_post3d_f64.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi
from .engine import _require_cuda, _stream

_WS = {}


def _ws(dev) -> torch.Tensor:
    key = (dev.index, _stream())
    w = _WS.get(key)
    if w is None:
        w = torch.empty(max(1024, _capi.lib().fmb_workspace_bytes_pathpost()), dtype=torch.uint8, device=dev)
        _WS[key] = w
    return w


def stitch_rover_paths_device(pathS: torch.Tensor, countS: torch.Tensor, pathG: torch.Tensor, countG: torch.Tensor,
                              resolution: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """pathS / pathG (np, cap, 2) float64 tracer slabs with their row counts -> (out (np, 2 * cap, 2), count (np,))."""
    for t, nme in ((pathS, "pathS"), (pathG, "pathG"), (countS, "countS"), (countG, "countG")):
        _require_cuda(t, nme)
    if pathS.dtype != torch.float64 or pathG.dtype != torch.float64 or pathS.shape != pathG.shape or pathS.dim() != 3 or pathS.shape[2] != 2:
        raise ValueError("pathS and pathG must be float64 tensors of the same (np, cap, 2) shape")
    pathS, pathG = pathS.contiguous(), pathG.contiguous()
    cS, cG = countS.to(torch.int32).contiguous(), countG.to(torch.int32).contiguous()
    npairs, cap, _ = pathS.shape
    dev = pathS.device
    out = torch.empty((npairs, 2 * cap, 2), dtype=torch.float64, device=dev)
    cnt = torch.empty(npairs, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().fmb_path_stitch2d_f64(pathS.data_ptr(), cS.data_ptr(), pathG.data_ptr(), cG.data_ptr(), cap,
                                                      npairs, float(resolution), out.data_ptr(), cnt.data_ptr(), _stream()))
    return out, cnt


def stitch_rover_path(pathS, pathG, resolution: float) -> np.ndarray:
    """NumPy in / out form of one pair (the planner's lines :1232-1234)."""
    pS = np.ascontiguousarray(np.asarray(pathS, dtype=np.float64).reshape(-1, 2))
    pG = np.ascontiguousarray(np.asarray(pathG, dtype=np.float64).reshape(-1, 2))
    cap = max(len(pS), len(pG), 1)
    dev = torch.device("cuda", torch.cuda.current_device())
    S = torch.zeros((1, cap, 2), dtype=torch.float64, device=dev)
    G = torch.zeros((1, cap, 2), dtype=torch.float64, device=dev)
    S[0, :len(pS)] = torch.from_numpy(pS).to(dev)
    G[0, :len(pG)] = torch.from_numpy(pG).to(dev)
    out, cnt = stitch_rover_paths_device(S, torch.tensor([len(pS)], dtype=torch.int32, device=dev), G,
                                         torch.tensor([len(pG)], dtype=torch.int32, device=dev), resolution)
    return out[0, :int(cnt[0])].cpu().numpy()


def smooth_resample_arm_device(paths: torch.Tensor, count: torch.Tensor, res3: Sequence[float], offset3: Sequence[float],
                               last: Optional[torch.Tensor], m: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """paths (np, cap, 3) float64 tracer slab, count (np,) -> (out (np, m, 3), status (np,) int32; 1 = fewer than 11
    rows, where scipy raises ValueError: those rows are NaN)."""
    _require_cuda(paths, "paths")
    _require_cuda(count, "count")
    if paths.dtype != torch.float64 or paths.dim() != 3 or paths.shape[2] != 3:
        raise ValueError("paths must be a float64 (np, cap, 3) tensor")
    if m < 1:
        raise ValueError("m must be positive")
    paths = paths.contiguous()
    cnt = count.to(torch.int32).contiguous()
    npaths, cap, _ = paths.shape
    dev = paths.device
    if last is not None:
        _require_cuda(last, "last")
        last = last.to(torch.float64).reshape(npaths, 3).contiguous()
    out = torch.empty((npaths, m, 3), dtype=torch.float64, device=dev)
    status = torch.empty(npaths, dtype=torch.int32, device=dev)
    sc = (C.c_double * 3)(*[float(v) for v in res3])
    of = (C.c_double * 3)(*[float(v) for v in offset3])
    ws = _ws(dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().fmb_path_post3d_f64(paths.data_ptr(), cnt.data_ptr(), cap, npaths, sc, of,
                                                    last.data_ptr() if last is not None else None, int(m), out.data_ptr(),
                                                    status.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return out, status


def smooth_resample_arm(path3d, res3, offset3, last3, m: int) -> np.ndarray:
    """NumPy in / out form of one path (gamma3D -> resizedGamma3D, :1641-1671); ValueError like scipy's when the path
    is shorter than the 11-tap window."""
    p = np.ascontiguousarray(np.asarray(path3d, dtype=np.float64).reshape(-1, 3))
    dev = torch.device("cuda", torch.cuda.current_device())
    P = torch.from_numpy(p).to(dev).unsqueeze(0)
    last = None if last3 is None else torch.tensor([float(v) for v in last3], dtype=torch.float64, device=dev).reshape(1, 3)
    out, status = smooth_resample_arm_device(P, torch.tensor([len(p)], dtype=torch.int32, device=dev), res3, offset3, last, m)
    if int(status[0]) != 0:
        raise ValueError("If mode is 'interp', window_length must be less than or equal to the size of x.")
    return out[0].cpu().numpy()
