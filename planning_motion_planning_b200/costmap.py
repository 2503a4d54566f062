"""2D cost map of the rover planner, built on the device.

Replaces the inline pipeline of ``main()`` in the reference,
``src/Coupled_motion_planner.py:1144-1216`` (with its helpers ``surface_normal`` :37-80,
``image_filling`` :83-95, ``structural_disk`` :97-109): slope obstacles from the DEM, hole
filling, opening / closing by disks, distance band, 50x50 box blur, +inf map limits.  The
result stays on the GPU in ``[y, x]`` order, i.e. exactly the array the planner hands to
``FM.biComputeTmap`` (``cMap.T``, :1226), so the solve can start without a host round trip.

No CPU fallback: without CUDA or the library every call raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np
import torch

from . import _capi

# constants of the reference's main() (:1129 diagonal, :1154 slope threshold, :1167 opening radius, :1187 band width)
ROVER_DIAGONAL = 0.9
SLOPE_MAX = 0.20
R_OPEN = 10
EXPANSION = 1

_WS = {}


def radii(resolution: float, diagonal: float = ROVER_DIAGONAL) -> Tuple[int, int, int]:
    """Structuring-element radii in cells (:1167, :1172-1173, :1187-1188)."""
    return R_OPEN, int(round(diagonal / 2 / resolution)), int(round(EXPANSION / resolution))


def build_costmap_device(dem: torch.Tensor, resolution: float, size: float, diagonal: float = ROVER_DIAGONAL,
                         stages: bool = False, sync: bool = True):
    """``dem``: (n, n) float64 CUDA tensor, zero-based heights (the planner's ``Zs`` after :1101).
    Returns the cost map as an (n, n) float64 CUDA tensor in ``[y, x]`` order (== ``cMap.T`` of the
    reference); with ``stages`` also a dict of intermediate device tensors (``raw``, ``obst``
    uint8; ``pre_blur`` float64, ``[y, x]`` order)."""
    if not dem.is_cuda:
        raise RuntimeError("dem must be a CUDA tensor (this path has no CPU implementation)")
    if dem.dtype != torch.float64 or dem.dim() != 2 or dem.shape[0] != dem.shape[1]:
        raise TypeError("dem must be a square float64 map")
    dem = dem.contiguous()
    n = dem.shape[0]
    if n != int(round(size / resolution)):
        raise ValueError("dem shape does not match round(size / resolution) (:41-43)")
    dev = dem.device
    L = _capi.lib()
    nbytes = L.fmb_workspace_bytes_costmap2d(n)
    if nbytes == 0:
        raise ValueError("map too small")
    st = torch.cuda.current_stream().cuda_stream
    key = (dev.index, st)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _WS[key] = ws
    grid = torch.from_numpy(np.linspace(0, size, n)).to(dev)
    cost = torch.empty((n, n), dtype=torch.float64, device=dev)
    raw = obst = pre = None
    if stages:
        raw = torch.empty((n, n), dtype=torch.uint8, device=dev)
        obst = torch.empty((n, n), dtype=torch.uint8, device=dev)
        pre = torch.empty((n, n), dtype=torch.float64, device=dev)
    r_open, r_close, r_exp = radii(resolution, diagonal)
    with torch.cuda.device(dev):
        _capi.check(L.fmb_costmap2d_f64(dem.data_ptr(), grid.data_ptr(), n, float(resolution), SLOPE_MAX, r_open, r_close,
                                        r_exp, cost.data_ptr(), raw.data_ptr() if stages else None,
                                        obst.data_ptr() if stages else None, pre.data_ptr() if stages else None,
                                        ws.data_ptr(), ws.numel(), st))
        if sync:
            npos = C.c_int32(0)
            _capi.check(L.fmb_costmap2d_finish(ws.data_ptr(), ws.numel(), st, C.byref(npos)))
            if npos.value == 0:       # np.min of an empty selection at :1198
                raise ValueError("zero-size array to reduction operation minimum which has no identity")
    if stages:
        return cost, {"raw": raw, "obst": obst, "pre_blur": pre}
    return cost


def build_costmap(Zs: np.ndarray, resolution: float, size: float, diagonal: float = ROVER_DIAGONAL) -> np.ndarray:
    """NumPy in / NumPy out: the reference's ``cMap`` (indexed ``[x, y]``; returned as the transposed
    view of the device result, so ``cMap.T`` -- what the planner passes on -- is C-contiguous)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if dev is None:
        raise RuntimeError("no CUDA device (this path has no CPU implementation)")
    dem = torch.from_numpy(np.ascontiguousarray(Zs, dtype=np.float64)).to(dev)
    return build_costmap_device(dem, resolution, size, diagonal).cpu().numpy().T
