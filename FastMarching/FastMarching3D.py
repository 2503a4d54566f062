# -*- coding: utf-8 -*-
"""3D Fast Marching -- drop-in for the reference module ``FastMarching/FastMarching3D.py``.

Same names and signatures as ``/root/reference/src/FastMarching/FastMarching3D.py``.
Volumes are ``[y, x, z]`` (z contiguous), nodes are ``[x, y, z]``.

Exported (reference line in brackets): computeTmap [126]  getPathGDM [198]
interpolatePoint [275].  Not exported: updateNode / getMinNB / sumlist (heap-loop helpers).
"""
from __future__ import annotations

import numpy as np
import torch

from planning_motion_planning_b200 import engine
from . import _compat as _c


def computeTmap(costMap, goal, start):
    """Total-cost volume from ``goal``; the reference stops as soon as ``start`` is accepted
    (FastMarching3D.py:141-142), so cells beyond that front are +inf.  Pass an unreachable
    ``start`` (e.g. ``[-1, -1, -1]``) for the full field."""
    c = np.ascontiguousarray(np.asarray(costMap, dtype=np.float64))
    if c.ndim != 3:
        raise ValueError("costMap must be a 3D array [y, x, z]")
    ny, nx, nz = c.shape
    g = [int(v) for v in goal]
    if not (0 <= g[0] < nx and 0 <= g[1] < ny and 0 <= g[2] < nz):
        raise IndexError(f"goal {g} is outside the {ny}x{nx}x{nz} volume")
    dev = _c.device()
    cd = _c.to_device(c, dev)
    s = [int(np.int64(v)) for v in start]
    s = [v if -2 ** 31 <= v < 2 ** 31 else -1 for v in s]          # (np.uint32(-1) and friends: outside the volume)
    # One library call (fmb_solve3d_until_f64): the solve in the reference's own arithmetic (libm pow for `**2` on
    # scalars, so exactly tied values are tied here too), pop ranks with the reference's LIFO order inside the tie
    # group of `start` (only that group decides which cells are accepted when it pops), replay of the first
    # rank[start] pops.  start == goal / outside / unreached: the full field (the goal is closed before the loop).
    T, info, ws = _c.solve3d_until(cd, g, s)
    out = _c.to_host(T, keep_device=True)               # synchronises
    _c.finish(ws, dev)
    _c.check_info(info.tolist())
    return out


def getPathGDM(totalCostMap, initWaypoint, endWaypoint, tau):
    """Gradient-descent path over a 3D field, FastMarching3D.py:198-271; ``(K, 3)`` float64
    rows ``[x, y, z]``.  Raises what the reference raises when a waypoint degenerates
    (OverflowError / ValueError / IndexError)."""
    Tn = np.ascontiguousarray(np.asarray(totalCostMap, dtype=np.float64))
    init = np.asarray(initWaypoint, dtype=np.float64).reshape(-1)[:3]
    end = np.asarray(endWaypoint, dtype=np.float64).reshape(-1)[:3]
    dev = _c.device()
    out, count, status = _c.trace_field(Tn, dev, lambda Td: engine.trace3d(Td, init[None, :], end[None, :], tau))
    n, st = int(count[0]), int(status[0])
    _c.raise_trace(st)
    return out[0, :n].cpu().numpy()


def interpolatePoint(point, mapI):
    """The reference's tri-linear-like interpolant (FastMarching3D.py:275-314), including its
    non-standard eighth coefficient (:290).  Host-side helper for API parity."""
    px, py, pz = point[0], point[1], point[2]
    i, j, k = np.uint32(np.fix(px)), np.uint32(np.fix(py)), np.uint32(np.fix(pz))
    a, b, c = px - i, py - j, pz - k
    o = mapI[j, i, k]
    ex, ey, ez = mapI[j, i + 1, k], mapI[j + 1, i, k], mapI[j, i, k + 1]
    exy, exz, eyz = mapI[j + 1, i + 1, k], mapI[j, i + 1, k + 1], mapI[j + 1, i, k + 1]
    exyz = mapI[j + 1, i + 1, k + 1]
    coef = (o, ex - o, ey - o, ez - o, exy + o - ex - ey, exz + o - ex - ez, eyz + o - ey - ez,
            exyz + o - ey - ez - ex)
    return (coef[0] + coef[1] * a + coef[2] * b + coef[3] * c + coef[4] * a * b + coef[5] * a * c
            + coef[6] * b * c + coef[7] * a * b * c)
