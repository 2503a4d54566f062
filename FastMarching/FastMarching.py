# -*- coding: utf-8 -*-
"""2D Fast Marching -- drop-in for the reference module ``FastMarching/FastMarching.py``.

Same function names, argument order and return types as the reference
(``/root/reference/src/FastMarching/FastMarching.py``); the sorted-list FMM loop is
replaced by the B200 tile solver and the per-step full-map gradient by the warp tracer
(``planning_motion_planning_b200``).  Arrays are ``[y, x]``, nodes are ``[x, y]``.

Exported (reference line in brackets):
  biComputeTmap [114]  computeTmap [92]  getPathGDM [164]  computeGradient [242]
  interpolatePoint [305]  getEikonal [17]
Not exported: updateNode / getMinNB / getNeighbours -- helpers of the heap loop, which
has no counterpart in a tile solver.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from planning_motion_planning_b200 import engine
from . import _compat as _c


def getEikonal(Thor, Tver, cost):
    """Scalar first-order upwind update of one node from its best horizontal and vertical
    neighbour values (reference: FastMarching.py:17-29).  Kept for API parity; the solver
    evaluates the same expression, in the same rounding order, on the device."""
    gap = np.abs(Thor - Tver)
    if not (gap <= cost):                 # one neighbour unusable (+inf) or too far: one-sided
        lowest = np.minimum(Thor, Tver)
        return lowest + cost if np.isfinite(lowest) else np.inf
    return .5 * (Thor + Tver + math.sqrt(2 * np.square(cost) - np.square(Thor - Tver)))


def _to_numpy_field(Tt: torch.Tensor, swap: bool) -> np.ndarray:
    a = _c.to_host(Tt, keep_device=True)
    return a.T if swap else a          # .T of a C array is F-ordered, like zeros_like() of the planner's view


def computeTmap(costMap, goal, start):
    """Single-front total-cost map from ``goal``, stopping when ``start`` is accepted.

    The shipped reference function raises ValueError on its first iteration
    (FastMarching.py:107 unpacks three values into two); this implements the evident
    intent, i.e. the semantics of the working 3D driver (FastMarching3D.py:126-145).
    One library call (fmb_solve2d_until_f64): full solve, pop ranks in the reference's order, replay of the first
    rank[start] pops.  start == goal, or an unreached / outside start: the FULL field (the goal is closed before the
    loop and never popped, FastMarching3D.py:127-142).
    """
    c, swap = _c.as_c_field(costMap)
    c = np.ascontiguousarray(c)
    rows, cols = c.shape
    g, s = _c.node2(goal, swap), _c.node2(start, swap)
    _c.check_node2(g, rows, cols)
    dev = _c.device()
    pinned = _c.is_page_locked(c)
    full = g == s or not (0 <= s[0] < cols and 0 <= s[1] < rows)          # known on the host: the loop never stops early
    if full and pinned and _c.H2D_OVERLAP:
        # the map's upload runs in bands behind the solve that already consumes it (fmb_solve2d_h2d_f64)
        T, ws = _c.solve2d_h2d(c, g, dev)
        out = _to_numpy_field(T, swap)                  # synchronises
        _c.finish(ws, dev)
        return out
    cd = _c.to_device(c, dev, pinned=pinned)
    T, info, ws = _c.solve2d_until(cd, g, s, swap)
    out = _to_numpy_field(T, swap)                      # synchronises
    _c.finish(ws, dev)
    _c.check_info(info.tolist())
    return out


def biComputeTmap(costMap, goal, start):
    """Two fronts (G from ``goal``, S from ``start``) advanced alternately until they meet,
    FastMarching.py:114-162.  Returns ``(TmapG, TmapS, nodeJoin)`` with ``nodeJoin`` a
    ``np.uint32[2]`` ``[x, y]``.  Raises ``NameError`` when the fronts never meet, like the
    reference (:161).  One library call (fmb_bisolve2d_f64); the S front's ranks and replay run on a second stream."""
    c, swap = _c.as_c_field(costMap)
    c = np.ascontiguousarray(c)
    g, s = _c.node2(goal, swap), _c.node2(start, swap)
    _c.check_node2(g, *c.shape)
    _c.check_node2(s, *c.shape)
    dev = _c.device()
    pinned = _c.is_page_locked(c)
    if pinned and _c.H2D_OVERLAP:
        TG, TS, info, ws = _c.bisolve2d(c, g, s, swap, dev)          # the map's upload runs in bands behind the two solves
    else:
        TG, TS, info, ws = _c.bisolve2d(_c.to_device(c, dev, pinned=pinned), g, s, swap)
    _c.finish(ws, dev)                                  # synchronises; device-side failures of the solve surface here
    inf = info.tolist()
    k, j = inf[0], inf[1]
    if k == 0x7fffffff:
        raise NameError("name 'nodeJoin' is not defined")
    _c.check_info(inf, fronts=2)
    jy, jx = divmod(j, c.shape[1])
    node = (jy, jx) if swap else (jx, jy)
    return _to_numpy_field(TG, swap), _to_numpy_field(TS, swap), np.uint32(node)


def getPathGDM(totalCostMap, initWaypoint, endWaypoint, tau):
    """Gradient-descent path from ``initWaypoint`` to ``endWaypoint`` over ``totalCostMap``,
    FastMarching.py:164-236; returns an ``(K, 2)`` float64 array of ``[x, y]`` rows."""
    c, swap = _c.as_c_field(totalCostMap)
    init = np.asarray(initWaypoint, dtype=np.float64).reshape(-1)[:2]
    end = np.asarray(endWaypoint, dtype=np.float64).reshape(-1)[:2]
    dev = _c.device()

    out, count, status = _c.trace_field2d(np.ascontiguousarray(c), swap, dev, init, end, tau)
    n, st = int(count[0]), int(status[0])
    _c.raise_trace(st)
    return out[0, :n].cpu().numpy()


def computeGradient(cost, point=[]):
    """Normalised inf-aware gradient, FastMarching.py:242-300: whole map when ``point`` is
    empty, else the 6x6 window around ``point`` (zeros elsewhere).  Not used by the planner;
    provided for API parity (vectorised NumPy on the host, not part of the hot path)."""
    cost = np.asarray(cost, dtype=np.float64)
    m, n = cost.shape
    if len(point) == 0:
        jmin, imin, jmax, imax = 0, 0, m, n
    else:
        jmax, imax = min(m, int(point[1]) + 3), min(n, int(point[0]) + 3)
        jmin, imin = max(0, int(point[1] - 3)), max(0, int(point[0] - 3))
    Gnx = np.zeros_like(cost)
    Gny = np.zeros_like(cost)
    if jmax <= jmin or imax <= imin:
        return Gnx, Gny
    with np.errstate(all="ignore"):
        def axis_grad(a):            # along axis 0 of `a`
            L = a.shape[0]
            g = np.zeros_like(a)
            up, dn, mid = a[:-2], a[2:], a[1:-1]
            iu, idn = np.isinf(up), np.isinf(dn)
            inner = np.where(idn, np.where(iu, 0.0, mid - up), np.where(iu, dn - mid, (dn - up) / 2))
            g[1:-1] = inner
            g[0] = a[1] - a[0]
            g[L - 1] = a[L - 1] - a[L - 2]
            return g
        Gy = axis_grad(cost)
        Gx = axis_grad(cost.T).T
        nrm = np.sqrt(Gx ** 2 + Gy ** 2)
        w = (slice(jmin, jmax), slice(imin, imax))
        Gnx[w] = (Gx / nrm)[w]
        Gny[w] = (Gy / nrm)[w]
    return Gnx, Gny


def interpolatePoint(point, mapI):
    """Bilinear interpolation of ``mapI`` at ``point`` = [x, y] (reference:
    FastMarching.py:305-338).  A zero fractional part short-circuits the corresponding
    corners so that NaN/inf nodes the point does not actually touch cannot leak in."""
    px, py = point[0], point[1]
    col, row = np.uint32(np.fix(px)), np.uint32(np.fix(py))
    fx, fy = px - col, py - row
    nrows, ncols = mapI.shape
    if col == ncols or row == nrows:      # degenerate edge forms of the reference (they index past the array)
        if col == ncols and row == nrows:
            return mapI[row, col]
        if col == ncols:
            return fy * mapI[row + 1, col] + (1 - fy) * mapI[row, col]
        return fx * mapI[row, col + 1] + (1 - fx) * mapI[row, col]
    base = mapI[row, col]
    value = base
    if fx != 0:
        value = value + (mapI[row, col + 1] - base) * fx
    if fy != 0:
        value = value + (mapI[row + 1, col] - base) * fy
    if fx != 0 and fy != 0:
        value = value + (mapI[row + 1, col + 1] + base - mapI[row, col + 1] - mapI[row + 1, col]) * fx * fy
    return value
