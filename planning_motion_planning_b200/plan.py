"""Batch entry for the host of the planner: N planning queries in ONE call, owned results.

The reference's native host (src/MotionPlanning.cpp:31-91) calls ``main()`` of the planner once per query through the
embedded interpreter and then reads four module globals (Coupled_motion_planner.py:1402-1406, 1678-1693) through raw
``PyArray_DATA`` pointers whose owners it never releases.  ``plan_batch`` is the entry such a host calls instead
(``PyObject_CallObject`` on this function, or -- without any interpreter -- ``fmb_plan_batch2d_host`` of
include/fm_b200.h, which this function is a thin ctypes front of): host arrays in, a list of freshly allocated NumPy
arrays out, no module-level state, no borrowed buffers.

Per query: full-field solve from the goal (FastMarching.py:92-112), path from the start to the goal
(FastMarching.py:164-236), waypoints in metres, ``resolution * (cell + 1)`` (Coupled_motion_planner.py:1234).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _capi


def plan_batch(cost_map, goals: Sequence, starts: Sequence, tau: float = 0.5, resolution: float = 1.0,
               max_steps: Optional[int] = None, device: int = -1) -> Tuple[List[np.ndarray], np.ndarray, dict]:
    """cost_map (rows, cols) float64 (any strides; +inf = obstacle); goals / starts (N, 2) [x, y] integer nodes.

    Returns (paths, status, info): ``paths[q]`` is an owned (K_q, 2) float64 array (rows from the start to the goal,
    metres), ``status[q]`` the tracer status (0 = the goal was reached and appended), ``info`` the device times.
    Raises ``_capi.FmbError`` on bad arguments or any CUDA / device-side failure (never falls back to the CPU).
    """
    c = np.asarray(cost_map, dtype=np.float64)
    if c.ndim != 2:
        raise ValueError("cost_map must be 2-D")
    if c.strides[1] != c.itemsize or c.strides[0] % c.itemsize or c.strides[0] < c.shape[1] * c.itemsize:
        c = np.ascontiguousarray(c)
    g = np.ascontiguousarray(np.asarray(goals, dtype=np.int32).reshape(-1, 2))
    s = np.ascontiguousarray(np.asarray(starts, dtype=np.int32).reshape(-1, 2))
    if len(g) != len(s) or len(g) < 1:
        raise ValueError("goals and starts must be non-empty and of the same length")
    res = C.POINTER(_capi.FmbPlan2DResult)()
    L = _capi.lib()
    _capi.check(L.fmb_plan_batch2d_host(c.ctypes.data, c.strides[0] // c.itemsize, c.shape[0], c.shape[1], g.ctypes.data,
                                        s.ctypes.data, len(g), float(tau), int(max_steps or 0), float(resolution), int(device),
                                        C.byref(res)))
    try:
        r = res.contents
        nq = int(r.nq)
        off = np.ctypeslib.as_array(r.offsets, shape=(nq + 1,)).copy()
        status = np.ctypeslib.as_array(r.status, shape=(nq,)).copy()
        total = int(off[-1])
        way = np.ctypeslib.as_array(r.waypoints, shape=(total, 2)).copy() if total else np.zeros((0, 2))
        info = {"solve_ms": float(r.solve_ms), "trace_ms": float(r.trace_ms)}
    finally:
        L.fmb_plan2d_free(res)
    return [way[off[q]:off[q + 1]].copy() for q in range(nq)], status, info
