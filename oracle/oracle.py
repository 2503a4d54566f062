"""ctypes front-end of the CPU oracle (oracle/fmm_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs.  The product path
(``FastMarching/``, ``planning_motion_planning_b200/``) never imports this module.

Function names mirror the reference modules so parity tests read like calls to
``/root/reference/src/FastMarching/FastMarching.py`` / ``FastMarching3D.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TRACE_OK, TRACE_EARLY, TRACE_VALUEERROR, TRACE_INDEXERROR, TRACE_OVERFLOW = 0, 1, 2, 3, 4


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("fmm_oracle.c", "costvol_oracle.c", "Makefile")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        ip64 = C.POINTER(C.c_int64)
        ip32 = C.POINTER(C.c_int32)
        L.orc_eikonal2d.restype = C.c_double
        L.orc_eikonal2d.argtypes = [C.c_double] * 3
        L.orc_solve3d.restype = C.c_double
        L.orc_solve3d.argtypes = [C.c_double] * 4
        L.orc_fmm2d.restype = C.c_int
        L.orc_fmm2d.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, ip64, ip64, ip64]
        L.orc_bifmm2d.restype = C.c_int
        L.orc_bifmm2d.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, ip32, ip64]
        L.orc_gradient2d.restype = None
        L.orc_gradient2d.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, dp, dp]
        L.orc_interp2d.restype = C.c_double
        L.orc_interp2d.argtypes = [dp, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int)]
        L.orc_trace2d.restype = C.c_int64
        L.orc_trace2d.argtypes = [dp, C.c_int, C.c_int, dp, dp, C.c_double, dp, C.c_int64, C.POINTER(C.c_int)]
        L.orc_fmm3d.restype = C.c_int
        L.orc_pow2_array.argtypes = [dp, dp, C.c_int64]
        L.orc_pow2_array.restype = None
        L.orc_fmm3d.argtypes = [dp, C.c_int, C.c_int, C.c_int, ip32, ip32, dp, ip64, ip64, ip64]
        L.orc_interp3d.restype = C.c_double
        L.orc_interp3d.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int)]
        L.orc_trace3d.restype = C.c_int64
        L.orc_trace3d.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, dp, C.c_double, dp, C.c_int64, C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_EXC = {TRACE_VALUEERROR: ValueError, TRACE_INDEXERROR: IndexError, TRACE_OVERFLOW: OverflowError}


# ---------------------------------------------------------------- 2D -------
def getEikonal(Thor, Tver, cost):
    return lib().orc_eikonal2d(float(Thor), float(Tver), float(cost))


def computeTmap(costMap, goal, start=None, return_stats=False):
    """Single-front 2D FMM from ``goal`` (intended semantics of FastMarching.py:92-112).
    ``start=None`` -> full field."""
    c = _c64(costMap)
    rows, cols = c.shape
    T = np.empty_like(c)
    sx, sy = (-1, -1) if start is None else (int(start[0]), int(start[1]))
    npop = C.c_int64(0)
    evals = C.c_int64(0)
    order = np.empty(c.size, dtype=np.int64) if return_stats else None
    rc = lib().orc_fmm2d(_dp(c), rows, cols, int(goal[0]), int(goal[1]), sx, sy, _dp(T),
                         order.ctypes.data_as(C.POINTER(C.c_int64)) if order is not None else None,
                         C.byref(npop), C.byref(evals))
    if rc:
        raise MemoryError("oracle allocation failed")
    if return_stats:
        return T, order[: npop.value], evals.value
    return T


def biComputeTmap(costMap, goal, start):
    """FastMarching.py:114-162."""
    c = _c64(costMap)
    rows, cols = c.shape
    TG = np.empty_like(c)
    TS = np.empty_like(c)
    join = np.zeros(2, dtype=np.int32)
    npop = C.c_int64(0)
    rc = lib().orc_bifmm2d(_dp(c), rows, cols, int(goal[0]), int(goal[1]), int(start[0]), int(start[1]),
                           _dp(TG), _dp(TS), join.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(npop))
    if rc == 1:
        raise NameError("name 'nodeJoin' is not defined")
    if rc:
        raise MemoryError("oracle allocation failed")
    return TG, TS, np.uint32(join)


def computeGradient(cost, point=()):
    c = _c64(cost)
    m, n = c.shape
    Gnx = np.empty_like(c)
    Gny = np.empty_like(c)
    if len(point) == 0:
        lib().orc_gradient2d(_dp(c), m, n, 0, 0.0, 0.0, _dp(Gnx), _dp(Gny))
    else:
        lib().orc_gradient2d(_dp(c), m, n, 1, float(point[0]), float(point[1]), _dp(Gnx), _dp(Gny))
    return Gnx, Gny


def interpolatePoint(point, mapI):
    m = _c64(mapI)
    oob = C.c_int(0)
    v = lib().orc_interp2d(_dp(m), m.shape[0], m.shape[1], float(point[0]), float(point[1]), C.byref(oob))
    if oob.value:
        raise IndexError("interpolation stencil outside the map")
    return v


def getPathGDM(totalCostMap, initWaypoint, endWaypoint, tau, return_status=False):
    """FastMarching.py:164-236."""
    T = _c64(totalCostMap)
    m, n = T.shape
    cap = int(round(15000 / tau)) + 2
    out = np.empty((cap, 2), dtype=np.float64)
    init = _c64(np.asarray(initWaypoint, dtype=np.float64).ravel())
    end = _c64(np.asarray(endWaypoint, dtype=np.float64).ravel())
    st = C.c_int(0)
    k = lib().orc_trace2d(_dp(T), m, n, _dp(init), _dp(end), float(tau), _dp(out), cap, C.byref(st))
    if return_status:
        return out[:k].copy(), st.value
    if st.value in _EXC:
        raise _EXC[st.value]("reference getPathGDM raises here")
    return out[:k].copy()


# ---------------------------------------------------------------- 3D -------
def solve3d(Tx, Ty, Tz, cost):
    return lib().orc_solve3d(float(Tx), float(Ty), float(Tz), float(cost))


def computeTmap3D(costMap, goal, start=None, return_stats=False):
    """FastMarching3D.py:126-145; ``start=None`` -> full field."""
    c = _c64(costMap)
    ny, nx, nz = c.shape
    T = np.empty_like(c)
    g = np.asarray([int(v) for v in goal], dtype=np.int32)
    s = np.asarray([-1, -1, -1] if start is None else [int(np.int64(v)) for v in start], dtype=np.int32)
    npop = C.c_int64(0)
    evals = C.c_int64(0)
    order = np.empty(c.size, dtype=np.int64) if return_stats else None
    ip32 = C.POINTER(C.c_int32)
    rc = lib().orc_fmm3d(_dp(c), ny, nx, nz, g.ctypes.data_as(ip32), s.ctypes.data_as(ip32), _dp(T),
                         order.ctypes.data_as(C.POINTER(C.c_int64)) if order is not None else None,
                         C.byref(npop), C.byref(evals))
    if rc:
        raise MemoryError("oracle allocation failed")
    if return_stats:
        return T, order[: npop.value], evals.value
    return T


def interpolatePoint3D(point, mapI):
    m = _c64(mapI)
    oob = C.c_int(0)
    v = lib().orc_interp3d(_dp(m), m.shape[0], m.shape[1], m.shape[2],
                           float(point[0]), float(point[1]), float(point[2]), C.byref(oob))
    if oob.value:
        raise IndexError("interpolation stencil outside the volume")
    return v


def getPathGDM3D(totalCostMap, initWaypoint, endWaypoint, tau, return_status=False):
    """FastMarching3D.py:198-271."""
    T = _c64(totalCostMap)
    ny, nx, nz = T.shape
    cap = int(round(15000 / tau)) + 2
    out = np.empty((cap, 3), dtype=np.float64)
    init = _c64(np.asarray(initWaypoint, dtype=np.float64).ravel())
    end = _c64(np.asarray(endWaypoint, dtype=np.float64).ravel())
    st = C.c_int(0)
    k = lib().orc_trace3d(_dp(T), ny, nx, nz, _dp(init), _dp(end), float(tau), _dp(out), cap, C.byref(st))
    if return_status:
        return out[:k].copy(), st.value
    if st.value in _EXC:
        raise _EXC[st.value]("reference getPathGDM (3D) raises here")
    return out[:k].copy()


def pow2(x):
    """libm ``pow(x, 2.0)`` element-wise: the value of ``x**2`` for a NumPy *scalar* x (FastMarching3D.py:68-71)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib().orc_pow2_array(_dp(x), _dp(out), x.size)
    return out
