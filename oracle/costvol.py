"""ctypes front-end of oracle/costvol_oracle.c -- TEST INFRASTRUCTURE ONLY (see its header).

Function names and argument order mirror the reference (src/Coupled_motion_planner.py:319
``GetObstMap``, :505 ``TunnelCost``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as _O

_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)


def _lib():
    L = _O.lib()
    L.cv_obst_map.restype = C.c_int
    L.cv_obst_map.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _dp,
                              C.c_double, C.c_double, _dp]
    L.cv_tunnel_cost.restype = C.c_int
    L.cv_tunnel_cost.argtypes = [C.c_double, C.c_double, C.c_double, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.c_double, C.c_double, _dp, _ip, _ip, _dp]
    return L


def GetObstMap(ZsMap, resX, resY, resZ, sX, sY, sZ, newObstMap, xm, ym):
    """finalMap only (the planner discards the other two return values, :1577)."""
    Z = np.ascontiguousarray(ZsMap, dtype=np.float64)
    ob = np.ascontiguousarray(newObstMap, dtype=np.float64)
    m, n = Z.shape
    out = np.empty((sX, sY, sZ))
    rc = _lib().cv_obst_map(Z.ctypes.data_as(_dp), m, n, resX, resY, resZ, sX, sY, sZ, ob.ctypes.data_as(_dp), xm, ym,
                            out.ctypes.data_as(_dp))
    if rc:
        raise IndexError("index out of bounds")
    return out


def TunnelCost(rlim, rO, rm, gamma2D, sX, sY, sZ, resX, resY, resZ, finalBaseHeading, finalWayPointArm, initialWayPointArm):
    g = np.ascontiguousarray(gamma2D, dtype=np.float64)
    h = np.ascontiguousarray(finalBaseHeading, dtype=np.float64)
    assert g.shape == h.shape and g.shape[1] == 3
    fin = np.ascontiguousarray(finalWayPointArm, dtype=np.int64)
    ini = np.ascontiguousarray(initialWayPointArm, dtype=np.int64)
    out = np.empty((sY, sX, sZ))
    rc = _lib().cv_tunnel_cost(rlim, rO, rm, g.ctypes.data_as(_dp), g.shape[0], sX, sY, sZ, resX, resY, resZ,
                               h.ctypes.data_as(_dp), fin.ctypes.data_as(_ip), ini.ctypes.data_as(_ip), out.ctypes.data_as(_dp))
    assert rc == 0
    return out
