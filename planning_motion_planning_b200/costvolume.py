"""3D arm-workspace cost volume of the planner, built on the device.

Mirrors the reference's ``GetObstMap`` (``src/Coupled_motion_planner.py:319-358``) and
``TunnelCost`` (:505-725) -- same names, argument order and return arrays -- and adds
``build_cost_volume`` for the product ``Cmap1*Cmap2`` of :1627, which is what
``FM3D.computeTmap`` consumes.  The host side prepares a few thousand table entries with the
reference's own scalar expressions (frames, linspace axes, per-(i, k) costs); the voxel work
-- an order-dependent scatter in the reference -- runs in ``csrc/costvolume.cuh``.

No CPU fallback: without CUDA or the library every call raises.
"""
from __future__ import annotations

import ctypes as C
import functools
import math

import numpy as np
import torch

from . import _capi

GRADIENT = 15          # :513


class _Desc(C.Structure):
    _fields_ = [("d_Zs", C.c_void_p), ("zs_rows", C.c_int32), ("zs_cols", C.c_int32),
                ("resX", C.c_double), ("resY", C.c_double), ("resZ", C.c_double), ("xm", C.c_double), ("ym", C.c_double),
                ("sX", C.c_int32), ("sY", C.c_int32), ("sZ", C.c_int32),
                ("d_frames", C.c_void_p), ("npose", C.c_int32),
                ("d_li", C.c_void_p), ("d_lk", C.c_void_p), ("nX", C.c_int32), ("nZ", C.c_int32),
                ("d_norm", C.c_void_p), ("d_val", C.c_void_p), ("rlim", C.c_double),
                ("d_lr", C.c_void_p), ("d_hval", C.c_void_p), ("nK", C.c_int32),
                ("d_angles", C.c_void_p), ("shell", C.c_double),
                ("fin", C.c_int64 * 3), ("ini", C.c_int64 * 3)]


def _frame(alpha, beta, gamma, p):
    """Rows 0..2 of the homogeneous base frame (:530-533)."""
    ca, cb, cg = math.cos(alpha), math.cos(beta), math.cos(gamma)
    sa, sb, sg = math.sin(alpha), math.sin(beta), math.sin(gamma)
    return [ca * cb, ca * sb * sg - sa * cg, ca * sb * cg + sa * sg, p[0],
            sa * cb, sa * sb * sg + ca * cg, sa * sb * cg - ca * sg, p[1],
            -sb, cb * sg, cb * cg, p[2]]


@functools.lru_cache(maxsize=8)
def _static_tables(rlim, rO, rm, resX, resZ):
    """The part of TunnelCost's tables that depends on the arm and the grid resolution only -- not on the base path --
    as the reference computes it (numpy scalars from np.linspace, ``**2`` on scalars = libm pow, math.sqrt / cos / sin).
    A pure function of five floats: cached, a planner that re-plans with the same arm pays the Python loops once
    (2.7 ms of a 3.2 ms call at 256^3)."""
    tunnelRad = rlim + 2 * resX
    nX = int(round(2 * tunnelRad / resX) + 1)
    nZ = int(round(2 * tunnelRad / resZ) + 1)
    li = np.linspace(-tunnelRad, tunnelRad, nX, endpoint=True)
    lk = np.linspace(-tunnelRad, tunnelRad, nZ, endpoint=True)
    norm = np.empty((nX, nZ))
    val = np.empty((nX, nZ))
    k2 = [k ** 2 for k in lk]
    for a, i in enumerate(li):
        i2 = i ** 2
        ramp = 4 * (i + rlim + 2 * resZ)
        for b in range(nZ):
            nr = math.sqrt(i2 + k2[b])
            norm[a, b] = nr
            val[a, b] = GRADIENT * (nr - (rO + rm) / 2) ** 2 + 2 + ramp
    lr = np.linspace(0, tunnelRad, round(nZ / 2) + 1, endpoint=True)
    hval = np.array([GRADIENT * (k - (rO + rm) / 2) ** 2 + 2 for k in lr])
    th = [math.pi * i / 180 for i in range(-100, 100, 2)]
    sg = [math.pi * j / 180 for j in range(-90, 90, 2)]
    angles = np.array([math.cos(t) for t in th] + [math.sin(t) for t in th] + [math.cos(s) for s in sg] + [math.sin(s) for s in sg])
    return dict(li=li, lk=lk, norm=norm, val=val, lr=lr, hval=hval, angles=angles, shell=rlim + 2 * resZ, nX=nX, nZ=nZ)


def _tables(rlim, rO, rm, gamma2D, resX, resZ, heading):
    """Everything of TunnelCost that does not depend on the voxel grid: the cached static tables plus the base frames of
    this path (math.cos / math.sin per pose, as the reference)."""
    t = dict(_static_tables(float(rlim), float(rO), float(rm), float(resX), float(resZ)))
    m = gamma2D.shape[0]
    frames = [_frame(heading[j, 2] - math.pi / 2, heading[j, 1], heading[j, 0], gamma2D[j]) for j in range(m)]
    frames.append(_frame(heading[m - 1, 2], heading[m - 1, 1], heading[m - 1, 0], gamma2D[m - 1]))      # :649 (no -pi/2)
    t["frames"] = np.array(frames, dtype=np.float64)
    return t


_WS = {}


def _run(Zs, resX, resY, resZ, sX, sY, sZ, xm, ym, tun, want, device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device (this path has no CPU implementation)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    L = _capi.lib()
    f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)     # noqa: E731
    Zd = f64(Zs)
    keep = [Zd]
    d = _Desc()
    d.d_Zs, d.zs_rows, d.zs_cols = Zd.data_ptr(), Zs.shape[0], Zs.shape[1]
    d.resX, d.resY, d.resZ, d.xm, d.ym = float(resX), float(resY), float(resZ), float(xm), float(ym)
    d.sX, d.sY, d.sZ = int(sX), int(sY), int(sZ)
    for name, key in (("d_frames", "frames"), ("d_li", "li"), ("d_lk", "lk"), ("d_norm", "norm"), ("d_val", "val"),
                      ("d_lr", "lr"), ("d_hval", "hval"), ("d_angles", "angles")):
        t = f64(tun[key])
        keep.append(t)
        setattr(d, name, t.data_ptr())
    d.npose, d.nX, d.nZ, d.nK = tun["frames"].shape[0] - 1, tun["nX"], tun["nZ"], len(tun["lr"])
    d.rlim, d.shell = float(tun["rlim"]), float(tun["shell"])
    for k in range(3):
        d.fin[k], d.ini[k] = int(tun["fin"][k]), int(tun["ini"][k])
    cells = sX * sY * sZ
    outs = {k: torch.empty(cells, dtype=torch.float64, device=dev) for k in want}
    nbytes = L.fmb_workspace_bytes_costvolume(sX, sY, sZ)
    st = torch.cuda.current_stream().cuda_stream
    ws = _WS.get((dev.index, st))
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _WS[(dev.index, st)] = ws
    ptr = lambda k: outs[k].data_ptr() if k in outs else None      # noqa: E731
    with torch.cuda.device(dev):
        _capi.check(L.fmb_costvolume_f64(C.byref(d), ptr("cmap"), ptr("tunnel"), ptr("terrain"), ws.data_ptr(), ws.numel(), st))
    return outs


def _neutral_tunnel(resX, resZ):
    """A tunnel description that writes nothing (for GetObstMap alone)."""
    t = _tables(0.0, 0.0, 0.0, np.zeros((1, 3)), resX, resZ, np.zeros((1, 3)))
    t.update(rlim=0.0, fin=[-1, -1, -1], ini=[-1, -1, -1])
    return t


def GetObstMap(ZsMap, resX, resY, resZ, sX, sY, sZ, newObstMap, xm, ym):
    """Reference signature (:319).  Returns ``(finalMap, obstMap, groundMap)``; the two partial maps
    are derived on the host from finalMap and ``newObstMap`` (the planner ignores them, :1577)."""
    Zs = np.asarray(ZsMap, dtype=np.float64)
    if Zs.shape[0] > sX or Zs.shape[1] > sY:
        raise IndexError("index out of bounds for the (sX, sY, sZ) volume")
    out = _run(Zs, resX, resY, resZ, sX, sY, sZ, xm, ym, _neutral_tunnel(resX, resZ), ("terrain",))
    final = out["terrain"].cpu().numpy().reshape(sX, sY, sZ)
    ob = np.zeros((sX, sY), dtype=bool)
    m, n = Zs.shape
    ob[:m, :n] = np.asarray(newObstMap)[:m, :n] == 1
    inner = np.zeros_like(final, dtype=bool)
    inner[1:-1, 1:-1, 1:-1] = np.isinf(final[1:-1, 1:-1, 1:-1])
    # border voxels of finalMap are +inf by construction; the partial maps keep the terrain voxel there too
    col = np.zeros_like(final, dtype=bool)
    with np.errstate(invalid="ignore"):
        iz = np.rint(Zs / resZ)
    jj, ii = np.nonzero((resX * np.arange(n)[None, :] != xm) & (resY * np.arange(m)[:, None] != ym) & (iz < sZ)
                        & (np.arange(n)[None, :] < sX) & (np.arange(m)[:, None] < sY))
    col[jj, ii, iz[jj, ii].astype(np.int64)] = True
    obst = np.where(col & ob[:, :, None], np.inf, 1.0)
    ground = np.where(col & ~ob[:, :, None], np.inf, 1.0)
    return final, obst, ground


def TunnelCost(rlim, rO, rm, gamma2D, sX, sY, sZ, resX, resY, resZ, finalBaseHeading, finalWayPointArm, initialWayPointArm):
    """Reference signature (:505).  Returns Cmap of shape (sY, sX, sZ)."""
    g = np.asarray(gamma2D, dtype=np.float64)
    h = np.asarray(finalBaseHeading, dtype=np.float64)
    if g.shape != h.shape or g.ndim != 2 or g.shape[1] < 3:
        raise ValueError("gamma2D and finalBaseHeading must both be (m, 3)")
    tun = _tables(rlim, rO, rm, g, resX, resZ, h)
    tun.update(rlim=rlim, fin=[int(v) for v in finalWayPointArm], ini=[int(v) for v in initialWayPointArm])
    # the terrain part needs a DEM: a 1 x 1 dummy that marks nothing inside the tunnel output
    out = _run(np.zeros((1, 1)), resX, resY, resZ, sX, sY, sZ, math.nan, math.nan, tun, ("tunnel",))
    return out["tunnel"].cpu().numpy().reshape(sY, sX, sZ)


def build_cost_volume_device(ZsMap, resX, resY, resZ, sX, sY, sZ, xm, ym, rlim, rO, rm, gamma2D, finalBaseHeading,
                             finalWayPointArm, initialWayPointArm) -> torch.Tensor:
    """``GetObstMap(...)[0] * TunnelCost(...)`` (:1577, :1623, :1627) as one device tensor of shape
    (sX, sY, sZ), ready for ``engine.solve3d`` / ``FM3D.computeTmap``."""
    if sX != sY:
        raise ValueError(f"operands could not be broadcast together with shapes ({sX},{sY},{sZ}) ({sY},{sX},{sZ})")
    g = np.asarray(gamma2D, dtype=np.float64)
    h = np.asarray(finalBaseHeading, dtype=np.float64)
    tun = _tables(rlim, rO, rm, g, resX, resZ, h)
    tun.update(rlim=rlim, fin=[int(v) for v in finalWayPointArm], ini=[int(v) for v in initialWayPointArm])
    out = _run(np.asarray(ZsMap, dtype=np.float64), resX, resY, resZ, sX, sY, sZ, xm, ym, tun, ("cmap",))
    return out["cmap"].reshape(sX, sY, sZ)


def build_cost_volume(*args) -> np.ndarray:
    return build_cost_volume_device(*args).cpu().numpy()
