"""Test helper: builds and drives tools/host_emu (CPU execution of the CUDA kernel source
under a SIMT emulator).  Test tooling only -- see tools/host_emu/cuda_emu.h."""
import ctypes as C
import os
import subprocess

import numpy as np

from conftest import ROOT

_DIR = os.path.join(ROOT, "tools", "host_emu")
_SO = os.path.join(_DIR, "libfm_emu.so")
_L = None
dp, fp, ip, up = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)


def lib():
    global _L
    if _L is None:
        srcs = [os.path.join(_DIR, f) for f in ("emu_kernels.cpp", "cuda_emu.h")]
        csrc = os.path.join(ROOT, "planning_motion_planning_b200", "csrc")
        srcs += [os.path.join(csrc, f) for f in os.listdir(csrc)]
        if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
            subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-DFMB_HOST_EMU", "-ffp-contract=off", "-fPIC",
                                   "-shared", "-I", _DIR, "-o", _SO, os.path.join(_DIR, "emu_kernels.cpp")])
        L = C.CDLL(_SO)
        L.emu_solve2d_f64.argtypes = [dp, C.c_longlong, dp, C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_int, up]
        L.emu_solve2d_f32.argtypes = [fp, C.c_longlong, fp, C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_int, up]
        L.emu_solve2d_cta_f64.argtypes = [dp, C.c_longlong, dp, C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, up]
        L.emu_solve2d_cta_f32.argtypes = [fp, C.c_longlong, fp, C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, up]
        L.emu_solve3d_f64.argtypes = [dp, C.c_longlong, dp, C.c_int, C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_int, up]
        L.emu_solve3d_f32.argtypes = [fp, C.c_longlong, fp, C.c_int, C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_int, up]
        L.emu_trace2d_f64.argtypes = [dp, C.c_int, C.c_int, C.c_int, ip, dp, dp, C.c_double, C.c_int, dp, C.c_longlong, ip, ip]
        L.emu_trace3d_f64.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, ip, dp, dp, C.c_double, C.c_int, dp, C.c_longlong, ip, ip]
        _L = L
    return _L


def solve2d(cost, seeds, tw=32, nblocks=2, shared=True):
    dt = cost.dtype
    cost = np.ascontiguousarray(cost)
    seeds = np.ascontiguousarray(seeds, dtype=np.int32).reshape(-1, 2)
    nq = len(seeds)
    rows, cols = cost.shape[-2:]
    T = np.empty((nq, rows, cols), dtype=dt)
    st = np.zeros(8, dtype=np.uint64)
    P = dp if dt == np.float64 else fp
    fn = lib().emu_solve2d_f64 if dt == np.float64 else lib().emu_solve2d_f32
    rc = fn(cost.ctypes.data_as(P), 0 if shared else rows * cols, T.ctypes.data_as(P), rows, cols, nq,
            seeds.ctypes.data_as(ip), tw, nblocks, st.ctypes.data_as(up))
    assert rc == 0, f"emulated solve2d failed rc={rc}"
    return T, dict(visits=int(st[0]), steps=int(st[1]), evals=int(st[2]), pushes=int(st[3]), written=int(st[4]))


def solve2d_cta(cost, seeds, R=2, nblocks=2, shared=True, best_first=0, windowed=0, window=2):
    """CTA-per-tile engine (csrc/eikonal2d_cta.cuh) under the emulator."""
    dt = cost.dtype
    cost = np.ascontiguousarray(cost)
    seeds = np.ascontiguousarray(seeds, dtype=np.int32).reshape(-1, 2)
    nq = len(seeds)
    rows, cols = cost.shape[-2:]
    T = np.empty((nq, rows, cols), dtype=dt)
    st = np.zeros(8, dtype=np.uint64)
    P = dp if dt == np.float64 else fp
    fn = lib().emu_solve2d_cta_f64 if dt == np.float64 else lib().emu_solve2d_cta_f32
    rc = fn(cost.ctypes.data_as(P), 0 if shared else rows * cols, T.ctypes.data_as(P), rows, cols, nq,
            seeds.ctypes.data_as(ip), R, nblocks, best_first, windowed, window, st.ctypes.data_as(up))
    assert rc == 0, f"emulated solve2d_cta failed rc={rc}"
    return T, dict(visits=int(st[0]), steps=int(st[1]), evals=int(st[2]), pushes=int(st[3]), written=int(st[4]))


def solve3d(cost, seeds, tz=32, nblocks=2, shared=True):
    dt = cost.dtype
    cost = np.ascontiguousarray(cost)
    seeds = np.ascontiguousarray(seeds, dtype=np.int32).reshape(-1, 3)
    nq = len(seeds)
    ny, nx, nz = cost.shape[-3:]
    T = np.empty((nq, ny, nx, nz), dtype=dt)
    st = np.zeros(8, dtype=np.uint64)
    P = dp if dt == np.float64 else fp
    fn = lib().emu_solve3d_f64 if dt == np.float64 else lib().emu_solve3d_f32
    rc = fn(cost.ctypes.data_as(P), 0 if shared else ny * nx * nz, T.ctypes.data_as(P), ny, nx, nz, nq,
            seeds.ctypes.data_as(ip), tz, nblocks, st.ctypes.data_as(up))
    assert rc == 0, f"emulated solve3d failed rc={rc}"
    return T, dict(visits=int(st[0]), steps=int(st[1]), evals=int(st[2]))


def _trace(fn, T, dim, init, end, tau, field_of_path=None):
    T = np.ascontiguousarray(T, dtype=np.float64)
    init = np.ascontiguousarray(init, dtype=np.float64).reshape(-1, dim)
    end = np.ascontiguousarray(end, dtype=np.float64).reshape(-1, dim)
    npaths = len(init)
    ms = int(round(15000 / tau))
    cap = ms + 2
    out = np.zeros((npaths, cap, dim))
    cnt = np.zeros(npaths, np.int32)
    st = np.zeros(npaths, np.int32)
    fop = None if field_of_path is None else np.ascontiguousarray(field_of_path, dtype=np.int32)
    shape = T.shape[-dim:]
    fn(T.ctypes.data_as(dp), *shape, npaths, None if fop is None else fop.ctypes.data_as(ip), init.ctypes.data_as(dp),
       end.ctypes.data_as(dp), tau, ms, out.ctypes.data_as(dp), cap, cnt.ctypes.data_as(ip), st.ctypes.data_as(ip))
    return [out[p, :cnt[p]].copy() for p in range(npaths)], [int(s) for s in st]


def trace2d(T, init, end, tau=0.5, field_of_path=None):
    return _trace(lib().emu_trace2d_f64, T, 2, init, end, tau, field_of_path)


def trace3d(T, init, end, tau=0.5, field_of_path=None):
    return _trace(lib().emu_trace3d_f64, T, 3, init, end, tau, field_of_path)


def _ranks(F):
    """Stable ascending sort of the full field -> int32 pop rank per cell (source 0)."""
    flat = np.ascontiguousarray(F).reshape(-1)
    order = np.argsort(flat, kind="stable")
    rank = np.empty(flat.size, dtype=np.int32)
    rank[order] = np.arange(flat.size, dtype=np.int32)
    rank[~np.isfinite(flat)] = np.iinfo(np.int32).max
    return rank


def truncate(F, cost, k, rank=None):
    F = np.ascontiguousarray(F, dtype=np.float64)
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    rank = _ranks(F) if rank is None else np.ascontiguousarray(rank, dtype=np.int32)
    out = np.empty_like(F)
    L = lib()
    if F.ndim == 2:
        L.emu_truncate2d_f64.argtypes = [dp, dp, ip, C.c_int, C.c_int, C.c_int, dp]
        ov = L.emu_truncate2d_f64(F.ctypes.data_as(dp), cost.ctypes.data_as(dp), rank.ctypes.data_as(ip), *F.shape, int(k), out.ctypes.data_as(dp))
    else:
        L.emu_truncate3d_f64.argtypes = [dp, dp, ip, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        ov = L.emu_truncate3d_f64(F.ctypes.data_as(dp), cost.ctypes.data_as(dp), rank.ctypes.data_as(ip), *F.shape, int(k), out.ctypes.data_as(dp))
    return out, ov


def costmap2d(Zs, resolution, size, diagonal=0.9):
    """Emulated csrc/costmap2d.cuh pipeline; returns (cost [y,x], raw, obst, pre_blur [y,x], n_positive)."""
    Zs = np.ascontiguousarray(Zs, dtype=np.float64)
    n = Zs.shape[0]
    grid = np.linspace(0, size, n)
    cost, pre = np.empty((n, n)), np.empty((n, n))
    raw, obst = np.empty((n, n), np.uint8), np.empty((n, n), np.uint8)
    npos = C.c_int(0)
    u8 = C.POINTER(C.c_ubyte)
    fn = lib().emu_costmap2d_f64
    fn.argtypes = [dp, dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, dp, u8, u8, dp, ip]
    rc = fn(Zs.ctypes.data_as(dp), grid.ctypes.data_as(dp), n, resolution, 0.20, 10, int(round(diagonal / 2 / resolution)),
            int(round(1 / resolution)), cost.ctypes.data_as(dp), raw.ctypes.data_as(u8), obst.ctypes.data_as(u8),
            pre.ctypes.data_as(dp), C.byref(npos))
    assert rc == 0
    return cost, raw, obst, pre, npos.value


def costvolume(Zs, resX, resY, resZ, sX, sY, sZ, xm, ym, rlim, rO, rm, gamma2D, heading, fin, ini):
    """Emulated csrc/costvolume.cuh; the small host tables come from the product wrapper
    (planning_motion_planning_b200.costvolume._tables).  Returns (cmap, tunnel, terrain)."""
    from planning_motion_planning_b200 import costvolume as CVP
    tun = CVP._tables(rlim, rO, rm, np.asarray(gamma2D, dtype=np.float64), resX, resZ, np.asarray(heading, dtype=np.float64))
    keep = {k: np.ascontiguousarray(tun[k], dtype=np.float64) for k in ("frames", "li", "lk", "norm", "val", "lr", "hval", "angles")}
    Zs = np.ascontiguousarray(Zs, dtype=np.float64)
    d = CVP._Desc()
    d.d_Zs, d.zs_rows, d.zs_cols = Zs.ctypes.data, Zs.shape[0], Zs.shape[1]
    d.resX, d.resY, d.resZ, d.xm, d.ym = resX, resY, resZ, xm, ym
    d.sX, d.sY, d.sZ = sX, sY, sZ
    for name, key in (("d_frames", "frames"), ("d_li", "li"), ("d_lk", "lk"), ("d_norm", "norm"), ("d_val", "val"),
                      ("d_lr", "lr"), ("d_hval", "hval"), ("d_angles", "angles")):
        setattr(d, name, keep[key].ctypes.data)
    d.npose, d.nX, d.nZ, d.nK = keep["frames"].shape[0] - 1, tun["nX"], tun["nZ"], len(keep["lr"])
    d.rlim, d.shell = rlim, tun["shell"]
    for k in range(3):
        d.fin[k], d.ini[k] = int(fin[k]), int(ini[k])
    cells = sX * sY * sZ
    cmap, tunnel, terrain = np.empty(cells), np.empty(cells), np.empty(cells)
    fn = lib().emu_costvolume_f64
    fn.argtypes = [C.c_void_p, dp, dp, dp]
    assert fn(C.byref(d), cmap.ctypes.data_as(dp), tunnel.ctypes.data_as(dp), terrain.ctypes.data_as(dp)) == 0
    return cmap.reshape(sX, sY, sZ), tunnel.reshape(sY, sX, sZ), terrain.reshape(sX, sY, sZ)


def tie_groups(F, tol=None):
    """(members, gstart, gsize) of the T-sorted tie groups, as FastMarching/_compat.py builds them
    (values closer than `tol`, relative, are tied; default = the 2D / 3D tolerance of _compat)."""
    import ranks_ref as _compat
    if tol is None:
        tol = _compat.TIE_TOL_2D if np.ndim(F) == 2 else _compat.TIE_TOL_3D
    flat = np.ascontiguousarray(F, dtype=np.float64).ravel()
    order = np.argsort(flat, kind="stable")
    ts = flat[order]
    with np.errstate(invalid="ignore"):
        new = np.concatenate([[True], (ts[1:] - ts[:-1]) > tol * ts[1:]]) | ~np.isfinite(ts)
    starts = np.nonzero(new)[0]
    sizes = np.diff(np.concatenate([starts, [flat.size]]))
    grp = np.cumsum(new) - 1
    group = np.empty(flat.size, np.int64)
    group[order] = grp
    gstart = starts[group].astype(np.int32)
    gsize = np.where(np.isfinite(flat), sizes[group], 1).astype(np.int32)
    return order.astype(np.int32), gstart, gsize


def tie_order2d(F, cost, seed, transposed=False):
    """Emulated tie_sweep_kernel<2>: int32 pop ranks incl. the LIFO order among equal values."""
    return _tie_order(F, cost, int(seed[1]) * F.shape[1] + int(seed[0]), transposed)


def tie_order3d(F, cost, seed):
    """3D form; seed = [x, y, z], F indexed [y, x, z]."""
    return _tie_order(F, cost, (int(seed[1]) * F.shape[1] + int(seed[0])) * F.shape[2] + int(seed[2]))


def _tie_order(F, cost, seed_idx, transposed=False):
    F = np.ascontiguousarray(F, dtype=np.float64)
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    members, gstart, gsize = tie_groups(F)
    rank, tau = np.empty(F.size, np.int32), np.empty(F.size, np.int32)
    if F.ndim == 2:
        fn = lib().emu_tie_order2d
        fn.argtypes = [dp, dp, ip, ip, ip, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
        extra = (int(bool(transposed)),)
    else:
        extra = ()
        fn = lib().emu_tie_order3d
        fn.argtypes = [dp, dp, ip, ip, ip, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
    failed = fn(F.ctypes.data_as(dp), cost.ctypes.data_as(dp), members.ctypes.data_as(ip), gstart.ctypes.data_as(ip),
                gsize.ctypes.data_as(ip), *F.shape, seed_idx, *extra, rank.ctypes.data_as(ip), tau.ctypes.data_as(ip))
    assert failed == 0
    return rank.reshape(F.shape)


def bi_join(rankG, rankS):
    """Emulated fmb_bi_join: (k, flat join index) or (None, None)."""
    rG = np.ascontiguousarray(rankG, dtype=np.int32).ravel()
    rS = np.ascontiguousarray(rankS, dtype=np.int32).ravel()
    out = np.zeros(2, np.int32)
    fn = lib().emu_bi_join
    fn.argtypes = [ip, ip, C.c_longlong, ip]
    fn(rG.ctypes.data_as(ip), rS.ctypes.data_as(ip), rG.size, out.ctypes.data_as(ip))
    return (None, None) if out[0] == 2**31 - 1 else (int(out[0]), int(out[1]))
