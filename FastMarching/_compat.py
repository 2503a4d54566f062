"""Helpers shared by the two drop-in modules: argument normalisation and the
emulation of the reference's early-exit ("partial field") semantics."""
from __future__ import annotations

import os
import weakref

import numpy as np
import torch

from planning_motion_planning_b200 import _capi, engine

_EXC = {
    _capi.TRACE_VALUEERROR: (ValueError, "cannot convert float NaN to integer"),
    _capi.TRACE_INDEXERROR: (IndexError, "index out of bounds for the field"),
    _capi.TRACE_OVERFLOW: (OverflowError, "cannot convert float infinity to integer"),
}


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("FastMarching (B200 build) needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# ---- host <-> device plumbing of the NumPy boundary ------------------------------------------------
# The planner hands pageable NumPy arrays over and gets NumPy arrays back.  A pageable copy of a 128 MiB field moves
# at 3-4 GB/s; through page-locked memory it moves at PCIe speed.  So: uploads are staged chunk-wise through four small
# pinned buffers (the memcpy of chunk k+1 overlaps the DMA of chunk k), and results are downloaded straight into a
# pinned tensor whose memory backs the returned NumPy array (owned by the caller through the array's base).  An array
# that came from here is recognised on its way back in (getPathGDM on a field biComputeTmap returned): the tracer runs
# on the device copy that was kept of it and only the windows the path read are checked against the array, in place
# (trace_field2d below); where it has to be uploaded after all, one DMA does it.
# fmb_solve2d_h2d_f64 (upload overlapped with the solve) can be switched off, e.g. under a profiler that replays kernels
# and restores device memory in between (the solve waits on flags the copy stream sets): FMB_H2D_OVERLAP=0
H2D_OVERLAP = os.environ.get("FMB_H2D_OVERLAP", "1") != "0"
_STAGE = {}
_STAGE_ELEMS = 2 << 20            # 16 MiB of float64 per staging buffer
_STAGE_BUFS = 4


def is_page_locked(a: np.ndarray) -> bool:
    """True for an array this module returned and for a caller's array it has page-locked (a second sighting does that)."""
    return a.size >= (1 << 18) and (_is_ours(a) or _page_locked_on_reuse(a))


def to_device(a: np.ndarray, dev: torch.device, pinned=None) -> torch.Tensor:
    """C-contiguous float64 NumPy array -> device tensor of the same shape (pinned: is_page_locked(a), when the caller
    already asked)."""
    t = torch.from_numpy(a)
    if a.size < (1 << 18):
        return t.to(dev)
    if is_page_locked(a) if pinned is None else pinned:         # one DMA
        return t.to(dev, non_blocking=True)
    out = torch.empty(a.shape, dtype=torch.float64, device=dev)
    flat_src, flat_dst = t.reshape(-1), out.reshape(-1)
    key = dev.index
    if key not in _STAGE:
        _STAGE[key] = ([torch.empty(_STAGE_ELEMS, dtype=torch.float64).pin_memory() for _ in range(_STAGE_BUFS)],
                       [torch.cuda.Event() for _ in range(_STAGE_BUFS)])
    bufs, evs = _STAGE[key]
    n = flat_src.numel()
    stream = torch.cuda.current_stream(dev)
    for k, lo in enumerate(range(0, n, _STAGE_ELEMS)):
        hi = min(n, lo + _STAGE_ELEMS)
        b = k % len(bufs)
        if k >= len(bufs):
            evs[b].synchronize()                       # the DMA that last read this staging buffer is done
        bufs[b][:hi - lo].copy_(flat_src[lo:hi])
        flat_dst[lo:hi].copy_(bufs[b][:hi - lo], non_blocking=True)
        evs[b].record(stream)
    return out


# A caller's array that comes in a SECOND time (goal sweeps over one cost map, a planner that re-plans on the same map)
# is page-locked in place with cudaHostRegister (one-off cost: 10-16 ms for 128 MiB, tools/gpu_hostregister_cost.py -- more
# than the 5.8 ms of one staged upload, which is why a first sighting is not registered) and from then on
# uploaded with one DMA at PCIe speed instead of through the staging copy (4096^2: 5.8 -> 2.5 ms per call).  The
# registration ends when the array is garbage-collected; at most _REG_MAX arrays / _REG_BYTES_MAX bytes are held.
_SEEN = {}                         # id(root array) -> weakref to it
_REGISTERED = {}                   # id(root array) -> (address, bytes)
_REG_MAX, _REG_BYTES_MAX = 4, 2 << 30


def _unregister(key: int, ptr: int):
    _REGISTERED.pop(key, None)
    _SEEN.pop(key, None)
    try:
        torch.cuda.cudart().cudaHostUnregister(ptr)
    except Exception:
        pass


def _page_locked_on_reuse(a: np.ndarray) -> bool:
    root = a
    while isinstance(root.base, np.ndarray):
        root = root.base
    key = id(root)
    if key in _REGISTERED:
        return True
    r = _SEEN.get(key)
    if r is None or r() is not root:
        try:
            _SEEN[key] = weakref.ref(root, lambda _r, k=key: _SEEN.pop(k, None))
        except TypeError:
            pass
        return False
    if not (root.flags.c_contiguous or root.flags.f_contiguous) or len(_REGISTERED) >= _REG_MAX \
            or sum(b for _, b in _REGISTERED.values()) + root.nbytes > _REG_BYTES_MAX:
        return False
    ptr = root.ctypes.data
    try:
        rc = torch.cuda.cudart().cudaHostRegister(ptr, root.nbytes, 0)
    except Exception:
        return False
    if int(rc) != 0:
        return False
    _REGISTERED[key] = (ptr, root.nbytes)
    weakref.finalize(root, _unregister, key, ptr)
    return True


# Result arrays: page-locked memory from a small pool.  cudaHostAlloc of a 128 MiB field costs ~18 ms, more than the
# solve; a buffer goes back to the pool when the NumPy array handed to the caller (and with it every view of it) is
# garbage-collected, so a planner that calls in a loop allocates once.
_POOL = {}                         # nbytes -> [pinned uint8 tensors]
_POOL_KEEP = 4
_OURS = {}                         # id(array handed out) -> weakref to it


def _is_ours(a: np.ndarray) -> bool:
    base = a
    while isinstance(base.base, np.ndarray):
        base = base.base
    r = _OURS.get(id(base))
    return r is not None and r() is base


def _give_back(key: int, nbytes: int, buf: torch.Tensor):
    _OURS.pop(key, None)
    _DEVCOPY.pop(key, None)
    free = _POOL.setdefault(nbytes, [])
    if len(free) < _POOL_KEEP:
        free.append(buf)


def to_host(t: torch.Tensor, keep_device: bool = False) -> np.ndarray:
    """Device tensor -> fresh NumPy array (page-locked memory for large fields, see above).  keep_device: remember the
    device tensor next to the array (see trace_field)."""
    t = t.contiguous()
    if t.numel() < (1 << 18):
        return t.cpu().numpy()
    nbytes = t.numel() * t.element_size()
    free = _POOL.get(nbytes)
    buf = free.pop() if free else torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h = buf.view(t.dtype).view(t.shape)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    a = h.numpy()
    _OURS[id(a)] = weakref.ref(a)
    weakref.finalize(a, _give_back, id(a), nbytes, buf)
    if keep_device:
        _DEVCOPY[id(a)] = t
        while len(_DEVCOPY) > _DEVCOPY_KEEP:
            _DEVCOPY.pop(next(iter(_DEVCOPY)))
    return a


# The planner hands the fields it just received straight to getPathGDM (Coupled_motion_planner.py:1226-1230,
# 1636-1639).  The device tensor a large field was downloaded from is kept (the few most recent ones, released with the
# host array); when that array comes back, the tracer starts on the device copy at once while the array is uploaded
# again on a second stream, and a bitwise comparison of the two (fmb_fields_differ_f64, ~0.05 ms for 128 MiB) decides
# whether the path stands -- a caller who changed the array in between gets the path of the changed array.
_DEVCOPY = {}                      # id(array handed out) -> device tensor it was downloaded from (insertion-ordered)
_DEVCOPY_KEEP = 4
TRACE_STATS = {"reused": 0, "retraced": 0}      # tracer calls served from a kept device copy / repeated on the upload


def _device_copy_of(c: np.ndarray):
    base = c
    while isinstance(base.base, np.ndarray):
        base = base.base
    r = _OURS.get(id(base))
    if r is None or r() is not base:
        return None
    t = _DEVCOPY.get(id(base))
    if t is None or tuple(t.shape) != c.shape or not c.flags.c_contiguous or c.ctypes.data != base.ctypes.data:
        return None
    return t


_LOG_BLOCKS = 4096                 # block log of the 2D tracer (a block per ~30 steps)


def trace_field2d(c: np.ndarray, swap: bool, dev: torch.device, init, end, tau):
    """2D tracer on the C-contiguous host field `c` (swap: the caller's array is its transpose; the tracer is not
    symmetric in x and y -- the reference normalises dx first and reuses it for dy, FastMarching.py:226-227 -- so an
    F-ordered field is put back into [y, x] order on the device).  When the device copy `c` was downloaded from is still
    kept, the path is traced on it at once and the cells it READ -- the windows of its gradient blocks, logged by the
    tracer -- are compared bitwise with the caller's array, which the comparison kernel reads in place over the bus
    (fmb_trace2d_logged_f64 + fmb_windows_differ_f64): nothing of the 128 MiB array moves.  A difference (the caller
    edited the array where the path looked at it) or an overflowing log: upload and trace again."""
    def run(Td, log=0):
        Tt = Td.T.contiguous() if swap else Td
        return Tt, engine.trace2d(Tt, init[None, :], end[None, :], tau, log_blocks=log)
    cached = _device_copy_of(c)
    if cached is None:
        return run(to_device(c, dev))[1]
    Tt, (out, count, status, blocks, nblocks) = run(cached, _LOG_BLOCKS)
    rows, cols = Tt.shape
    hs_y, hs_x = (1, c.shape[1]) if swap else (c.shape[1], 1)          # element (y, x) of the traced field inside `c`
    flag = torch.empty(1, dtype=torch.int32, device=dev)
    rc = _capi.lib().fmb_windows_differ_f64(Tt.data_ptr(), cols, rows, cols, c.ctypes.data, hs_y, hs_x, blocks.data_ptr(),
                                            nblocks.data_ptr(), _LOG_BLOCKS, flag.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0 or int(flag[0]):            # (rc: the array is not addressable by the device -- compare the slow way)
        TRACE_STATS["retraced"] += 1
        return run(to_device(c, dev))[1]
    TRACE_STATS["reused"] += 1
    return out, count, status


def trace_field(c: np.ndarray, dev: torch.device, run):
    """run(Td) -> (paths, count, status) for the C-contiguous float64 host field `c`; returns what run returns."""
    cached = _device_copy_of(c)
    if cached is None:
        return run(to_device(c, dev))
    cur, side = torch.cuda.current_stream(dev), _side_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        Tu = to_device(c, dev)                          # page-locked: one DMA
    Tu.record_stream(cur)
    res = run(cached)
    cur.wait_stream(side)
    flag = torch.empty(1, dtype=torch.int32, device=dev)
    _capi.check(_capi.lib().fmb_fields_differ_f64(cached.data_ptr(), Tu.data_ptr(), cached.numel(), flag.data_ptr(), cur.cuda_stream))
    if int(flag[0]):
        TRACE_STATS["retraced"] += 1
        return run(Tu)
    TRACE_STATS["reused"] += 1
    return res


def as_c_field(a):
    """Return (C-contiguous float64 view or copy, transposed?) for an array of any order.

    The planner passes an F-ordered view (``cMap.T``, Coupled_motion_planner.py:1226).  The
    4-neighbour update is symmetric in its two axes (FastMarching.py:17-29), so an F-ordered
    map is solved as its C-ordered transpose with x and y swapped -- no copy.
    """
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2 and a.flags.f_contiguous and not a.flags.c_contiguous:
        return a.T, True
    return np.ascontiguousarray(a), False


def node2(p, swap):
    x, y = int(p[0]), int(p[1])
    return (y, x) if swap else (x, y)


def check_node2(p, rows, cols):
    x, y = p
    if not (0 <= x < cols and 0 <= y < rows):
        # the reference wraps negative indices / raises IndexError at the array edge
        raise IndexError(f"node {list(p)} is outside the {rows}x{cols} map")


def raise_trace(status: int):
    if status in _EXC:
        exc, msg = _EXC[status]
        raise exc(msg)


# --------------------------------------------------------------------------
# Early-exit emulation.  The reference stops popping as soon as the start node is accepted
# (FastMarching3D.py:141-142) or the two fronts meet (FastMarching.py:150-155) and returns a PARTIAL field: accepted
# cells hold final values, narrow-band cells hold their last tentative value, the rest is +inf.  The whole emulation
# -- full solve(s), pop ranks incl. the reference's LIFO order among exactly equal values, join, replay -- runs
# inside libfm_b200 (csrc/fm_capi_ranks.inc: fmb_bisolve2d_f64, fmb_solve2d_until_f64, fmb_solve3d_until_f64,
# fmb_pop_ranks*): this module only normalises arguments, provides device memory and reads the result back.
import ctypes as C
import warnings

# Relative gap below which two values count as tied.  Zero = exact equality (a tolerance was tried in round 1 and
# misorders far more cells on plateau maps, which hold many DISTINCT values one or two ulp apart).
TIE_TOL_2D = 0.0
TIE_TOL_3D = 0.0

_WS = {}
_SIDE = {}


def _ws(nbytes: int, dev: torch.device, tag: str) -> torch.Tensor:
    key = (dev.index, tag)
    w = _WS.get(key)
    if w is None or w.numel() < nbytes:
        w = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _WS[key] = w
    return w


def _side_stream(dev: torch.device) -> torch.cuda.Stream:
    if dev.index not in _SIDE:
        _SIDE[dev.index] = torch.cuda.Stream(device=dev)
    return _SIDE[dev.index]


def _i32(vals):
    return (C.c_int32 * len(vals))(*[int(v) for v in vals])


def check_info(info, fronts=1):
    """info: the int32[16] block of the one-call entries, already on the host."""
    for f in range(fronts):
        if info[4 + 4 * f + 2] or info[12 + f]:
            raise RuntimeError("early-exit emulation: a dependency wait hit its safety limit")
        if info[4 + 4 * f + 3]:
            warnings.warn("a tie group of more than 4096 cells: the order inside it is the plain stable order, "
                          "not necessarily the reference's (degenerate map)", RuntimeWarning, stacklevel=3)


def pop_ranks2d(T: torch.Tensor, cost: torch.Tensor, seed, transposed: bool = False) -> torch.Tensor:
    """int32 pop ranks of a full 2D field (source 0, unreached INT32_MAX) in the reference's order, ties included
    (fmb_pop_ranks2d_f64).  seed = [x, y] in the field's own orientation."""
    T, cost = T.contiguous(), cost.contiguous()
    rows, cols = T.shape
    L = _capi.lib()
    ws = _ws(L.fmb_workspace_bytes_pop_ranks(rows * cols), T.device, "ranks")
    rank = torch.empty((rows, cols), dtype=torch.int32, device=T.device)
    with torch.cuda.device(T.device):
        _capi.check(L.fmb_pop_ranks2d_f64(T.data_ptr(), cost.data_ptr(), rows, cols, int(seed[1]) * cols + int(seed[0]),
                                          int(bool(transposed)), rank.data_ptr(), ws.data_ptr(), ws.numel(),
                                          torch.cuda.current_stream().cuda_stream))
    return rank


def pop_ranks3d(T: torch.Tensor, cost: torch.Tensor, seed, focus=None) -> torch.Tensor:
    """3D form (fmb_pop_ranks3d_f64); focus = [x, y, z] restricts the tie question to that cell's group."""
    T, cost = T.contiguous(), cost.contiguous()
    ny, nx, nz = T.shape
    idx = lambda p: (int(p[1]) * nx + int(p[0])) * nz + int(p[2])
    L = _capi.lib()
    ws = _ws(L.fmb_workspace_bytes_pop_ranks(ny * nx * nz), T.device, "ranks")
    rank = torch.empty((ny, nx, nz), dtype=torch.int32, device=T.device)
    with torch.cuda.device(T.device):
        _capi.check(L.fmb_pop_ranks3d_f64(T.data_ptr(), cost.data_ptr(), ny, nx, nz, idx(seed), -1 if focus is None else idx(focus),
                                          rank.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    return rank


def ranks_status(dev: torch.device):
    """(ties exist, largest group, waits at the limit, group too large) of the last pop_ranks* call on this device."""
    out = (C.c_int32 * 4)()
    ws = _WS[(dev.index, "ranks")]
    _capi.check(_capi.lib().fmb_pop_ranks_status(ws.data_ptr(), torch.cuda.current_stream().cuda_stream, out))
    return tuple(int(v) for v in out)


def truncate(T: torch.Tensor, cost: torch.Tensor, rank: torch.Tensor, k: int) -> torch.Tensor:
    """Partial field after k pops (accepted final, narrow band tentative, rest +inf), rebuilt on the
    device by libfm_b200's fmb_truncate{2d,3d}_f64 (csrc/truncate.cuh)."""
    T = T.contiguous()
    cost = cost.contiguous()
    rank = rank.contiguous()
    out = torch.empty_like(T)
    scratch = torch.empty(T.numel(), dtype=torch.int32, device=T.device)      # compacted narrow-band cells
    counters = torch.zeros(2, dtype=torch.int32, device=T.device)
    memo = torch.empty(T.numel() * (4 if T.dim() == 2 else 6), dtype=torch.float64, device=T.device)   # shared replay memo
    L = _capi.lib()
    stream = torch.cuda.current_stream().cuda_stream
    if T.dim() == 2:
        rc = L.fmb_truncate2d_f64(T.data_ptr(), cost.data_ptr(), rank.data_ptr(), T.shape[0], T.shape[1], int(k),
                                  out.data_ptr(), scratch.data_ptr(), counters.data_ptr(), memo.data_ptr(), stream)
    else:
        rc = L.fmb_truncate3d_f64(T.data_ptr(), cost.data_ptr(), rank.data_ptr(), T.shape[0], T.shape[1], T.shape[2],
                                  int(k), out.data_ptr(), scratch.data_ptr(), counters.data_ptr(), memo.data_ptr(), stream)
    _capi.check(rc)
    return out


def bi_join(rankG: torch.Tensor, rankS: torch.Tensor):
    """(k, flat index of the join cell) of the two fronts, or (None, None) when they never meet
    (FastMarching.py:141-161): libfm_b200's fmb_bi_join on the int32 pop ranks."""
    out = torch.empty(4, dtype=torch.int32, device=rankG.device)
    _capi.check(_capi.lib().fmb_bi_join(rankG.contiguous().data_ptr(), rankS.contiguous().data_ptr(), rankG.numel(),
                                        out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    k, j = (int(v) for v in out[:2].tolist())
    big = torch.iinfo(torch.int32).max
    return (None, None) if k == big else (k, j)


def bisolve2d(cd, goal, start, transposed: bool, dev: torch.device = None):
    """fmb_bisolve2d_f64: (TG, TS, info) -- both partial fields of biComputeTmap on the device and the int32[16] info
    tensor (k, join cell, statuses), still on the device: nothing here synchronises.  cd: the map on the device, or -- a
    page-locked C-contiguous NumPy array -- still on the host: its upload then runs in bands on the side stream behind
    the two solves that consume it (fmb_bisolve2d_h2d_f64)."""
    on_host = isinstance(cd, np.ndarray)
    rows, cols = cd.shape
    dev = cd.device if not on_host else dev
    L = _capi.lib()
    ws = _ws(L.fmb_workspace_bytes_bisolve2d(rows, cols), dev, "bisolve")
    out = torch.empty((2, rows, cols), dtype=torch.float64, device=dev)
    info = torch.empty(16, dtype=torch.int32, device=dev)
    cur = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.device(dev):
        if on_host:
            cdev = torch.empty((rows, cols), dtype=torch.float64, device=dev)
            _capi.check(L.fmb_bisolve2d_h2d_f64(cd.ctypes.data, cdev.data_ptr(), rows, cols, _i32(goal), _i32(start), int(bool(transposed)),
                                                out[0].data_ptr(), out[1].data_ptr(), info.data_ptr(), ws.data_ptr(), ws.numel(),
                                                cur.cuda_stream, side.cuda_stream))
            cdev.record_stream(side)
        else:
            _capi.check(L.fmb_bisolve2d_f64(cd.data_ptr(), rows, cols, _i32(goal), _i32(start), int(bool(transposed)),
                                            out[0].data_ptr(), out[1].data_ptr(), info.data_ptr(), ws.data_ptr(), ws.numel(),
                                            cur.cuda_stream, side.cuda_stream))
    return out[0], out[1], info, ws


def solve2d_h2d(c: np.ndarray, goal, dev: torch.device):
    """fmb_solve2d_h2d_f64: the full field of the page-locked host map `c`, its upload (bands of rows, nearest to the goal
    first, on the side stream) overlapped with the solve.  Returns (T on the device, workspace)."""
    rows, cols = c.shape
    L = _capi.lib()
    ws = _ws(L.fmb_workspace_bytes_2d_h2d(rows, cols), dev, "h2d")
    cd = torch.empty((rows, cols), dtype=torch.float64, device=dev)
    out = torch.empty((rows, cols), dtype=torch.float64, device=dev)
    cur, side = torch.cuda.current_stream(dev), _side_stream(dev)
    with torch.cuda.device(dev):
        _capi.check(L.fmb_solve2d_h2d_f64(c.ctypes.data, cd.data_ptr(), rows, cols, _i32(goal), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                          cur.cuda_stream, side.cuda_stream))
    cd.record_stream(side)
    return out, ws


def solve2d_until(cd: torch.Tensor, goal, start, transposed: bool):
    rows, cols = cd.shape
    dev = cd.device
    L = _capi.lib()
    ws = _ws(L.fmb_workspace_bytes_until2d(rows, cols), dev, "until")
    out = torch.empty((rows, cols), dtype=torch.float64, device=dev)
    info = torch.zeros(16, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(L.fmb_solve2d_until_f64(cd.data_ptr(), rows, cols, _i32(goal), _i32(start), int(bool(transposed)),
                                            out.data_ptr(), info.data_ptr(), ws.data_ptr(), ws.numel(),
                                            torch.cuda.current_stream(dev).cuda_stream))
    return out, info, ws


def solve3d_until(cd: torch.Tensor, goal, start):
    ny, nx, nz = cd.shape
    dev = cd.device
    L = _capi.lib()
    ws = _ws(L.fmb_workspace_bytes_until3d(ny, nx, nz), dev, "until")
    out = torch.empty((ny, nx, nz), dtype=torch.float64, device=dev)
    info = torch.zeros(16, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(L.fmb_solve3d_until_f64(cd.data_ptr(), ny, nx, nz, _i32(goal), _i32(start), out.data_ptr(), info.data_ptr(),
                                            ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    return out, info, ws


def finish(ws: torch.Tensor, dev: torch.device):
    """Synchronise and raise on a device-side failure of the solve that ran from this workspace."""
    _capi.check(_capi.lib().fmb_finish(ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream, None))
