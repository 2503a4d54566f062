"""Helpers shared by the two drop-in modules: argument normalisation and the
emulation of the reference's early-exit ("partial field") semantics."""
from __future__ import annotations

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
# at 3-4 GB/s; through page-locked memory it moves at PCIe speed.  So: uploads are staged chunk-wise through two small
# pinned buffers (the memcpy of chunk k+1 overlaps the DMA of chunk k), and results are downloaded straight into a
# pinned tensor whose memory backs the returned NumPy array (owned by the caller through the array's base).  An array
# that came from here is recognised as pinned on its way back in (getPathGDM on a field biComputeTmap returned) and
# uploaded with one DMA.
_STAGE = {}
_STAGE_ELEMS = 4 << 20            # 32 MiB of float64 per staging buffer


def to_device(a: np.ndarray, dev: torch.device) -> torch.Tensor:
    """C-contiguous float64 NumPy array -> device tensor of the same shape."""
    t = torch.from_numpy(a)
    if a.size < (1 << 18):
        return t.to(dev)
    try:
        if t.is_pinned():
            return t.to(dev, non_blocking=True)
    except RuntimeError:
        pass
    out = torch.empty(a.shape, dtype=torch.float64, device=dev)
    flat_src, flat_dst = t.reshape(-1), out.reshape(-1)
    key = dev.index
    if key not in _STAGE:
        _STAGE[key] = ([torch.empty(_STAGE_ELEMS, dtype=torch.float64).pin_memory() for _ in range(2)],
                       [torch.cuda.Event(), torch.cuda.Event()])
    bufs, evs = _STAGE[key]
    n = flat_src.numel()
    stream = torch.cuda.current_stream(dev)
    for k, lo in enumerate(range(0, n, _STAGE_ELEMS)):
        hi = min(n, lo + _STAGE_ELEMS)
        b = k & 1
        if k >= 2:
            evs[b].synchronize()                       # the DMA that last read this staging buffer is done
        bufs[b][:hi - lo].copy_(flat_src[lo:hi])
        flat_dst[lo:hi].copy_(bufs[b][:hi - lo], non_blocking=True)
        evs[b].record(stream)
    return out


def to_host(t: torch.Tensor) -> np.ndarray:
    """Device tensor -> fresh NumPy array (page-locked memory for large fields, see above)."""
    t = t.contiguous()
    if t.numel() < (1 << 18):
        return t.cpu().numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


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
# Early-exit emulation.  The reference stops popping as soon as
# the start node is accepted (FastMarching3D.py:141-142) or the two fronts meet
# (FastMarching.py:150-155) and returns a PARTIAL field: accepted cells hold final
# values, narrow-band cells hold their last tentative value, the rest is +inf.
# The solver always produces the full field; pop ranks are recovered by a stable
# sort of T (ties in row-major order) and the partial field is rebuilt from them.
def pop_ranks(T: torch.Tensor) -> torch.Tensor:
    """int32 rank[c] = number of nodes popped before-or-with c (source = 0, unreached = INT32_MAX).
    Stable ascending sort of T: exact whenever no two cells carry exactly the same value."""
    flat = T.reshape(-1)
    order = torch.sort(flat, stable=True).indices
    rank = torch.empty(flat.numel(), dtype=torch.int32, device=T.device)
    rank[order] = torch.arange(flat.numel(), dtype=torch.int32, device=T.device)
    rank[~torch.isfinite(flat)] = torch.iinfo(torch.int32).max
    return rank.reshape(T.shape)


_BIG = torch.iinfo(torch.int64).max
# Relative gap below which two values count as tied.  Zero = exact equality: a tolerance was tried (the device
# field is within 3 ulp of the reference's in 2D, so a pair the reference ties exactly can differ by an ulp
# here) but on plateau maps the reference itself holds many DISTINCT values one or two ulp apart, and merging
# those misorders far more cells (154 vs 1 on the case that motivated it).
TIE_TOL_2D = 0.0
EXACT_3D = True          # FastMarching3D.computeTmap: follow the solve with the exact polish pass (fmb_polish3d_f64): the field then
                         # carries the reference's own rounding (libm pow for `**2` on scalars) and exact ties are the reference's:
                         # tools/gpu_fuzz_3d.py 300 volumes: 0 pattern mismatches with it, 13 without
TIE_TOL_3D = 0.0


def _shift(a, dy, dx, fill):
    out = torch.full_like(a, fill)
    H, W = a.shape
    ys, yd = slice(max(0, dy), H + min(0, dy)), slice(max(0, -dy), H + min(0, -dy))
    xs, xd = slice(max(0, dx), W + min(0, dx)), slice(max(0, -dx), W + min(0, -dx))
    out[yd, xd] = a[ys, xs]
    return out


def _lex_order(T, k2, k3):
    """argsort by (T, k2, k3) ascending via successive stable sorts."""
    order = torch.sort(k3.reshape(-1), stable=True).indices
    order = order[torch.sort(k2.reshape(-1)[order], stable=True).indices]
    return order[torch.sort(T.reshape(-1)[order], stable=True).indices]


def pop_ranks_lifo2d(T: torch.Tensor, cost: torch.Tensor, seed, max_iters: int = 96, transposed: bool = False) -> torch.Tensor:
    """Pop ranks of the reference's 2D front INCLUDING its order among exactly equal values.

    The reference keeps the narrow band sorted with bisect_left + insert (FastMarching.py:65-67,
    76-78): among equal T the node (re)inserted LAST pops first.  A node's final value is inserted
    at the first pop of one of its neighbours at which the update, fed with the neighbour values
    that are final by then, already yields the final value (:57-62), and within one updateNode
    call in child order (:46-54).  So the pop order is the ascending order of
    (T, -insertion time, -child index), where insertion times depend on the ranks themselves:
    iterate to the fixed point.  Maps without exact ties return after the plain sort.
    ``transposed``: T is the transpose of the caller's map (an F-ordered input solved as its C-ordered
    transpose, see as_c_field); the child order is not symmetric in x and y, so it is mapped back.
    Measured against the reference's true pop order (oracle): 0 misplaced cells on every uniform,
    plateau and random map tried (the plain sort misplaces thousands on tie-heavy maps)."""
    H, W = T.shape
    fin = torch.isfinite(T)
    flat = T.reshape(-1)
    seed_idx = int(seed[1]) * W + int(seed[0])
    if T.is_cuda:
        return _pop_ranks_lifo2d_cuda(T.contiguous(), cost.contiguous(), seed_idx, max_iters, transposed)
    nfin = int(fin.sum())
    if nfin == int(torch.unique(flat[fin.reshape(-1)]).numel()):
        return pop_ranks(T)                                       # no ties: the sort is already exact
    idx = torch.arange(H * W, device=T.device).reshape(H, W)
    order = _lex_order(T, torch.zeros_like(idx), idx)
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=T.device)
    rank = rank.reshape(H, W)
    tau = rank.clone()
    INF = float("inf")
    TL, TR, TU, TD = _shift(T, 0, -1, INF), _shift(T, 0, 1, INF), _shift(T, -1, 0, INF), _shift(T, 1, 0, INF)
    for _ in range(max_iters):
        r = torch.where(fin, rank, torch.full_like(rank, _BIG))
        t0 = torch.where(fin, tau, torch.full_like(tau, _BIG))
        r.view(-1)[seed_idx] = 0
        t0.view(-1)[seed_idx] = -1                                # the source is final before anything pops
        RL, RR, RU, RD = _shift(r, 0, -1, _BIG), _shift(r, 0, 1, _BIG), _shift(r, -1, 0, _BIG), _shift(r, 1, 0, _BIG)
        AL, AR, AU, AD = _shift(t0, 0, -1, _BIG), _shift(t0, 0, 1, _BIG), _shift(t0, -1, 0, _BIG), _shift(t0, 1, 0, _BIG)
        # insertion time = the earliest neighbour pop at which the update, fed only with neighbour values
        # that are already final by then (the others count as +inf), reproduces the final value
        limit = T * (1.0 + 1e-14)
        tau_new = torch.full_like(r, _BIG)
        cidx = torch.zeros_like(r)
        inf_t = torch.full_like(T, INF)
        for R, ci in (((RL, 2), (RR, 1), (RU, 4), (RD, 3)) if transposed else ((RL, 4), (RR, 3), (RU, 2), (RD, 1))):   # child index w.r.t. the popped neighbour
            lt = torch.where(AL <= R, TL, inf_t)
            rt = torch.where(AR <= R, TR, inf_t)
            ut = torch.where(AU <= R, TU, inf_t)
            dt = torch.where(AD <= R, TD, inf_t)
            a, b = torch.minimum(lt, rt), torch.minimum(ut, dt)
            dd = a - b
            one = torch.minimum(a, b) + cost
            two = 0.5 * (a + b + torch.sqrt((2.0 * (cost * cost) - dd * dd).clamp_min(0.0)))
            v = torch.where(dd.abs() <= cost, two, one)
            ok = (R < _BIG) & (R < tau_new) & (v <= limit)
            tau_new = torch.where(ok, R, tau_new)
            cidx = torch.where(ok, torch.full_like(r, ci), cidx)
        tau_new = torch.where(fin, tau_new, torch.full_like(tau_new, _BIG))
        tau_new.view(-1)[seed_idx] = -1
        k2 = torch.where(fin, -tau_new, torch.zeros_like(tau_new))
        k3 = torch.where(fin, -cidx, torch.zeros_like(cidx))
        order = _lex_order(T, k2, k3)
        new_rank = torch.empty_like(order)
        new_rank[order] = torch.arange(order.numel(), device=T.device)
        new_rank = new_rank.reshape(H, W)
        done = torch.equal(new_rank, rank) and torch.equal(tau_new, tau)
        rank, tau = new_rank, tau_new
        if done:
            break
    out = rank.to(torch.int32)
    out[~fin] = torch.iinfo(torch.int32).max
    return out


def _pop_ranks_lifo2d_cuda(T, cost, seed_idx: int, max_iters: int, transposed: bool = False) -> torch.Tensor:
    """Device path of :func:`pop_ranks_lifo2d`.  One sort of T gives the tie groups; the order inside
    them is then settled by ONE kernel of libfm_b200 (csrc/tiekeys.cuh, tie_sweep_kernel<2|3>): only
    strictly upwind neighbours take part in a cell's final update and they pop before the cell's
    group starts, so the fixed point of the iteration above is reached group by group in ascending T
    without iterating.  Maps with a tie group of more than 4096 cells use the iterated form."""
    shape = tuple(T.shape)
    n = T.numel()
    dev = T.device
    flat = T.reshape(-1)
    fin = torch.isfinite(flat)
    order = torch.sort(flat, stable=True).indices
    ts = flat[order]
    # tie groups (TIE_TOL_* = 0: exact equality, see the note at their definition)
    tol = TIE_TOL_2D if len(shape) == 2 else TIE_TOL_3D
    new_grp = torch.cat([torch.ones(1, dtype=torch.bool, device=dev), (ts[1:] - ts[:-1]) > tol * ts[1:]])
    new_grp = new_grp | ~torch.isfinite(ts)
    rank = torch.empty(n, dtype=torch.int32, device=dev)
    rank[order] = torch.arange(n, dtype=torch.int32, device=dev)
    if not bool((~new_grp).any()):                               # no two reached cells share a value: the sort is the pop order
        rank[~fin] = torch.iinfo(torch.int32).max
        return rank.reshape(shape)
    grp_sorted = torch.cumsum(new_grp.to(torch.int32), 0) - 1
    starts = torch.nonzero(new_grp).reshape(-1).to(torch.int32)                 # first sorted position of each group
    sizes = torch.diff(torch.cat([starts, torch.tensor([n], dtype=torch.int32, device=dev)]))
    group = torch.empty(n, dtype=torch.int32, device=dev)
    group[order] = grp_sorted.to(torch.int32)
    gstart = starts[group.long()].contiguous()
    gsize = torch.where(fin, sizes[group.long()], torch.ones_like(group)).to(torch.int32).contiguous()   # unreached cells: no re-ranking
    members = order.to(torch.int32).contiguous()
    if int(gsize.max()) > 4096:           # a degenerate map: the quadratic in-group count would dominate
        if len(shape) == 2:
            return _pop_ranks_lifo2d_sort(T, cost, seed_idx, max_iters, group, rank, transposed)
        rank[~fin] = torch.iinfo(torch.int32).max                 # 3D: the plain sort (ties in sorted order)
        return rank.reshape(shape)
    tau = torch.empty_like(rank)
    key = torch.empty(n, dtype=torch.int64, device=dev)
    scratch = torch.empty(2 * n + 2, dtype=torch.int32, device=dev)
    L = _capi.lib()
    tail = (rank.data_ptr(), tau.data_ptr(), key.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if len(shape) == 2:
        _capi.check(L.fmb_tie_order2d_f64(T.data_ptr(), cost.data_ptr(), members.data_ptr(), gstart.data_ptr(), gsize.data_ptr(),
                                          *shape, seed_idx, int(bool(transposed)), *tail))
    else:
        _capi.check(L.fmb_tie_order3d_f64(T.data_ptr(), cost.data_ptr(), members.data_ptr(), gstart.data_ptr(), gsize.data_ptr(),
                                          *shape, seed_idx, *tail))
    if int(scratch[-1]) != 0:
        raise RuntimeError("tie-order sweep: a dependency wait hit its safety limit")
    return rank.reshape(shape)


def pop_ranks_lifo3d(T: torch.Tensor, cost: torch.Tensor, seed) -> torch.Tensor:
    """3D pop ranks incl. the reference's LIFO order among equal values (FastMarching3D.py:22-33 child
    order, :77-95 bisect_left insertion): the same ordered sweep as in 2D on the device; CPU tensors
    (no product path uses them) get the plain stable sort."""
    if not T.is_cuda:
        return pop_ranks(T)
    ny, nx, nz = T.shape
    seed_idx = (int(seed[1]) * nx + int(seed[0])) * nz + int(seed[2])
    return _pop_ranks_lifo2d_cuda(T.contiguous(), cost.contiguous(), seed_idx, 96)


def _pop_ranks_lifo2d_sort(T, cost, seed_idx: int, max_iters: int, group, rank, transposed: bool = False) -> torch.Tensor:
    """Fallback of the device path for maps with a huge tie group: one global stable sort per step."""
    H, W = T.shape
    n = H * W
    dev = T.device
    fin = torch.isfinite(T.reshape(-1))
    ar = torch.arange(n, dtype=torch.int32, device=dev)
    tau = rank.clone()
    tau_new = torch.empty_like(tau)
    key = torch.empty(n, dtype=torch.int64, device=dev)
    L = _capi.lib()
    stream = torch.cuda.current_stream().cuda_stream
    for it in range(max_iters):
        _capi.check(L.fmb_tie_keys2d_f64(T.data_ptr(), cost.data_ptr(), rank.data_ptr(), tau.data_ptr(), group.data_ptr(),
                                         H, W, seed_idx, int(bool(transposed)), tau_new.data_ptr(), key.data_ptr(), stream))
        order = torch.sort(key, stable=True).indices
        new_rank = torch.empty_like(rank)
        new_rank[order] = ar
        # the convergence test synchronises with the device: only every fourth iteration
        done = (it & 3) == 3 and torch.equal(new_rank, rank) and torch.equal(tau_new, tau)
        rank, tau, tau_new = new_rank, tau_new, tau
        if done:
            break
    out = rank.clone()
    out[~fin] = torch.iinfo(torch.int32).max
    return out.reshape(H, W)


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


_SIDE = {}


def both_fronts(fn_g, fn_s):
    """Run the G-front and the S-front piece of biComputeTmap concurrently: each on its own CUDA
    stream and host thread (the rank computation has host syncs; they release the GIL), joined back
    into the current stream.  Returns (result_g, result_s)."""
    from concurrent.futures import ThreadPoolExecutor
    cur = torch.cuda.current_stream()
    dev = torch.cuda.current_device()
    if dev not in _SIDE:
        _SIDE[dev] = (torch.cuda.Stream(), torch.cuda.Stream(), ThreadPoolExecutor(max_workers=2))
    sa, sb, pool = _SIDE[dev]

    def run(stream, fn):
        torch.cuda.set_device(dev)
        with torch.cuda.stream(stream):
            stream.wait_stream(cur)
            return fn()
    fa, fb = pool.submit(run, sa, fn_g), pool.submit(run, sb, fn_s)
    ra, rb = fa.result(), fb.result()
    cur.wait_stream(sa)
    cur.wait_stream(sb)
    return ra, rb
