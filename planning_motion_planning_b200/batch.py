"""Batches of independent planning queries, sharded across the GPUs of one node.

SURVEY.md 8e: goal sweeps / many costmaps are independent units -> query q belongs to exactly
one rank (contiguous block partition), every rank runs the same single-GPU path on its shard,
and there is NO collective on the data path.  ``torch.distributed`` is only used, optionally,
to gather the (small) waypoint lists on rank 0 afterwards (NCCL or gloo, whatever the
process group was initialised with).  The reference has no counterpart: it plans one query
per interpreter (Coupled_motion_planner.py:1092).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

_LANE_STREAMS = {}            # (device, lanes) -> streams of solve_queries' lanes (kept: their memory pools stay warm)
FIRST_PASS_CROSSINGS = 4      # path slab of the first tracing pass, in crossings of the map (solve_chunk_gpu)


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition: ranks < n % world get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist_info():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), dist
    except Exception:
        pass
    return 0, 1, None


class PackedPaths:
    """Paths of a chunk of queries as three arrays: ``flat`` (K, D) all waypoints back to back, ``counts`` (n,) rows per
    query, ``status`` (n,) tracer status.  ``unpack()`` gives the ``[(path, status), ...]`` list (views into ``flat``)."""
    __slots__ = ("flat", "counts", "status", "dev")

    def __init__(self, flat, counts, status, dev=None):
        self.flat, self.counts, self.status = flat, counts, status
        self.dev = dev                 # the device tensor `flat` was downloaded from, when there is one (gather over NCCL)

    def __len__(self):
        return len(self.counts)

    def unpack(self):
        ends = np.cumsum(self.counts)
        return [(self.flat[ends[i] - self.counts[i]:ends[i]], int(self.status[i])) for i in range(len(self.counts))]


def solve_chunk_gpu(cost, goals, starts, tau: float = 0.5):
    """Default per-chunk worker: full-field solve + one path per query on the current CUDA
    device.  ``cost`` is one shared (rows, cols) map or (n, rows, cols) maps (numpy or torch).
    Only the rows that were written travel to the host: the tracer's slab is (n, 30002, 2) fp64 = 480 KB per query; the
    written rows are packed on the device (fmb_path_pack_f64) and copied in one piece."""
    import torch
    from . import _capi, engine
    dev = torch.device("cuda", torch.cuda.current_device())
    c = cost if isinstance(cost, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(cost))
    c = c.to(dev, non_blocking=True)
    n = len(goals)
    T = engine.solve2d(c, np.asarray(goals, dtype=np.int32), nq=n, sync=False)
    # Path slabs: the reference's step cap (round(15000 / tau) = 30000 steps) would make every query's slab 480 KB -- 2 GB
    # for 4096 queries.  The first pass gives every path room for four crossings of the map; the few paths that use up
    # that room (they ran into the smaller cap, not into the goal) are traced again with the reference's own cap.
    full_steps = int(round(15000 / tau))
    rows, cols = T.shape[-2:]
    first_steps = min(full_steps, int(FIRST_PASS_CROSSINGS * (rows + cols) / tau) + 64)
    init_h, end_h = np.asarray(starts, dtype=np.float64), np.asarray(goals, dtype=np.float64)
    out, cnt, st = engine.trace2d(T, init_h, end_h, tau, max_steps=first_steps)
    engine.finish(dev)
    cnt_h = cnt.cpu().numpy()
    redo = np.nonzero(cnt_h >= first_steps + 2)[0] if first_steps < full_steps else np.zeros(0, dtype=np.int64)
    extra = None
    if len(redo):
        extra = engine.trace2d(T, init_h[redo], end_h[redo], tau, field_of_path=redo.astype(np.int32), max_steps=full_steps)
        cnt[torch.from_numpy(redo).to(dev)] = 0                   # packed separately below
        cnt_h = cnt.cpu().numpy()
    off = torch.cumsum(cnt.to(torch.int64), 0) - cnt.to(torch.int64)
    total = int(cnt_h.sum())
    packed = torch.empty((max(total, 1), 2), dtype=torch.float64, device=dev)
    _capi.check(_capi.lib().fmb_path_pack_f64(out.data_ptr(), cnt.data_ptr(), off.data_ptr(), out.shape[1], n, 2, 1.0, 0.0,
                                              packed.data_ptr(), torch.cuda.current_stream().cuda_stream))
    if extra is not None:
        # splice the re-traced paths back in, in query order
        flat, st_h = packed[:total].cpu().numpy(), st.cpu().numpy()
        e_out, e_cnt, e_st = (t.cpu().numpy() for t in extra)
        ends = np.cumsum(cnt_h)
        pieces, counts = [], cnt_h.copy()
        where = {int(q): k for k, q in enumerate(redo)}
        for q in range(n):
            if q in where:
                k = where[q]
                pieces.append(e_out[k, :e_cnt[k]]); counts[q] = e_cnt[k]; st_h[q] = e_st[k]
            else:
                pieces.append(flat[ends[q] - cnt_h[q]:ends[q]])
        return PackedPaths(np.concatenate(pieces) if pieces else np.zeros((0, 2)), counts, st_h)
    return PackedPaths(packed[:total].cpu().numpy(), cnt_h, st.cpu().numpy(), dev=packed[:total])


def solve_queries(cost, goals: Sequence, starts: Sequence, tau: float = 0.5, chunk: int = 64,
                  gather: bool = False, solve_fn: Optional[Callable] = None, lanes: int = 2):
    """Plan ``len(goals)`` independent queries (path from starts[q] to goals[q]).

    cost: (rows, cols) shared by all queries, or (Q, rows, cols) one map per query.
    Each rank processes its contiguous shard in chunks of ``chunk`` queries (bounds device
    memory: chunk * rows * cols * 8 B of fields).  Returns ``(lo, results)`` with
    ``results[i] = (path ndarray (K,2), status)`` for global query ``lo + i``; with
    ``gather=True`` rank 0 instead gets ``(0, all results in query order)`` and the other
    ranks ``(lo, their own)``.  ``lanes``: chunks in flight per rank on the default GPU worker (each lane a host thread
    with its own CUDA stream and workspace: the download and host-side unpacking of one chunk overlap the solve of the
    next; device memory = lanes * chunk fields).
    """
    rank, world, dist = _dist_info()
    Q = len(goals)
    if len(starts) != Q:
        raise ValueError("goals and starts must have the same length")
    per_query = hasattr(cost, "ndim") and cost.ndim == 3 or (hasattr(cost, "dim") and cost.dim() == 3)
    if per_query and cost.shape[0] != Q:
        raise ValueError("per-query costmaps must have one map per query")
    lo, hi = shard_bounds(Q, rank, world)
    fn = solve_fn or solve_chunk_gpu
    parts: List = []                      # per chunk: a PackedPaths or a plain [(path, status), ...] list (custom solve_fn)
    spans = [(a, min(hi, a + max(1, chunk))) for a in range(lo, hi, max(1, chunk))]

    def one(a, b):
        cc = cost[a:b] if per_query else cost
        return fn(cc, [list(g) for g in goals[a:b]], [list(s) for s in starts[a:b]], tau)
    if solve_fn is None and lanes > 1 and len(spans) > 1:
        import torch
        from concurrent.futures import ThreadPoolExecutor
        dev = torch.cuda.current_device()
        if (dev, lanes) not in _LANE_STREAMS:
            _LANE_STREAMS[(dev, lanes)] = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
        streams = _LANE_STREAMS[(dev, lanes)]
        ready = torch.cuda.Event()
        ready.record()                     # whatever the caller queued (the cost map's upload) comes first

        import itertools
        import threading
        tls, next_lane = threading.local(), itertools.count()

        def lane_job(k):
            if not hasattr(tls, "stream"):             # one stream (hence one engine workspace) per worker thread
                tls.stream = streams[next(next_lane)]
                tls.stream.wait_event(ready)
            with torch.cuda.device(dev), torch.cuda.stream(tls.stream):
                return one(*spans[k])
        with ThreadPoolExecutor(max_workers=lanes) as pool:
            parts = list(pool.map(lane_job, range(len(spans))))
    else:
        parts = [one(a, b) for a, b in spans]

    def as_packed(ps):
        """all chunks of one rank as ONE PackedPaths (three arrays pickle / travel much faster than thousands of small ones)"""
        if not ps:
            return PackedPaths(np.zeros((0, 2)), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))
        if all(isinstance(p, PackedPaths) for p in ps):
            dev = None
            if all(p.dev is not None for p in ps):
                import torch
                dev = ps[0].dev if len(ps) == 1 else torch.cat([p.dev for p in ps])
            return PackedPaths(np.concatenate([p.flat for p in ps]), np.concatenate([p.counts for p in ps]),
                               np.concatenate([p.status for p in ps]), dev=dev)
        lst = [r for p in ps for r in (p.unpack() if isinstance(p, PackedPaths) else p)]
        D = lst[0][0].shape[1] if lst and lst[0][0].ndim == 2 else 2
        return PackedPaths(np.concatenate([np.asarray(r[0], dtype=np.float64).reshape(-1, D) for r in lst]) if lst else np.zeros((0, D)),
                           np.array([len(r[0]) for r in lst], dtype=np.int32), np.array([r[1] for r in lst], dtype=np.int32))
    mine = as_packed(parts)
    if gather and dist is not None and world > 1:
        # Three arrays per rank travel as TENSORS through the process group's own transport (NCCL: device buffers over
        # NVLink, the waypoints straight from the device copy the chunks kept; gloo: host buffers) -- no pickling of tens
        # of megabytes of waypoints.  Shards are contiguous blocks in rank order, so rank order is query order.
        import torch
        on_gpu = dist.get_backend() == "nccl"
        tdev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
        D = mine.flat.shape[1] if mine.flat.ndim == 2 else 2
        meta = torch.tensor([lo, len(mine.counts), mine.flat.shape[0], D], dtype=torch.int64, device=tdev)
        metas = [torch.zeros_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta)
        metas = [[int(v) for v in m.tolist()] for m in metas]
        if rank == 0:
            bufs, reqs = [], []
            for r in range(1, world):
                _, n_r, k_r, d_r = metas[r]
                f = torch.empty((k_r, d_r), dtype=torch.float64, device=tdev)
                cs = torch.empty((2, n_r), dtype=torch.int32, device=tdev)
                if k_r:
                    reqs.append(dist.irecv(f, src=r))
                if n_r:
                    reqs.append(dist.irecv(cs, src=r))
                bufs.append((f, cs))
            for q in reqs:
                q.wait()
            pos = len(mine.counts)
            flats, cnts, sts = [mine.flat.reshape(-1, D)], [mine.counts], [mine.status]
            for r, (f, cs) in enumerate(bufs, start=1):
                assert metas[r][0] == pos
                pos += metas[r][1]
                cs_h = cs.cpu().numpy()
                flats.append(f.cpu().numpy()); cnts.append(cs_h[0]); sts.append(cs_h[1])
            merged = PackedPaths(np.concatenate(flats), np.concatenate(cnts), np.concatenate(sts))
            return 0, merged.unpack()
        if len(mine.counts):
            f = mine.dev if (on_gpu and mine.dev is not None) else torch.from_numpy(np.ascontiguousarray(mine.flat, dtype=np.float64)).to(tdev)
            cs = torch.from_numpy(np.stack([np.asarray(mine.counts, dtype=np.int32), np.asarray(mine.status, dtype=np.int32)])).to(tdev)
            if f.is_cuda:
                f.record_stream(torch.cuda.current_stream())
            if f.shape[0]:
                dist.send(f.contiguous(), dst=0)
            dist.send(cs.contiguous(), dst=0)
    return lo, mine.unpack()
