"""TEST INFRASTRUCTURE: torch reference implementations of the reference's pop order among exactly equal values
(moved out of FastMarching/_compat.py: the product computes ranks inside libfm_b200, fmb_pop_ranks*).

pop_ranks_lifo2d on CPU tensors is the iterated fixed-point form the device kernels were derived from; it is compared
with the oracle's true pop order (tests/test_host_side.py) and with the library (tests/test_gpu_parity.py).
pop_ranks_lifo2d_sort drives the iterated device form fmb_tie_keys2d_f64 with one global sort per step."""
from __future__ import annotations

import numpy as np
import torch

from planning_motion_planning_b200 import _capi

TIE_TOL_2D = 0.0
TIE_TOL_3D = 0.0


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



def pop_ranks_lifo2d_sort(T, cost, seed_idx: int, max_iters: int, group, rank, transposed: bool = False) -> torch.Tensor:
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


