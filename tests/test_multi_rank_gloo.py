"""CPU, world_size 2 over gloo: the N>1 path of the batched-query API (sharding + optional
gather).  The per-chunk solver is injected (the CUDA solver cannot run here); the real
multi-GPU run is exercised by bench.py under torchrun on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from planning_motion_planning_b200 import batch


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = batch.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n and hi - lo in (n // world, n // world + 1)
                cover.extend(range(lo, hi))
            assert cover == list(range(n))


def _fake_solver(cost, goals, starts, tau):
    # deterministic function of the query alone (what independence of world size requires)
    out = []
    for i, (g, s) in enumerate(zip(goals, starts)):
        base = float(cost[i].sum()) if getattr(cost, "ndim", 2) == 3 else float(cost.sum())
        out.append((np.array([s, g], dtype=np.float64) + base, 0))
    return out


def _worker(rank, world, port, Q, per_query, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    goals = rng.integers(1, 30, size=(Q, 2)).tolist()
    starts = rng.integers(1, 30, size=(Q, 2)).tolist()
    cost = rng.random((Q, 4, 4)) if per_query else rng.random((4, 4))
    lo, res = batch.solve_queries(cost, goals, starts, chunk=3, gather=True, solve_fn=_fake_solver)
    if rank == 0:
        ret.put([(p.tolist(), st) for p, st in res])
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("per_query", [False, True])
def test_two_ranks_give_the_single_rank_answer(per_query):
    Q = 11
    rng = np.random.default_rng(0)
    goals = rng.integers(1, 30, size=(Q, 2)).tolist()
    starts = rng.integers(1, 30, size=(Q, 2)).tolist()
    cost = rng.random((Q, 4, 4)) if per_query else rng.random((4, 4))
    _, single = batch.solve_queries(cost, goals, starts, chunk=4, solve_fn=_fake_solver)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, Q, per_query, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(got) == Q
    for (p2, s2), (p1, s1) in zip(got, single):
        assert s2 == s1 and np.array_equal(np.array(p2), p1)
