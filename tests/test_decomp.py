"""Row-slab domain decomposition (optional config 5 path): the exchange/convergence logic with the
CUDA kernel SOURCE run under the CPU emulator (single process, and 2 ranks over gloo)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import emu
from conftest import rand_map, rel_err
from oracle import oracle as O
from planning_motion_planning_b200 import decomp


def emu_resolve(cost, T, seed, activate, halo_rows):
    import torch
    is_t = isinstance(T, torch.Tensor)
    c = np.ascontiguousarray(cost.numpy() if is_t else cost)
    t = np.ascontiguousarray(T.numpy() if is_t else T).copy()
    sd = np.asarray(seed, dtype=np.int32)
    L = emu.lib()
    L.emu_resolve2d_f64.argtypes = [emu.dp, emu.dp, C.c_int, C.c_int, emu.ip, C.c_int, C.c_int, C.c_int]
    rc = L.emu_resolve2d_f64(c.ctypes.data_as(emu.dp), t.ctypes.data_as(emu.dp), t.shape[0], t.shape[1],
                             sd.ctypes.data_as(emu.ip), int(activate), int(halo_rows), 2)
    assert rc == 0
    if is_t:
        T.copy_(torch.from_numpy(t))
    else:
        T[...] = t


def test_slab_bounds_partition():
    for rows in (1, 31, 32, 100, 8192):
        for n in (1, 2, 3, 8):
            cover = []
            for s in range(n):
                lo, hi = decomp.slab_bounds(rows, n, s)
                assert lo % 32 == 0 or lo == rows
                cover.extend(range(lo, hi))
            assert cover == list(range(rows))


@pytest.mark.parametrize("nslabs,goal", [(2, [20, 10]), (3, [50, 90]), (4, [30, 64])])
def test_local_slabs_match_single_solve(nslabs, goal):
    c = rand_map((130, 70), 4)
    c[64, 5:40] = np.inf                    # a wall along a slab cut: the front has to go round it
    T, rounds = decomp.solve2d_slabs_local(c, goal, nslabs, resolve_fn=emu_resolve)
    assert rounds >= 2
    assert rel_err(T, O.computeTmap(c, goal)) < 1e-9


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = rand_map((96, 50), 6)
    lo, hi, T, rounds = decomp.solve2d_slabs_dist(torch.from_numpy(c), [25, 80], resolve_fn=emu_resolve)
    ret.put((rank, lo, hi, T.numpy().copy(), rounds))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_halo_exchange():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    parts = sorted([ret.get(timeout=300) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = rand_map((96, 50), 6)
    ref = O.computeTmap(c, [25, 80])
    T = np.concatenate([p[3] for p in parts], 0)
    assert [p[1] for p in parts] == [0, 64] and parts[1][2] == 96
    assert rel_err(T, ref) < 1e-9
