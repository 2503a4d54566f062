"""CPU: the CUDA kernel SOURCE (csrc/*.cuh) executed under the host SIMT emulator
(tools/host_emu) and compared with the oracle.  This checks kernel logic -- tile staging,
masks, queue/tile-state protocol, tracer quirks -- where no GPU exists; the real parity
tests (-m gpu) run the compiled sm_100a code through the C ABI."""
import numpy as np
import pytest

import emu
from conftest import GOLDEN, plateau_map, rand_map, rel_err
from oracle import oracle as O

TOL64 = 1e-9          # north_star: fp64 field within 1e-9 relative
TOL32 = 1e-4          # fp32 variant within 1e-4
TOLP = 1e-3           # waypoints within 1e-3 cell


@pytest.mark.parametrize("shape,goal,tw", [((9, 9), [4, 4], 32), ((64, 64), [40, 12], 32), ((64, 64), [40, 12], 16),
                                           ((45, 97), [95, 1], 32), ((33, 65), [32, 31], 16)])
def test_solve2d_random(shape, goal, tw):
    c = rand_map(shape, 7)
    T, st = emu.solve2d(c, [goal], tw=tw)
    assert rel_err(T[0], O.computeTmap(c, goal)) < TOL64
    assert st["visits"] > 0


@pytest.mark.parametrize("tw", [32, 16])
def test_solve2d_interior_tiles_take_the_cp_async_path(tw):
    """100x100 has fully interior tiles (16-byte aligned rows): exercises the bulk-staging path."""
    c = rand_map((100, 100), 0)
    T, _ = emu.solve2d(c, [[25, 25]], tw=tw)
    assert rel_err(T[0], O.computeTmap(c, [25, 25])) < TOL64
    c = rand_map((101, 99), 1)                 # odd pitch: register path everywhere
    T, _ = emu.solve2d(c, [[70, 40]], tw=tw)
    assert rel_err(T[0], O.computeTmap(c, [70, 40])) < TOL64


def test_solve2d_plateau_and_walls():
    c = plateau_map(80, 3)
    c[30:50, 40] = np.inf              # a wall
    c[10, 10:30] = np.inf
    T, _ = emu.solve2d(c, [[20, 60]])
    assert rel_err(T[0], O.computeTmap(c, [20, 60])) < TOL64


def test_solve2d_enclosed_region_stays_inf():
    c = rand_map((40, 40), 1)
    c[10:20, 10] = c[10:20, 19] = c[10, 10:20] = c[19, 10:20] = np.inf
    T, _ = emu.solve2d(c, [[3, 3]])
    assert np.all(np.isinf(T[0][11:19, 11:19]))
    assert rel_err(T[0], O.computeTmap(c, [3, 3])) < TOL64


def test_solve2d_batch_shared_and_per_query_maps():
    c = rand_map((48, 70), 2)
    goals = [[5, 5], [60, 40], [31, 32], [32, 31]]      # incl. seeds on tile edges
    T, _ = emu.solve2d(c, goals, nblocks=3)
    for q, g in enumerate(goals):
        assert rel_err(T[q], O.computeTmap(c, g)) < TOL64
    cs = np.stack([rand_map((40, 40), s) for s in (1, 2, 3)])
    T, _ = emu.solve2d(cs, [[20, 20]] * 3, shared=False)
    for q in range(3):
        assert rel_err(T[q], O.computeTmap(cs[q], [20, 20])) < TOL64


def test_solve2d_fp32_variant():
    c = rand_map((100, 100), 4)
    T, _ = emu.solve2d(c.astype(np.float32), [[10, 50]])
    ref = O.computeTmap(c.astype(np.float32).astype(np.float64), [10, 50])
    assert rel_err(T[0].astype(np.float64), ref) < TOL32


@pytest.mark.parametrize("shape,goal,tz", [((9, 9, 9), [4, 4, 4], 32), ((20, 20, 20), [4, 4, 4], 32), ((13, 26, 64), [20, 3, 50], 32),
                                           ((13, 21, 40), [10, 5, 33], 16), ((10, 12, 28), [8, 7, 6], 32)])
def test_solve3d_random(shape, goal, tz):
    c = rand_map(shape, 5)
    if shape == (20, 20, 20):
        c[8:10, 3:15, 3:15] = np.inf
    T, _ = emu.solve3d(c, [goal], tz=tz)
    assert rel_err(T[0], O.computeTmap3D(c, goal)) < TOL64


def test_solve3d_fp32_variant():
    c = rand_map((12, 16, 20), 6)
    T, _ = emu.solve3d(c.astype(np.float32), [[5, 5, 5]])
    ref = O.computeTmap3D(c.astype(np.float32).astype(np.float64), [5, 5, 5])
    assert rel_err(T[0].astype(np.float64), ref) < TOL32


def test_trace2d_matches_oracle_paths():
    g = np.load(f"{GOLDEN}/ref2d.npz")
    c = rand_map((100, 100), 0)
    TG, TS, j = O.biComputeTmap(c, [10, 10], [90, 90])
    paths, st = emu.trace2d(np.stack([TG, TS]), [j, j], [[10, 10], [90, 90]])
    assert st == [0, 0]
    assert np.abs(paths[0] - g["kat3_pathG"]).max() < TOLP and np.abs(paths[1] - g["kat3_pathS"]).max() < TOLP
    assert np.array_equal(paths[0][0], j.astype(float)) and np.array_equal(paths[0][-1], [10, 10])


def test_trace2d_nan_fallback_and_errors():
    c = rand_map((30, 30), 2)
    c[15, 5:25] = np.inf
    T = O.computeTmap(c, [5, 5], [25, 25])            # truncated field: start sits on the front
    for init in ([25, 25], [25, 24], [14, 16]):
        po, so = O.getPathGDM(T, np.array(init), [5, 5], 0.5, return_status=True)
        p, s = emu.trace2d(T, [init], [[5, 5]])
        assert s[0] == so and p[0].shape == po.shape
        if len(po):
            assert np.abs(p[0] - po).max() < TOLP
    # stencil leaving the map -> IndexError status
    p, s = emu.trace2d(T, [[29.2, 29.2]], [[5, 5]])
    assert s[0] == 3 == O.getPathGDM(T, np.array([29.2, 29.2]), [5, 5], 0.5, return_status=True)[1]


def test_trace3d_matches_oracle_paths():
    g = np.load(f"{GOLDEN}/ref3d.npz")
    c = rand_map((24, 24, 24), 0)
    for tag, start in (("full", None), ("trunc", [18, 17, 16])):
        F = O.computeTmap3D(c, [5, 6, 7], start)
        p, s = emu.trace3d(F, [[18, 17, 16]], [[5, 6, 7]])
        assert s == [0]
        assert np.abs(p[0] - g[f"kat4_path_{tag}"]).max() < TOLP


@pytest.mark.parametrize("replay", ["dense", "sparse"])
def test_truncate_rebuilds_reference_partial_fields(replay, monkeypatch):
    """csrc/truncate.cuh under the emulator: early-exit (2D/3D) and bidirectional partial fields
    equal the reference's, including every narrow-band tentative value -- with the dense replay (every relaxation of
    the first k pops) and with the sparse one (only the dependency cone of the narrow band: cone_* kernels, sorted
    tickets, list sweep; emu.truncate returns -1 if that form gave up)."""
    monkeypatch.setenv("FMB_REPLAY_SPARSE", "1" if replay == "sparse" else "0")
    for c, g, s in ((rand_map((64, 64), 7), [40, 12], [57, 9]), (plateau_map(80, 3), [20, 60], [70, 71])):
        F, Tt = O.computeTmap(c, g), O.computeTmap(c, g, s)
        rank = emu._ranks(F)
        out, ovf = emu.truncate(F, c, rank[s[1] * c.shape[1] + s[0]], rank)
        assert ovf == 0 and np.array_equal(out, Tt)
    c = rand_map((100, 100), 0)
    TG, TS, j = O.biComputeTmap(c, [10, 10], [90, 90])
    FG, FS = O.computeTmap(c, [10, 10]), O.computeTmap(c, [90, 90])
    rG, rS = emu._ranks(FG), emu._ranks(FS)
    m = np.maximum(rG, rS)
    k = int(m.min())
    jj = int(np.argmin(m))
    assert [jj % 100, jj // 100] == list(j)
    assert np.array_equal(emu.truncate(FG, c, k, rG)[0], TG) and np.array_equal(emu.truncate(FS, c, k, rS)[0], TS)
    c3 = rand_map((24, 24, 24), 0)
    F3, T3 = O.computeTmap3D(c3, [5, 6, 7]), O.computeTmap3D(c3, [5, 6, 7], [18, 17, 16])
    r3 = emu._ranks(F3)
    out, ovf = emu.truncate(F3, c3, r3[(17 * 24 + 18) * 24 + 16], r3)
    # 3D: the reference squares NumPy scalars with libm pow, which differs from the exactly rounded
    # product in about 1 of 1200 cases, so tentative values agree to rounding, not always bit for bit
    fin = np.isfinite(T3)
    assert ovf == 0 and np.array_equal(np.isfinite(out), fin)
    assert np.max(np.abs(out[fin] - T3[fin]) / np.maximum(T3[fin], 1.0)) < 1e-12 and np.mean(out[fin] == T3[fin]) > 0.999


def test_div3_is_correctly_rounded():
    """num<double>::div3 (FMA-corrected multiply used by the 3D update) == IEEE x / 3."""
    import ctypes as C
    L = emu.lib()
    L.emu_div3.argtypes = [emu.dp, emu.dp, C.c_longlong]
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(2_000_000) * 10.0 ** rng.integers(-8, 9, 2_000_000),
                        -rng.random(100_000) * 1e4, np.arange(0, 4097, dtype=np.float64),
                        np.array([0.0, -0.0, np.inf, -np.inf, 1e308, 1e-308, 5e-324, 3.0, 1.0, 2.0 / 3.0])])
    # neighbours of multiples of 3 (the hard rounding cases)
    m = np.arange(1, 50_000, dtype=np.float64) * 3.0
    x = np.concatenate([x, np.nextafter(m, np.inf), np.nextafter(m, 0.0)])
    out = np.empty_like(x)
    L.emu_div3(np.ascontiguousarray(x).ctypes.data_as(emu.dp), out.ctypes.data_as(emu.dp), x.size)
    assert np.array_equal(out, x / 3.0)


def test_edge_sizes_and_maps_without_inf_border():
    """Ragged / minimal shapes and maps with no inf border (out-of-domain is treated as +inf)."""
    rng = np.random.default_rng(0)
    for shape, g in (((1, 1), [0, 0]), ((1, 7), [3, 0]), ((5, 1), [0, 2]), ((2, 2), [1, 1]), ((32, 32), [31, 31]),
                     ((33, 33), [32, 0]), ((64, 32), [0, 63])):
        c = 1 + 4 * rng.random(shape)
        T, _ = emu.solve2d(c, [g])
        assert rel_err(T[0], O.computeTmap(c, g)) < TOL64
    for shape, g in (((1, 1, 1), [0, 0, 0]), ((2, 3, 4), [1, 0, 3]), ((4, 8, 16), [7, 3, 15]), ((5, 9, 17), [8, 4, 16])):
        c = 1 + 4 * rng.random(shape)
        T, _ = emu.solve3d(c, [g])
        assert rel_err(T[0], O.computeTmap3D(c, g)) < TOL64


def test_tracers_fuzz_against_oracle_including_failure_modes():
    """Seeded fuzz: partial fields with walls, integer / fractional / arbitrary start points.  Paths
    (atol 1e-9) and status codes (early return, IndexError, ValueError, OverflowError) must match."""
    rng = np.random.default_rng(123)
    for _ in range(30):
        n = int(rng.integers(12, 36))
        c = rand_map((n, n + 3), int(rng.integers(0, 1000)))
        for _w in range(int(rng.integers(0, 4))):
            y, x = int(rng.integers(1, n - 1)), int(rng.integers(1, n))
            c[y, x:x + int(rng.integers(1, 8))] = np.inf
        free = np.argwhere(np.isfinite(c))
        gy, gx = free[int(rng.integers(0, len(free)))]
        sy, sx = free[int(rng.integers(0, len(free)))]
        T = O.computeTmap(c, [int(gx), int(gy)], [int(sx), int(sy)] if rng.random() < .6 else None)
        inits = [[float(sx), float(sy)], [sx + rng.random() * .9, sy + rng.random() * .9],
                 [float(rng.random() * (n + 2)), float(rng.random() * (n - 1))]]
        end = [float(gx), float(gy)]
        p, s = emu.trace2d(T, inits, [end] * 3, field_of_path=[0, 0, 0])
        for k, init in enumerate(inits):
            po, so = O.getPathGDM(T, np.array(init), end, 0.5, return_status=True)
            assert s[k] == so and p[k].shape == po.shape
            assert len(po) == 0 or np.allclose(p[k], po, rtol=0, atol=1e-9, equal_nan=True)
    for _ in range(12):
        shp = tuple(int(v) for v in rng.integers(8, 15, size=3))
        c = rand_map(shp, int(rng.integers(0, 1000)))
        if rng.random() < .5:
            c[shp[0] // 2, 2:shp[1] - 2, 2:shp[2] - 3] = np.inf
        free = np.argwhere(np.isfinite(c))
        g, s0 = free[int(rng.integers(0, len(free)))], free[int(rng.integers(0, len(free)))]
        goal, start = [int(g[1]), int(g[0]), int(g[2])], [int(s0[1]), int(s0[0]), int(s0[2])]
        T = O.computeTmap3D(c, goal, start if rng.random() < .6 else None)
        inits = [[float(v) for v in start], [start[0] + rng.random() * .9, start[1] + rng.random() * .9, start[2] + rng.random() * .9]]
        p, s = emu.trace3d(T, inits, [[float(v) for v in goal]] * 2, field_of_path=[0, 0])
        for k, init in enumerate(inits):
            po, so = O.getPathGDM3D(T, np.array(init), np.array(goal, dtype=float), 0.5, return_status=True)
            assert s[k] == so and p[k].shape == po.shape
            assert len(po) == 0 or np.allclose(p[k], po, rtol=0, atol=1e-9, equal_nan=True)


def test_tie_order_sweep_matches_reference_pop_order():
    """csrc/tiekeys.cuh tie_sweep2d_kernel (one ordered sweep, no iteration) under the emulator against
    the reference's true pop order on tie-heavy maps and on the planner's own cost map."""
    uniform = np.pad(np.ones((60, 60)), 1, constant_values=np.inf)
    g = np.load(f"{GOLDEN}/planner_calls.npz")
    pc = np.ascontiguousarray(g["bi_cost"])
    for c, s, max_bad in ((uniform, [30, 30], 0), (uniform, [12, 40], 0), (plateau_map(80, 1), [8, 8], 0), (plateau_map(80, 2), [8, 8], 0),
                          (plateau_map(120, 5), [60, 8], 0), (rand_map((60, 60), 2), [9, 40], 0),
                          (pc, [int(v) for v in g["bi_goal"]], 0),
                          # one mirror pair tied to the last bit whose order hinges on an ulp-level coincidence of a
                          # NON-final neighbour value: the plain sort misplaces 100+ cells here
                          (pc, [int(v) for v in g["bi_start"]], 2)):
        T, order, _ = O.computeTmap(c, s, return_stats=True)
        r = emu.tie_order2d(T, c, s)
        mine = np.argsort(r.ravel(), kind="stable")[1:1 + len(order)]
        assert int((mine != order).sum()) <= max_bad


@pytest.mark.parametrize("window", ["1", "4"])
def test_windowed_order_under_the_emulator(window, monkeypatch):
    """solve2d_kernel with the windowed FIFO order (per-level counters, deferral) == oracle."""
    monkeypatch.setenv("FMB_WINDOWED", "1")
    monkeypatch.setenv("FMB_WINDOW", window)
    for c, g in ((rand_map((130, 170), 5), [160, 5]), (plateau_map(160, 2), [8, 8])):
        T, st = emu.solve2d(c, [g], nblocks=3)
        ref = O.computeTmap(c, g)
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(T[0]), fin)
        assert np.max(np.abs(T[0][fin] - ref[fin]) / np.maximum(ref[fin], 1e-300)) < 1e-12


def test_tie_order_sweep_3d_matches_reference_pop_order():
    """tie_sweep_kernel<3> under the emulator: uniform-cost volumes (where the plain sort misplaces
    most of the cells), the planner-like arm volume and a tie-free random volume."""
    from planning_motion_planning_b200 import synth

    def uniform_vol(n, val):
        c = np.full((n, n, n), val)
        c[0] = c[-1] = np.inf
        c[:, 0] = c[:, -1] = np.inf
        c[:, :, 0] = c[:, :, -1] = np.inf
        return c
    c3, g3, _ = synth.arm_volume((48, 48, 28), 1)
    for c, g in ((uniform_vol(24, 20.0), [5, 6, 7]), (uniform_vol(22, 1.0), [11, 11, 11]), (c3, list(g3)), (rand_map((16, 16, 16), 1), [4, 5, 6])):
        F, order, _ = O.computeTmap3D(c, g, [-1, -1, -1], return_stats=True)
        r = emu.tie_order3d(F, c, g)
        mine = np.argsort(r.ravel(), kind="stable")[1:1 + len(order)]
        assert int((mine != order).sum()) == 0


def test_tie_order_of_a_transposed_field_follows_the_callers_orientation():
    """An F-ordered input is solved as its C-ordered transpose; the child order of updateNode
    (FastMarching.py:46-54) is not symmetric in x and y, so the tie order must be mapped back."""
    import torch
    import ranks_ref as _compat
    uniform = np.pad(np.full((60, 60), 3.0), 1, constant_values=np.inf)
    uniform[30, 8:25] = np.inf
    for c, s in ((uniform, [18, 40]), (plateau_map(96, 7), [70, 66])):
        T, order, _ = O.computeTmap(c, s, return_stats=True)
        ct, Tt, st = np.ascontiguousarray(c.T), np.ascontiguousarray(T.T), s[::-1]
        H, W = c.shape
        for ranks in (emu.tie_order2d(Tt, ct, st, transposed=True),
                      _compat.pop_ranks_lifo2d(torch.from_numpy(Tt), torch.from_numpy(ct), st, transposed=True).numpy()):
            mt = np.argsort(ranks.ravel(), kind="stable")[1:1 + len(order)]
            assert int((((mt % H) * W + mt // H) != order).sum()) == 0
        wrong = np.argsort(emu.tie_order2d(Tt, ct, st).ravel(), kind="stable")[1:1 + len(order)]
        assert int((((wrong % H) * W + wrong // H) != order).sum()) > 100          # the flag matters


def test_emulated_field_carries_the_references_bits():
    """With the 'a few ulp higher also replaces' rule the emulated kernel's field is an exact fixed point of
    the update: bit-identical to the reference on uniform, plateau and random maps."""
    uniform = np.pad(np.full((110, 110), 7.0), 1, constant_values=np.inf)
    uniform[40, 20:60] = np.inf
    for c, g in ((uniform, [17, 18]), (plateau_map(128, 7), [96, 90]), (rand_map((150, 170), 3), [20, 30])):
        T, _ = emu.solve2d(c, [g], nblocks=3)
        ref = O.computeTmap(c, g)
        assert np.array_equal(T[0], ref)


def test_device_pow2_reproduces_libm_pow_bit_for_bit():
    """csrc/pow2_glibc.cuh (the port of glibc's pow for y = 2 that the exact 3D update uses) == libm pow(x, 2.0), the
    value of `x**2` on a NumPy scalar (FastMarching3D.py:68-71), on 1e7 inputs: wide exponents, cost-like values,
    integers and halves, sevenths, [1, 2), T-like sums of square roots, the neighbourhood of 1, subnormals / 0 / inf."""
    rng = np.random.default_rng(7)
    n = 1_500_000
    xs = [np.ldexp(1.0 + rng.random(n), rng.integers(-200, 200, n)),
          1.0 + 304.0 * rng.random(n),
          rng.integers(0, 100000, n) * 0.5,
          rng.integers(0, 100000, n) / 7.0,
          1.0 + rng.random(n),
          20.0 * np.sqrt(rng.integers(1, 4000, n).astype(np.float64)) + rng.integers(0, 100, n),
          np.abs(np.float64(1.0) + (rng.integers(-1000, 1000, n) * np.finfo(np.float64).eps)),
          np.array([0.0, np.inf, 5e-324, 1e-310, 2.2250738585072014e-308, 1.0, 2.0, 0.5, 1e-100, 1e100])]
    x = np.ascontiguousarray(np.concatenate(xs))
    assert x.size >= 10_000_000
    L = emu.lib()
    import ctypes as C
    L.emu_pow2.argtypes = [emu.dp, emu.dp, C.c_longlong]
    out = np.empty_like(x)
    L.emu_pow2(x.ctypes.data_as(emu.dp), out.ctypes.data_as(emu.dp), x.size)
    ref = O.pow2(x)
    assert np.array_equal(out.view(np.uint64), ref.view(np.uint64))
    assert int((ref != x * x).sum()) > 1000          # the inputs do exercise the cases where pow(x, 2) != x*x
    # and the reference really evaluates numpy-scalar squares through that function
    for v in x[::997_003]:
        assert np.float64(v) ** 2 == ref[np.where(x == v)[0][0]] or not np.isfinite(v)


# ------------------------------------------------------------------ 3D sweep engine (csrc/eikonal3d_sweep.cuh)
@pytest.mark.parametrize("shape,goal", [((9, 9, 9), [4, 4, 4]), ((20, 20, 20), [4, 4, 4]), ((13, 21, 40), [10, 5, 33]),
                                        ((10, 12, 28), [8, 7, 6])])
def test_solve3d_sweep_engine_random(shape, goal):
    """tz=0 selects the eight-warp sweep visits with the local causal order (the library's default 3D engine)."""
    c = rand_map(shape, 5)
    if shape == (20, 20, 20):
        c[8:10, 3:15, 3:15] = np.inf
    T, st = emu.solve3d(c, [goal], tz=0, nblocks=3)
    assert rel_err(T[0], O.computeTmap3D(c, goal)) < TOL64
    assert st["visits"] > 0


def test_solve3d_sweep_engine_uniform_batch_and_fp32():
    c = np.full((16, 20, 36), 20.0)
    for d in range(3):
        for e in (0, -1):
            sl = [slice(None)] * 3
            sl[d] = e
            c[tuple(sl)] = np.inf
    goals = [[5, 5, 5], [10, 8, 30]]
    T, _ = emu.solve3d(c, goals, tz=0, nblocks=2)
    for q, g in enumerate(goals):
        assert rel_err(T[q], O.computeTmap3D(c, g)) < TOL64
    T32, _ = emu.solve3d(c.astype(np.float32), [goals[0]], tz=0)
    assert rel_err(T32[0].astype(np.float64), O.computeTmap3D(c, goals[0])) < TOL32


def test_update3d_branch_free_form_is_bitwise_the_branching_form():
    """solve3d_update_sel<EXACT> (one quadratic chosen by selects) == solve3d_update / solve3d_update_exact bit for bit,
    incl. inf patterns, ties and dropped dimensions; the exact form == the oracle's libm-pow arithmetic."""
    import ctypes as C
    L = emu.lib()
    rng = np.random.default_rng(0)
    n = 200000
    t = np.empty((n, 4))
    base = rng.random(n) * 1000
    for k in range(3):
        t[:, k] = base + rng.random(n) * 8
    t[:, 3] = np.where(rng.random(n) < 0.5, 20.0, 1 + rng.random(n) * 30)
    m = rng.random((n, 3)) < 0.15
    t[:, :3][m] = np.inf
    tie = rng.random(n) < 0.1; t[tie, 1] = t[tie, 0]
    tie = rng.random(n) < 0.05; t[tie, 2] = t[tie, 1]
    far = rng.random(n) < 0.1; t[far, 0] += 100
    t = np.ascontiguousarray(t)
    for exact in (0, 1):
        a, b, sl = np.empty(n), np.empty(n), np.zeros(n, dtype=np.int32)
        L.emu_update3d(t.ctypes.data_as(emu.dp), C.c_longlong(n), exact, a.ctypes.data_as(emu.dp), b.ctypes.data_as(emu.dp),
                       sl.ctypes.data_as(emu.ip))
        assert np.array_equal(a.view(np.uint64)[sl == 0], b.view(np.uint64)[sl == 0])
        assert sl.sum() == 0
        if exact == 1:
            some = ~np.all(np.isinf(t[:5000, :3]), axis=1)           # all-inf: the reference raises, the device returns inf
            ref = np.array([O.solve3d(*row) for row in t[:5000]])
            assert np.array_equal(ref.view(np.uint64)[some], a[:5000].view(np.uint64)[some])


# ------------------------------------------------------------------ 2D sweep engines (csrc/eikonal2d_sweep.cuh, eikonal2d_wsweep.cuh)
@pytest.mark.parametrize("variant,ring2", [("0", "1.0"), ("1", "1.0"), ("0", "0"), ("1", "3.0")])
def test_solve2d_sweep_engine_causal_order(variant, ring2, monkeypatch):
    """The library's default 2D engine (four sweep warps per tile visit, local causal order with the second-ring wait
    rule) under the emulator, both forms of the sweep step: == oracle on random, obstacle and plateau maps."""
    monkeypatch.setenv("FMB_EMU_VARIANT", variant)
    monkeypatch.setenv("FMB_EMU_RING2", ring2)
    cases = [(rand_map((100, 100), 0), [25, 25]), (rand_map((70, 133), 2), [120, 5]), (plateau_map(128, 3), [100, 90])]
    c = rand_map((130, 130), 1)
    c[40:44, 10:100] = np.inf
    c[80:84, 30:129] = np.inf
    cases.append((c, [64, 64]))
    for c, g in cases:
        T, st = emu.solve2d_cta(c, [g], R=0, nblocks=3, best_first=0, windowed=2)
        assert rel_err(T[0], O.computeTmap(c, g)) < TOL64
        assert st["visits"] > 0
    T, _ = emu.solve2d_cta(cases[2][0], [[10, 10], [100, 90], [64, 20]], R=0, nblocks=2, best_first=1)
    for q, g in enumerate([[10, 10], [100, 90], [64, 20]]):
        assert rel_err(T[q], O.computeTmap(cases[2][0], g)) < TOL64


@pytest.mark.parametrize("mode", ["1", "2"])
def test_solve2d_warp_sweep_engine_for_batches(mode, monkeypatch):
    """Warp-per-tile sweep visits (engine2d = 4 / 5: costs staged in shared memory / read from global memory)."""
    monkeypatch.setenv("FMB_EMU_WSWEEP", mode)
    c = rand_map((100, 100), 0)
    for bf in (0, 1):
        T, _ = emu.solve2d_cta(c, [[25, 25], [70, 40]], R=0, nblocks=3, best_first=bf, windowed=0)
        for q, g in enumerate([[25, 25], [70, 40]]):
            assert rel_err(T[q], O.computeTmap(c, g)) < TOL64
    c = plateau_map(96, 3)
    c[30:34, 10:70] = np.inf
    T, _ = emu.solve2d_cta(c, [[10, 10]], R=0, nblocks=2, best_first=0, windowed=0)
    assert rel_err(T[0], O.computeTmap(c, [10, 10])) < TOL64
    T32, _ = emu.solve2d_cta(c.astype(np.float32), [[10, 10]], R=0, nblocks=2, best_first=0)
    assert rel_err(T32[0].astype(np.float64), O.computeTmap(c.astype(np.float32).astype(np.float64), [10, 10])) < TOL32


def test_solve3d_sweep_engine_octant_rule(monkeypatch):
    """variant bit 3: every cell is evaluated by the one sweep whose upwind side carries its lower neighbours, sweeps without
    such a cell skip the round (12 instead of 36 evaluations per cell on the bench volume): same field."""
    monkeypatch.setenv("FMB_EMU_VARIANT", "9")
    c = rand_map((20, 20, 20), 5)
    c[8:10, 3:15, 3:15] = np.inf
    T, st = emu.solve3d(c, [[4, 4, 4]], tz=0, nblocks=3)
    assert rel_err(T[0], O.computeTmap3D(c, [4, 4, 4])) < TOL64


def test_warp_engine_with_costs_from_global_memory(monkeypatch):
    """engine2d = 6 (the default of best-first batches): the armed-cell visit without the shared-memory cost tile -- same
    fields as the oracle on interior, edge and ragged tiles, fp64 and fp32."""
    monkeypatch.setenv("FMB_BEST_FIRST", "1")
    monkeypatch.setenv("FMB_COST_GLOBAL", "1")
    rng = np.random.default_rng(11)
    for shape, goals in (((100, 130), [[25, 25], [120, 90], [64, 3]]), ((97, 65), [[1, 1], [60, 90]]), ((33, 31), [[15, 16]])):
        c = 1.0 + 4.0 * rng.random(shape)
        c[rng.random(shape) < 0.08] = np.inf
        for g in goals:
            c[g[1], g[0]] = 1.0
        T, st = emu.solve2d(c, goals, nblocks=3)
        for q, g in enumerate(goals):
            assert rel_err(T[q], O.computeTmap(c, g)) < 1e-12
        T32, _ = emu.solve2d(c.astype(np.float32), goals[:1], nblocks=2)
        assert rel_err(T32[0].astype(np.float64), O.computeTmap(c.astype(np.float32).astype(np.float64), goals[0])) < 1e-4
