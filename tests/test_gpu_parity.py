"""GPU (-m gpu): the compiled sm_100a path, called through the C ABI (ctypes) via the torch
plumbing, against the CPU oracle and the golden vectors frozen from the reference.

Tolerances are the north-star's: fp64 field 1e-9 relative (fp32 1e-4), waypoints 1e-3 cell."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, plateau_map, rand_map, rel_err

pytestmark = pytest.mark.gpu

TOL64, TOL32, TOLP = 1e-9, 1e-4, 1e-3


@pytest.fixture(scope="module")
def eng():
    import torch
    from planning_motion_planning_b200 import engine
    assert torch.cuda.is_available()
    return engine


def _gpu2d(eng, c, seeds, **kw):
    import torch
    T = eng.solve2d(torch.from_numpy(np.ascontiguousarray(c)).cuda(), seeds, **kw)
    return T.cpu().numpy()


def _gpu3d(eng, c, seeds, **kw):
    import torch
    T = eng.solve3d(torch.from_numpy(np.ascontiguousarray(c)).cuda(), seeds, **kw)
    return T.cpu().numpy()


# ------------------------------------------------------------------ 2D solve
@pytest.mark.parametrize("engine", ["sweep", "cta1", "cta2", "cta4", "warp32", "warp16"])
@pytest.mark.parametrize("shape,goal", [((9, 9), [4, 4]), ((100, 100), [25, 25]), ((257, 300), [290, 3]),
                                        ((33, 65), [32, 31]), ((512, 512), [100, 400])])
def test_solve2d_random_vs_oracle(eng, shape, goal, engine, fmb_opts):
    from conftest import ENGINES_2D
    from oracle import oracle as O
    fmb_opts(**ENGINES_2D[engine])
    c = rand_map(shape, 0)
    T = _gpu2d(eng, c, [goal])[0]
    assert rel_err(T, O.computeTmap(c, goal)) < TOL64
    st = eng.last_stats()
    assert st["tile_visits"] > 0 and st["evals"] >= np.isfinite(T).sum() - 1


def test_branch_free_sqrt_is_correctly_rounded(eng):
    """sqrt_rn_fast (csrc/fm_common.cuh) == sqrt.rn.f64 bit for bit: 3e7 random mantissas over the whole exponent
    range it accepts, values next to perfect squares and powers of two, and the discriminants 2c^2 - d^2 of the
    update for planner-like costs."""
    import torch
    from planning_motion_planning_b200 import _capi
    g = torch.Generator(device="cuda").manual_seed(1)
    n = 10_000_000
    mant = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) + 1.0
    expo = torch.randint(-960, 1020, (n,), device="cuda", generator=g).to(torch.float64)
    xs = [mant * torch.exp2(expo)]
    k = torch.arange(1, n + 1, dtype=torch.float64, device="cuda")
    sq = k * k
    xs.append(torch.cat([sq, torch.nextafter(sq, sq * 2), torch.nextafter(sq, sq * 0)])[:n])
    c = 1.0 + 304.0 * torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    d = c * torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    xs.append(2.0 * (c * c) - d * d)
    p2 = torch.exp2(torch.arange(-960, 1020, dtype=torch.float64, device="cuda"))
    xs.append(torch.cat([p2, torch.nextafter(p2, p2 * 2), torch.nextafter(p2, p2 * 0)]))
    for x in xs:
        bad = torch.zeros(1, dtype=torch.int64, device="cuda")
        _capi.check(_capi.lib().fmb_debug_sqrt_check(x.data_ptr(), x.numel(), bad.data_ptr(), None))
        torch.cuda.synchronize()
        assert int(bad.item()) == 0


def test_kat1_and_kat3_values(eng):
    c = np.pad(np.ones((7, 7)), 1, constant_values=np.inf)
    T = _gpu2d(eng, c, [[4, 4]])[0]
    assert list(T[4, 1:8]) == [3, 2, 1, 0, 1, 2, 3]
    assert abs(T[5, 5] - 1.7071067811865475) < 1e-15 and abs(T[6, 6] - 3.25243570661267) < 1e-14
    c = rand_map((100, 100), 0)
    T = _gpu2d(eng, c, [[25, 25]])[0]
    assert abs(np.sum(T[np.isfinite(T)]) - 1224641.4437929934) / 1224641.4437929934 < 1e-12
    assert abs(T[98, 98] - 280.68674463626854) < 1e-9 * 280.7


def test_solve2d_plateau_walls_enclosed(eng):
    from oracle import oracle as O
    c = plateau_map(160, 3)
    c[30:90, 40] = np.inf
    c[100:120, 100] = c[100:120, 119] = c[100, 100:120] = c[119, 100:120] = np.inf
    T = _gpu2d(eng, c, [[20, 60]])[0]
    assert rel_err(T, O.computeTmap(c, [20, 60])) < TOL64
    assert np.all(np.isinf(T[101:119, 101:119]))


def test_solve2d_planner_like_map(eng):
    from oracle import oracle as O
    from planning_motion_planning_b200 import synth
    c = synth.mars_costmap(768, 1)
    g = synth.free_cell_near(c, 200, 190)
    T = _gpu2d(eng, c, [g])[0]
    assert rel_err(T, O.computeTmap(c, g)) < TOL64


def test_solve2d_batches(eng):
    import torch
    from oracle import oracle as O
    c = rand_map((96, 130), 2)
    goals = [[5, 5], [120, 90], [31, 32], [32, 31], [64, 64], [63, 63]]
    T = _gpu2d(eng, c, goals)
    for q, g in enumerate(goals):
        assert rel_err(T[q], O.computeTmap(c, g)) < TOL64
    cs = np.stack([rand_map((80, 80), s) for s in range(5)])
    T = eng.solve2d(torch.from_numpy(cs).cuda(), [[40, 40]] * 5).cpu().numpy()
    for q in range(5):
        assert rel_err(T[q], O.computeTmap(cs[q], [40, 40])) < TOL64


def test_solve2d_fp32(eng):
    from oracle import oracle as O
    c32 = rand_map((300, 300), 4).astype(np.float32)
    T = _gpu2d(eng, c32, [[10, 250]])[0]
    assert T.dtype == np.float32
    assert rel_err(T.astype(np.float64), O.computeTmap(c32.astype(np.float64), [10, 250])) < TOL32


def test_solve2d_is_exact_fixed_point_4096(eng):
    """Size-independent property (KAT-5) at BASELINE's full size: the returned field is a
    bitwise fixed point of the reference update (FastMarching.py:17-29), evaluated here
    with plain torch fp64 ops; and it matches the C oracle within 1e-9."""
    import torch
    from oracle import oracle as O
    from planning_motion_planning_b200 import synth
    n = 4096
    c = synth.mars_costmap(n, 0)
    g = synth.free_cell_near(c, n // 4, n // 4)
    cd = torch.from_numpy(c).cuda()
    T = eng.solve2d(cd, [g])[0]
    inf = float("inf")
    P = torch.nn.functional.pad(T, (1, 1, 1, 1), value=inf)
    a = torch.minimum(P[1:-1, :-2], P[1:-1, 2:])
    b = torch.minimum(P[:-2, 1:-1], P[2:, 1:-1])
    d = a - b
    one_sided = ~(d.abs() <= cd)
    m = torch.minimum(a, b)
    two = 0.5 * ((a + b) + torch.sqrt(2 * (cd * cd) - d * d))
    U = torch.where(one_sided, m + cd, two)
    U = torch.where(torch.isinf(cd), torch.full_like(U, inf), U)
    free = torch.isfinite(cd)
    free[g[1], g[0]] = False
    both_inf = torch.isinf(T) & torch.isinf(U)
    # no free cell can be improved by one more relaxation (T <= U), and T equals U up to the
    # last bit (the min-iteration keeps an older value when rounding makes the update non-monotone)
    gap = torch.where(both_inf, torch.zeros_like(T), U - T)
    assert float(gap[free].min()) >= 0.0
    assert float((gap / T.clamp_min(1.0))[free].max()) < 1e-14
    assert float(T[g[1], g[0]]) == 0.0
    ref = O.computeTmap(c, g)
    assert rel_err(T.cpu().numpy(), ref) < TOL64


@pytest.mark.parametrize("seed", range(6))
def test_solve2d_differential_random_obstacles(eng, seed):
    """Seeded differential test: random size, random walls/blocks (inf), mixed cost scales."""
    from oracle import oracle as O
    rng = np.random.default_rng(1000 + seed)
    rows, cols = int(rng.integers(40, 300)), int(rng.integers(40, 300))
    c = rand_map((rows, cols), seed)
    c[np.isfinite(c)] *= rng.choice([1.0, 10.0, 300.0], size=int(np.isfinite(c).sum()), p=[.8, .15, .05])
    for _ in range(int(rng.integers(3, 12))):
        y, x = int(rng.integers(1, rows - 1)), int(rng.integers(1, cols - 1))
        if rng.random() < .5:
            c[y, x:min(cols - 1, x + int(rng.integers(5, 80)))] = np.inf
        else:
            c[y:min(rows - 1, y + int(rng.integers(5, 80))), x] = np.inf
    free = np.argwhere(np.isfinite(c))
    gy, gx = free[int(rng.integers(0, len(free)))]
    T = _gpu2d(eng, c, [[int(gx), int(gy)]])[0]
    assert rel_err(T, O.computeTmap(c, [int(gx), int(gy)])) < TOL64


@pytest.mark.parametrize("engine", ["sweep", "cta2", "warp32", "warp32g"])
def test_best_first_and_fifo_agree(eng, engine, fmb_opts):
    """The two work orders of the persistent kernel reach the same fixed point."""
    import torch
    from oracle import oracle as O
    c = rand_map((160, 200), 9)
    goals = [[int(5 + 7 * i) % 190 + 2, int(3 + 11 * i) % 150 + 2] for i in range(16)]
    outs = []
    from conftest import ENGINES_2D
    for bf in (0, 1):
        fmb_opts(best_first=bf, **ENGINES_2D[engine])
        outs.append(_gpu2d(eng, c, goals))
    for q, g in enumerate(goals):
        ref = O.computeTmap(c, g)
        assert rel_err(outs[0][q], ref) < TOL64 and rel_err(outs[1][q], ref) < TOL64


@pytest.mark.parametrize("engine", ["sweep", "cta2", "warp32"])
@pytest.mark.parametrize("window", [1, 2, 8])
def test_windowed_order_reaches_the_same_field(eng, window, engine, fmb_opts):
    """Windowed FIFO (deferral of tiles far above the lowest queued level; the default for one map of
    >= 16384 tiles) forced on for smaller maps: same fixed point as the oracle, whatever the window."""
    from oracle import oracle as O
    from planning_motion_planning_b200 import synth
    from conftest import ENGINES_2D
    fmb_opts(windowed=1, window=window, **ENGINES_2D[engine])
    for c, g in ((rand_map((257, 300), 3), [290, 3]), (plateau_map(400, 2), [8, 8]), (synth.mars_costmap(768, 4), None)):
        if g is None:
            g = list(synth.free_cell_near(c, 100, 650))
        T = _gpu2d(eng, c, [g])[0]
        assert rel_err(T, O.computeTmap(c, g)) < TOL64
    assert eng.last_stats()["deferrals"] >= 0


@pytest.mark.parametrize("share", [0, 3])
def test_concurrent_solves_on_separate_streams(eng, share, fmb_opts):
    """Three independent queries in flight on their own streams (what bench.py does): each lane has its
    own workspace, the results equal the one-at-a-time results bit for bit in their inf pattern and to
    rounding in value (iteration order differs run to run) -- with the default grid per solve and with
    fmb_options.concurrent_solves = 3 (a third of the resident CTA slots each)."""
    import torch
    from planning_motion_planning_b200 import synth
    c = synth.mars_costmap(2048, 5)
    cd = torch.from_numpy(c).cuda()
    goals = [list(synth.free_cell_near(c, 300, 300)), list(synth.free_cell_near(c, 1700, 400)), list(synth.free_cell_near(c, 1000, 1800))]
    serial = [eng.solve2d(cd, [g]).clone() for g in goals]
    fmb_opts(concurrent_solves=share)
    streams = [torch.cuda.Stream() for _ in goals]
    outs = [torch.empty((1, 2048, 2048), dtype=torch.float64, device="cuda") for _ in goals]
    torch.cuda.synchronize()
    for rep in range(3):
        for s, g, o in zip(streams, goals, outs):
            with torch.cuda.stream(s):
                eng.solve2d(cd, [g], out=o, sync=False)
        for s in streams:
            with torch.cuda.stream(s):
                eng.finish()
        for a, b in zip(serial, outs):
            fin = torch.isfinite(a)
            assert torch.equal(fin, torch.isfinite(b))
            assert float(((a - b).abs() / a.clamp_min(1e-300))[fin].max()) < 1e-12


def test_solve2d_field_is_an_exact_fixed_point_like_the_references(eng):
    """The relaxation accepts a value a few ulp ABOVE the stored one (not only lower ones), so the field
    ends as an exact fixed point of the update -- like the reference's -- instead of the minimum over a
    history of roundings: on these maps every cell carries the reference's bits, which is what keeps
    exact ties between mirror-image cells (and the early-exit patterns that depend on them) intact."""
    from oracle import oracle as O
    uniform = np.pad(np.full((110, 110), 7.0), 1, constant_values=np.inf)
    uniform[40, 20:60] = np.inf
    for c, g, min_equal in ((uniform, [17, 18], 1.0), (plateau_map(128, 7), [96, 90], 1.0), (rand_map((150, 170), 3), [20, 30], 0.999)):
        T = _gpu2d(eng, c, [g])[0]
        ref = O.computeTmap(c, g)
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(T), fin)
        assert np.mean(T[fin] == ref[fin]) >= min_equal
        assert rel_err(T, ref) < 1e-14


def test_batch_api_single_rank(eng):
    from oracle import oracle as O
    from planning_motion_planning_b200 import batch, synth
    c = synth.mars_costmap(256, 5)
    rng = np.random.default_rng(3)
    ok = np.argwhere(np.isfinite(c) & (c <= 2.0))
    goals = ok[rng.integers(0, len(ok), size=20)][:, ::-1].tolist()
    starts = ok[rng.integers(0, len(ok), size=20)][:, ::-1].tolist()
    lo, res = batch.solve_queries(c, goals, starts, chunk=8)
    assert lo == 0 and len(res) == 20
    for (p, st), g, s in zip(res, goals, starts):
        po, so = O.getPathGDM(O.computeTmap(c, g), np.array(s, dtype=np.float64), g, 0.5, return_status=True)
        assert st == so and p.shape == po.shape and (len(po) == 0 or np.abs(p - po).max() < TOLP)


def test_solve2d_8192_fixed_point_property(eng):
    """Config 5 size on one GPU (1 GiB of fields): no oracle run (tens of seconds on the host),
    only the size-independent property that no free cell can be improved by one more update."""
    import torch
    from planning_motion_planning_b200 import synth
    n = 8192
    c = synth.random_costmap((n, n), 11)
    cd = torch.from_numpy(c).cuda()
    del c
    T = eng.solve2d(cd, [[n // 3, n // 5]])[0]
    inf = float("inf")
    P = torch.nn.functional.pad(T, (1, 1, 1, 1), value=inf)
    a = torch.minimum(P[1:-1, :-2], P[1:-1, 2:])
    b = torch.minimum(P[:-2, 1:-1], P[2:, 1:-1])
    del P
    d = a - b
    U = torch.where(~(d.abs() <= cd), torch.minimum(a, b) + cd, 0.5 * ((a + b) + torch.sqrt(2 * (cd * cd) - d * d)))
    del a, b, d
    free = torch.isfinite(cd)
    free[n // 5, n // 3] = False
    assert bool(torch.isfinite(T[free]).all())
    gap = (U - T)[free]
    assert float(gap.min()) >= 0.0 and float((gap / T[free].clamp_min(1.0)).max()) < 1e-14


# ------------------------------------------------------------------ 3D solve
@pytest.mark.parametrize("tz", [32, 16])
@pytest.mark.parametrize("shape,goal", [((9, 9, 9), [4, 4, 4]), ((24, 24, 24), [5, 6, 7]), ((13, 21, 40), [10, 5, 33]),
                                        ((44, 44, 28), [35, 27, 6]), ((64, 64, 64), [10, 50, 30])])
def test_solve3d_random_vs_oracle(eng, shape, goal, tz, fmb_opts):
    from oracle import oracle as O
    fmb_opts(tile_z3d=tz)
    c = rand_map(shape, 0)
    T = _gpu3d(eng, c, [goal])[0]
    assert rel_err(T, O.computeTmap3D(c, goal)) < TOL64


def test_solve3d_planner_like_volume_and_batch(eng):
    from oracle import oracle as O
    from planning_motion_planning_b200 import synth
    c, goal, start = synth.arm_volume((90, 90, 28), 0)
    T = _gpu3d(eng, c, [goal, start])
    assert rel_err(T[0], O.computeTmap3D(c, goal)) < TOL64
    assert rel_err(T[1], O.computeTmap3D(c, start)) < TOL64


def test_solve3d_fp32(eng):
    from oracle import oracle as O
    c32 = rand_map((40, 40, 40), 6).astype(np.float32)
    T = _gpu3d(eng, c32, [[5, 5, 5]])[0]
    assert rel_err(T.astype(np.float64), O.computeTmap3D(c32.astype(np.float64), [5, 5, 5])) < TOL32


# ------------------------------------------------------------------ tracers
def test_trace2d_vs_oracle_and_golden(eng):
    import torch
    from oracle import oracle as O
    g = np.load(f"{GOLDEN}/ref2d.npz")
    c = rand_map((100, 100), 0)
    TG, TS, j = O.biComputeTmap(c, [10, 10], [90, 90])
    out, cnt, st = eng.trace2d(torch.from_numpy(np.stack([TG, TS])).cuda(), [j, j], [[10, 10], [90, 90]])
    out, cnt, st = out.cpu().numpy(), cnt.cpu().numpy(), st.cpu().numpy()
    assert list(st) == [0, 0]
    pG, pS = out[0, :cnt[0]], out[1, :cnt[1]]
    assert pG.shape == g["kat3_pathG"].shape and np.abs(pG - g["kat3_pathG"]).max() < TOLP
    assert pS.shape == g["kat3_pathS"].shape and np.abs(pS - g["kat3_pathS"]).max() < TOLP


def test_trace2d_step_cap_case(eng):
    """plateau80: the reference tracer never reaches `end` and runs all 30000 steps."""
    import torch
    from oracle import oracle as O
    g = np.load(f"{GOLDEN}/ref2d.npz")
    c = plateau_map(80, 3)
    T = O.computeTmap(c, [20, 60])
    out, cnt, st = eng.trace2d(torch.from_numpy(T).cuda(), [[70, 71]], [[20, 60]])
    n = int(cnt[0])
    ref = g["plateau80_pathF"]
    assert n == len(ref) and int(st[0]) == 0
    assert np.abs(out[0, :n].cpu().numpy() - ref).max() < TOLP


def test_trace2d_fallback_statuses(eng):
    import torch
    from oracle import oracle as O
    c = rand_map((30, 30), 2)
    c[15, 5:25] = np.inf
    T = O.computeTmap(c, [5, 5], [25, 25])
    inits = [[25, 25], [25, 24], [14, 16], [29.2, 29.2]]
    out, cnt, st = eng.trace2d(torch.from_numpy(T).cuda(), inits, [[5, 5]] * len(inits))
    for p, init in enumerate(inits):
        po, so = O.getPathGDM(T, np.array(init), [5, 5], 0.5, return_status=True)
        assert int(st[p]) == so and int(cnt[p]) == len(po)
        if len(po):
            assert np.abs(out[p, :len(po)].cpu().numpy() - po).max() < TOLP


def test_trace3d_vs_golden(eng):
    import torch
    from oracle import oracle as O
    g = np.load(f"{GOLDEN}/ref3d.npz")
    c = rand_map((24, 24, 24), 0)
    for tag, start in (("full", None), ("trunc", [18, 17, 16])):
        F = O.computeTmap3D(c, [5, 6, 7], start)
        out, cnt, st = eng.trace3d(torch.from_numpy(F).cuda(), [[18, 17, 16]], [[5, 6, 7]])
        ref = g[f"kat4_path_{tag}"]
        assert int(st[0]) == 0 and int(cnt[0]) == len(ref)
        assert np.abs(out[0, :len(ref)].cpu().numpy() - ref).max() < TOLP


# ------------------------------------------------------------------ drop-in API
def test_dropin_bi_and_paths_kat3b():
    import FastMarching.FastMarching as FM
    from oracle import oracle as O
    g = np.load(f"{GOLDEN}/ref2d.npz")
    for name, mk in (("kat3", lambda: rand_map((100, 100), 0)), ("plateau80", lambda: plateau_map(80, 3))):
        c = mk()
        gg, ss = list(g[f"{name}_g2"]), list(g[f"{name}_s2"])
        TG, TS, j = FM.biComputeTmap(c, gg, ss)
        assert j.dtype == np.uint32 and np.array_equal(j, g[f"{name}_join"])
        oTG, oTS, _ = O.biComputeTmap(c, gg, ss)
        assert rel_err(TG, oTG) < TOL64 and rel_err(TS, oTS) < TOL64
        pG = FM.getPathGDM(TG, j, gg, 0.5)
        assert pG.shape == g[f"{name}_pathG"].shape and np.abs(pG - g[f"{name}_pathG"]).max() < TOLP


def test_dropin_replays_planner_calls():
    """The five calls the unmodified planner makes (Coupled_motion_planner.py:1226-1230,
    1636-1639), with the planner's own arguments (incl. the F-ordered costmap view)."""
    import FastMarching.FastMarching as FM
    import FastMarching.FastMarching3D as FM3D
    from oracle import oracle as O
    g = np.load(f"{GOLDEN}/planner_calls.npz")
    cost = np.asfortranarray(g["bi_cost"])           # the planner passes cMap.T (F-ordered)
    goal, start = [int(v) for v in g["bi_goal"]], [int(v) for v in g["bi_start"]]
    TG, TS, j = FM.biComputeTmap(cost, goal, start)
    assert np.array_equal(j, g["bi_join"]) and j.dtype == np.uint32
    assert TG.flags.f_contiguous and TG.shape == cost.shape
    oTG, oTS, _ = O.biComputeTmap(g["bi_cost"], goal, start)
    assert rel_err(np.ascontiguousarray(TG), oTG) < TOL64 and rel_err(np.ascontiguousarray(TS), oTS) < TOL64
    pG = FM.getPathGDM(TG, j, goal, 0.5)
    pS = FM.getPathGDM(TS, j, start, 0.5)
    assert pG.shape == g["pathG"].shape and np.abs(pG - g["pathG"]).max() < TOLP
    assert pS.shape == g["pathS"].shape and np.abs(pS - g["pathS"]).max() < TOLP
    T3 = FM3D.computeTmap(g["c3"], np.uint32(g["g3"]), np.uint32(g["s3"]))
    assert rel_err(T3, O.computeTmap3D(g["c3"], list(g["g3"]), list(g["s3"]))) < TOL64
    p3 = FM3D.getPathGDM(T3, np.uint32(g["path3d_init"]), np.uint32(g["path3d_end"]), 0.5)
    assert p3.shape == g["path3d"].shape and np.abs(p3 - g["path3d"]).max() < TOLP


@pytest.mark.parametrize("replay", ["dense", "sparse"])
def test_dropin_early_exit_fields_match_reference_bitwise_pattern(replay, fmb_opts):
    fmb_opts(replay_sparse=1 if replay == "sparse" else 0)        # both forms of the replay (truncate.cuh)
    """Partial fields (accepted / narrow band / far) of the early-exit calls, SURVEY 8a a-5."""
    import FastMarching.FastMarching as FM
    import FastMarching.FastMarching3D as FM3D
    from oracle import oracle as O
    g3 = np.load(f"{GOLDEN}/ref3d.npz")
    c = rand_map((64, 64), 7)
    T = FM.computeTmap(c, [40, 12], [57, 9])
    ref = O.computeTmap(c, [40, 12], [57, 9])
    assert rel_err(T, ref) < TOL64 and np.isfinite(ref).sum() < np.isfinite(c).sum()
    c3 = rand_map((24, 24, 24), 0)
    T3 = FM3D.computeTmap(c3, np.uint32([5, 6, 7]), np.uint32([18, 17, 16]))
    r3 = O.computeTmap3D(c3, [5, 6, 7], [18, 17, 16])
    assert rel_err(T3, r3) < TOL64 and int(np.isfinite(T3).sum()) == int(g3["kat4_nfin_trunc"]) == 9830
    p = FM3D.getPathGDM(T3, np.uint32([18, 17, 16]), np.uint32([5, 6, 7]), 0.5)
    assert p.shape == g3["kat4_path_trunc"].shape and np.abs(p - g3["kat4_path_trunc"]).max() < TOLP


def test_domain_decomposition_local_slabs(eng):
    """Row-slab decomposition with halo exchange (one process holding all slabs) == single solve."""
    import torch
    from planning_motion_planning_b200 import decomp, synth
    c = synth.mars_costmap(1024, 2)
    g = synth.free_cell_near(c, 700, 300)
    cd = torch.from_numpy(c).cuda()
    single = eng.solve2d(cd, [g])[0]
    for nslabs in (2, 5):
        T, rounds = decomp.solve2d_slabs_local(cd, g, nslabs)
        assert rounds >= 2 and T.shape == single.shape
        fin = torch.isfinite(single)
        assert bool(torch.equal(torch.isfinite(T), fin))
        assert float(((T - single).abs() / single.clamp_min(1e-300))[fin].max()) < 1e-12


def test_device_side_failure_is_reported_not_hidden(eng, fmb_opts):
    """A solve that cannot finish (in-tile iteration cap hit) raises; the next solve is unaffected."""
    import torch
    from oracle import oracle as O
    from planning_motion_planning_b200 import _capi
    c = rand_map((120, 120), 3)
    cd = torch.from_numpy(c).cuda()
    fmb_opts(step_cap=3)
    with pytest.raises(_capi.FmbError) as ei:
        eng.solve2d(cd, [[60, 60]])
    assert ei.value.code == _capi.FMB_E_STEPCAP
    fmb_opts(step_cap=1 << 20)
    T = eng.solve2d(cd, [[60, 60]])[0].cpu().numpy()
    assert rel_err(T, O.computeTmap(c, [60, 60])) < TOL64
    with pytest.raises(_capi.FmbError):           # bad arguments are rejected by the C ABI
        _capi.check(_capi.lib().fmb_solve2d_f64(cd.data_ptr(), 10, 0, cd.data_ptr(), 120, 0, 120, 120, 1, None, None, 0, None))
    # a failure of a solve that was QUEUED (sync=False) is not erased by the next solve's init kernels: it is reported by
    # the next finish(), once
    fmb_opts(step_cap=3)
    eng.solve2d(cd, [[60, 60]], sync=False)
    fmb_opts(step_cap=1 << 20)
    eng.solve2d(cd, [[60, 60]], sync=False)
    with pytest.raises(_capi.FmbError) as ei:
        eng.finish()
    assert ei.value.code == _capi.FMB_E_STEPCAP
    T = eng.solve2d(cd, [[60, 60]])[0].cpu().numpy()          # ... and the workspace is clean again
    assert rel_err(T, O.computeTmap(c, [60, 60])) < TOL64


def test_tie_order_kernel_matches_torch_reference_and_oracle_order():
    """fmb_pop_ranks2d_f64 (radix sort + tie groups + the ordered sweep of csrc/tiekeys.cuh, one library call) == the
    torch implementation of the same rule on the CPU (tests/ranks_ref.py), and both reproduce the reference's true pop
    order on tie-heavy maps."""
    import torch
    import ranks_ref
    from FastMarching import _compat
    from oracle import oracle as O
    uniform = np.pad(np.ones((60, 60)), 1, constant_values=np.inf)
    for c, g, max_bad in ((uniform, [30, 30], 0), (uniform, [12, 40], 0), (plateau_map(80, 1), [8, 8], 0),
                          (plateau_map(80, 2), [8, 8], 0)):
        T, order, _ = O.computeTmap(c, g, return_stats=True)
        r_cpu = ranks_ref.pop_ranks_lifo2d(torch.from_numpy(T), torch.from_numpy(c), g)
        r_gpu = _compat.pop_ranks2d(torch.from_numpy(T).cuda(), torch.from_numpy(c).cuda(), g).cpu()
        st = _compat.ranks_status(torch.device("cuda", torch.cuda.current_device()))
        assert st[0] == 1 and st[1] > 1 and st[2] == 0 and st[3] == 0          # ties seen, no wait at the limit, no oversized group
        assert torch.equal(r_cpu, r_gpu)                  # iterated torch form == one-pass device sweep
        Td, cd = torch.from_numpy(T).cuda(), torch.from_numpy(c).cuda()
        flat = Td.reshape(-1)
        order0 = torch.sort(flat, stable=True).indices
        ts = flat[order0]
        grp = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.int32, device="cuda"), (ts[1:] != ts[:-1]).to(torch.int32)]), 0)
        group = torch.empty(flat.numel(), dtype=torch.int32, device="cuda")
        group[order0] = grp.to(torch.int32)
        rank0 = torch.empty(flat.numel(), dtype=torch.int32, device="cuda")
        rank0[order0] = torch.arange(flat.numel(), dtype=torch.int32, device="cuda")
        r_it = ranks_ref.pop_ranks_lifo2d_sort(Td, cd, g[1] * T.shape[1] + g[0], 96, group, rank0).cpu()
        assert torch.equal(r_it, r_gpu)                   # and == the iterated device fallback (huge tie groups)
        mine = np.argsort(r_gpu.numpy().ravel(), kind="stable")[1:1 + len(order)]
        assert int((mine != order).sum()) <= max_bad


@pytest.mark.parametrize("replay", ["dense", "sparse"])
def test_dropin_partial_fields_on_tie_heavy_maps(replay, fmb_opts):
    fmb_opts(replay_sparse=1 if replay == "sparse" else 0)        # both forms of the replay (truncate.cuh)
    """Uniform-cost and block-plateau maps are full of exactly equal T values; the reference pops
    those LIFO, which decides the join node and which cells are accepted when the fronts meet."""
    import FastMarching.FastMarching as FM
    from oracle import oracle as O
    uniform = np.pad(np.ones((40, 40)), 1, constant_values=np.inf)
    for c, g, s in ((uniform, [5, 5], [35, 30]), (uniform, [5, 20], [35, 20]), (plateau_map(80, 1), [8, 8], [70, 71]),
                    (plateau_map(80, 3), [8, 8], [70, 71])):
        TG, TS, j = FM.biComputeTmap(c, g, s)
        oTG, oTS, oj = O.biComputeTmap(c, g, s)
        assert np.array_equal(j, oj)
        assert rel_err(TG, oTG) < TOL64 and rel_err(TS, oTS) < TOL64
        T1 = FM.computeTmap(c, g, s)
        assert rel_err(T1, O.computeTmap(c, g, s)) < TOL64


@pytest.mark.parametrize("replay", ["dense", "sparse"])
def test_dropin_bisolve_fuzz_join_and_patterns_exact(replay, fmb_opts):
    fmb_opts(replay_sparse=1 if replay == "sparse" else 0)        # both forms of the replay (truncate.cuh)
    """Seeded random / plateau / uniform maps with walls through the public drop-in API: the join
    node and the accepted / narrow-band / +inf pattern of both partial fields equal the heap
    loop's exactly, the values to 1e-9 (FastMarching.py:114-162)."""
    import FastMarching.FastMarching as FM
    from oracle import oracle as O
    rng = np.random.default_rng(11)
    checked = 0
    for it in range(36):
        kind = it % 3
        if kind == 0:
            m = int(rng.integers(20, 90))
            c = rand_map((m, m + 5), int(rng.integers(0, 999)))
        elif kind == 1:
            c = plateau_map(64, int(rng.integers(0, 999)))
        else:
            m = int(rng.integers(20, 60))
            c = np.pad(np.ones((m, m)), 1, constant_values=np.inf)
        for _ in range(int(rng.integers(0, 4))):
            y, x = int(rng.integers(1, c.shape[0] - 1)), int(rng.integers(1, c.shape[1] - 1))
            c[y, x:x + int(rng.integers(1, 12))] = np.inf
        free = np.argwhere(np.isfinite(c))
        gy, gx = free[int(rng.integers(0, len(free)))]
        sy, sx = free[int(rng.integers(0, len(free)))]
        g, s = [int(gx), int(gy)], [int(sx), int(sy)]
        try:
            oTG, oTS, oj = O.biComputeTmap(c, g, s)
        except NameError:
            with pytest.raises(NameError):
                FM.biComputeTmap(c, g, s)
            continue
        if it % 2:                  # the planner passes an F-ordered view (cMap.T): solved as the transpose
            c = np.asfortranarray(c)
        TG, TS, j = FM.biComputeTmap(c, g, s)
        assert np.array_equal(j, oj), (it, kind)
        assert np.array_equal(np.isfinite(TG), np.isfinite(oTG)) and np.array_equal(np.isfinite(TS), np.isfinite(oTS)), (it, kind)
        assert rel_err(TG, oTG) < TOL64 and rel_err(TS, oTS) < TOL64
        checked += 1
    assert checked >= 24


def test_dropin_3d_early_exit_on_tie_heavy_volumes():
    """Uniform-cost volumes are full of exactly equal T values; which of them are accepted when the
    start pops follows the reference's LIFO order (FastMarching3D.py:77-95): identical inf pattern."""
    import FastMarching.FastMarching3D as FM3D
    from oracle import oracle as O

    def uniform_vol(n, val):
        c = np.full((n, n, n), val)
        c[0] = c[-1] = np.inf
        c[:, 0] = c[:, -1] = np.inf
        c[:, :, 0] = c[:, :, -1] = np.inf
        return c
    for c, g, s in ((uniform_vol(30, 20.0), [5, 6, 7], [22, 20, 18]), (uniform_vol(30, 20.0), [15, 15, 15], [22, 15, 15]),
                    (uniform_vol(26, 1.0), [3, 12, 20], [20, 12, 3])):
        T = FM3D.computeTmap(c, np.uint32(g), np.uint32(s))
        ref = O.computeTmap3D(c, g, s)
        assert np.array_equal(np.isfinite(T), np.isfinite(ref))
        assert rel_err(T, ref) < TOL64


def test_dropin_errors():
    import FastMarching.FastMarching as FM
    c = rand_map((40, 40), 0)
    c[:, 20] = np.inf                                 # two disconnected halves: fronts never meet
    with pytest.raises(NameError):
        FM.biComputeTmap(c, [5, 5], [30, 30])
    with pytest.raises(IndexError):
        FM.biComputeTmap(c, [5, 5], [40, 3])


# ---- 2D cost-map builder --------------------------------------------------
def _box_sum_50(a, fill):
    """Row sums over x-25..x+24 then column sums over y-25..y+24 (`fill` outside), added one tap at a
    time in ascending order -- the kernel's own order, so the comparison is exact at any size (scipy's
    2500-tap convolve2d takes minutes at 2048^2 and beyond)."""
    n = a.shape[0]
    p = np.pad(a, ((0, 0), (25, 24)), constant_values=fill)
    h = np.zeros_like(a)
    for u in range(50):
        h = h + p[:, u:u + n]
    p = np.pad(h, ((25, 24), (0, 0)), constant_values=50 * fill)
    v = np.zeros_like(a)
    for u in range(50):
        v = v + p[u:u + n, :]
    return v


@pytest.mark.parametrize("kind,n,res,seed,sigma", [("planner", 200, 0.05, 0, None), ("craters", 512, 0.05, 1, None),
                                                  ("rough", 384, 0.06, 2, 0.09), ("coarse", 300, 0.1, 4, 0.12)])
def test_costmap2d_vs_oracle(kind, n, res, seed, sigma):
    """csrc/costmap2d.cuh through the C ABI against the cv2/scipy restatement of
    Coupled_motion_planner.py:1144-1216: obstacle maps and the pre-blur cost bit-exact, the
    blurred map to 1e-12 (summation order), identical +inf limits."""
    import torch
    from oracle import costmap_oracle as CO
    from planning_motion_planning_b200 import costmap, synth
    if kind == "planner":
        Z = CO.planner_dem(n, res)
    else:
        Z = synth.crater_dem(n, res, seed)
        if sigma:
            Z = Z + np.random.default_rng(seed + 100).normal(0.0, sigma * res, Z.shape)
            Z = Z - Z.min()
    c, st = CO.costmap2d(Z, res, n * res, stages=True)
    cost, dv = costmap.build_costmap_device(torch.from_numpy(Z).cuda(), res, n * res, stages=True)
    assert np.array_equal(dv["raw"].cpu().numpy(), st["raw"])
    assert np.array_equal(dv["obst"].cpu().numpy(), st["obst"].astype(np.uint8))
    assert np.array_equal(dv["pre_blur"].cpu().numpy(), st["pre_blur"].T)
    got = cost.cpu().numpy().T
    fin = np.isfinite(c)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.max(np.abs(got[fin] - c[fin]) / c[fin]) < 1e-12
    if kind == "planner":     # and against the unmodified planner's own array
        d = np.load(os.path.join(GOLDEN, "planner_calls.npz"), allow_pickle=True)
        ref = d["bi_cost"]
        f2 = np.isfinite(ref)
        assert np.max(np.abs(cost.cpu().numpy()[f2] - ref[f2]) / ref[f2]) < 1e-12
        assert np.array_equal(costmap.build_costmap(Z, res, n * res), got)


def test_costmap2d_2048_stages_exact_and_blur_property():
    import torch
    from oracle import costmap_oracle as CO
    from planning_motion_planning_b200 import costmap, synth
    n, res = 2048, 0.05
    Z = synth.crater_dem(n, res, 1)
    _, st = CO.costmap2d(Z, res, n * res, stages=True, blur=False)
    cost, dv = costmap.build_costmap_device(torch.from_numpy(Z).cuda(), res, n * res, stages=True)
    assert np.array_equal(dv["raw"].cpu().numpy(), st["raw"])
    assert np.array_equal(dv["obst"].cpu().numpy(), st["obst"].astype(np.uint8))
    pre = dv["pre_blur"].cpu().numpy()
    assert np.array_equal(pre, st["pre_blur"].T)
    want = _box_sum_50(pre, 300.0) * (1.0 / 2500.0)
    got = cost.cpu().numpy()
    inner = (slice(1, -1), slice(1, -1))
    assert np.all(np.isinf(got[0])) and np.all(np.isinf(got[-1])) and np.all(np.isinf(got[:, 0])) and np.all(np.isinf(got[:, -1]))
    assert np.array_equal(got[inner], want[inner])


def test_costmap2d_feeds_the_solver_like_the_planner():
    """DEM -> device cost map -> biComputeTmap: same join node and paths as the unmodified planner run."""
    import FastMarching.FastMarching as FM
    from oracle import costmap_oracle as CO
    from planning_motion_planning_b200 import costmap
    d = np.load(os.path.join(GOLDEN, "planner_calls.npz"), allow_pickle=True)
    n, res = int(d["n"]), float(d["res"])
    cMap = costmap.build_costmap(CO.planner_dem(n, res), res, n * res)
    TG, TS, join = FM.biComputeTmap(cMap.T, [int(v) for v in d["bi_goal"]], [int(v) for v in d["bi_start"]])
    assert np.array_equal(join, d["bi_join"])
    pG = FM.getPathGDM(TG, join, [int(v) for v in d["bi_goal"]], 0.5)
    assert pG.shape == d["pathG"].shape and np.max(np.abs(pG - d["pathG"])) < TOLP


def test_costmap2d_all_obstacle_map_raises_valueerror():
    import torch
    from planning_motion_planning_b200 import costmap, synth
    n, res = 96, 0.03
    with pytest.raises(ValueError):
        costmap.build_costmap_device(torch.from_numpy(synth.crater_dem(n, res, 3)).cuda(), res, n * res)


# ---- 3D cost-volume builder ------------------------------------------------
def _cv_cases():
    d = np.load(os.path.join(GOLDEN, "costvolume.npz"))
    out = {}
    for tag in ("p", "d0", "d1"):
        g = lambda k: d[f"{tag}_{k}"]      # noqa: E731
        out[tag] = dict(Zs=g("Zs"), res=[float(v) for v in g("res")], shape=[int(v) for v in g("shape")], obst=g("obst"),
                        xm_ym=[float(v) for v in g("xm_ym")], final=g("final"), radii=[float(v) for v in g("radii")],
                        path=g("path"), heading=g("heading"), fin=g("fin"), ini=g("ini"), tunnel=g("tunnel"))
    return out


@pytest.mark.parametrize("tag", ["p", "d0", "d1"])
def test_costvolume_vs_golden(tag):
    """csrc/costvolume.cuh through the C ABI, called with the reference's own function signatures,
    against arrays captured from the unmodified GetObstMap / TunnelCost: bit-exact."""
    from planning_motion_planning_b200 import costvolume as CVP
    c = _cv_cases()[tag]
    (rx, ry, rz), (sX, sY, sZ) = c["res"], c["shape"]
    final, obst, ground = CVP.GetObstMap(c["Zs"], rx, ry, rz, sX, sY, sZ, c["obst"], *c["xm_ym"])
    assert np.array_equal(final, c["final"])
    inner = (slice(1, -1),) * 3
    assert np.array_equal((obst + ground)[inner], c["final"][inner])
    assert not np.any(np.isinf(obst) & np.isinf(ground))
    tunnel = CVP.TunnelCost(*c["radii"], c["path"], sX, sY, sZ, rx, ry, rz, c["heading"], c["fin"], c["ini"])
    assert np.array_equal(tunnel, c["tunnel"])
    cmap = CVP.build_cost_volume(c["Zs"], rx, ry, rz, sX, sY, sZ, *c["xm_ym"], *c["radii"], c["path"], c["heading"], c["fin"], c["ini"])
    assert np.array_equal(cmap, c["final"] * c["tunnel"])


def test_costvolume_feeds_the_3d_solver_like_the_planner(eng):
    """DEM crop + base path -> device cost volume -> 3D solve + path == the planner's own volume and path."""
    import FastMarching.FastMarching3D as FM3D
    from planning_motion_planning_b200 import costvolume as CVP
    p = np.load(os.path.join(GOLDEN, "planner_calls.npz"), allow_pickle=True)
    c = _cv_cases()["p"]
    (rx, ry, rz), (sX, sY, sZ) = c["res"], c["shape"]
    vol = CVP.build_cost_volume_device(c["Zs"], rx, ry, rz, sX, sY, sZ, *c["xm_ym"], *c["radii"], c["path"], c["heading"],
                                       c["fin"], c["ini"])
    assert np.array_equal(vol.cpu().numpy(), p["c3"])
    T = FM3D.computeTmap(vol.cpu().numpy(), np.uint32(p["g3"]), np.uint32(p["s3"]))
    path = FM3D.getPathGDM(T, np.uint32(p["path3d_init"]), np.uint32(p["path3d_end"]), float(p["path3d_tau"]))
    assert path.shape == p["path3d"].shape and np.max(np.abs(path - p["path3d"])) < TOLP


def test_costvolume_large_vs_oracle():
    """A 160^3 volume with a long, banking base path (2 x 10^6 scatter events) against the C restatement."""
    from oracle import costvol as CV
    from planning_motion_planning_b200 import costvolume as CVP
    rng = np.random.default_rng(5)
    sX = sY = sZ = 160
    rx = ry = 0.0125
    rz = 0.02
    m = 120
    Zs = 0.2 + 0.1 * rng.random((sX, sY))
    s = np.linspace(0, 1, m)
    path = np.stack([(0.2 + 0.6 * s) * sX * rx, (0.3 + 0.4 * s ** 2) * sY * ry, 1.2 + 0.3 * np.sin(3 * s)], axis=1)
    head = np.stack([0.2 * np.sin(5 * s), 0.15 * np.cos(4 * s), 0.3 + 1.2 * s], axis=1)
    fin, ini = np.uint32([120, 90, 70]), np.uint32([40, 50, 65])
    rad = (0.527, 0.2673, 0.1105)
    want = CV.GetObstMap(Zs, rx, ry, rz, sX, sY, sZ, np.zeros((sX, sY)), 0.3, 0.4) * \
        CV.TunnelCost(*rad, path, sX, sY, sZ, rx, ry, rz, head, fin, ini)
    got = CVP.build_cost_volume(Zs, rx, ry, rz, sX, sY, sZ, 0.3, 0.4, *rad, path, head, fin, ini)
    assert np.array_equal(got, want)
    assert 0.02 < np.mean((want != 20) & np.isfinite(want)) < 0.6        # the tunnel is really there


# ------------------------------------------------------------------ BASELINE-size parity (configs 2, 3, 4) and regressions of round 2
def _path_close(p, po):
    return p.shape == po.shape and (len(po) == 0 or float(np.abs(p - po).max()) < TOLP)


def test_config3_volume_256_cubed_solve_and_path_vs_oracle(eng):
    """BASELINE config 3 at full size: 256^3 planner-like arm volume, field within 1e-9 of the oracle (C port of the
    reference, ~9 s on one core), identical inf pattern, and the extracted path within 1e-3 cell."""
    import torch
    from oracle import oracle as O
    from planning_motion_planning_b200 import synth
    c, g, s = synth.arm_volume((256, 256, 256), 0)
    T = eng.solve3d(torch.from_numpy(c).cuda(), [g])
    ref = O.computeTmap3D(c, g)
    assert rel_err(T[0].cpu().numpy(), ref) < TOL64
    out, cnt, st = eng.trace3d(T, np.asarray([s], dtype=np.float64), np.asarray([g], dtype=np.float64), 0.5)
    po, so = O.getPathGDM3D(ref, np.uint32(s), np.uint32(g), 0.5, return_status=True)
    assert int(st[0]) == so
    assert _path_close(out[0, :int(cnt[0])].cpu().numpy(), po)


def test_config4_batch_4096_queries_512_sampled_vs_oracle(eng):
    """BASELINE config 4 at full size on one GPU: 4096 goal queries on a 512^2 costmap through batch.solve_queries;
    24 sampled queries are checked against the oracle (field 1e-9 through a second solve of the same goals, path 1e-3)."""
    import torch
    from oracle import oracle as O
    from planning_motion_planning_b200 import batch, synth
    n, Q = 512, 4096
    c = synth.mars_costmap(n, 1)
    rng = np.random.default_rng(100)
    ok = np.argwhere(np.isfinite(c) & (c <= 2.0))
    goals = ok[rng.integers(0, len(ok), size=Q)][:, ::-1].tolist()
    starts = ok[rng.integers(0, len(ok), size=Q)][:, ::-1].tolist()
    lo, res = batch.solve_queries(c, goals, starts, chunk=1024)
    assert lo == 0 and len(res) == Q
    pick = rng.choice(Q, size=24, replace=False)
    fields = eng.solve2d(torch.from_numpy(c).cuda(), [goals[i] for i in pick]).cpu().numpy()
    for k, i in enumerate(pick):
        ref = O.computeTmap(c, goals[i])
        assert rel_err(fields[k], ref) < TOL64
        po, so = O.getPathGDM(ref, np.array(starts[i], dtype=np.float64), goals[i], 0.5, return_status=True)
        p, st = res[i]
        assert st == so and _path_close(p, po)


@pytest.mark.parametrize("n,order", [(1024, "C"), (1024, "F"), (4096, "C"), (4096, "F")])
def test_dropin_bicomputetmap_and_paths_at_config_sizes(n, order):
    """The drop-in call the planner makes (FM.biComputeTmap + two FM.getPathGDM, Coupled_motion_planner.py:1226-1230) at
    config 2 (1024^2) and at the metric's 4096^2, C- and F-ordered input (the planner passes cMap.T): join node exact,
    inf patterns identical, partial fields 1e-9, both half paths 1e-3 cell against the oracle."""
    import FastMarching.FastMarching as FM
    from oracle import oracle as O
    from planning_motion_planning_b200 import synth
    c = synth.mars_costmap(n, 0)
    g = list(synth.free_cell_near(c, n // 4, n // 4))
    s = list(synth.free_cell_near(c, 3 * n // 4, 3 * n // 4))
    cin = c if order == "C" else np.asfortranarray(c)
    TG, TS, j = FM.biComputeTmap(cin, g, s)
    oTG, oTS, oj = O.biComputeTmap(c, g, s)
    assert np.array_equal(j, oj)
    assert rel_err(np.asarray(TG), oTG) < TOL64 and rel_err(np.asarray(TS), oTS) < TOL64
    for Tm, oTm, endp in ((TG, oTG, g), (TS, oTS, s)):
        po = O.getPathGDM(oTm, np.array(oj, dtype=np.float64), endp, 0.5)
        assert _path_close(FM.getPathGDM(Tm, j, endp, 0.5), po)


def test_start_equals_goal_returns_what_the_reference_returns():
    """start == goal: the reference never pops the goal (it is closed before the loop), so computeTmap returns the FULL
    field in 2D-as-intended and in 3D, and biComputeTmap joins at the first popped node with k = 1."""
    import FastMarching.FastMarching as FM
    import FastMarching.FastMarching3D as FM3D
    from oracle import oracle as O
    for seed, shape, g in ((1, (40, 50), [10, 12]), (2, (64, 33), [5, 40])):
        c = rand_map(shape, seed)
        for cin in (c, np.asfortranarray(c)):
            assert rel_err(FM.computeTmap(cin, g, g), O.computeTmap(c, g, g)) < TOL64
            TG, TS, j = FM.biComputeTmap(cin, g, g)
            oTG, oTS, oj = O.biComputeTmap(c, g, g)
            assert np.array_equal(j, oj)
            assert rel_err(np.asarray(TG), oTG) < TOL64 and rel_err(np.asarray(TS), oTS) < TOL64
    c3 = rand_map((12, 14, 16), 3)
    assert rel_err(FM3D.computeTmap(c3, np.uint32([5, 6, 7]), np.uint32([5, 6, 7])), O.computeTmap3D(c3, [5, 6, 7], [5, 6, 7])) < TOL64


def test_polish3d_reaches_the_fixed_point_of_the_reference_arithmetic(eng):
    """fmb_polish3d_f64: after the polish pass every free cell equals, bit for bit, the value the reference's own update
    (libm pow for the squares it takes on NumPy scalars, oracle.solve3d) assigns from the cell's six neighbours, and the
    field stays within 1e-9 of the oracle's."""
    import torch
    from oracle import oracle as O
    from planning_motion_planning_b200 import synth
    for c, g in ((rand_map((20, 22, 24), 1), [5, 6, 7]), (synth.arm_volume((40, 40, 28), 0)[0], synth.arm_volume((40, 40, 28), 0)[1])):
        T = eng.solve3d(torch.from_numpy(c).cuda(), [list(g)], exact=True)[0].cpu().numpy()
        assert rel_err(T, O.computeTmap3D(c, g)) < TOL64
        P = np.pad(T, 1, constant_values=np.inf)
        rng = np.random.default_rng(0)
        free = np.argwhere(np.isfinite(T) & (T > 0))
        for y, x, z in free[rng.choice(len(free), size=min(400, len(free)), replace=False)]:
            tx = min(P[y + 1, x, z + 1], P[y + 1, x + 2, z + 1])
            ty = min(P[y, x + 1, z + 1], P[y + 2, x + 1, z + 1])
            tz = min(P[y + 1, x + 1, z], P[y + 1, x + 1, z + 2])
            v = O.solve3d(tx, ty, tz, c[y, x, z])
            assert v == T[y, x, z] or abs(v - T[y, x, z]) <= 4 * np.spacing(T[y, x, z])
            # (exact equality for all but cells whose inputs are themselves within the acceptance threshold)


@pytest.mark.parametrize("replay", ["dense", "sparse"])
def test_dropin_3d_early_exit_fuzz_slice(replay, fmb_opts):
    fmb_opts(replay_sparse=1 if replay == "sparse" else 0)        # both forms of the replay (truncate.cuh)
    """120 volumes of tools/gpu_fuzz_3d.py (70 % uniform-cost with obstacles: the tie-heavy class the planner's real
    volume belongs to; 30 % random) through FM3D.computeTmap: the accepted / narrow-band / far pattern is the
    reference's in EVERY case and the values agree to 1e-9.  This needs the field to carry the reference's own rounding
    (libm pow for `**2` on NumPy scalars, csrc/pow2_glibc.cuh + fmb_polish3d_f64): without the polish pass 13 of 300
    such volumes differ in 1-10 cells around T[start] (tools/gpu_fuzz_3d.py <n> <seed> 0)."""
    import importlib.util
    import FastMarching.FastMarching3D as FM3D
    from conftest import ROOT
    from oracle import oracle as O
    spec = importlib.util.spec_from_file_location("gpu_fuzz_3d", os.path.join(ROOT, "tools", "gpu_fuzz_3d.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    rng = np.random.default_rng(0)
    for it in range(120):
        c, g, s, uniform = fz.case(rng)
        ref = O.computeTmap3D(c, g, s)
        got = FM3D.computeTmap(c, np.uint32(g), np.uint32(s))
        assert np.array_equal(np.isfinite(got), np.isfinite(ref)), (it, uniform, int((np.isfinite(got) != np.isfinite(ref)).sum()))
        f = np.isfinite(ref)
        assert float(np.max(np.abs(got[f] - ref[f]) / np.maximum(ref[f], 1.0))) < 1e-9


def test_one_host_thread_two_devices_timing_events():
    """the timing events of the C ABI are kept per (thread, device, stream); one host thread that solves on a
    second GPU must not fail (needs >= 2 GPUs, skipped otherwise)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import oracle as O
    from planning_motion_planning_b200 import engine
    c = rand_map((120, 140), 2)
    ref = O.computeTmap(c, [60, 60])
    for dev in (0, 1, 0, 1):
        with torch.cuda.device(dev):
            T = engine.solve2d(torch.from_numpy(c).to(f"cuda:{dev}"), [[60, 60]])[0].cpu().numpy()
            assert engine.last_stats()["solve_kernel_ms"] > 0
        assert rel_err(T, ref) < TOL64


def test_c_abi_alone_reproduces_bicomputetmap_and_the_3d_early_exit(fmb_opts):
    """What INTEGRATION.md promises a C / C++ host: the reference's RETURN VALUES from include/fm_b200.h alone.  torch
    only provides device memory here -- no torch op runs between the library calls: fmb_bisolve2d_f64 (KAT-3b: join
    node, both partial fields), fmb_solve2d_until_f64, fmb_solve3d_until_f64 (KAT-4 truncated field), each bitwise
    (2D) / 1e-9 (3D) against the oracle's early-exit loops."""
    import ctypes as C
    import torch
    from oracle import oracle as O
    from planning_motion_planning_b200 import _capi
    L = _capi.lib()
    fmb_opts(replay_sparse=1)               # (auto = dense below 2^21 cells; the sparse form is asserted on below)
    st = torch.cuda.current_stream().cuda_stream
    i32 = lambda v: (C.c_int32 * len(v))(*v)
    c = rand_map((100, 100), 0)
    g, s = [10, 10], [90, 90]
    cd = torch.from_numpy(c).cuda()
    ws = torch.empty(L.fmb_workspace_bytes_bisolve2d(100, 100), dtype=torch.uint8, device="cuda")
    out = torch.empty((2, 100, 100), dtype=torch.float64, device="cuda")
    info = torch.empty(16, dtype=torch.int32, device="cuda")
    _capi.check(L.fmb_bisolve2d_f64(cd.data_ptr(), 100, 100, i32(g), i32(s), 0, out[0].data_ptr(), out[1].data_ptr(), info.data_ptr(),
                                    ws.data_ptr(), ws.numel(), st, None))
    _capi.check(L.fmb_finish(ws.data_ptr(), ws.numel(), st, None))
    TG, TS, j = O.biComputeTmap(c, g, s)
    inf = info.tolist()
    assert [inf[1] % 100, inf[1] // 100] == [int(j[0]), int(j[1])] == [64, 36]
    assert inf[6] == 0 and inf[10] == 0 and inf[12] == 0 and inf[13] == 0
    # the replay ran in its sparse form (the dependency cone of the narrow band) on both fronts
    assert (inf[15] & 3) == 0 and 0 < inf[14] <= 4 * (inf[0] + 1), inf
    for a, ref in ((out[0], TG), (out[1], TS)):
        assert rel_err(a.cpu().numpy(), ref) < TOL64
    # single front, early exit when `start` is accepted
    ws2 = torch.empty(L.fmb_workspace_bytes_until2d(100, 100), dtype=torch.uint8, device="cuda")
    T2 = torch.empty((100, 100), dtype=torch.float64, device="cuda")
    _capi.check(L.fmb_solve2d_until_f64(cd.data_ptr(), 100, 100, i32([25, 25]), i32([60, 70]), 0, T2.data_ptr(), info.data_ptr(),
                                        ws2.data_ptr(), ws2.numel(), st))
    _capi.check(L.fmb_finish(ws2.data_ptr(), ws2.numel(), st, None))
    assert rel_err(T2.cpu().numpy(), O.computeTmap(c, [25, 25], [60, 70])) < TOL64
    assert (info.tolist()[15] & 3) == 0, info.tolist()
    # 3D (KAT-4)
    c3 = rand_map((24, 24, 24), 0)
    g3, s3 = [5, 6, 7], [18, 17, 16]
    c3d = torch.from_numpy(c3).cuda()
    ws3 = torch.empty(L.fmb_workspace_bytes_until3d(24, 24, 24), dtype=torch.uint8, device="cuda")
    T3 = torch.empty((24, 24, 24), dtype=torch.float64, device="cuda")
    _capi.check(L.fmb_solve3d_until_f64(c3d.data_ptr(), 24, 24, 24, i32(g3), i32(s3), T3.data_ptr(), info.data_ptr(), ws3.data_ptr(),
                                        ws3.numel(), st))
    _capi.check(L.fmb_finish(ws3.data_ptr(), ws3.numel(), st, None))
    ref3 = O.computeTmap3D(c3, g3, s3)
    assert int(np.isfinite(ref3).sum()) == 9830 and rel_err(T3.cpu().numpy(), ref3) < TOL64
    assert (info.tolist()[15] & 3) == 0 and (info.tolist()[15] >> 8) >= 1, info.tolist()      # ([14] = the exact solve ran, here)


def test_batch_long_paths_are_retraced_with_the_reference_cap(eng, monkeypatch):
    """batch.solve_chunk_gpu gives every path room for a few crossings of the map and re-traces the ones that use it up
    with the reference's own 30000-step cap: a serpentine map with a long path next to short ones (room cut to one
    crossing here so that the long path needs the second pass)."""
    from oracle import oracle as O
    from planning_motion_planning_b200 import batch
    monkeypatch.setattr(batch, "FIRST_PASS_CROSSINGS", 1)
    n = 64
    c = np.ones((n, n))
    c[0, :] = c[-1, :] = c[:, 0] = c[:, -1] = np.inf
    for k, y in enumerate(range(4, n - 4, 4)):                    # walls with a gap alternating left / right
        c[y, 1:n - 1] = np.inf
        c[y, (n - 4) if k % 2 == 0 else 2:(n - 2) if k % 2 == 0 else 4] = 1.0
    goals = [[2, 2], [2, 2], [40, 2]]
    starts = [[n - 3, n - 3], [10, 2], [3, n - 3]]
    lo, res = batch.solve_queries(c, goals, starts, chunk=8)
    assert lo == 0 and len(res) == 3
    long_rows = 0
    for (p, st), g, s in zip(res, goals, starts):
        T = O.computeTmap(c, g)
        ref, rst = O.getPathGDM(T, np.array(s, dtype=np.float64), np.array(g, dtype=np.float64), 0.5, return_status=True)
        assert st == rst and p.shape == ref.shape and np.abs(p - ref).max() < TOLP
        long_rows = max(long_rows, len(p))
    assert long_rows > 2 * (n + n) + 66          # at least one path exceeded the first-pass room


def test_batch_default_engine_reads_costs_from_global_memory(eng, fmb_opts):
    """Best-first batches run the armed-cell visit WITHOUT the shared-memory cost tile by default (engine2d = 6, 20 resident
    warps per SM): ragged map with obstacles, per-query cost maps, fp64 against the oracle and fp32 within 1e-4."""
    import torch
    from oracle import oracle as O
    from planning_motion_planning_b200 import _capi
    assert _capi.get_options()["engine2d"] == 0
    rng = np.random.default_rng(21)
    shape = (257, 300)
    cs = 1.0 + 4.0 * rng.random((12,) + shape)
    cs[rng.random(cs.shape) < 0.06] = np.inf
    goals = [[int(rng.integers(1, shape[1] - 1)), int(rng.integers(1, shape[0] - 1))] for _ in range(12)]
    for q, g in enumerate(goals):
        cs[q, g[1], g[0]] = 1.0
    T = eng.solve2d(torch.from_numpy(cs).cuda(), goals).cpu().numpy()           # (nq, rows, cols) costs: one map per query
    for q in (0, 5, 11):
        assert rel_err(T[q], O.computeTmap(cs[q], goals[q])) < TOL64
    Ts = eng.solve2d(torch.from_numpy(cs[0]).cuda(), goals).cpu().numpy()        # one shared map
    for q in (1, 7):
        c0 = cs[0].copy()
        ref = O.computeTmap(c0, goals[q])
        if np.isfinite(c0[goals[q][1], goals[q][0]]):
            assert rel_err(Ts[q], ref) < TOL64
    T32 = eng.solve2d(torch.from_numpy(cs[0].astype(np.float32)).cuda(), goals).cpu().numpy()
    assert rel_err(T32[0].astype(np.float64), O.computeTmap(cs[0].astype(np.float32).astype(np.float64), goals[0])) < 1e-4


@pytest.mark.parametrize("order", ["C", "F"])
def test_dropin_tracer_reuses_the_device_field_only_if_unchanged(order):
    """getPathGDM on an array computeTmap / biComputeTmap returned starts on the device copy the drop-in kept while the
    array is uploaded again; the bitwise comparison of the two decides (fmb_fields_differ_f64).  Unchanged array: the
    oracle's path.  Array changed in place by the caller (a wall of +inf across the old path, one value nudged by an
    ulp): the oracle's path over the CHANGED array."""
    import gc
    import FastMarching.FastMarching as FM
    from FastMarching import _compat as C
    from oracle import oracle as O
    from planning_motion_planning_b200 import synth
    n = 640
    c = synth.mars_costmap(n, 6)
    goal = list(synth.free_cell_near(c, 80, 90)); start = list(synth.free_cell_near(c, 560, 500))
    cm = np.asfortranarray(c.T).T if order == "C" else np.asfortranarray(c)       # same values, C- or F-ordered storage
    T = FM.computeTmap(cm, goal, [-1, -1])
    assert C._device_copy_of(T.T if order == "F" else T) is not None
    Tref = O.computeTmap(c, goal)
    assert rel_err(np.ascontiguousarray(T), Tref) < TOL64
    s = np.array(start, dtype=np.float64)
    before = dict(C.TRACE_STATS)
    p = FM.getPathGDM(T, s, goal, 0.5)
    ref = O.getPathGDM(np.ascontiguousarray(T), s, np.array(goal, dtype=np.float64), 0.5)
    assert p.shape == ref.shape and np.abs(p - ref).max() < TOLP
    assert C.TRACE_STATS["reused"] == before["reused"] + 1 and C.TRACE_STATS["retraced"] == before["retraced"]
    # an edit the path never looked at (far from every window of its gradient blocks) changes nothing the tracer reads:
    # the kept device copy still serves, and the result is the oracle's path over the EDITED array
    yy, xx = np.mgrid[20:n - 20:40, 20:n - 20:40]
    far = np.array([np.min(np.hypot(p[:, 0] - x, p[:, 1] - y)) for y, x in zip(yy.ravel(), xx.ravel())])
    fy, fx = int(yy.ravel()[far.argmax()]), int(xx.ravel()[far.argmax()])
    assert far.max() > 60
    T[fy - 3:fy + 3, fx - 3:fx + 3] += 1.0
    p1 = FM.getPathGDM(T, s, goal, 0.5)
    ref1 = O.getPathGDM(np.ascontiguousarray(T), s, np.array(goal, dtype=np.float64), 0.5)
    assert p1.shape == ref1.shape and np.abs(p1 - ref1).max() < TOLP and np.array_equal(p1, p)
    assert C.TRACE_STATS["reused"] == before["reused"] + 2 and C.TRACE_STATS["retraced"] == before["retraced"]
    # the caller edits the array in place: a wall with one gap far from the old path
    mid = int(p[len(p) // 2, 1])
    T[mid, :] = np.inf
    T[mid, 5:8] = Tref[mid, 5:8]
    p2 = FM.getPathGDM(T, s, goal, 0.5)
    ref2, st2 = O.getPathGDM(np.ascontiguousarray(T), s, np.array(goal, dtype=np.float64), 0.5, return_status=True)
    assert p2.shape == ref2.shape and (len(ref2) == 0 or np.abs(p2 - ref2).max() < TOLP)
    assert p2.shape != p.shape or np.abs(p2 - p).max() > 1.0
    assert C.TRACE_STATS["retraced"] == before["retraced"] + 1
    # a one-ulp change anywhere is seen as well
    T2 = FM.computeTmap(cm, goal, [-1, -1])
    j, i = int(p[3, 1]), int(p[3, 0])
    T2[j, i] = np.nextafter(T2[j, i], np.inf)
    flag_before = C._device_copy_of(T2.T if order == "F" else T2)
    assert flag_before is not None
    p3 = FM.getPathGDM(T2, s, goal, 0.5)
    ref3 = O.getPathGDM(np.ascontiguousarray(T2), s, np.array(goal, dtype=np.float64), 0.5)
    assert p3.shape == ref3.shape and np.abs(p3 - ref3).max() < 1e-9
    assert C.TRACE_STATS["retraced"] == before["retraced"] + 2
    # the kept device copies are released with the host arrays
    del T, T2, flag_before
    gc.collect()
    assert len(C._DEVCOPY) == 0


def test_dropin_pagelocks_a_callers_array_on_reuse():
    """A cost map that comes in a second time is page-locked in place and uploaded by one DMA from then on; every call
    still uploads the CURRENT contents (an in-place edit between calls is seen); the registration ends with the array."""
    import gc
    import FastMarching.FastMarching as FM
    from FastMarching import _compat as C
    from oracle import oracle as O
    from planning_motion_planning_b200 import synth
    n = 640
    c = synth.mars_costmap(n, 8)
    goal = list(synth.free_cell_near(c, 500, 120))
    n0 = len(C._REGISTERED)
    T1 = FM.computeTmap(c, goal, [-1, -1])
    assert len(C._REGISTERED) == n0
    T2 = FM.computeTmap(c, goal, [-1, -1])                  # second sighting: registered
    assert len(C._REGISTERED) == n0 + 1
    ref = O.computeTmap(c, goal)
    assert rel_err(T1, ref) < TOL64 and rel_err(T2, ref) < TOL64
    c[300:340, 100:500] = np.inf                            # the caller edits its map in place
    T3 = FM.computeTmap(c, goal, [-1, -1])
    assert rel_err(T3, O.computeTmap(c, goal)) < TOL64 and not np.array_equal(np.isfinite(T3), np.isfinite(T2))
    start = list(synth.free_cell_near(c, 100, 560))
    TG, TS, jn = FM.biComputeTmap(c, goal, start)          # page-locked map: uploaded in bands behind the two solves
    oTG, oTS, oj = O.biComputeTmap(c, goal, start)
    assert list(jn) == list(oj) and rel_err(TG, oTG) < TOL64 and rel_err(TS, oTS) < TOL64
    cf = np.asfortranarray(c)                               # the planner's F-ordered view, twice: the second call overlaps
    for _ in range(2):
        TGf, TSf, jf = FM.biComputeTmap(cf, goal, start)
        assert list(jf) == list(oj) and rel_err(np.ascontiguousarray(TGf), oTG) < TOL64 and rel_err(np.ascontiguousarray(TSf), oTS) < TOL64
    del cf
    del c
    gc.collect()
    assert len(C._REGISTERED) == n0


def test_branch_free_division_is_correctly_rounded(eng):
    """ddiv_rn_fast (csrc/fm_common.cuh, the tracer's fast step) == div.rn.f64 bit for bit wherever it accepts its
    operands: random mantissas over the accepted exponent range and beyond it, quotients next to 1 and to powers of two,
    the tracer's own operand shapes (a gradient component over a hypotenuse), zeros, infinities and NaNs."""
    import torch
    from planning_motion_planning_b200 import _capi
    g = torch.Generator(device="cuda").manual_seed(2)
    n = 10_000_000
    def rnd(lo, hi):
        m = (torch.rand(n, dtype=torch.float64, device="cuda", generator=g) + 1.0)
        e = torch.randint(lo, hi, (n,), device="cuda", generator=g).to(torch.float64)
        sgn = torch.where(torch.rand(n, device="cuda", generator=g) < 0.5, -1.0, 1.0).to(torch.float64)
        return sgn * m * torch.exp2(e)
    cases = [(rnd(-520, 520), rnd(-520, 520)), (rnd(-1070, 1023), rnd(-1070, 1023)), (rnd(-3, 3), rnd(-3, 3))]
    b = rnd(-2, 2).abs()
    k = torch.randint(-3, 4, (n,), device="cuda", generator=g).to(torch.float64)
    cases.append((b * (1.0 + k * 2.0 ** -52), b))                             # quotients within a few ulp of 1
    dx, dy = rnd(-30, 1), rnd(-30, 1)
    cases.append((dx, torch.sqrt(dx * dx + dy * dy)))                          # the tracer's nx = dx / hypot(dx, dy)
    sp = torch.tensor([0.0, -0.0, 1.0, -1.0, float("inf"), float("-inf"), float("nan"), 5e-324, 1e-310, 1e308, 3.0, 0.01],
                      dtype=torch.float64, device="cuda")
    cases.append((sp.repeat_interleave(len(sp)), sp.repeat(len(sp))))
    accepted = 0
    for a, b in cases:
        bad = torch.zeros(2, dtype=torch.int64, device="cuda")
        _capi.check(_capi.lib().fmb_debug_div_check(a.contiguous().data_ptr(), b.contiguous().data_ptr(), a.numel(), bad.data_ptr(), None))
        torch.cuda.synchronize()
        assert int(bad[0]) == 0
        accepted += int(bad[1])
    assert accepted > 25_000_000


@pytest.mark.parametrize("shape,goal", [((2048, 2048), [300, 1900]), ((1000, 1100), [1050, 20]), ((97, 130), [5, 90])])
def test_solve_overlapped_with_the_upload_of_its_cost_map(eng, shape, goal):
    """fmb_solve2d_h2d_f64: the cost map is uploaded in bands of rows (nearest to the goal first) on a second stream
    while the solve already runs and waits per band on device flags.  Same field as the solve of the resident map (to
    rounding: the visit order differs) and as the oracle; the device copy of the map is complete afterwards; a pageable
    host array is refused."""
    import ctypes as C
    import torch
    from oracle import oracle as O
    from planning_motion_planning_b200 import _capi, synth
    L = _capi.lib()
    rows, cols = shape
    c = synth.mars_costmap(max(rows, cols), 12)[:rows, :cols].copy() if rows >= 512 else rand_map(shape, 5)
    c[0, :] = c[-1, :] = c[:, 0] = c[:, -1] = np.inf
    c[goal[1], goal[0]] = 1.0
    h = torch.empty(shape, dtype=torch.float64).pin_memory()
    h.copy_(torch.from_numpy(c))
    cd = torch.full(shape, float("nan"), dtype=torch.float64, device="cuda")
    T = torch.empty(shape, dtype=torch.float64, device="cuda")
    ws = torch.empty(L.fmb_workspace_bytes_2d_h2d(rows, cols), dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    st = torch.cuda.current_stream()
    g32 = (C.c_int32 * 2)(*goal)
    for rep in range(3):
        cd.fill_(float("nan"))
        _capi.check(L.fmb_solve2d_h2d_f64(h.data_ptr(), cd.data_ptr(), rows, cols, g32, T.data_ptr(), ws.data_ptr(), ws.numel(),
                                          st.cuda_stream, side.cuda_stream))
        _capi.check(L.fmb_finish(ws.data_ptr(), ws.numel(), st.cuda_stream, None))
        assert torch.equal(cd.cpu(), h)                                   # every band arrived
        got = T.cpu().numpy()
        if rep == 0:
            ref = O.computeTmap(c, goal)
        assert rel_err(got, ref) < TOL64
    res = eng.solve2d(cd, [goal])[0].cpu().numpy()
    assert np.array_equal(np.isfinite(res), np.isfinite(got)) and rel_err(got, res) < 1e-12
    pageable = np.ascontiguousarray(c)
    rc = L.fmb_solve2d_h2d_f64(pageable.ctypes.data, cd.data_ptr(), rows, cols, g32, T.data_ptr(), ws.data_ptr(), ws.numel(),
                               st.cuda_stream, side.cuda_stream)
    assert rc != 0 and b"page-locked" in L.fmb_last_error()
