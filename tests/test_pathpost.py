"""Path post-processing and the batch entry for the native host (rank 4).

CPU: oracle/pathpost_oracle.py against tests/golden/pathpost.npz -- frozen from the unmodified planner main() (rover
path) and from the reference's own source lines 1641-1671 executed unmodified (arm path), oracle/gen_golden.py pathpost.
GPU (-m gpu): csrc/pathpost.cuh and fmb_plan_batch2d_host through the C ABI against the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rand_map


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "pathpost.npz"), allow_pickle=True)


def _rows_in_order(sub, full):
    j = 0
    for row in sub:
        while j < len(full) and not np.array_equal(full[j], row[:2]):
            j += 1
        if j == len(full):
            return False
        j += 1
    return True


def test_oracle_stitch_matches_planner_roverpath(gold):
    from oracle import pathpost_oracle as PO
    st = PO.stitch_rover_path(gold["pathS"], gold["pathG"], float(gold["resolution"]))
    assert st.shape == (len(gold["pathS"]) + len(gold["pathG"]) - 1, 2)
    assert _rows_in_order(gold["roverPath_trimmed"], st)          # bit for bit, the rows the planner kept
    assert len(gold["roverPath_trimmed"]) >= len(st) - 16


def test_oracle_post3d_matches_reference_lines(gold):
    from oracle import pathpost_oracle as PO
    for k in range(int(gold["npost"])):
        ref = gold[f"post{k}_resized"]
        got = PO.smooth_resample_arm(gold[f"post{k}_path"], gold[f"post{k}_res3"], gold[f"post{k}_off3"], gold[f"post{k}_last3"], len(ref))
        assert np.array_equal(got, ref), k


def test_oracle_post3d_short_path_raises_like_the_planner(gold):
    from oracle import pathpost_oracle as PO
    p = gold["planner_path3d"]
    assert len(p) < 11 and "ValueError" in str(gold["status"])
    with pytest.raises(ValueError):
        PO.smooth_resample_arm(p, [0.03, 0.03, 0.02], [0, 0, 0], [1, 2, 3], 20)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_stitch2d_device_vs_golden(gold):
    from oracle import pathpost_oracle as PO
    from planning_motion_planning_b200 import pathpost
    st = pathpost.stitch_rover_path(gold["pathS"], gold["pathG"], float(gold["resolution"]))
    ref = PO.stitch_rover_path(gold["pathS"], gold["pathG"], float(gold["resolution"]))
    assert np.array_equal(st, ref)                                  # one add + one multiply per number: bit-exact
    assert _rows_in_order(gold["roverPath_trimmed"], st)


@pytest.mark.gpu
def test_post3d_device_vs_golden(gold):
    from planning_motion_planning_b200 import pathpost
    for k in range(int(gold["npost"])):
        ref = gold[f"post{k}_resized"]
        got = pathpost.smooth_resample_arm(gold[f"post{k}_path"], gold[f"post{k}_res3"], gold[f"post{k}_off3"], gold[f"post{k}_last3"], len(ref))
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))), k
    with pytest.raises(ValueError):
        pathpost.smooth_resample_arm(gold["planner_path3d"], [0.03, 0.03, 0.02], [0, 0, 0], [1, 2, 3], 20)


@pytest.mark.gpu
def test_pathpost_batched_from_tracer_outputs(gold):
    """The batch forms read the tracer's own slabs and device-side counts: 2D bi-solve halves stitched, 3D paths
    smoothed and resampled, each against the oracle on the same traced rows."""
    import torch
    from oracle import oracle as O
    from oracle import pathpost_oracle as PO
    from planning_motion_planning_b200 import engine, pathpost
    c = rand_map((160, 160), 3)
    goals = [[20, 25], [130, 40], [80, 140]]
    starts = [[140, 130], [15, 150], [30, 12]]
    joins = [[80, 80], [70, 90], [60, 70]]
    cd = torch.from_numpy(c).cuda()
    TG = engine.solve2d(cd, goals, nq=3)
    TS = engine.solve2d(cd, starts, nq=3)
    pG, cG, sG = engine.trace2d(TG, joins, goals, 0.5)
    pS, cS, sS = engine.trace2d(TS, joins, starts, 0.5)
    out, cnt = pathpost.stitch_rover_paths_device(pS, cS, pG, cG, 0.05)
    for q in range(3):
        ref = PO.stitch_rover_path(pS[q, :int(cS[q])].cpu().numpy(), pG[q, :int(cG[q])].cpu().numpy(), 0.05)
        assert int(cnt[q]) == len(ref) and np.array_equal(out[q, :len(ref)].cpu().numpy(), ref)
    # 3D: low costs -> short steps -> enough rows for the 11-tap window
    c3 = 0.3 + 0.2 * np.random.default_rng(5).random((40, 40, 40))
    for d in range(3):
        for e in (0, -1):
            sl = [slice(None)] * 3; sl[d] = e; c3[tuple(sl)] = np.inf
    g3, s3 = [[5, 6, 7], [30, 8, 20]], [[33, 31, 30], [6, 33, 9]]
    T3 = engine.solve3d(torch.from_numpy(c3).cuda(), g3, nq=2)
    p3, c3n, s3n = engine.trace3d(T3, s3, g3, 0.5)
    last = torch.tensor([[1.0, 2.0, 3.0], [-4.0, 0.5, 9.0]], dtype=torch.float64, device="cuda")
    res3, off3 = [0.031, 0.029, 0.02], [-1.5, 2.25, 0.125]
    out3, st3 = pathpost.smooth_resample_arm_device(p3, c3n, res3, off3, last, 50)
    for q in range(2):
        n = int(c3n[q])
        assert n >= 11 and int(st3[q]) == 0
        ref = PO.smooth_resample_arm(p3[q, :n].cpu().numpy(), res3, off3, last[q].cpu().numpy(), 50)
        assert np.max(np.abs(out3[q].cpu().numpy() - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.gpu
def test_plan_batch_host_entry_64_queries_vs_oracle():
    """fmb_plan_batch2d_host: host pointers in, owned result out, >= 64 queries, each checked against the oracle
    (full field from the goal + path from the start, metres = resolution * (cell + 1))."""
    from oracle import oracle as O
    from planning_motion_planning_b200 import plan
    n = 96
    c = rand_map((n, n), 11)
    c[30:34, 10:70] = np.inf                                     # a wall the paths go around
    rng = np.random.default_rng(2)
    free = np.argwhere(np.isfinite(c))
    pick = free[rng.choice(len(free), size=(72, 2), replace=False)]
    goals = [[int(p[0][1]), int(p[0][0])] for p in pick]
    starts = [[int(p[1][1]), int(p[1][0])] for p in pick]
    res = 0.05
    paths, status, info = plan.plan_batch(c, goals, starts, tau=0.5, resolution=res)
    assert len(paths) == 72 and info["solve_ms"] > 0
    # strided input (a view with a larger pitch) goes through cudaMemcpy2D
    big = np.full((n, n + 7), np.inf); big[:, :n] = c
    paths_v, status_v, _ = plan.plan_batch(big[:, :n], goals[:5], starts[:5], tau=0.5, resolution=res)
    worst = 0.0
    for q in range(72):
        T = O.computeTmap(c, goals[q])
        ref, rst = O.getPathGDM(T, np.array(starts[q], dtype=np.float64), np.array(goals[q], dtype=np.float64), 0.5, return_status=True)
        assert status[q] == rst and paths[q].shape == ref.shape, q      # incl. the reference's own failure modes
        worst = max(worst, float(np.max(np.abs(paths[q] / res - 1 - ref))))
        if q < 5:
            assert np.array_equal(paths_v[q], paths[q])
    assert worst < 1e-3 and int((status == 0).sum()) >= 48
    with pytest.raises(Exception):
        plan.plan_batch(c, [[n + 5, 3]], [[4, 4]])
