"""One-off GPU fuzz of the two builders and of the windowed solve order against their oracles.
   python tools/gpu_fuzz_builders.py [cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import costmap_oracle as CO, costvol as CV, oracle as O
from planning_motion_planning_b200 import costmap, costvolume as CVP, engine, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
t0 = time.time()
for it in range(N):
    # ---- 2D cost map
    n = int(rng.integers(60, 420)); res = float(rng.choice([0.03, 0.05, 0.08, 0.1, 0.15]))
    Z = synth.crater_dem(n, res, int(rng.integers(0, 999)), craters=int(rng.integers(0, 4)), rocks=int(rng.integers(0, 12)))
    Z = Z + rng.normal(0, float(rng.choice([0.0, 0.05, 0.1, 0.2])) * res, Z.shape); Z -= Z.min()
    try:
        with np.errstate(all="ignore"):
            c, st = CO.costmap2d(Z, res, n * res, stages=True)
        want_err = False
    except ValueError:
        want_err = True
    try:
        cost, dv = costmap.build_costmap_device(torch.from_numpy(Z).cuda(), res, n * res, stages=True)
        got_err = False
    except ValueError:
        got_err = True
    ok = want_err == got_err
    if ok and not want_err:
        fin = np.isfinite(c); g = cost.cpu().numpy().T
        ok = (np.array_equal(dv["raw"].cpu().numpy(), st["raw"]) and np.array_equal(dv["obst"].cpu().numpy(), st["obst"].astype(np.uint8))
              and np.array_equal(dv["pre_blur"].cpu().numpy(), st["pre_blur"].T) and np.array_equal(np.isfinite(g), fin)
              and float(np.max(np.abs(g[fin] - c[fin]) / c[fin])) < 1e-12)
    if not ok:
        bad += 1; print("COSTMAP MISMATCH", it, n, res, want_err, got_err, flush=True)
    # ---- 3D cost volume
    sX = sY = int(rng.integers(24, 90)); sZ = int(rng.integers(20, 60)); m = int(rng.integers(2, 40))
    rx = ry = float(rng.uniform(0.008, 0.03)); rz = 0.02
    Zs = 0.1 + 0.3 * rng.random((sX, sY))
    s_ = np.linspace(0, 1, m)
    path = np.stack([(0.1 + 0.8 * s_) * sX * rx, (0.2 + 0.6 * s_ ** 2) * sY * ry, 0.3 * sZ * rz + 0.2 * np.sin(3 * s_)], axis=1)
    head = np.stack([0.3 * rng.normal(size=m), 0.3 * rng.normal(size=m), rng.uniform(-3, 3) + 1.5 * s_], axis=1)
    fin3, ini3 = np.uint32(rng.integers(1, [sX - 1, sY - 1, sZ - 1])), np.uint32(rng.integers(1, [sX - 1, sY - 1, sZ - 1]))
    rad = (float(rng.uniform(0.2, 0.6)), 0.2673, 0.1105)
    xm, ym = rx * int(rng.integers(0, sX)), ry * int(rng.integers(0, sY))
    want = CV.GetObstMap(Zs, rx, ry, rz, sX, sY, sZ, np.zeros((sX, sY)), xm, ym) * CV.TunnelCost(*rad, path, sX, sY, sZ, rx, ry, rz, head, fin3, ini3)
    got = CVP.build_cost_volume(Zs, rx, ry, rz, sX, sY, sZ, xm, ym, *rad, path, head, fin3, ini3)
    if not np.array_equal(got, want):
        bad += 1; print("COSTVOLUME MISMATCH", it, (sX, sY, sZ), m, int((got != want).sum()), "cells", flush=True)
    # ---- windowed solve order forced on a mid-size map
    if it % 4 == 0:
        os.environ["FMB_WINDOWED"] = "1"; os.environ["FMB_WINDOW"] = str(int(rng.integers(1, 6)))
        nn = int(rng.integers(300, 1200))
        cm = synth.mars_costmap(nn, int(rng.integers(0, 99))) if rng.random() < 0.5 else synth.random_costmap((nn, nn), int(rng.integers(0, 99)))
        gl = list(synth.free_cell_near(cm, int(rng.integers(10, nn - 10)), int(rng.integers(10, nn - 10))))
        T = engine.solve2d(torch.from_numpy(cm).cuda(), [gl])[0].cpu().numpy()
        ref = O.computeTmap(cm, gl); f = np.isfinite(ref)
        if not (np.array_equal(np.isfinite(T), f) and float(np.max(np.abs(T[f] - ref[f]) / np.maximum(ref[f], 1.0))) < 1e-9):
            bad += 1; print("WINDOWED SOLVE MISMATCH", it, nn, flush=True)
        os.environ.pop("FMB_WINDOWED"); os.environ.pop("FMB_WINDOW")
print(f"cases {N} bad {bad} in {time.time() - t0:.1f} s")
