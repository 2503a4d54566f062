"""One-off GPU fuzz of the solver / tracer entry points against the oracle: odd shapes, batches (shared and
per-query maps, both work orders), 3D volumes, fp32 variants, traced paths with their failure statuses.
   python tools/gpu_fuzz_solvers.py [cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import plateau_map, rand_map
from oracle import oracle as O
from planning_motion_planning_b200 import engine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
t0 = time.time()

def rel(a, b):
    f = np.isfinite(b)
    if not np.array_equal(np.isfinite(a), f):
        return np.inf
    return float(np.max(np.abs(a[f] - b[f]) / np.maximum(b[f], 1.0))) if f.any() else 0.0

for it in range(N):
    # ---- 2D batch
    rows, cols = int(rng.integers(3, 300)), int(rng.integers(3, 300))
    nq = int(rng.choice([1, 2, 3, 9, 17, 40]))
    shared = bool(rng.random() < 0.5)
    maps = [rand_map((rows, cols), int(rng.integers(0, 9999))) if rng.random() < 0.7 or min(rows, cols) < 40 else
            np.pad(np.full((rows - 2, cols - 2), 2.0), 1, constant_values=np.inf) for _ in range(1 if shared else nq)]
    for m in maps:
        for _ in range(int(rng.integers(0, 4))):
            y, x = int(rng.integers(0, rows)), int(rng.integers(0, cols)); m[y, x:x + int(rng.integers(1, 30))] = np.inf
    seeds = []
    for q in range(nq):
        m = maps[0 if shared else q]
        free = np.argwhere(np.isfinite(m))
        if len(free) == 0:
            seeds.append([0, 0]); continue
        y, x = free[int(rng.integers(0, len(free)))]; seeds.append([int(x), int(y)])
    os.environ["FMB_BEST_FIRST"] = str(int(rng.integers(0, 2)))
    f32 = rng.random() < 0.25
    cd = torch.from_numpy(maps[0] if shared else np.stack(maps)).cuda()
    if f32:
        cd = cd.float()
    T = engine.solve2d(cd, seeds, nq=nq).cpu().numpy().astype(np.float64)
    for q in range(nq):
        m = maps[0 if shared else q]
        if not np.isfinite(m[seeds[q][1], seeds[q][0]]):
            continue
        ref = O.computeTmap(m.astype(np.float32).astype(np.float64) if f32 else m, seeds[q])
        e = rel(T[q], ref)
        if e > (1e-4 if f32 else 1e-9):
            bad += 1; print("SOLVE2D MISMATCH", it, (rows, cols), nq, shared, f32, q, e, flush=True); break
    os.environ.pop("FMB_BEST_FIRST")
    # ---- traced paths on the first field (fp64 only)
    if not f32:
        m = maps[0]; ref = O.computeTmap(m, seeds[0])
        free = np.argwhere(np.isfinite(ref))
        if len(free) > 4 and rows > 4 and cols > 4:
            pts = free[rng.integers(0, len(free), 6)]
            init = np.stack([pts[:, 1], pts[:, 0]], axis=1).astype(np.float64) + rng.choice([0.0, 0.25, 0.5], (6, 2))
            end = np.tile(np.array(seeds[0], dtype=np.float64), (6, 1))
            out, cnt, stt = engine.trace2d(torch.from_numpy(ref).cuda(), init, end, 0.5)
            out, cnt, stt = out.cpu().numpy(), cnt.cpu().numpy(), stt.cpu().numpy()
            for p in range(6):
                po, so = O.getPathGDM(ref, init[p], seeds[0], 0.5, return_status=True)
                if so != int(stt[p]) or (so in (0, 1) and (len(po) != int(cnt[p]) or (len(po) and np.max(np.abs(out[p, :cnt[p]] - po)) > 1e-3))):
                    bad += 1; print("TRACE2D MISMATCH", it, (rows, cols), init[p], so, int(stt[p]), len(po), int(cnt[p]), flush=True); break
    # ---- 3D
    if it % 2 == 0:
        sh = tuple(int(v) for v in rng.integers(3, 40, 3))
        c3 = rand_map(sh, int(rng.integers(0, 9999)))
        free = np.argwhere(np.isfinite(c3))
        if len(free):
            y, x, z = free[int(rng.integers(0, len(free)))]
            g3 = [int(x), int(y), int(z)]
            os.environ["FMB_TZ3D"] = str(int(rng.choice([16, 32])))
            T3 = engine.solve3d(torch.from_numpy(c3).cuda(), [g3], nq=1)[0].cpu().numpy()
            os.environ.pop("FMB_TZ3D")
            e = rel(T3, O.computeTmap3D(c3, g3, [-1, -1, -1]))
            if e > 1e-9:
                bad += 1; print("SOLVE3D MISMATCH", it, sh, g3, e, flush=True)
print(f"cases {N} bad {bad} in {time.time() - t0:.1f} s")
