"""One-off GPU fuzz of the drop-in early-exit calls against the oracle (join node, accepted / narrow-band /
far pattern, values): python tools/gpu_fuzz_dropin.py [cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import FastMarching.FastMarching as FM
import FastMarching.FastMarching3D as FM3D
from conftest import plateau_map, rand_map
from oracle import oracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = checked = 0
paths_checked = [0]
worst = 0.0
t0 = time.time()
for it in range(N):
    kind = it % 4
    if kind == 3:       # 3D early exit
        n = int(rng.integers(12, 34))
        c = rand_map((n, n, n), int(rng.integers(0, 999))) if rng.random() < 0.5 else np.pad(np.full((n - 2,) * 3, float(rng.integers(1, 30))), 1, constant_values=np.inf)
        for _ in range(int(rng.integers(0, 3))):
            a = rng.integers(1, n - 1, 3); c[a[0], a[1], a[2]:a[2] + int(rng.integers(1, 8))] = np.inf
        free = np.argwhere(np.isfinite(c))
        gy, gx, gz = free[int(rng.integers(0, len(free)))]; sy, sx, sz = free[int(rng.integers(0, len(free)))]
        g, s = [int(gx), int(gy), int(gz)], [int(sx), int(sy), int(sz)]
        ref = O.computeTmap3D(c, g, s)
        got = FM3D.computeTmap(c, np.uint32(g), np.uint32(s))
        ok = np.array_equal(np.isfinite(got), np.isfinite(ref))
        if not ok:
            print("   3D pattern differs in", int((np.isfinite(got) != np.isfinite(ref)).sum()), "cells of", got.size,
                  "; cells tied with T[start] in the reference field:", int((O.computeTmap3D(c, g) == O.computeTmap3D(c, g)[s[1], s[0], s[2]]).sum()))
        f = np.isfinite(ref) & np.isfinite(got)
        e = float(np.max(np.abs(got[f] - ref[f]) / np.maximum(ref[f], 1.0))) if f.any() else 0.0
    else:
        if kind == 0:
            m = int(rng.integers(20, 200)); c = rand_map((m, m + int(rng.integers(0, 40))), int(rng.integers(0, 999)))
        elif kind == 1:
            c = plateau_map(int(rng.integers(2, 7)) * 32, int(rng.integers(0, 999)))
        else:
            m = int(rng.integers(20, 120)); c = np.pad(np.full((m, m), float(rng.integers(1, 9))), 1, constant_values=np.inf)
        for _ in range(int(rng.integers(0, 5))):
            y, x = int(rng.integers(1, c.shape[0] - 1)), int(rng.integers(1, c.shape[1] - 1)); c[y, x:x + int(rng.integers(1, 20))] = np.inf
        free = np.argwhere(np.isfinite(c))
        gy, gx = free[int(rng.integers(0, len(free)))]; sy, sx = free[int(rng.integers(0, len(free)))]
        g, s = [int(gx), int(gy)], [int(sx), int(sy)]
        if rng.random() < 0.3:
            c = np.asfortranarray(c)
        try:
            oTG, oTS, oj = O.biComputeTmap(np.ascontiguousarray(c), g, s)
        except NameError:
            try:
                FM.biComputeTmap(c, g, s); ok = False
            except NameError:
                ok = True
            e = 0.0
            checked += 1; bad += (not ok)
            continue
        TG, TS, j = FM.biComputeTmap(c, g, s)
        ok = np.array_equal(j, oj) and np.array_equal(np.isfinite(TG), np.isfinite(oTG)) and np.array_equal(np.isfinite(TS), np.isfinite(oTS))
        if not ok:
            print("   2D: join", list(j), "reference", list(oj), "; pattern differs in", int((np.isfinite(TG) != np.isfinite(oTG)).sum()), "+",
                  int((np.isfinite(TS) != np.isfinite(oTS)).sum()), "cells of", oTG.size, "; distinct cost values", np.unique(c[np.isfinite(c)]).size)
        e = 0.0
        if ok:
            for a, b in ((TG, oTG), (TS, oTS)):
                f = np.isfinite(b)
                e = max(e, float(np.max(np.abs(np.asarray(a)[f] - b[f]) / np.maximum(b[f], 1.0))))
            # single-front early exit (intended semantics of the reference's broken computeTmap)
            T1, o1 = FM.computeTmap(c, g, s), O.computeTmap(np.ascontiguousarray(c), g, s)
            if not np.array_equal(np.isfinite(T1), np.isfinite(o1)):
                ok = False
                print("   2D single front: pattern differs in", int((np.isfinite(T1) != np.isfinite(o1)).sum()), "cells of", o1.size,
                      "; distinct cost values", np.unique(c[np.isfinite(c)]).size)
            f = np.isfinite(o1) & np.isfinite(T1)
            e = max(e, float(np.max(np.abs(np.asarray(T1)[f] - o1[f]) / np.maximum(o1[f], 1.0))))
            # the two half paths from the join node over the partial fields, incl. the failure modes
            for Tm, oTm, endp in ((TG, oTG, g), (TS, oTS, s)):
                try:
                    po, so = O.getPathGDM(oTm, np.array(oj, dtype=np.float64), endp, 0.5, return_status=True)
                except Exception as ex:      # noqa: BLE001
                    po, so = None, type(ex).__name__
                try:
                    pg = FM.getPathGDM(Tm, j, endp, 0.5)
                    sg = 0
                except Exception as ex:      # noqa: BLE001
                    pg, sg = None, type(ex).__name__
                if isinstance(so, int) and so in (2, 3, 4):
                    ok = ok and sg == {2: "ValueError", 3: "IndexError", 4: "OverflowError"}[so]
                elif pg is None or po is None or pg.shape != po.shape:
                    ok = False
                else:
                    pe = float(np.max(np.abs(pg - po))) if len(po) else 0.0
                    ok = ok and pe < 1e-3
                    paths_checked[0] += 1
    checked += 1
    worst = max(worst, e)
    if not ok or e > 1e-9:
        bad += 1
        print("MISMATCH case", it, "kind", kind, "shape", c.shape, "goal", g, "start", s, "ok", ok, "err", e, flush=True)
print(f"cases {checked} bad {bad} worst rel err {worst:.2e} paths compared {paths_checked[0]} in {time.time() - t0:.1f} s")
