"""GPU fuzz of the 3D early exit (FastMarching3D.computeTmap through the drop-in) against the oracle:
accepted / narrow-band / far pattern and values.   python tools/gpu_fuzz_3d.py [cases] [seed] [exact 0|1]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import FastMarching.FastMarching3D as FM3D
from FastMarching import _compat
from conftest import rand_map
from oracle import oracle as O


def case(rng):
    n = int(rng.integers(12, 34))
    uniform = rng.random() < 0.7
    c = np.pad(np.full((n - 2,) * 3, float(rng.integers(1, 30))), 1, constant_values=np.inf) if uniform else rand_map((n, n, n), int(rng.integers(0, 999)))
    for _ in range(int(rng.integers(0, 3))):
        a = rng.integers(1, n - 1, 3); c[a[0], a[1], a[2]:a[2] + int(rng.integers(1, 8))] = np.inf
    free = np.argwhere(np.isfinite(c))
    gy, gx, gz = free[int(rng.integers(0, len(free)))]; sy, sx, sz = free[int(rng.integers(0, len(free)))]
    return c, [int(gx), int(gy), int(gz)], [int(sx), int(sy), int(sz)], uniform


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    EXACT_3D = True   # (the drop-in always solves in the exact arithmetic now: fmb_solve3d_until_f64)
    bad = 0; worst = 0.0; t0 = time.time()
    for it in range(N):
        c, g, s, uniform = case(rng)
        ref = O.computeTmap3D(c, g, s)
        got = FM3D.computeTmap(c, np.uint32(g), np.uint32(s))
        ok = np.array_equal(np.isfinite(got), np.isfinite(ref))
        f = np.isfinite(ref) & np.isfinite(got)
        e = float(np.max(np.abs(got[f] - ref[f]) / np.maximum(ref[f], 1.0))) if f.any() else 0.0
        worst = max(worst, e)
        if not ok or e > 1e-9:
            bad += 1
            print("MISMATCH case", it, "uniform" if uniform else "random", c.shape, g, s, "pattern differs in",
                  int((np.isfinite(got) != np.isfinite(ref)).sum()), "cells; err", e, flush=True)
    print(f"exact={EXACT_3D} cases {N} bad {bad} worst rel err {worst:.2e} in {time.time() - t0:.1f} s")
