"""biComputeTmap + the two half paths through the drop-in on the bench map (NumPy in / out), wall-clock phases.
    python tools/gpu_bicompute_once.py [size] [reps]      (run under `ncu --metrics gpu__time_duration.sum` for the launch list)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import FastMarching.FastMarching as FM
from bench import make_map
from planning_motion_planning_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = make_map(n, "mars")
goal = synth.free_cell_near(c, n // 4, n // 4)
start = synth.free_cell_near(c, 3 * n // 4, 3 * n // 4)
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    TG, TS, j = FM.biComputeTmap(c, goal, start)
    t1 = time.perf_counter()
    pG = FM.getPathGDM(TG, j, goal, 0.5)
    pS = FM.getPathGDM(TS, j, start, 0.5)
    t2 = time.perf_counter()
    print(f"rep {r}: biComputeTmap {1e3 * (t1 - t0):.1f} ms, two paths {1e3 * (t2 - t1):.1f} ms, join {j.tolist()}, rows {len(pG)} + {len(pS)}", flush=True)
