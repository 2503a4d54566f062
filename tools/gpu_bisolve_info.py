"""fmb_bisolve2d_f64 on the bench map: the info block (k, join, statuses, sparse-replay counters) and the wall time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from FastMarching import _compat as C
from bench import make_map
from planning_motion_planning_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c = make_map(n, "mars")
goal = synth.free_cell_near(c, n // 4, n // 4); start = synth.free_cell_near(c, 3 * n // 4, 3 * n // 4)
dev = C.device()
cd = C.to_device(c, dev)
for rep in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    TG, TS, info, ws = C.bisolve2d(cd, goal, start, False)
    C.finish(ws, dev)
    t1 = time.perf_counter()
    print(f"rep {rep}: {1e3 * (t1 - t0):.2f} ms info {info.tolist()}", flush=True)
