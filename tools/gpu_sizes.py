"""Solve (+ path) times of the default engines over the BASELINE sizes, for the tables of DESIGN.md."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import make_map
from planning_motion_planning_b200 import engine, synth
for n in (512, 1024, 2048, 4096, 8192):
    c = make_map(n, "mars")
    goal = synth.free_cell_near(c, n // 4, n // 4); start = synth.free_cell_near(c, 3 * n // 4, 3 * n // 4)
    cd = torch.from_numpy(c).cuda()
    T = torch.empty((1, n, n), dtype=torch.float64, device="cuda")
    best = None
    for rep in range(3):
        engine.solve2d(cd, [goal], out=T, nq=1, sync=False)
        s = engine.finish()
        best = s if best is None or s["solve_kernel_ms"] < best["solve_kernel_ms"] else best
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, cnt, st = engine.trace2d(T, [start], [goal], 0.5); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"n": n, "solve_ms": round(best["solve_kernel_ms"], 3), "Mcells/s": round(n * n / best["solve_kernel_ms"] / 1e3),
                      "frac_of_hbm_roofline": round(16 * n * n / (best["solve_kernel_ms"] * 1e-3) / 6538.6e9, 5),
                      "visits/tile": round(best["tile_visits"] / ((n / 32) ** 2), 2), "evals/cell": round(best["evals"] / n / n, 1),
                      "trace_ms": round(e0.elapsed_time(e1), 3), "path_rows": int(cnt[0])}), flush=True)
    del cd, T
