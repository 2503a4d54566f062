"""Tuning probe (GPU): the 3D solve in the reference's own arithmetic (fmb_solve3d_exact_f64) on the bench volume."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from planning_motion_planning_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
c, goal, start = synth.arm_volume((n, n, n), 0)
cd = torch.from_numpy(c).cuda()
T = torch.empty((1, n, n, n), dtype=torch.float64, device="cuda")
for mode in (False, True):
    best = None
    for rep in range(2):
        engine.solve3d(cd, [goal], out=T, nq=1, sync=False, exact=mode)
        s = engine.finish()
        best = s if best is None or s["solve_kernel_ms"] < best["solve_kernel_ms"] else best
    print(json.dumps({"exact": mode, "solve_ms": round(best["solve_kernel_ms"], 2), "evals/cell": round(best["evals"] / n ** 3, 1),
                      "visits/tile": round(best["tile_visits"] / (n ** 3 / 512), 2),
                      "cyc/step": round(best["cyc_relax"] / max(1, best["steps"]))}), flush=True)
