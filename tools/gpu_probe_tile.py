"""One-tile probe (GPU): cycles per sweep step / per check pass of the sweep engine on a map that is ONE tile."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from planning_motion_planning_b200 import _capi, engine
for n, kind in ((32, "uniform"), (32, "random"), (64, "uniform"), (256, "uniform")):
    rng = np.random.default_rng(0)
    c = np.ones((n, n)) if kind == "uniform" else 1 + 4 * rng.random((n, n))
    c[0, :] = c[-1, :] = c[:, 0] = c[:, -1] = np.inf
    cd = torch.from_numpy(c).cuda()
    for cfg in (dict(engine2d=3), dict(engine2d=1)):
        _capi.set_options(**cfg)
        for rep in range(3):
            engine.solve2d(cd, [[1, 1]], sync=False)
            s = engine.finish()
        print(n, kind, cfg, "ms %.3f visits %d steps %d cyc/step %.0f check/round %.0f load/visit %.0f store/visit %.0f wait/visit %.0f evals/cell %.1f" % (
            s["solve_kernel_ms"], s["tile_visits"], s["steps"], s["cyc_relax"] / max(1, s["steps"]),
            s.get("cyc_check", 0) / max(1, s["steps"] / 64), s["cyc_load"] / s["tile_visits"], s["cyc_store"] / s["tile_visits"],
            s["cyc_wait"] / s["tile_visits"], s["evals"] / (n * n)))
