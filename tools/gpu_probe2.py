"""Tuning probe (GPU): one 2D map solved under several fmb_options settings; prints time, work counters and
per-phase cycles.  Not part of the product or the tests.
    python tools/gpu_probe2.py [size] [map] "k=v,k=v" "k=v" ...      (keys = fields of fmb_options)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bench import make_map
from planning_motion_planning_b200 import _capi, engine, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kind = sys.argv[2] if len(sys.argv) > 2 else "mars"
configs = sys.argv[3:] or [""]
c = make_map(n, kind)
goal = synth.free_cell_near(c, n // 4, n // 4)
cd = torch.from_numpy(c).cuda()
T = torch.empty((1, n, n), dtype=torch.float64, device="cuda")
ref = None
base = _capi.get_options()
for cfg in configs:
    kv = {k: int(v) for k, v in (x.split("=") for x in cfg.split(",") if x)}
    _capi.set_options(**base)
    _capi.set_options(**kv)
    best = None
    for rep in range(3):
        engine.solve2d(cd, [goal], out=T, nq=1, sync=False)
        s = engine.finish()
        if best is None or s["solve_kernel_ms"] < best["solve_kernel_ms"]:
            best = s
    if ref is None:
        ref = T.clone()
        same = True
    else:
        same = bool(torch.equal(torch.isfinite(ref), torch.isfinite(T))) and float(((ref - T).abs() / ref.clamp_min(1e-300))[torch.isfinite(ref)].max()) < 1e-12
    cells = n * n
    tot = max(1, best["cyc_wait"] + best["cyc_load"] + best["cyc_relax"] + best["cyc_store"])
    print(json.dumps({"cfg": cfg, "ms": round(best["solve_kernel_ms"], 3), "evals/cell": round(best["evals"] / cells, 2),
                      "visits/tile": round(best["tile_visits"] / ((n / 32) ** 2), 2), "steps/visit": round(best["steps"] / max(1, best["tile_visits"]), 1),
                      "cyc/step(relax)": round(best["cyc_relax"] / max(1, best["steps"]), 1),
                      "noop%": round(100 * best.get("noop_visits", 0) / max(1, best["tile_visits"]), 1), "rounds/visit": round(best.get("rounds", 0) / max(1, best["tile_visits"]), 2), "cont/visit": round(best.get("continuations", 0) / max(1, best["tile_visits"]), 2),
                      "cyc_check/round": round(best.get("cyc_check", 0) / max(1, best["steps"] / 64)),
                      "cyc_wait/visit": round(best["cyc_wait"] / max(1, best["tile_visits"])),
                      "cyc_load/visit": round(best["cyc_load"] / max(1, best["tile_visits"])),
                      "cyc_store/visit": round(best["cyc_store"] / max(1, best["tile_visits"])),
                      "phase%": {k: round(100 * best["cyc_" + k] / tot, 1) for k in ("wait", "load", "relax", "store")},
                      "deferrals": best.get("deferrals", 0), "consistent": same}), flush=True)
