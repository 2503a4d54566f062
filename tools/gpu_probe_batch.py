"""Tuning probe (GPU): batched goal queries on one map under several fmb_options settings (k=v,k=v ...)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import make_map
from planning_motion_planning_b200 import _capi, engine
n = int(sys.argv[1]); Q = int(sys.argv[2]); kind = sys.argv[3]
configs = sys.argv[4:] or [""]
c = make_map(n, kind, seed=1)
rng = np.random.default_rng(100)
ok = np.argwhere(np.isfinite(c) & (c <= 2.0))
seeds = torch.tensor(ok[rng.integers(0, len(ok), size=Q)][:, ::-1].copy(), dtype=torch.int32, device="cuda")
cd = torch.from_numpy(c).cuda()
T = torch.empty((Q, n, n), dtype=torch.float64, device="cuda")
ref = None
base = _capi.get_options()
for cfg in configs:
    kv = {k: int(v) for k, v in (x.split("=") for x in cfg.split(",") if x)}
    _capi.set_options(**base)
    _capi.set_options(**kv)
    best = None
    for rep in range(3):
        engine.solve2d(cd, seeds, out=T, nq=Q, sync=False)
        s = engine.finish()
        if best is None or s["solve_kernel_ms"] < best["solve_kernel_ms"]:
            best = s
    if ref is None:
        ref = T.clone(); same = True
    else:
        fin = torch.isfinite(ref)
        same = bool(torch.equal(fin, torch.isfinite(T))) and float(((ref - T).abs() / ref.clamp_min(1e-300))[fin].max()) < 1e-12
    cells = Q * n * n
    tot = max(1, best["cyc_wait"] + best["cyc_load"] + best["cyc_relax"] + best["cyc_store"])
    print(json.dumps({"cfg": cfg, "ms": round(best["solve_kernel_ms"], 3), "queries/s": round(Q / best["solve_kernel_ms"] * 1e3),
                      "Gcells/s": round(cells / best["solve_kernel_ms"] / 1e6, 2), "evals/cell": round(best["evals"] / cells, 2),
                      "visits/tile": round(best["tile_visits"] / (Q * (n / 32) ** 2), 2),
                      "cyc/step": round(best["cyc_relax"] / max(1, best["steps"]), 1),
                      "phase%": {k: round(100 * best["cyc_" + k] / tot, 1) for k in ("wait", "load", "relax", "store")},
                      "consistent": same}), flush=True)
