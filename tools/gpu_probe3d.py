"""Tuning probe (GPU): 3D solve (+ path) on a planner-like arm-workspace volume."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from planning_motion_planning_b200 import _capi, engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
check = len(sys.argv) > 2 and sys.argv[2] == "check"
configs = sys.argv[3:] or [""]
cache = f"/tmp/fmb_vol_{n}.npz"
if os.path.exists(cache):
    z = np.load(cache); c, goal, start = z["c"], z["g"].tolist(), z["s"].tolist()
else:
    c, goal, start = synth.arm_volume((n, n, n), 0)
    np.savez(cache, c=c, g=np.array(goal), s=np.array(start))
cd = torch.from_numpy(c).cuda()
T = torch.empty((1, n, n, n), dtype=torch.float64, device="cuda")
base = _capi.get_options()
for cfg in configs:
    kv = {k: int(v) for k, v in (x.split("=") for x in cfg.split(",") if x)}
    _capi.set_options(**base)
    _capi.set_options(**kv)
    best = None
    for rep in range(3):
        engine.solve3d(cd, [goal], out=T, nq=1, sync=False)
        s = engine.finish()
        if best is None or s["solve_kernel_ms"] < best["solve_kernel_ms"]:
            best = s
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, cnt, st = engine.trace3d(T, [start], [goal], 0.5); e1.record(); torch.cuda.synchronize()
    cells = n ** 3
    tot = max(1, best["cyc_wait"] + best["cyc_load"] + best["cyc_relax"] + best["cyc_store"])
    print(json.dumps({"cfg": cfg, "n": n, "solve_ms": round(best["solve_kernel_ms"], 3), "init_ms": round(best["init_kernel_ms"], 3),
                      "trace_ms": round(e0.elapsed_time(e1), 3), "path_rows": int(cnt[0]), "path_status": int(st[0]),
                      "Mcells/s": round(cells / best["solve_kernel_ms"] / 1e3, 1), "evals/cell": round(best["evals"] / cells, 2),
                      "visits": best["tile_visits"], "steps/visit": round(best["steps"] / max(1, best["tile_visits"]), 1),
                      "cyc/step": round(best["cyc_relax"] / max(1, best["steps"]), 1),
                      "cyc_load/visit": round(best["cyc_load"] / max(1, best["tile_visits"])),
                      "visits/tile": round(best["tile_visits"] / (n ** 3 / 512), 2), "rounds/visit": round(best.get("rounds", 0) / max(1, best["tile_visits"]), 2),
                      "cont/visit": round(best.get("continuations", 0) / max(1, best["tile_visits"]), 2), "deferrals": best.get("deferrals", 0),
                      "cyc_wait/visit": round(best["cyc_wait"] / max(1, best["tile_visits"])), "cyc_relax/visit": round(best["cyc_relax"] / max(1, best["tile_visits"])),
                      "cyc_store/visit": round(best["cyc_store"] / max(1, best["tile_visits"])),
                      "phase%": {k: round(100 * best["cyc_" + k] / tot, 1) for k in ("wait", "load", "relax", "store")}}), flush=True)
if check:
    from oracle import oracle as O
    t0 = time.time(); ref = O.computeTmap3D(c, goal); dt = time.time() - t0
    Th = T[0].cpu().numpy(); fin = np.isfinite(ref)
    print(json.dumps({"oracle_s": round(dt, 2), "same_inf": bool(np.array_equal(np.isfinite(Th), fin)),
                      "max_rel": float(np.max(np.abs(Th[fin] - ref[fin]) / np.maximum(ref[fin], 1e-300)))}))
    p, pst = O.getPathGDM3D(ref, np.uint32(start), np.uint32(goal), 0.5, return_status=True)
    gp = out[0, :int(cnt[0])].cpu().numpy()
    print(json.dumps({"path_rows": [len(gp), len(p)], "path_dev": float(np.abs(gp - p).max()) if gp.shape == p.shape else None}))
