"""Where the time of the plugin call at the headline size goes (GPU box): FM.computeTmap + FM.getPathGDM on a pageable
4096^2 NumPy map; the phases are timed INSIDE the calls (wrappers with synchronising timers), the total without them."""
import collections, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import FastMarching.FastMarching as FM
from FastMarching import _compat as C
from planning_motion_planning_b200 import engine, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c = synth.mars_costmap(n, 0)
goal = synth.free_cell_near(c, n // 8, n // 8)
start = synth.free_cell_near(c, n - n // 8, n - n // 8)
dev = C.device()
acc = collections.defaultdict(float)
orig = {}
def wrap(mod, name, tag):
    f = getattr(mod, name); orig[(mod, name)] = f
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); acc[tag] += 1e3 * (time.perf_counter() - t0)
        return r
    setattr(mod, name, g)
def unwrap():
    for (mod, name), f in orig.items(): setattr(mod, name, f)
    orig.clear()

def call():
    Tn = FM.computeTmap(c, goal, [-1, -1])
    return FM.getPathGDM(Tn, np.array(start, dtype=np.float64), goal, 0.5)

def total(reps=6):
    call(); call(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): p = call()
    torch.cuda.synchronize()
    return round(1e3 * (time.perf_counter() - t0) / reps, 2), len(p)

out = {"n": n, "threads": torch.get_num_threads(), "cpus": len(os.sched_getaffinity(0))}
out["plugin_call_ms"], out["path_rows"] = total()
wrap(C, "to_device", "to_device"); wrap(C, "to_host", "to_host"); wrap(C, "solve2d_until", "solve2d_until")
wrap(engine, "trace2d", "trace2d"); wrap(FM, "computeTmap", "computeTmap"); wrap(FM, "getPathGDM", "getPathGDM")
call(); call(); acc.clear()
for _ in range(6): call()
out["phases_ms"] = {k: round(v / 6, 2) for k, v in acc.items()}
unwrap()
for var in sys.argv[2:]:                      # e.g. stage=2:4  (MiElems per staging buffer : buffers)
    k, v = var.split("=")
    if k == "stage":
        a, b = v.split(":")
        C._STAGE_ELEMS, C._STAGE_BUFS = int(float(a) * (1 << 20)), int(b); C._STAGE.clear()
    out[var] = total()[0]
print(json.dumps(out))
