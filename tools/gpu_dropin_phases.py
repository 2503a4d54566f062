"""Where the time of the drop-in calls goes (GPU box): wraps the phases of FM.biComputeTmap /
FM3D.computeTmap with synchronising timers."""
import collections, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import FastMarching.FastMarching as FM
import FastMarching.FastMarching3D as FM3D
from FastMarching import _compat as C
from planning_motion_planning_b200 import engine, synth

acc = collections.defaultdict(float)
def wrap(mod, name, tag):
    f = getattr(mod, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); acc[tag] += 1e3 * (time.perf_counter() - t0)
        return r
    setattr(mod, name, g)
wrap(engine, "solve2d", "solve2d"); wrap(engine, "solve3d", "solve3d")
wrap(C, "bisolve2d", "bisolve2d"); wrap(C, "solve3d_until", "solve3d_until"); wrap(C, "to_device", "to_device"); wrap(C, "to_host", "to_host")

def run(label, f, reps=5):
    f(); acc.clear()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    tot = 1e3 * (time.perf_counter() - t0) / reps
    print(json.dumps({"call": label, "total_ms": round(tot, 2), "phases_ms": {k: round(v / reps, 2) for k, v in acc.items()}}))

g = np.load(os.path.join(ROOT, "tests", "golden", "planner_calls.npz"))
cost = np.asfortranarray(g["bi_cost"]); goal = [int(v) for v in g["bi_goal"]]; start = [int(v) for v in g["bi_start"]]
run("biComputeTmap captured 200^2", lambda: FM.biComputeTmap(cost, goal, start))
run("computeTmap3D captured 44x44x28", lambda: FM3D.computeTmap(g["c3"], np.uint32(g["g3"]), np.uint32(g["s3"])))
c = synth.mars_costmap(400, 3).T
goal = synth.free_cell_near(c.T, 320, 320)[::-1]; start = synth.free_cell_near(c.T, 80, 80)[::-1]
run("biComputeTmap 400^2", lambda: FM.biComputeTmap(c, goal, start))
c3, g3, s3 = synth.arm_volume((90, 90, 28), 0)
run("computeTmap3D 90x90x28", lambda: FM3D.computeTmap(c3, np.uint32(g3), np.uint32(s3)))
