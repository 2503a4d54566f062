"""Wall-clock of the five FastMarching calls the unmodified planner makes, through the drop-in
package (numpy in / numpy out, includes H2D/D2H and the early-exit emulation), on the call
arguments captured from the reference planner (tests/golden/planner_calls.npz) and on a
planner-scale synthetic case (400^2 map, 90x90x28 volume)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import FastMarching.FastMarching as FM
import FastMarching.FastMarching3D as FM3D
from planning_motion_planning_b200 import synth

def timed(f, reps=5):
    f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return r, 1e3 * float(np.median(ts))

g = np.load(os.path.join(ROOT, "tests", "golden", "planner_calls.npz"))
cost = np.asfortranarray(g["bi_cost"]); goal = [int(v) for v in g["bi_goal"]]; start = [int(v) for v in g["bi_start"]]
(TG, TS, j), t_bi = timed(lambda: FM.biComputeTmap(cost, goal, start))
_, t_pg = timed(lambda: FM.getPathGDM(TG, j, goal, 0.5))
_, t_ps = timed(lambda: FM.getPathGDM(TS, j, start, 0.5))
T3, t_3d = timed(lambda: FM3D.computeTmap(g["c3"], np.uint32(g["g3"]), np.uint32(g["s3"])))
_, t_p3 = timed(lambda: FM3D.getPathGDM(T3, np.uint32(g["path3d_init"]), np.uint32(g["path3d_end"]), 0.5))
print(json.dumps({"case": "captured planner calls (200^2 map, 44x44x28 volume)", "ms": {"biComputeTmap": round(t_bi, 2), "getPathGDM_G": round(t_pg, 2),
      "getPathGDM_S": round(t_ps, 2), "computeTmap3D": round(t_3d, 2), "getPathGDM3D": round(t_p3, 2), "total": round(t_bi + t_pg + t_ps + t_3d + t_p3, 2)}}))
c = synth.mars_costmap(400, 3).T          # F-ordered view like the planner's cMap.T
goal = synth.free_cell_near(c.T, 320, 320)[::-1]; start = synth.free_cell_near(c.T, 80, 80)[::-1]
(TG, TS, j), t_bi = timed(lambda: FM.biComputeTmap(c, goal, start))
_, t_pg = timed(lambda: FM.getPathGDM(TG, j, goal, 0.5))
_, t_ps = timed(lambda: FM.getPathGDM(TS, j, start, 0.5))
c3, g3, s3 = synth.arm_volume((90, 90, 28), 0)
T3, t_3d = timed(lambda: FM3D.computeTmap(c3, np.uint32(g3), np.uint32(s3)))
_, t_p3 = timed(lambda: FM3D.getPathGDM(T3, np.uint32(s3), np.uint32(g3), 0.5))
print(json.dumps({"case": "planner scale (400^2 map, 90x90x28 volume); reference measured 4.6 s + 1.0 s + 3.3 s (SURVEY 3.3)", "ms": {"biComputeTmap": round(t_bi, 2),
      "getPathGDM_G": round(t_pg, 2), "getPathGDM_S": round(t_ps, 2), "computeTmap3D": round(t_3d, 2), "getPathGDM3D": round(t_p3, 2),
      "total": round(t_bi + t_pg + t_ps + t_3d + t_p3, 2)}}))
