"""3D exact arithmetic: from scratch vs fast solve + polish pass, per volume size."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from planning_motion_planning_b200 import engine, synth
for shape in ((44, 44, 28), (90, 90, 28), (128, 128, 64), (160, 160, 160), (256, 256, 256)):
    c, goal, start = synth.arm_volume(shape, 0)
    cd = torch.from_numpy(c).cuda()
    res = {}
    ref = None
    for mode in (False, True, "polish"):
        ts = []
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            T = engine.solve3d(cd, [goal], nq=1, exact=mode)
            torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
        res[str(mode)] = round(min(ts), 2)
        if mode is True: ref = T.clone()
        if mode == "polish": res["polish==exact"] = bool(torch.equal(ref, T))
    print(json.dumps({"shape": shape, "ms": res}), flush=True)
