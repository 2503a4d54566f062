"""GPU probe: per-kernel timing of the 2D cost-map builder (csrc/costmap2d.cuh) at several sizes.
   python tools/gpu_costmap_probe.py [n ...]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from planning_motion_planning_b200 import costmap, synth  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [1024, 4096]:
    res = 0.05
    Z = torch.from_numpy(synth.crater_dem(n, res, 1)).cuda()
    for _ in range(2):
        c = costmap.build_costmap_device(Z, res, n * res)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record()
        c = costmap.build_costmap_device(Z, res, n * res, sync=False)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(json.dumps({"n": n, "ms": float(np.median(ts)), "all_ms": ts, "cells_per_s": n * n / (np.median(ts) * 1e-3),
                      "obstacle_frac": float((c > 100).float().mean())}))
