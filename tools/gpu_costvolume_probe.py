"""GPU probe: kernel times of the 3D cost-volume builder (csrc/costvolume.cuh).  python tools/gpu_costvolume_probe.py [size]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from planning_motion_planning_b200 import costvolume as CVP  # noqa: E402

sv = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rxy, rz, mpose = 2.0 / sv, 0.02, 200
rng = np.random.default_rng(5)
Zv = 0.2 + 0.1 * rng.random((sv, sv))
sp = np.linspace(0, 1, mpose)
path = np.stack([(0.2 + 0.6 * sp) * sv * rxy, (0.3 + 0.4 * sp ** 2) * sv * rxy, 0.4 * sv * rz + 0.3 * np.sin(3 * sp)], axis=1)
head = np.stack([0.2 * np.sin(5 * sp), 0.15 * np.cos(4 * sp), 0.3 + 1.2 * sp], axis=1)
fin, ini = np.uint32([int(0.75 * sv), int(0.6 * sv), int(0.4 * sv)]), np.uint32([int(0.25 * sv), int(0.3 * sv), int(0.4 * sv)])
for _ in range(3):
    t0 = time.perf_counter()
    v = CVP.build_cost_volume_device(Zv, rxy, rxy, rz, sv, sv, sv, 0.3, 0.4, 0.527, 0.2673, 0.1105, path, head, fin, ini)
    torch.cuda.synchronize()
    print("call ms", 1e3 * (time.perf_counter() - t0))
