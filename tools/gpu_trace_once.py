"""One 4096^2 solve + the bench path, for an ncu capture of trace2d_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import make_map
from planning_motion_planning_b200 import engine, synth
n = 4096
c = make_map(n, "mars")
goal = synth.free_cell_near(c, n // 4, n // 4)
start = synth.free_cell_near(c, 3 * n // 4, 3 * n // 4)
T = engine.solve2d(torch.from_numpy(c).cuda(), [goal])
for _ in range(2):
    out, cnt, st = engine.trace2d(T, [start], [goal], 0.5)
torch.cuda.synchronize()
print(int(cnt[0]), int(st[0]))
