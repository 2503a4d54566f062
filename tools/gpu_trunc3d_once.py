import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import FastMarching.FastMarching3D as FM3D
import FastMarching.FastMarching as FM
from planning_motion_planning_b200 import synth
c3, g3, s3 = synth.arm_volume((90, 90, 28), 0)
for _ in range(2):
    T3 = FM3D.computeTmap(c3, np.uint32(g3), np.uint32(s3))
g = np.load(os.path.join(ROOT, "tests", "golden", "planner_calls.npz"))
cost = np.asfortranarray(g["bi_cost"]); goal = [int(v) for v in g["bi_goal"]]; start = [int(v) for v in g["bi_start"]]
for _ in range(2):
    FM.biComputeTmap(cost, goal, start)
torch.cuda.synchronize()
