"""The planner-scale biComputeTmap (400^2) and FM3D.computeTmap (90x90x28) once each after a warm-up: run under
ncu --metrics gpu__time_duration.sum for the launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import FastMarching.FastMarching as FM
import FastMarching.FastMarching3D as FM3D
from planning_motion_planning_b200 import synth
c = synth.mars_costmap(400, 3).T
goal = synth.free_cell_near(c.T, 320, 320)[::-1]; start = synth.free_cell_near(c.T, 80, 80)[::-1]
c3, g3, s3 = synth.arm_volume((90, 90, 28), 0)
for rep in range(2):
    FM.biComputeTmap(c, goal, start)
    FM3D.computeTmap(c3, np.uint32(g3), np.uint32(s3))
print("done")
