"""CPU restatement of the planner's path post-processing (TEST INFRASTRUCTURE: only tests/, smoke() and the
cpu_baseline leg of bench.py may import this).

Follows Coupled_motion_planner.py with the same third-party calls the reference makes (scipy.signal.savgol_filter,
scipy.interpolate.interp1d, numpy.linspace):

  stitch_rover_path   :1232-1234   roverPath = resolution * (vstack(flipud(pathS), pathG[1:]) + 1)
  smooth_resample_arm :1641-1671   per-axis scale, savgol_filter(., 11, 3), shift to the global frame, last row := the
                                   sample pose, interp1d(range(n), .)(linspace(0, n - 1, m))

Pinned by tests/golden/pathpost.npz: the locals of the UNMODIFIED planner main() at its return (oracle/gen_golden.py
pathpost) -- this file reproduces roverPath and resizedGamma3D of that run bit for bit.
"""
from __future__ import annotations

import numpy as np


def stitch_rover_path(pathS, pathG, resolution):
    """Coupled_motion_planner.py:1232-1234."""
    roverPath = np.vstack((np.flipud(np.asarray(pathS, dtype=np.float64)), np.asarray(pathG, dtype=np.float64)[1:, :]))
    return np.dot(resolution, roverPath + 1)


def smooth_resample_arm(path3d, res3, offset3, last3, m):
    """Coupled_motion_planner.py:1641-1671 (gamma3D -> resizedGamma3D).  Raises ValueError like the reference when
    the path has fewer than 11 rows (scipy: window_length must be <= the size of x)."""
    from scipy import interpolate, signal
    g = np.array(path3d, dtype=np.float64)
    for d in range(3):
        g[:, d] = g[:, d] * res3[d]
    for d in range(3):
        g[:, d] = signal.savgol_filter(g[:, d], 11, 3)
    for d in range(3):
        g[:, d] = g[:, d] + offset3[d]
    if last3 is not None:
        g[-1, :] = last3
    out = np.zeros([m, 3])
    for d in range(3):
        f = interpolate.interp1d(range(0, len(g)), g[:, d])
        out[:, d] = f(np.linspace(0, len(g) - 1, m, endpoint=True))
    return out
