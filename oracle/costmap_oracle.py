"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the planner's 2D cost-map construction
(reference: src/Coupled_motion_planner.py:37-80 surface_normal,
:83-95 image_filling, :97-109 structural_disk, :1144-1216 the inline pipeline of main()).

It calls the same third-party routines the reference calls (cv2 4.13 erode / dilate /
floodFill, scipy 1.18 ndimage.distance_transform_edt and signal.convolve2d, numpy 2.3), so it
is pinned by the unmodified planner's own output: ``tests/golden/planner_calls.npz`` holds
the cost map main() handed to biComputeTmap for a synthetic DEM, and
``tests/test_costmap_oracle.py`` checks this restatement reproduces it bit for bit.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.
"""
import math

import numpy as np


def disk(r: int) -> np.ndarray:
    """Binary disk of radius r (:97-109): offsets with di^2 + dj^2 <= r^2."""
    k = np.arange(-r, r + 1)
    return (np.sqrt(np.add.outer(k * k, k * k).astype(np.float64)) <= r).astype(np.uint8)


def fill_holes(im: np.ndarray) -> np.ndarray:
    """Zero regions not 4-connected to pixel (0,0) become 1 (:83-95)."""
    import cv2
    h, w = im.shape
    ff = im.copy()
    cv2.floodFill(ff, np.zeros((h + 2, w + 2), np.uint8), (0, 0), 1)
    return im | (cv2.bitwise_not(ff) - 254)


def normals_z(resolution: float, size: float, z: np.ndarray) -> np.ndarray:
    """z component of the unit surface normal of the DEM on its meshgrid (:37-80)."""
    from scipy import signal
    n = int(round(size / resolution))
    g = np.linspace(0, size, n)
    x, y = np.meshgrid(g, g)

    def pad(a):
        m = a.shape[0]
        a = np.vstack((3 * a[0] - 3 * a[1] + a[2], a, 3 * a[m - 1] - 3 * a[m - 2] + a[m - 3]))
        k = a.shape[1]
        return np.hstack(((3 * a[:, 0] - 3 * a[:, 1] + a[:, 2])[:, None], a,
                          (3 * a[:, k - 1] - 3 * a[:, k - 2] + a[:, k - 3])[:, None]))
    xx, yy, zz = pad(x), pad(y), pad(z)
    s1 = np.array([[0, 0, 0], [1, 0, -1], [0, 0, 0]]) / 2
    s2 = np.array([[0, -1, 0], [0, 0, 0], [0, 1, 0]]) / 2
    ax, ay, az = (-signal.convolve2d(v, np.flipud(s1), mode="valid") for v in (xx, yy, zz))
    bx, by, bz = (signal.convolve2d(v, np.flipud(s2), mode="valid") for v in (xx, yy, zz))
    nx = -(ay * bz - az * by)
    ny = -(az * bx - ax * bz)
    nz = -(ax * by - ay * bx)
    mag = np.sqrt(nx * nx + ny * ny + nz * nz)
    mag[mag == 0] = np.finfo(float).eps
    return nz / mag


def costmap2d(Zs: np.ndarray, resolution: float, size: float, diagonal: float = 0.9, stages: bool = False,
              blur: bool = True):
    """The cost map main() builds from the zero-based DEM ``Zs`` (:1144-1216).  Returns cMap
    (indexed [x, y] -- the planner passes cMap.T to biComputeTmap); with ``stages`` also the
    intermediate maps.  ``blur=False`` stops before the 2500-tap convolution (tens of seconds at
    4096^2 on a CPU) and returns only the stages."""
    import cv2
    from scipy import ndimage, signal
    slope = np.arccos(normals_z(resolution, size, Zs))
    obst = np.zeros(Zs.shape)
    obst[slope > 0.20] = 1
    obst[0, :] = 0; obst[-1, :] = 0; obst[:, 0] = 0; obst[:, -1] = 0
    obst = fill_holes(np.uint8(obst))
    raw = obst.copy()
    se = disk(10)
    obst = cv2.dilate(cv2.erode(obst, se, iterations=1), se, iterations=1)
    se = disk(int(round(diagonal / 2 / resolution)))
    obst = cv2.erode(fill_holes(cv2.dilate(obst, se, iterations=1)), se, iterations=1)
    obst[0, :] = 1; obst[-1, :] = 1; obst[:, 0] = 1; obst[:, -1] = 1
    obst = np.float64(obst)
    dil = cv2.dilate(obst, disk(int(round(1 / resolution))), iterations=1)
    dist = resolution * ndimage.distance_transform_edt(obst == 0)
    od = dil * (1 - dist / np.max(dist))
    pos = od > 0
    od[pos] = od[pos] - np.min(od[pos])
    pre = 1 + (obst * 300 + od * 10).T
    if not blur:
        return None, {"raw": raw, "obst": obst, "dilated": dil, "dist": dist, "pre_blur": pre}
    cmap = signal.convolve2d(pre, np.flipud(np.ones((50, 50)) / 50 ** 2), mode="same", fillvalue=300)
    cmap[0, :] = np.inf; cmap[-1, :] = np.inf; cmap[:, 0] = np.inf; cmap[:, -1] = np.inf
    if stages:
        return cmap, {"slope": slope, "raw": raw, "obst": obst, "dilated": dil, "dist": dist, "pre_blur": pre}
    return cmap


def planner_dem(n: int, res: float) -> np.ndarray:
    """The synthetic DEM oracle/gen_golden.py wrote for the planner run (after the text round
    trip and the planner's ``Zs - min``, Coupled_motion_planner.py:1101)."""
    size = n * res
    ax = (np.arange(n) + 0.5) * res
    X, Y = np.meshgrid(ax, ax)
    Z = 0.03 * np.sin(2 * np.pi * X / (0.5 * size)) * np.cos(2 * np.pi * Y / (0.7 * size))
    Z += 0.5 * np.exp(-((X - 0.5 * size) ** 2 + (Y - 0.45 * size) ** 2) / (2 * (0.06 * size) ** 2))
    return Z - np.min(Z)
