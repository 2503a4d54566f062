"""Seeded synthetic inputs for tests and bench.

Host-side NumPy/SciPy, not part of the hot path.  The maps follow the *semantics* of the
reference's own builders so that value ranges, plateaus and inf patterns look like what
``Coupled_motion_planner.main`` feeds the solver:

* 2D (``Coupled_motion_planner.py:1144-1216``): binary obstacles -> dilation by the rover
  half-diagonal -> cost = 1 + 300*obstacle + 10*graded distance band -> 50x50 box blur
  -> one-cell inf border.  Values in [1, ~305], a large exact-minimum plateau.
* 3D (``:319-358, :505-725, :1627``): free space 20.0, a graded "tunnel" of cost 4..8 from
  start to goal, inf ground sheet, inf obstacles, inf one-voxel border.
"""
from __future__ import annotations

import numpy as np


def random_costmap(shape, seed=0, lo=1.0, span=4.0, dtype=np.float64):
    """``lo + span*U(0,1)`` with a one-cell inf border: the survey's KAT maps (incoherent
    fronts: worst case for any iterative solver)."""
    rng = np.random.default_rng(seed)
    c = (lo + span * rng.random(shape)).astype(dtype)
    idx = [slice(None)] * len(shape)
    for d in range(len(shape)):
        for e in (0, -1):
            s = list(idx)
            s[d] = e
            c[tuple(s)] = np.inf
    return c


def _disk(r):
    y, x = np.ogrid[-r:r + 1, -r:r + 1]
    return (x * x + y * y) <= r * r


def mars_costmap(n, seed=0, rocks_per_mpx=60.0, dtype=np.float64):
    """Planner-like traversability costmap of side ``n``."""
    from scipy import ndimage

    rng = np.random.default_rng(seed)
    # 1/f terrain by spectral synthesis; slope threshold marks rough ground as obstacle
    kx = np.fft.fftfreq(n)[:, None]
    ky = np.fft.rfftfreq(n)[None, :]
    k = np.sqrt(kx * kx + ky * ky)
    k[0, 0] = 1.0
    spec = (rng.standard_normal(k.shape) + 1j * rng.standard_normal(k.shape)) / k ** 1.6
    spec[0, 0] = 0
    z = np.fft.irfft2(spec, s=(n, n))
    z = (z - z.min()) / (np.ptp(z) + 1e-30)
    gy, gx = np.gradient(z)
    slope = np.hypot(gx, gy)
    obst = slope > np.quantile(slope, 0.97)
    # rocks / craters: random disks
    nrocks = max(3, int(rocks_per_mpx * n * n / 1e6))
    scale = max(1.0, n / 400.0) ** 0.5
    for _ in range(nrocks):
        r = int(rng.integers(2, max(3, int(8 * scale))))
        cy, cx = rng.integers(r + 2, n - r - 2, size=2)
        d = _disk(r)
        obst[cy - r:cy + r + 1, cx - r:cx + r + 1] |= d
    obst[0, :] = obst[-1, :] = obst[:, 0] = obst[:, -1] = False
    obst = ndimage.binary_opening(obst, structure=_disk(2))
    half_diag = 9                                    # rover half-diagonal / resolution (0.45 m @ 5 cm)
    obst = ndimage.binary_dilation(obst, structure=_disk(3), iterations=half_diag // 3)
    obst[0, :] = obst[-1, :] = obst[:, 0] = obst[:, -1] = True
    obstf = obst.astype(np.float64)
    dist = ndimage.distance_transform_edt(~obst)
    band_r = 20.0                                    # 1 m expansion @ 5 cm
    band = np.where(dist <= band_r, 1.0 - dist / max(dist.max(), 1.0), 0.0)
    pos = band > 0
    if pos.any():
        band[pos] -= band[pos].min()
    cost = 1.0 + 300.0 * obstf + 10.0 * band
    cost = ndimage.uniform_filter(cost, size=50, mode="constant", cval=300.0)
    cost = np.maximum(cost, 1.0)
    cost[0, :] = cost[-1, :] = cost[:, 0] = cost[:, -1] = np.inf
    return cost.astype(dtype)


def free_cell_near(cost, x, y, max_cost=2.0):
    """Closest cell to (x, y) whose cost is finite and <= max_cost (so that goals do not
    sit inside an obstacle); returns [x, y]."""
    ok = np.argwhere(np.isfinite(cost) & (cost <= max_cost))
    if ok.size == 0:
        ok = np.argwhere(np.isfinite(cost))
    d = (ok[:, 0] - y) ** 2 + (ok[:, 1] - x) ** 2
    j, i = ok[int(np.argmin(d))]
    return [int(i), int(j)]


def arm_volume(shape, seed=0, dtype=np.float64):
    """Planner-like arm-workspace cost volume ``[y, x, z]``; returns (cost, goal, start)."""
    ny, nx, nz = shape
    rng = np.random.default_rng(seed)
    c = np.full(shape, 20.0)
    yy, xx, zz = np.meshgrid(np.arange(ny), np.arange(nx), np.arange(nz), indexing="ij")
    # ground sheet: inf below a gently undulating surface
    ground = 1 + (0.08 * nz * (1 + np.sin(2 * np.pi * xx[..., 0] / nx) * np.cos(2 * np.pi * yy[..., 0] / ny))).astype(int)
    c[zz < ground[..., None]] = np.inf
    # a few box obstacles
    for _ in range(max(2, int(ny * nx / 2000))):
        y0, x0 = rng.integers(2, ny - 8), rng.integers(2, nx - 8)
        h = int(rng.integers(nz // 4, nz // 2))
        c[y0:y0 + 5, x0:x0 + 5, :h] = np.inf
    start = np.array([nx // 5, ny // 4, min(nz - 3, int(0.55 * nz))])
    goal = np.array([4 * nx // 5, 3 * ny // 4, min(nz - 3, int(0.35 * nz) + 2)])
    # tunnel: graded cost 4..8 inside a tube around the start->goal segment, 20 outside
    p = np.stack([xx, yy, zz], axis=-1).astype(np.float64)
    a, b = start.astype(np.float64), goal.astype(np.float64)
    ab = b - a
    t = np.clip(((p - a) @ ab) / (ab @ ab), 0.0, 1.0)
    d = np.linalg.norm(p - (a + t[..., None] * ab), axis=-1)
    rad = max(3.0, 0.08 * min(ny, nx))
    inside = (d <= rad) & np.isfinite(c)
    c[inside] = 4.0 + 4.0 * d[inside] / rad
    for e in (0, -1):
        c[e] = np.inf
        c[:, e] = np.inf
        c[:, :, e] = np.inf
    c[start[1], start[0], start[2]] = 4.0
    c[goal[1], goal[0], goal[2]] = 4.0
    return c.astype(dtype), goal.tolist(), start.tolist()


def crater_dem(n, resolution, seed=0, craters=None, rocks=None):
    """Zero-based DEM (metres) with gentle 1/f-like undulation, steep-rimmed craters (ring
    obstacles whose floors the planner's hole filling closes, Coupled_motion_planner.py:1164)
    and small rocks (removed by the opening of :1167-1169 when thinner than 21 cells, kept
    otherwise).  Shapes scale with ``resolution`` so the slope threshold 0.20 rad (:1154)
    bites at any grid size."""
    rng = np.random.default_rng(seed)
    size = n * resolution
    ax = (np.arange(n) + 0.5) * resolution
    X, Y = np.meshgrid(ax, ax)
    Z = np.zeros((n, n))
    for k in range(1, 5):                                   # smooth relief, slope well below the threshold
        ph = rng.random(4) * 2 * np.pi
        Z += 0.02 * size / (3.0 * k * k) * np.sin(2 * np.pi * k * X / size + ph[0]) * np.cos(2 * np.pi * k * Y / size + ph[1])
    craters = max(1, n // 96) if craters is None else craters
    rocks = max(2, n // 32) if rocks is None else rocks
    for _ in range(craters):
        cx, cy = rng.uniform(0.15, 0.85, 2) * size
        R = rng.uniform(0.05, 0.12) * size
        d = np.hypot(X - cx, Y - cy)
        w = 0.18 * R
        Z += 0.6 * w * np.exp(-((d - R) / w) ** 2)          # raised rim: steep on both flanks, flat floor inside
    for _ in range(rocks):
        cx, cy = rng.uniform(0.05, 0.95, 2) * size
        R = rng.uniform(4, 30) * resolution
        Z += 0.8 * R * np.exp(-((X - cx) ** 2 + (Y - cy) ** 2) / (2 * (0.5 * R) ** 2))
    return Z - Z.min()
