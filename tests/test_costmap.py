"""2D cost-map builder.  CPU part: the oracle restatement against the
unmodified planner's own cost map (golden), and the CUDA kernel source of csrc/costmap2d.cuh
run under the host emulator against the oracle.  The GPU part is in test_gpu_parity.py."""
import os

import numpy as np
import pytest

import emu
from conftest import GOLDEN
from oracle import costmap_oracle as CO
from planning_motion_planning_b200 import synth


def _noisy(n, res, seed, sigma):
    Z = synth.crater_dem(n, res, seed)
    Z = Z + np.random.default_rng(seed + 100).normal(0.0, sigma * res, Z.shape)
    return Z - Z.min()


CASES = [("planner", 200, 0.05, None), ("craters", 128, 0.05, None), ("coarse", 160, 0.1, None),
         ("fine", 192, 0.03, None), ("rough", 144, 0.06, 0.09), ("very_rough", 128, 0.08, 0.16)]


def _dem(kind, n, res, sigma):
    if kind == "planner":
        return CO.planner_dem(n, res)
    return synth.crater_dem(n, res, 1) if sigma is None else _noisy(n, res, 2, sigma)


def test_oracle_reproduces_the_planners_cost_map_bitwise():
    """tests/golden/planner_calls.npz holds the array the UNMODIFIED planner main() passed to
    biComputeTmap (Coupled_motion_planner.py:1226) for the synthetic DEM of oracle/gen_golden.py."""
    d = np.load(os.path.join(GOLDEN, "planner_calls.npz"), allow_pickle=True)
    n, res = int(d["n"]), float(d["res"])
    c = CO.costmap2d(CO.planner_dem(n, res), res, n * res)
    assert np.array_equal(c.T, d["bi_cost"])


@pytest.mark.parametrize("kind,n,res,sigma", CASES)
def test_emulated_kernels_match_oracle(kind, n, res, sigma):
    Z = _dem(kind, n, res, sigma)
    c, st = CO.costmap2d(Z, res, n * res, stages=True)
    cost, raw, obst, pre, npos = emu.costmap2d(Z, res, n * res)
    assert npos > 0
    assert np.array_equal(raw, st["raw"])                               # slope threshold + first hole filling
    assert np.array_equal(obst, st["obst"].astype(np.uint8))            # opening, dilate / fill / erode, map limits
    assert np.array_equal(pre, st["pre_blur"].T)                        # distance band + composition: bit-exact
    fin = np.isfinite(c)
    assert np.array_equal(np.isfinite(cost.T), fin)
    assert np.max(np.abs(cost.T[fin] - c[fin]) / c[fin]) < 1e-12        # blur: summation order only
    if kind != "planner":
        assert 0 < obst.mean() < 0.9 and raw.sum() != obst.sum()        # the case exercises the morphology


def test_all_obstacle_map_raises_like_the_reference():
    """When the closing swallows the whole map no cell has a positive band value and the reference
    dies in np.min of an empty selection (:1198); the kernels report zero positive cells."""
    n, res = 96, 0.03
    Z = synth.crater_dem(n, res, 3)
    with pytest.raises(ValueError), np.errstate(all="ignore"):
        CO.costmap2d(Z, res, n * res)
    assert emu.costmap2d(Z, res, n * res)[4] == 0


@pytest.mark.parametrize("n,res,seed", [(40, 0.1, 1), (70, 0.08, 2), (33, 0.12, 4), (65, 0.1, 5), (97, 0.06, 3)])
def test_emulated_kernels_on_awkward_sizes(n, res, seed):
    """Maps smaller than the 50x50 blur window, sizes that are not multiples of the 32-pixel run chunks
    or the 64-row scan segments, nearly all-obstacle maps; (97, 0.06, 3) is one the reference dies on."""
    Z = synth.crater_dem(n, res, seed, craters=1, rocks=3)
    Z = Z + np.random.default_rng(seed).normal(0, 0.08 * res, Z.shape)
    Z -= Z.min()
    cost, raw, obst, pre, npos = emu.costmap2d(Z, res, n * res)
    try:
        with np.errstate(all="ignore"):
            c, st = CO.costmap2d(Z, res, n * res, stages=True)
    except ValueError:
        assert npos == 0
        return
    assert npos > 0
    assert np.array_equal(raw, st["raw"]) and np.array_equal(obst, st["obst"].astype(np.uint8))
    assert np.array_equal(pre, st["pre_blur"].T)
    fin = np.isfinite(c)
    assert np.array_equal(np.isfinite(cost.T), fin)
    assert np.max(np.abs(cost.T[fin] - c[fin]) / c[fin]) < 1e-12
