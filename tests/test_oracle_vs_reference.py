"""CPU, build container only: live differential test of the C oracle against the imported
reference (skipped where /root/reference does not exist, e.g. on the GPU box)."""
import numpy as np
import pytest

from conftest import rand_map
from oracle import oracle as O
from oracle import ref_loader as R

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not present")


@pytest.mark.parametrize("seed", [11, 12])
def test_2d_bitwise(seed):
    c = rand_map((40, 44), seed)
    assert np.array_equal(O.computeTmap(c, [7, 9]), R.computeTmap2D(c, [7, 9]))
    TG, TS, j = O.biComputeTmap(c, [5, 5], [38, 30])
    rTG, rTS, rj = R.biComputeTmap(c, [5, 5], [38, 30])
    assert np.array_equal(TG, rTG) and np.array_equal(TS, rTS) and np.array_equal(j, rj)
    assert np.array_equal(O.getPathGDM(TG, j, [5, 5], 0.5), R.getPathGDM2D(rTG, rj, [5, 5], 0.5))


def test_3d_bitwise():
    c = rand_map((14, 15, 16), 13)
    T, rT = O.computeTmap3D(c, [3, 4, 5], [11, 10, 9]), R.computeTmap3D(c, [3, 4, 5], np.uint32([11, 10, 9]))
    assert np.array_equal(T, rT)
    assert np.array_equal(O.getPathGDM3D(T, np.uint32([11, 10, 9]), np.uint32([3, 4, 5]), 0.5),
                          R.getPathGDM3D(rT, np.uint32([11, 10, 9]), np.uint32([3, 4, 5]), 0.5))
