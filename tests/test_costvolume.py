"""3D cost-volume builder.  CPU part: the C restatement against the golden
vectors captured from the unmodified planner / functions, and the CUDA kernel source of
csrc/costvolume.cuh under the host emulator against that oracle.  GPU part: test_gpu_parity.py."""
import os

import numpy as np
import pytest

import emu
from conftest import GOLDEN
from oracle import costvol as CV


def _cases():
    d = np.load(os.path.join(GOLDEN, "costvolume.npz"))
    for tag in ("p", "d0", "d1"):
        g = lambda k: d[f"{tag}_{k}"]      # noqa: E731
        yield tag, dict(Zs=g("Zs"), res=[float(v) for v in g("res")], shape=[int(v) for v in g("shape")], obst=g("obst"),
                        xm_ym=[float(v) for v in g("xm_ym")], final=g("final"), radii=[float(v) for v in g("radii")],
                        path=g("path"), heading=g("heading"), fin=g("fin"), ini=g("ini"), tunnel=g("tunnel"))


CASES = dict(_cases())


@pytest.mark.parametrize("tag", list(CASES))
def test_oracle_matches_golden(tag):
    """oracle/costvol_oracle.c == the unmodified GetObstMap / TunnelCost (bit for bit)."""
    c = CASES[tag]
    (rx, ry, rz), (sX, sY, sZ) = c["res"], c["shape"]
    assert np.array_equal(CV.GetObstMap(c["Zs"], rx, ry, rz, sX, sY, sZ, c["obst"], *c["xm_ym"]), c["final"])
    assert np.array_equal(CV.TunnelCost(*c["radii"], c["path"], sX, sY, sZ, rx, ry, rz, c["heading"], c["fin"], c["ini"]), c["tunnel"])


def test_planner_volume_is_the_product_of_the_golden_parts():
    """The volume the planner handed to FM3D.computeTmap (planner_calls.npz) is Cmap1*Cmap2 (:1627)."""
    p = np.load(os.path.join(GOLDEN, "planner_calls.npz"), allow_pickle=True)
    c = CASES["p"]
    assert np.array_equal(c["final"] * c["tunnel"], p["c3"])


@pytest.mark.parametrize("tag", list(CASES))
def test_emulated_kernels_match_golden(tag):
    c = CASES[tag]
    (rx, ry, rz), (sX, sY, sZ) = c["res"], c["shape"]
    cmap, tunnel, terrain = emu.costvolume(c["Zs"], rx, ry, rz, sX, sY, sZ, *c["xm_ym"], *c["radii"], c["path"], c["heading"],
                                           c["fin"], c["ini"])
    assert np.array_equal(terrain, c["final"])
    assert np.array_equal(tunnel, c["tunnel"])
    assert np.array_equal(cmap, c["final"] * c["tunnel"])
