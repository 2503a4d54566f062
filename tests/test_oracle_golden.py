"""CPU: pins the C oracle (oracle/fmm_oracle.c) to golden vectors frozen from the reference
(oracle/gen_golden.py ran /root/reference/src/FastMarching in the build container)."""
import hashlib

import numpy as np
import pytest

from conftest import GOLDEN, plateau_map, rand_map
from oracle import oracle as O


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def g2():
    return np.load(f"{GOLDEN}/ref2d.npz")


@pytest.fixture(scope="module")
def g3():
    return np.load(f"{GOLDEN}/ref3d.npz")


def test_kat1_values(g2):
    c = np.pad(np.ones((7, 7)), 1, constant_values=np.inf)
    T = O.computeTmap(c, [4, 4])
    assert np.array_equal(T, g2["kat1_T"])
    # SURVEY.md 8c KAT-1
    assert list(T[4, 1:8]) == [3, 2, 1, 0, 1, 2, 3]
    assert T[5, 5] == 1.7071067811865475 and T[6, 5] == 2.5453289254261224 and T[6, 6] == 3.25243570661267


def test_kat3_published_numbers():
    c = rand_map((100, 100), 0)
    assert c[50, 50] == 3.215565077017812
    T = O.computeTmap(c, [25, 25])
    assert np.sum(T[np.isfinite(T)]) == 1224641.4437929934
    assert T[50, 50] == 96.02583155444606 and T[98, 98] == 280.68674463626854
    assert sha(T)[:16] == "b23ad3237183a43c"


@pytest.mark.parametrize("name", ["kat3", "rand64", "plateau80"])
def test_2d_cases(g2, name):
    n, seed = int(g2[f"{name}_n"]), int(g2[f"{name}_seed"])
    c = rand_map((n, n), seed) if str(g2[f"{name}_kind"]) == "rand" else plateau_map(n, seed)
    goal, gg, ss = list(g2[f"{name}_goal"]), list(g2[f"{name}_g2"]), list(g2[f"{name}_s2"])
    T = O.computeTmap(c, goal)
    assert sha(T) == str(g2[f"{name}_full_sha"])
    assert sha(O.computeTmap(c, goal, ss)) == str(g2[f"{name}_early_sha"])
    TG, TS, j = O.biComputeTmap(c, gg, ss)
    assert sha(TG) == str(g2[f"{name}_TG_sha"]) and sha(TS) == str(g2[f"{name}_TS_sha"])
    assert np.array_equal(j, g2[f"{name}_join"])
    assert np.array_equal(O.getPathGDM(TG, j, gg, 0.5), g2[f"{name}_pathG"])
    assert np.array_equal(O.getPathGDM(TS, j, ss, 0.5), g2[f"{name}_pathS"])
    assert np.array_equal(O.getPathGDM(T, np.array(ss), goal, 0.5), g2[f"{name}_pathF"])


def test_kat2_values(g3):
    c = rand_map((9, 9, 9), 0)
    c[np.isfinite(c)] = 1.0
    T = O.computeTmap3D(c, [4, 4, 4])
    assert np.array_equal(T, g3["kat2_T"])
    assert T[4, 5, 4] == 1.0 and T[5, 5, 4] == 1.7071067811865475 and T[5, 5, 5] == 2.2844570503761727
    assert T[6, 5, 4] == 2.545328925426122 and T[6, 6, 6] == 4.243559040786821


@pytest.mark.parametrize("name", ["kat4", "slab20", "box"])
def test_3d_cases(g3, name):
    shape, seed = tuple(int(v) for v in g3[f"{name}_shape"]), int(g3[f"{name}_seed"])
    c = rand_map(shape, seed)
    slab = g3[f"{name}_slab"]
    if slab.size:
        c[tuple(slice(int(a), int(b)) for a, b in slab)] = np.inf
    goal, start = list(g3[f"{name}_goal"]), list(g3[f"{name}_start"])
    T = O.computeTmap3D(c, goal)
    assert sha(T) == str(g3[f"{name}_full_sha"])
    Tt = O.computeTmap3D(c, goal, start)
    assert sha(Tt) == str(g3[f"{name}_trunc_sha"])
    for tag, F in (("full", T), ("trunc", Tt)):
        exc = str(g3[f"{name}_exc_{tag}"])
        p, st = O.getPathGDM3D(F, np.uint32(start), np.uint32(goal), 0.5, return_status=True)
        assert {0: "", 2: "ValueError", 3: "IndexError", 4: "OverflowError"}[st] == exc
        if not exc:
            assert np.array_equal(p, g3[f"{name}_path_{tag}"])


def test_kat4_published_numbers():
    c = rand_map((24, 24, 24), 0)
    T = O.computeTmap3D(c, [5, 6, 7])
    assert np.sum(T[np.isfinite(T)]) == 385723.3651708018 and T[17, 18, 16] == 52.30754804936137
    Tt = O.computeTmap3D(c, [5, 6, 7], [18, 17, 16])
    assert int(np.isfinite(Tt).sum()) == 9830


def test_planner_calls():
    """What the UNMODIFIED planner main() passed to / got from the five FastMarching calls."""
    g = np.load(f"{GOLDEN}/planner_calls.npz")
    c = g["bi_cost"]
    TG, TS, j = O.biComputeTmap(c, list(g["bi_goal"]), list(g["bi_start"]))
    assert sha(TG) == str(g["bi_TG_sha"]) and sha(TS) == str(g["bi_TS_sha"])
    assert np.array_equal(j, g["bi_join"])
    assert np.array_equal(O.getPathGDM(TG, g["pathG_init"], g["pathG_end"], float(g["pathG_tau"])), g["pathG"])
    assert np.array_equal(O.getPathGDM(TS, g["pathS_init"], g["pathS_end"], float(g["pathS_tau"])), g["pathS"])
    T3 = O.computeTmap3D(g["c3"], list(g["g3"]), list(g["s3"]))
    assert sha(T3) == str(g["T3_sha"])
    assert np.array_equal(O.getPathGDM3D(T3, g["path3d_init"], g["path3d_end"], float(g["path3d_tau"])), g["path3d"])


def test_fixed_point_property():
    """KAT-5: a full field is a bitwise fixed point of the local update."""
    c = rand_map((50, 50), 3)
    T = O.computeTmap(c, [20, 30])
    P = np.pad(T, 1, constant_values=np.inf)
    for y in range(50):
        for x in range(50):
            if not np.isfinite(c[y, x]) or (x, y) == (20, 30):
                continue
            a = min(P[y + 1, x], P[y + 1, x + 2])
            b = min(P[y, x + 1], P[y + 2, x + 1])
            assert O.getEikonal(a, b, c[y, x]) == T[y, x]
