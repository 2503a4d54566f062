import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def rand_map(shape, seed):
    """1 + 4 U(0,1) with a one-cell inf border (the survey's KAT maps)."""
    rng = np.random.default_rng(seed)
    c = 1.0 + rng.random(shape) * 4
    sl = [slice(None)] * len(shape)
    for d in range(len(shape)):
        for e in (0, -1):
            s = list(sl)
            s[d] = e
            c[tuple(s)] = np.inf
    return c


def plateau_map(n, seed):
    rng = np.random.default_rng(seed)
    blocks = rng.choice([1.0, 1.0, 1.0, 150.5, 301.0], size=(n // 8 + 1, n // 8 + 1))
    c = np.kron(blocks, np.ones((8, 8)))[:n, :n].copy()
    c[0, :] = c[-1, :] = c[:, 0] = c[:, -1] = np.inf
    return c


def rel_err(a, ref):
    """max relative error over cells finite in ref; also asserts the inf patterns agree."""
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(a), fin), "inf pattern differs"
    pos = fin & (ref != 0)
    if not pos.any():
        return 0.0
    return float(np.max(np.abs(a[pos] - ref[pos]) / np.abs(ref[pos])))


@pytest.fixture
def fmb_opts():
    """Set solver tunables (fmb_options of include/fm_b200.h) for one test; restored afterwards."""
    from planning_motion_planning_b200 import _capi
    saved = _capi.get_options()

    def setter(**kw):
        _capi.set_options(**kw)
    yield setter
    _capi.set_options(**saved)


# 2D engines of the library: name -> fmb_options fields
ENGINES_2D = {"sweep": dict(engine2d=3), "cta1": dict(engine2d=2, cta_cells=1), "cta2": dict(engine2d=2, cta_cells=2),
              "cta4": dict(engine2d=2, cta_cells=4), "warp32": dict(engine2d=1, tile_w2d=32), "warp16": dict(engine2d=1, tile_w2d=16),
              "warp32g": dict(engine2d=6)}
