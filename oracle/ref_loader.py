"""Read-only loader for the *actual* reference (``/root/reference/src/FastMarching``).

TEST INFRASTRUCTURE ONLY, and only usable in the build container: the GPU box
has no ``/root/reference``.  ``oracle/gen_golden.py`` uses it to pin the C
restatement (``fmm_oracle.c``) and to freeze golden vectors under
``tests/golden/``; ``tests/test_oracle_vs_reference.py`` skips when it is absent.

The shipped 2D ``computeTmap`` (FastMarching.py:92-112) raises ``ValueError`` at
:107 (three return values unpacked into two), so :func:`computeTmap2D` drives
the reference's own ``updateNode`` / ``getMinNB`` from a corrected loop that is
otherwise the same as :92-106 / the working 3D driver (FastMarching3D.py:126-145).
"""
from __future__ import annotations

import importlib
import os
import sys
import warnings

import numpy as np

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "FastMarching", "FastMarching.py"))


def _import(name):
    """Import ``FastMarching.<name>`` from the reference tree without letting it shadow
    (or be shadowed by) this repo's own drop-in ``FastMarching`` package."""
    if not available():
        raise ImportError("reference tree not present (expected in the build container only)")
    sys.dont_write_bytecode = True
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "FastMarching" or k.startswith("FastMarching.")}
    sys.path.insert(0, REF_SRC)
    try:
        mod = importlib.import_module("FastMarching." + name)
    finally:
        sys.path.remove(REF_SRC)
        for k in list(sys.modules):
            if k == "FastMarching" or k.startswith("FastMarching."):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return mod


_FM = None
_FM3D = None


def FM():
    global _FM
    if _FM is None:
        _FM = _import("FastMarching")
    return _FM


def FM3D():
    global _FM3D
    if _FM3D is None:
        _FM3D = _import("FastMarching3D")
    return _FM3D


def computeTmap2D(costMap, goal, start=None):
    """Corrected driver around the reference's own 2D updateNode/getMinNB."""
    fm = FM()
    costMap = np.asarray(costMap, dtype=np.float64)
    closedMap = np.zeros_like(costMap)
    closedMap[np.where(costMap == np.inf)] = 1
    Tmap = np.ones_like(costMap) * np.inf
    nbT, nbNodes = [], []
    Tmap[goal[1], goal[0]] = 0
    closedMap[goal[1], goal[0]] = 1
    Tmap, nbT, nbNodes = fm.updateNode([goal[0], goal[1]], costMap, Tmap, nbT, nbNodes, closedMap)
    while nbT:
        node, nbT, nbNodes = fm.getMinNB(nbT, nbNodes)
        closedMap[node[1], node[0]] = 1
        Tmap, nbT, nbNodes = fm.updateNode(node, costMap, Tmap, nbT, nbNodes, closedMap)
        if start is not None and np.array_equal(node, start):
            break
    return Tmap


def biComputeTmap(costMap, goal, start):
    return FM().biComputeTmap(np.asarray(costMap, dtype=np.float64), goal, start)


def getPathGDM2D(T, init, end, tau):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return FM().getPathGDM(T, np.asarray(init), end, tau)


def computeTmap3D(costMap, goal, start=None):
    s = np.array([-1, -1, -1]) if start is None else np.asarray(start)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return FM3D().computeTmap(np.asarray(costMap, dtype=np.float64), np.asarray(goal), s)


def getPathGDM3D(T, init, end, tau):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return FM3D().getPathGDM(T, np.asarray(init), np.asarray(end), tau)
