"""Drop-in replacement of the reference package ``src/FastMarching`` (same package
and module names, same function signatures): put the directory that contains this
package first on ``sys.path`` and ``Coupled_motion_planner.py`` runs unchanged
(its only import site is Coupled_motion_planner.py:13-14).

The numerics run on a B200 through ``planning_motion_planning_b200`` (CUDA, C ABI);
there is no CPU implementation behind these functions.
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:          # make the engine importable when only this package's parent is on sys.path
    _sys.path.append(_root)
