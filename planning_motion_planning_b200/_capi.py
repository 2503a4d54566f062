"""ctypes binding of include/fm_b200.h -> libfm_b200.so.

Fails loudly when the CUDA library is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

FMB_OK = 0
FMB_E_INVALID, FMB_E_CUDA, FMB_E_WORKSPACE, FMB_E_WATCHDOG, FMB_E_STEPCAP, FMB_E_NOJOIN = 1, 2, 3, 4, 5, 6
TRACE_OK, TRACE_EARLY, TRACE_VALUEERROR, TRACE_INDEXERROR, TRACE_OVERFLOW = 0, 1, 2, 3, 4


class FmbStats(C.Structure):
    _fields_ = [("tile_visits", C.c_uint64), ("steps", C.c_uint64), ("evals", C.c_uint64),
                ("pushes", C.c_uint64), ("cells_written", C.c_uint64), ("solve_kernel_ms", C.c_double),
                ("init_kernel_ms", C.c_double), ("cyc_wait", C.c_uint64), ("cyc_load", C.c_uint64),
                ("cyc_relax", C.c_uint64), ("cyc_store", C.c_uint64), ("reserved", C.c_uint64 * 1),
                ("cyc_check", C.c_uint64), ("noop_visits", C.c_uint64), ("rounds", C.c_uint64),
                ("continuations", C.c_uint64)]

    def as_dict(self):
        d = {k: int(getattr(self, k)) for k in ("tile_visits", "steps", "evals", "pushes", "cells_written",
                                                "cyc_wait", "cyc_load", "cyc_relax", "cyc_store", "cyc_check", "noop_visits", "rounds", "continuations")}
        d["deferrals"] = int(self.reserved[0])
        d["solve_kernel_ms"] = float(self.solve_kernel_ms)
        d["init_kernel_ms"] = float(self.init_kernel_ms)
        return d


class FmbOptions(C.Structure):
    """fmb_options of include/fm_b200.h (process-wide solver tunables)."""
    _fields_ = [(k, C.c_int32) for k in ("engine2d", "cta_cells", "tile_w2d", "tile_z3d", "best_first", "windowed",
                                         "window", "worker_div", "max_blocks", "watchdog_ms", "step_cap", "engine3d",
                                         "level_div", "win_running", "check_passes", "pipeline", "precheck",
                                         "causal_slack", "tma", "ring2", "variant", "concurrent_solves", "replay_sparse")]


class FmbPlan2DResult(C.Structure):
    """fmb_plan2d_result of include/fm_b200.h (owned by the caller, released by fmb_plan2d_free)."""
    _fields_ = [("nq", C.c_int32), ("offsets", C.POINTER(C.c_int64)), ("waypoints", C.POINTER(C.c_double)),
                ("status", C.POINTER(C.c_int32)), ("solve_ms", C.c_double), ("trace_ms", C.c_double)]


class FmbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfm_b200 error {code}: {msg}")
        self.code = code


_vp, _i64, _i32, _sz, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_size_t, C.c_double

# name -> (restype, argtypes); every symbol include/fm_b200.h declares
SIGNATURES = {
    "fmb_version": (C.c_int, []),
    "fmb_last_error": (C.c_char_p, []),
    "fmb_sm_count": (C.c_int, []),
    "fmb_debug_sqrt_check": (C.c_int, [_vp, _i64, _vp, _vp]),
    "fmb_debug_div_check": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "fmb_get_options": (None, [C.POINTER(FmbOptions)]),
    "fmb_set_options": (C.c_int, [C.POINTER(FmbOptions)]),
    "fmb_workspace_bytes_2d": (_sz, [_i32, _i32, _i32]),
    "fmb_solve2d_f64": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "fmb_solve2d_f32": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "fmb_resolve2d_f64": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _i32, _i32, _vp, _sz, _vp]),
    "fmb_workspace_bytes_3d": (_sz, [_i32, _i32, _i32, _i32]),
    "fmb_solve3d_f64": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "fmb_solve3d_f32": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "fmb_solve3d_exact_f64": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "fmb_polish3d_f64": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "fmb_finish": (C.c_int, [_vp, _sz, _vp, C.POINTER(FmbStats)]),
    "fmb_trace2d_f64": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _dbl, _i32, _vp, _i64, _vp, _vp, _vp]),
    "fmb_trace3d_f64": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _dbl, _i32, _vp, _i64, _vp, _vp, _vp]),
    "fmb_truncate2d_f64": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "fmb_truncate3d_f64": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "fmb_tie_keys2d_f64": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "fmb_tie_order2d_f64": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "fmb_tie_order3d_f64": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "fmb_bi_join": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "fmb_workspace_bytes_costmap2d": (_sz, [_i32]),
    "fmb_costmap2d_f64": (C.c_int, [_vp, _vp, _i32, _dbl, _dbl, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "fmb_costmap2d_finish": (C.c_int, [_vp, _sz, _vp, C.POINTER(C.c_int32)]),
    "fmb_path_pack_f64": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _dbl, _dbl, _vp, _vp]),
    "fmb_workspace_bytes_pathpost": (_sz, []),
    "fmb_path_stitch2d_f64": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _dbl, _vp, _vp, _vp]),
    "fmb_path_post3d_f64": (C.c_int, [_vp, _vp, _i64, _i32, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "fmb_workspace_bytes_pop_ranks": (_sz, [_i64]),
    "fmb_pop_ranks2d_f64": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "fmb_pop_ranks3d_f64": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "fmb_pop_ranks_status": (C.c_int, [_vp, _vp, C.POINTER(C.c_int32)]),
    "fmb_workspace_bytes_bisolve2d": (_sz, [_i32, _i32]),
    "fmb_bisolve2d_f64": (C.c_int, [_vp, _i32, _i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i32, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "fmb_bisolve2d_h2d_f64": (C.c_int, [_vp, _vp, _i32, _i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i32, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "fmb_workspace_bytes_until2d": (_sz, [_i32, _i32]),
    "fmb_workspace_bytes_until3d": (_sz, [_i32, _i32, _i32]),
    "fmb_solve2d_until_f64": (C.c_int, [_vp, _i32, _i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i32, _vp, _vp, _vp, _sz, _vp]),
    "fmb_solve3d_until_f64": (C.c_int, [_vp, _i32, _i32, _i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _vp, _vp, _vp, _sz, _vp]),
    "fmb_plan_batch2d_host": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _i32, _dbl, _i32, _dbl, _i32,
                                        C.POINTER(C.POINTER(FmbPlan2DResult))]),
    "fmb_plan2d_free": (None, [C.POINTER(FmbPlan2DResult)]),
    "fmb_fields_differ_f64": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "fmb_trace2d_logged_f64": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _dbl, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp]),
    "fmb_windows_differ_f64": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp]),
    "fmb_workspace_bytes_2d_h2d": (_sz, [_i32, _i32]),
    "fmb_solve2d_h2d_f64": (C.c_int, [_vp, _vp, _i32, _i32, C.POINTER(C.c_int32), _vp, _vp, _sz, _vp, _vp]),
    "fmb_workspace_bytes_costvolume": (_sz, [_i32, _i32, _i32]),
    "fmb_costvolume_f64": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
}


def lib_path() -> str:
    return _build.LIB_PATH


def lib():
    """Load (building first if stale and nvcc is present) the CUDA library."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("FMB_LIB") or _build.LIB_PATH        # FMB_LIB: A/B builds of the same ABI
        if path == _build.LIB_PATH and _build.needs_build():
            try:
                _build.build()
            except Exception as e:  # no nvcc / compile error: loud, never a fallback
                if not os.path.exists(path):
                    raise ImportError(f"libfm_b200.so is missing and could not be built: {e}") from e
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def get_options() -> dict:
    o = FmbOptions()
    lib().fmb_get_options(C.byref(o))
    return {k: int(getattr(o, k)) for k, _ in FmbOptions._fields_ if k != "reserved"}


def set_options(**kw) -> dict:
    """Update solver tunables (see fmb_options); returns the previous values."""
    o = FmbOptions()
    lib().fmb_get_options(C.byref(o))
    prev = {k: int(getattr(o, k)) for k, _ in FmbOptions._fields_ if k != "reserved"}
    for k, v in kw.items():
        if k not in prev:
            raise KeyError(k)
        setattr(o, k, int(v))
    check(lib().fmb_set_options(C.byref(o)))
    return prev


def check(rc: int):
    if rc != FMB_OK:
        raise FmbError(rc, lib().fmb_last_error().decode())
