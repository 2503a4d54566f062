"""A/B of fmb_bisolve2d_f64 after an upload of the map against fmb_bisolve2d_h2d_f64 (upload overlapped), page-locked host map."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from FastMarching import _compat as C
from bench import make_map
from planning_motion_planning_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c = make_map(n, "mars")
goal = synth.free_cell_near(c, n // 4, n // 4); start = synth.free_cell_near(c, 3 * n // 4, 3 * n // 4)
dev = C.device()
h = torch.empty(c.shape, dtype=torch.float64).pin_memory(); h.copy_(torch.from_numpy(c)); hn = h.numpy()
for mode in ("upload, then bisolve", "bisolve_h2d", "upload, then bisolve", "bisolve_h2d"):
    ts = []
    for rep in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if mode == "bisolve_h2d":
            TG, TS, info, ws = C.bisolve2d(hn, goal, start, False, dev)
        else:
            cd = h.to(dev, non_blocking=True)
            TG, TS, info, ws = C.bisolve2d(cd, goal, start, False)
        import ctypes
        from planning_motion_planning_b200 import _capi
        st = _capi.FmbStats()
        _capi.check(_capi.lib().fmb_finish(ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream, ctypes.byref(st)))
        ts.append(1e3 * (time.perf_counter() - t0))
        last = st.as_dict()
    print(mode, [round(t, 2) for t in ts], "solve_kernel_ms", round(last["solve_kernel_ms"], 2), "wait%", round(100 * last["cyc_wait"] / max(1, last["cyc_wait"] + last["cyc_load"] + last["cyc_relax"] + last["cyc_store"]), 1), "visits", last["tile_visits"])
