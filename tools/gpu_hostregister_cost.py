"""Cost of page-locking a caller's array in place (cudaHostRegister / cudaHostUnregister), 128 MiB, touched memory."""
import time, numpy as np, torch
torch.cuda.init()
rt = torch.cuda.cudart()
for mib in (8, 32, 128):
    a = np.random.default_rng(0).random(mib * 131072)      # touched pages
    ts, tu = [], []
    for _ in range(5):
        t0 = time.perf_counter(); rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0); t1 = time.perf_counter()
        rt.cudaHostUnregister(a.ctypes.data); t2 = time.perf_counter()
        ts.append(1e3 * (t1 - t0)); tu.append(1e3 * (t2 - t1))
    print(mib, "MiB: register", [round(t, 2) for t in ts], "unregister", [round(t, 2) for t in tu], "rc", int(rc))
