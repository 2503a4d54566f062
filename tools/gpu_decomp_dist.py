"""Multi-GPU halo-exchange decomposition of one large map (launch with torch.distributed.run).
Prints, on rank 0, the decomposed solve time, the number of exchange rounds and the deviation
from a single-GPU solve of the same map."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from bench import make_map
from planning_motion_planning_b200 import decomp, engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
c = make_map(n, "mars")
goal = synth.free_cell_near(c, n // 4, n // 4)
cd = torch.from_numpy(c).to(dev)
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    lo, hi, T, rounds = decomp.solve2d_slabs_dist(cd, goal)
    dist.barrier(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
single = engine.solve2d(cd, [goal])[0]
t0 = time.perf_counter(); single = engine.solve2d(cd, [goal])[0]; ts = time.perf_counter() - t0
ref = single[lo:hi]
fin = torch.isfinite(ref)
ok = bool(torch.equal(torch.isfinite(T), fin))
err = float(((T - ref).abs() / ref.clamp_min(1e-300))[fin].max()) if fin.any() else 0.0
errs = [None] * dist.get_world_size()
dist.all_gather_object(errs, (lo, hi, ok, err))
if dist.get_rank() == 0:
    print(json.dumps({"n": n, "world": dist.get_world_size(), "decomposed_ms": round(dt * 1e3, 2), "rounds": rounds,
                      "single_gpu_ms": round(ts * 1e3, 2), "slabs": errs}))
dist.destroy_process_group()
