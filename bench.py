#!/usr/bin/env python
"""bench.py -- headline measurement of the Fast Marching hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # own arm (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm (CPU port, host cores)

Workload (BASELINE.json metric: "Eikonal solve ms + cell-updates/s (4096^2 map); batched
queries/s at 1/2/4/8 GPU"): one STEP = one full-field Eikonal solve of a 4096x4096 fp64
planner-like costmap (SURVEY.md 8d "metric map", seed 0) from one goal + one gradient-descent
path extraction over the result.  With N GPUs every rank runs that query on its own GPU
(replicas: identical per-GPU work, no data-path collective): weak scaling.  The batched section
(`batch`) gives every rank its own 4096 distinct goal queries.

  value = cells solved per second over the whole job = N * 4096^2 * K / t, t = device time of
          the K timed steps (CUDA events), max over ranks; inputs resident in HBM.
  e2e   = the same work through the reference-facing PLUGIN CALL, single caller thread like the planner
          (Coupled_motion_planner.py:1226-1230): FastMarching.FastMarching.computeTmap(costMap, goal, start) +
          getPathGDM(T, start, goal, tau) on NumPy arrays; host<->device copies inside the timed region.
          (`e2e_pipelined` keeps round 1's number: the engine API with pinned buffers, queries in flight.)
  roofline: dominant kernel = the persistent solve kernel (solve2d_sweep_kernel); algorithmic bytes = 2*8 B per
          cell (read cost once, write T once); duration = CUDA events recorded
          by the library around that kernel.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "eikonal_cell_updates_per_s_4096x4096_solve_plus_path"
UNIT = "cells/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=4096, help="map side (default: the metric's 4096)")
    ap.add_argument("--map", default="mars", choices=["mars", "random"])
    ap.add_argument("--no-batch", action="store_true", help="skip the batched-queries section")
    ap.add_argument("--batch-queries", type=int, default=4096, help="512^2 queries per GPU in the batched section")
    ap.add_argument("--no-3d", action="store_true", help="skip the 3D (arm-workspace volume) section")
    ap.add_argument("--no-costmap", action="store_true", help="skip the cost-map construction section")
    ap.add_argument("--inflight", type=int, default=4, help="independent queries solved concurrently per GPU (own stream each)")
    ap.add_argument("--size3d", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def make_map(n, kind, seed=0):
    from planning_motion_planning_b200 import synth
    cache = f"/tmp/fmb_bench_{kind}_{n}_{seed}.npy"
    if os.path.exists(cache):
        try:
            return np.load(cache)
        except Exception:
            pass
    c = synth.mars_costmap(n, seed) if kind == "mars" else synth.random_costmap((n, n), seed)
    try:        # atomic publish: several ranks may build the same cache concurrently
        tmp = f"{cache}.{os.getpid()}.tmp.npy"
        np.save(tmp, c)
        os.replace(tmp, cache)
    except Exception:
        pass
    return c


def make_config(args):
    """The workload both arms measure (identical dict in both lines)."""
    n = args.size
    return {"workload": f"{n}x{n} fp64 planner-like costmap ({args.map}, seed 0), goal at (n/4, n/4), path from (3n/4, 3n/4): "
                        "one step = one full-field Eikonal solve + 1 gradient-descent path per GPU",
            "size": n, "map": args.map, "tau": 0.5,
            "l2_policy": "inputs larger than L2 (cost + T = %d MiB)" % (2 * n * n * 8 >> 20),
            "parallelism": "one query per GPU per step (replicas; weak scaling)",
            "e2e_call": "FastMarching.computeTmap + getPathGDM on NumPy arrays, one caller thread per GPU"}


def goals_for(c, n_ranks):
    """Rank r solves from its own goal; rank 0's is the survey's metric-map source (n/4, n/4)
    with the path traced from (3n/4, 3n/4)."""
    from planning_motion_planning_b200 import synth
    n = c.shape[0]
    fr = [(0.25, 0.25), (0.75, 0.25), (0.25, 0.75), (0.5, 0.5), (0.6, 0.2), (0.2, 0.6), (0.4, 0.8), (0.8, 0.4)]
    goals, starts = [], []
    for r in range(n_ranks):
        fx, fy = fr[0]        # every rank solves the metric-map query: weak scaling = identical per-GPU work
                              # (distinct goals per rank are exercised by the batched section and batch.py)
        goals.append(synth.free_cell_near(c, int(fx * n), int(fy * n)))
        starts.append(synth.free_cell_near(c, int((1 - fx) * n) if fx != 0.5 else int(0.1 * n), int((1 - fy) * n) if fy != 0.5 else int(0.1 * n)))
    return goals, starts


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz, self.thread, self.ok = None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.ok:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------ CPU arms
def cpu_port_run(c, goal, start, threads, sample_n):
    """Oracle port (C restatement of the reference FMM + tracer) on `threads` host threads,
    each solving the top-left sample_n x sample_n crop of the map from its own goal."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    from planning_motion_planning_b200 import synth
    crop = np.ascontiguousarray(c[:sample_n, :sample_n]).copy()
    crop[0, :] = crop[-1, :] = crop[:, 0] = crop[:, -1] = np.inf
    rng = np.random.default_rng(1)
    jobs = []
    for t in range(threads):
        g = synth.free_cell_near(crop, int(rng.integers(sample_n // 8, sample_n // 3)), int(rng.integers(sample_n // 8, sample_n // 3)))
        s = synth.free_cell_near(crop, sample_n - g[0], sample_n - g[1])
        jobs.append((g, s))

    def one(job):
        g, s = job
        T = O.computeTmap(crop, g)
        p, _ = O.getPathGDM(T, np.array(s), g, 0.5, return_status=True)
        return int(np.isfinite(T).sum()), len(p)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(one, jobs))
    dt = time.perf_counter() - t0
    return threads * sample_n * sample_n / dt, dt, res


def reference_arm(args):
    """Reference arm: the CPU implementation of the SAME work (one full-size solve + path per
    query) on ALL the host threads of the box: one host thread per query (a heap FMM is sequential, a
    single query cannot use more), as many queries at a time as there are threads.
    It is the C port of the reference (oracle): the reference itself is pure Python (~2e4 cells/s,
    one thread) and has nothing to compile."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.build()
    cores = len(os.sched_getaffinity(0))
    n = args.size
    c = make_map(n, args.map)
    # Every host thread the box offers, one full-size query per thread at a time (a heap FMM is sequential: one query
    # cannot use more than one thread, a job of many queries uses them all).  One step = `threads` queries.
    threads = min(cores, 32)
    nq = threads
    goals, starts = goals_for(c, max(nq, max(1, args.gpus) * max(1, args.inflight)))

    def one(i):
        T = O.computeTmap(c, goals[i])
        p, _ = O.getPathGDM(T, np.array(starts[i]), goals[i], 0.5, return_status=True)
        return len(p)

    def step(count, workers):
        with ThreadPoolExecutor(max_workers=workers) as ex:
            return list(ex.map(one, range(count)))

    small = min(n, 1024)          # warm-up on a crop: page in the library, spin up the pool
    crop = np.ascontiguousarray(c[:small, :small]).copy()
    crop[0, :] = crop[-1, :] = crop[:, 0] = crop[:, -1] = np.inf
    for _ in range(max(1, args.warmup)):
        O.computeTmap(crop, [small // 4, small // 4])
    # K steps of `threads` full-size queries each -- unless that would take more than a few minutes (a driver that asks for
    # many steps): then the steps after the first run on a centred crop of the same map (same algorithm, same threads;
    # crops are cache-friendlier, which only flatters this arm), sized so that the whole run stays within the budget
    budget_s, cells_done, crop_note = float(os.environ.get("FMB_REF_BUDGET_S", "150")), 0, ""
    t0 = time.perf_counter()
    step(nq, threads)
    t_first = time.perf_counter() - t0
    cells_done += nq * n * n
    m = n
    if args.steps > 1 and t_first * args.steps > budget_s:
        while m > 256 and t_first * (m / n) ** 2 * (args.steps - 1) > budget_s - t_first:
            m //= 2
    if m < n:
        lo = (n - m) // 2
        cc = np.ascontiguousarray(c[lo:lo + m, lo:lo + m]).copy()
        cc[0, :] = cc[-1, :] = cc[:, 0] = cc[:, -1] = np.inf
        from planning_motion_planning_b200 import synth as _synth
        gq, sq = _synth.free_cell_near(cc, m // 4, m // 4), _synth.free_cell_near(cc, 3 * m // 4, 3 * m // 4)

        def one_crop(i):
            T = O.computeTmap(cc, gq)
            O.getPathGDM(T, np.array(sq), gq, 0.5, return_status=True)

        for _ in range(args.steps - 1):
            with ThreadPoolExecutor(max_workers=threads) as ex:
                list(ex.map(one_crop, range(nq)))
            cells_done += nq * m * m
        crop_note = f"; steps 2..{args.steps} on a centred {m}x{m} crop to stay within {int(budget_s)} s"
    else:
        for _ in range(args.steps - 1):
            step(nq, threads)
            cells_done += nq * n * n
    total = time.perf_counter() - t0
    value = cells_done / total
    # context 1: one query on one thread (the latency a single caller of the reference path sees)
    t0 = time.perf_counter()
    step(1, 1)
    one_thread_value = n * n / (time.perf_counter() - t0)
    # context 2: every core busy on crops
    allc, _, _ = cpu_port_run(c, None, None, min(cores, 32), small)
    # context 3: the planner's own call, biComputeTmap (two fronts, real early exit) + the two half paths
    s_bi = starts[0]
    t0 = time.perf_counter()
    TG, TS, j = O.biComputeTmap(c, goals[0], s_bi)
    O.getPathGDM(TG, np.array(j, dtype=np.float64), goals[0], 0.5)
    O.getPathGDM(TS, np.array(j, dtype=np.float64), s_bi, 0.5)
    bi_s = time.perf_counter() - t0
    sample = (f"{nq} full {n}x{n} solve(s) + path per step, one host thread per query, all at a time "
              f"({threads} of {cores} cores; C port of FastMarching.py heap FMM + tracer){crop_note}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "one_query_one_thread": {"value": one_thread_value, "unit": UNIT, "cores": 1,
                                                  "sample": "one full-size solve + path on one thread (latency of a single caller)"},
                         "throughput_all_cores_crops": {"value": allc, "unit": UNIT, "cores": min(cores, 32),
                                                        "sample": f"{min(cores, 32)} threads x one {small}x{small} crop each"},
                         "planner_call_bicompute_plus_2_paths_s": bi_s},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ own arm
def bind_to_gpu_numa(local_rank):
    """Pin this process (and with it the first-touch placement of its pinned host buffers) to the CPUs of the NUMA node
    its GPU hangs off: at N = 8 every rank otherwise runs on node 0 and all H2D / D2H traffic crosses one memory
    controller (round-1 SCALE record: e2e efficiency 0.58 at 8 GPUs).  Best effort; returns what was done."""
    info = {"bound": False}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dvc = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        if bus is None:
            return info
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dvc)
        node = int(open(path).read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as e:          # no sysfs / no permission: measured as is
        info["error"] = type(e).__name__
    return info


def own_arm(args):
    import torch
    import torch.distributed as dist
    from planning_motion_planning_b200 import _capi, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa(local) if world > 1 else {"bound": False, "note": "single rank: not bound"}
    if world > 1:
        # torch.distributed.run exports OMP_NUM_THREADS=1: the drop-in's pageable -> pinned staging copies (torch CPU copies)
        # would then run on ONE core per rank (measured at N = 2: e2e efficiency 0.58 with it); give every rank its share
        share = max(1, len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
        torch.set_num_threads(share)
        affinity["torch_threads"] = share
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _capi.lib()          # fail loudly if the CUDA library is missing

    n = args.size
    c = make_map(n, args.map)
    goals, starts = goals_for(c, max(world, 1))
    goal, start = goals[rank], starts[rank]
    cells = n * n
    K, W = args.steps, args.warmup
    tau = 0.5

    cost_h = torch.from_numpy(c).pin_memory()
    cost_d = cost_h.to(dev)
    T_d = torch.empty((1, n, n), dtype=torch.float64, device=dev)
    seeds_d = torch.tensor([goal], dtype=torch.int32, device=dev)
    init = torch.tensor([start], dtype=torch.float64, device=dev)
    end = torch.tensor([goal], dtype=torch.float64, device=dev)

    # One step = one query = full-field solve + its path.  Queries are independent, and a single solve
    # is bound by its chain of tile visits, not by throughput (in windowed order it uses a
    # third of the SM slots), so `--inflight` queries run concurrently, each on its own stream with its
    # own workspace and field buffers; the path of a query is traced on a second stream while the next
    # query of that lane already solves.  Every solve and every trace of the K steps lies inside the
    # timed region; `latency_ms_one_query` reports one query run alone.
    NF = max(1, args.inflight)
    T_bufs = [[T_d if (f == 0) else torch.empty_like(T_d), torch.empty_like(T_d)] for f in range(NF)]
    s_solve = [torch.cuda.Stream(device=dev) for _ in range(NF)]
    s_trace = [torch.cuda.Stream(device=dev) for _ in range(NF)]
    ev_solved2 = [[torch.cuda.Event(), torch.cuda.Event()] for _ in range(NF)]
    ev_traced2 = [[torch.cuda.Event(), torch.cuda.Event()] for _ in range(NF)]

    def run_resident(K_):
        cur = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(cur)
        res = None
        for f in range(NF):
            s_solve[f].wait_event(start)
            for b in (0, 1):
                ev_traced2[f][b].record(s_solve[f])
        for k in range(K_):
            f, b = k % NF, (k // NF) & 1
            if f == 0:          # a wave of min(NF, what is left) solves shares the resident CTA slots
                _capi.set_options(concurrent_solves=min(NF, K_ - k))
            with torch.cuda.stream(s_solve[f]):
                s_solve[f].wait_event(ev_traced2[f][b])          # the trace that last read this buffer is done
                engine.solve2d(cost_d, seeds_d, out=T_bufs[f][b], nq=1, sync=False)
                ev_solved2[f][b].record(s_solve[f])
            with torch.cuda.stream(s_trace[f]):
                s_trace[f].wait_event(ev_solved2[f][b])
                res = engine.trace2d(T_bufs[f][b], init, end, tau)
                ev_traced2[f][b].record(s_trace[f])
        for f in range(NF):
            cur.wait_stream(s_solve[f])
            cur.wait_stream(s_trace[f])
        return res

    def step_resident():
        engine.solve2d(cost_d, seeds_d, out=T_d, nq=1, sync=False)
        return engine.trace2d(T_d, init, end, tau)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also validates: device-side failure raises here)
    for _ in range(max(W, 3)):
        out, cnt, st = step_resident()
        stats = engine.finish(dev)
    # NF solves share the GPU: each persistent grid takes 1/NF of the resident CTA slots (fmb_options.concurrent_solves)
    _capi.set_options(concurrent_solves=NF)
    run_resident(2 * NF)
    torch.cuda.synchronize()
    for f in range(NF):
        with torch.cuda.stream(s_solve[f]):
            engine.finish(dev)
    path_len = int(cnt[0])
    path_status = int(st[0])

    # ---- device-timed region: K steps, inputs resident (cost+T = 2*n*n*8 B > L2 for n >= 4096)
    sampler = ClockSampler(local if os.environ.get("CUDA_VISIBLE_DEVICES") is None else 0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    run_resident(K)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    t_ms = ev0.elapsed_time(ev1)
    with torch.cuda.stream(s_solve[0]):
        stats = engine.finish(dev)
    if world > 1:
        tt = torch.tensor([t_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = float(tt[0])
    value = world * cells * K / (t_ms * 1e-3)
    _capi.set_options(concurrent_solves=0)                 # one solve alone from here on: the default grid

    # ---- per-kernel durations (library CUDA events around the persistent kernel), K more steps
    k_ms, i_ms, tr_ms = [], [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(max(3, min(K, 10))):
        engine.solve2d(cost_d, seeds_d, out=T_d, nq=1, sync=False)
        s = engine.finish(dev)
        k_ms.append(s["solve_kernel_ms"]); i_ms.append(s["init_kernel_ms"])
        e0.record()
        engine.trace2d(T_d, init, end, tau)
        e1.record()
        torch.cuda.synchronize()
        tr_ms.append(e0.elapsed_time(e1))
    solve_ms, init_ms, trace_ms = float(np.mean(k_ms)), float(np.mean(i_ms)), float(np.mean(tr_ms))
    peak, peak_src = measured_peak()
    alg_bytes = 2 * 8 * cells
    achieved = alg_bytes / (solve_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:        # DRAM bytes of one launch of this kernel on this workload, from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if n == 4096 and args.map == "mars":
            ent = tj["solve2d_sweep_kernel<double,false>|4096x4096 mars seed0"]
            traffic, traffic_src = ent["dram_bytes_read"] + ent["dram_bytes_write"], ent["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "solve2d_sweep_kernel<double,false>", "kernel_ms": solve_ms,
                "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                "note": "single-source solve is dependency-latency bound (DESIGN.md 5); see batch.roofline_frac_solve_kernel for the throughput regime"}

    # ---- end to end through the public API with host buffers
    # Every step copies its costmap from pinned host memory, solves, traces, and copies the field,
    # the path and its length back.  Copies run on two side streams so that step k's field
    # download overlaps its own trace and step k+1's upload overlaps step k's solve (two device
    # input buffers); all of it is inside the timed region, which ends with a full synchronize.
    cap = int(round(15000 / tau)) + 2
    T_hs = [torch.empty((1, n, n), dtype=torch.float64).pin_memory() for _ in range(NF)]
    path_hs = [torch.empty((1, cap, 2), dtype=torch.float64).pin_memory() for _ in range(NF)]
    cnt_hs = [torch.empty(1, dtype=torch.int32).pin_memory() for _ in range(NF)]
    T_h, path_h, cnt_h = T_hs[0], path_hs[0], cnt_hs[0]
    main = torch.cuda.current_stream()
    s_ins = [torch.cuda.Stream(device=dev) for _ in range(NF)]
    s_outs = [torch.cuda.Stream(device=dev) for _ in range(NF)]
    cbufs = [[torch.empty_like(cost_d), torch.empty_like(cost_d)] for _ in range(NF)]
    mk = lambda: [[torch.cuda.Event(), torch.cuda.Event()] for _ in range(NF)]      # noqa: E731
    ev_in, ev_free, ev_solved, ev_T_out, ev_traced = mk(), mk(), mk(), mk(), mk()

    def upload(f, j):
        b = j & 1
        with torch.cuda.stream(s_ins[f]):
            s_ins[f].wait_event(ev_free[f][b])          # the solve that last read this buffer is done
            cbufs[f][b].copy_(cost_h, non_blocking=True)
            ev_in[f][b].record(s_ins[f])

    def run_e2e(K_):
        start = torch.cuda.Event()
        start.record(main)
        for f in range(NF):
            for st_ in (s_solve[f], s_ins[f]):
                st_.wait_event(start)
            for b in (0, 1):
                ev_free[f][b].record(s_solve[f])
                ev_T_out[f][b].record(s_solve[f])
                ev_traced[f][b].record(s_solve[f])
            if f < K_:
                upload(f, 0)
        for k in range(K_):
            f, j = k % NF, k // NF
            b = j & 1
            if f == 0:
                _capi.set_options(concurrent_solves=min(NF, K_ - k))
            with torch.cuda.stream(s_solve[f]):
                s_solve[f].wait_event(ev_in[f][b])
                s_solve[f].wait_event(ev_T_out[f][b])   # download and trace of this lane's step j-2 are done with the buffer
                s_solve[f].wait_event(ev_traced[f][b])
                engine.solve2d(cbufs[f][b], seeds_d, out=T_bufs[f][b], nq=1, sync=False)
                ev_free[f][b].record(s_solve[f])
                ev_solved[f][b].record(s_solve[f])
            if k + NF < K_:
                upload(f, j + 1)
            with torch.cuda.stream(s_outs[f]):
                s_outs[f].wait_event(ev_solved[f][b])
                T_hs[f].copy_(T_bufs[f][b], non_blocking=True)
                ev_T_out[f][b].record(s_outs[f])
            with torch.cuda.stream(s_trace[f]):
                s_trace[f].wait_event(ev_solved[f][b])
                out, cnt, st = engine.trace2d(T_bufs[f][b], init, end, tau)
                path_hs[f].copy_(out, non_blocking=True)
                cnt_hs[f].copy_(cnt, non_blocking=True)
                ev_traced[f][b].record(s_trace[f])
        torch.cuda.synchronize()

    def finish_lanes():
        st_all = None
        for f in range(NF):
            with torch.cuda.stream(s_solve[f]):
                r = engine.finish(dev)
            st_all = st_all or r
        return st_all

    _capi.set_options(concurrent_solves=NF)
    run_e2e(2 * NF)
    finish_lanes()
    sampler2 = ClockSampler(local if os.environ.get("CUDA_VISIBLE_DEVICES") is None else 0)
    barrier()
    sampler2.start()
    t0 = time.perf_counter()
    run_e2e(K)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks2 = sampler2.stop()
    if clocks2["samples"]:            # both timed regions feed the clock record
        allm = sampler.samples + sampler2.samples
        clocks = {"sm_mhz": float(np.median(allm)), "sm_max_mhz": clocks["sm_max_mhz"] or clocks2["sm_max_mhz"],
                  "reasons": sorted(set(clocks["reasons"]) | set(clocks2["reasons"])), "samples": len(allm),
                  "samples_resident_region": clocks["samples"], "samples_e2e_region": clocks2["samples"]}
    finish_lanes()
    _capi.set_options(concurrent_solves=0)
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt[0])
    e2e_pipelined = {"value": world * cells * K / e2e_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / K,
                     "h2d_bytes_per_step": int(cost_h.numel() * 8), "d2h_bytes_per_step": int(T_h.numel() * 8 + path_h.numel() * 8 + 4),
                     "api": "planning_motion_planning_b200.engine (pinned host buffers)",
                     "pipelining": f"{NF} queries in flight, each lane with its own upload / solve / download / trace streams and double buffers"}

    # ---- end to end through the reference-facing plugin call: NumPy in, NumPy out, ONE caller thread (what the planner
    # does, Coupled_motion_planner.py:1226-1230).  computeTmap with an unreachable start returns the full field (the same
    # work as `value`'s step); every host<->device copy (cost map up, field down, path down; the tracer runs on the device
    # copy the drop-in kept and a kernel checks the windows it read against the caller's array in place: ~1 MB over the
    # bus, not counted) is inside the timed region.
    import FastMarching.FastMarching as FM
    far = [-1, -1]
    c_np = c                                    # pageable NumPy array, as a caller would hold it
    for _ in range(2):
        Tn = FM.computeTmap(c_np, goal, far)
        pn = FM.getPathGDM(Tn, np.array(start, dtype=np.float64), goal, tau)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        Tn = FM.computeTmap(c_np, goal, far)
        pn = FM.getPathGDM(Tn, np.array(start, dtype=np.float64), goal, tau)
    torch.cuda.synchronize()
    plug_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([plug_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        plug_s = float(tt[0])
    e2e = {"value": world * cells * K / plug_s, "unit": UNIT, "ms_per_step": 1e3 * plug_s / K,
           "h2d_bytes_per_step": int(cells * 8), "d2h_bytes_per_step": int(cells * 8 + pn.size * 8),
           "api": "FastMarching.FastMarching.computeTmap(costMap, goal, start) + getPathGDM(T, start, goal, tau): NumPy in, NumPy out, "
                  "one caller thread, pageable host arrays",
           "path_rows": int(len(pn))}
    # the planner's own call: biComputeTmap (two full solves + the early-exit emulation) + the two half paths
    bi = None
    if rank == 0:
        s_bi = start
        for _ in range(2):
            TG, TS, jn = FM.biComputeTmap(c_np, goal, s_bi)
        t0 = time.perf_counter()
        TG, TS, jn = FM.biComputeTmap(c_np, goal, s_bi)
        t_bi = time.perf_counter() - t0
        pG = FM.getPathGDM(TG, jn, goal, tau)
        pS = FM.getPathGDM(TS, jn, s_bi, tau)
        torch.cuda.synchronize()
        bi = {"call": "FastMarching.biComputeTmap(costMap, goal, start) + 2 x getPathGDM, NumPy in / out",
              "bicompute_ms": 1e3 * t_bi, "total_ms": 1e3 * (time.perf_counter() - t0), "join": [int(v) for v in jn],
              "path_rows": [int(len(pG)), int(len(pS))]}

    # ---- batched independent queries (config 4 style): Q goal queries on one 512^2 map per GPU
    batch = None
    if not args.no_batch:
        from planning_motion_planning_b200 import synth
        Q = args.batch_queries
        cb = make_map(512, args.map, seed=1)
        rng = np.random.default_rng(100 + rank)
        ok = np.argwhere(np.isfinite(cb) & (cb <= 2.0))
        pick = ok[rng.integers(0, len(ok), size=Q)]
        seeds_b = torch.tensor(pick[:, ::-1].copy(), dtype=torch.int32, device=dev)
        cb_d = torch.from_numpy(cb).to(dev)
        Tb = torch.empty((Q, 512, 512), dtype=torch.float64, device=dev)
        starts_b = torch.tensor(ok[rng.integers(0, len(ok), size=Q)][:, ::-1].copy(), dtype=torch.float64, device=dev)
        ends_b = seeds_b.to(torch.float64)
        for _ in range(2):
            engine.solve2d(cb_d, seeds_b, out=Tb, nq=Q, sync=False)
            engine.trace2d(Tb, starts_b, ends_b, tau)
            sb = engine.finish(dev)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        b0.record()
        for _ in range(reps):
            engine.solve2d(cb_d, seeds_b, out=Tb, nq=Q, sync=False)
            engine.trace2d(Tb, starts_b, ends_b, tau)
        b1.record()
        barrier()
        bms = b0.elapsed_time(b1) / reps
        sb = engine.finish(dev)
        if world > 1:
            tt = torch.tensor([bms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            bms = float(tt[0])
        bcells = Q * 512 * 512
        batch = {"workload": f"{Q} goal queries per GPU on one 512x512 fp64 map (solve + 1 path each)",
                 "queries_per_s": world * Q / (bms * 1e-3), "ms_per_batch": bms,
                 "cells_per_s": world * bcells / (bms * 1e-3),
                 "solve_kernel_ms": sb["solve_kernel_ms"],
                 "roofline_frac_solve_kernel": (2 * 8 * bcells / (sb["solve_kernel_ms"] * 1e-3) / 1e9) / peak,
                 "evals_per_cell": sb["evals"] / bcells}

        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            # parity of the batch at full size: sampled queries against the oracle (field 1e-9, path 1e-3 cell)
            from oracle import oracle as O
            O.build()
            pth, pcnt, pst = engine.trace2d(Tb, starts_b, ends_b, tau)
            wf, wp, same = 0.0, 0.0, True
            for qi in rng.choice(Q, size=8, replace=False):
                gq = [int(v) for v in seeds_b[qi].tolist()]
                ref = O.computeTmap(cb, gq)
                got = Tb[qi].cpu().numpy()
                fin = np.isfinite(ref)
                same = same and bool(np.array_equal(np.isfinite(got), fin))
                wf = max(wf, float(np.max(np.abs(got[fin] - ref[fin]) / np.maximum(ref[fin], 1e-300))))
                po, so = O.getPathGDM(ref, starts_b[qi].cpu().numpy(), gq, tau, return_status=True)
                pg = pth[qi, :int(pcnt[qi])].cpu().numpy()
                same = same and int(pst[qi]) == so and pg.shape == po.shape
                if pg.shape == po.shape and len(po):
                    wp = max(wp, float(np.abs(pg - po).max()))
            batch["parity"] = {"sampled_queries": 8, "field_max_rel_err": wf, "path_max_abs_dev_cells": wp, "same_inf_pattern_status_and_rows": same}
            del pth
        del Tb

    # ---- config 4 as BASELINE words it: ONE job of Q goal queries on a 512^2 map, sharded over the N ranks through
    # batch.solve_queries (contiguous blocks, no data-path collective), the paths gathered on rank 0's host: strong scaling
    sharded = None
    if not args.no_batch:
        from planning_motion_planning_b200 import batch as BT
        Qs = args.batch_queries
        cbs = make_map(512, args.map, seed=1)
        rs = np.random.default_rng(100)                      # the SAME job on every rank
        oks = np.argwhere(np.isfinite(cbs) & (cbs <= 2.0))
        goals_s = oks[rs.integers(0, len(oks), size=Qs)][:, ::-1].tolist()
        starts_s = oks[rs.integers(0, len(oks), size=Qs)][:, ::-1].tolist()
        cbs_d = torch.from_numpy(cbs).to(dev)
        BT.solve_queries(cbs_d, goals_s[:64 * world], starts_s[:64 * world], chunk=64, gather=True)
        BT.solve_queries(cbs_d, goals_s, starts_s, chunk=1024, gather=True)          # same chunking as the timed call: memory pools of both lanes warm
        barrier()
        t0 = time.perf_counter()
        lo_s, res_s = BT.solve_queries(cbs_d, goals_s, starts_s, chunk=1024, gather=True)
        barrier()
        dts = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dts], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dts = float(tt[0])
        sharded = {"workload": f"ONE job of {Qs} goal queries on a 512x512 fp64 map (solve + 1 path each) sharded over {world} GPU(s) "
                               "through batch.solve_queries(gather=True): waypoints of all queries on rank 0's host at the end",
                   "queries_per_s": Qs / dts, "seconds": dts, "scaling": "strong",
                   "results_on_rank0": len(res_s) if rank == 0 else None}
        del res_s

    # ---- coupled plan, 3D half (config 3): arm-workspace cost volume solve + path
    vol = None
    if not args.no_3d and rank == 0:
        from planning_motion_planning_b200 import synth
        m3 = args.size3d
        cache = f"/tmp/fmb_vol_{m3}.npz"
        if os.path.exists(cache):
            z = np.load(cache)
            c3, g3, s3 = z["c"], z["g"].tolist(), z["s"].tolist()
        else:
            c3, g3, s3 = synth.arm_volume((m3, m3, m3), 0)
            try:
                tmp = f"{cache}.{os.getpid()}.tmp.npz"
                np.savez(tmp, c=c3, g=np.array(g3), s=np.array(s3))
                os.replace(tmp, cache)
            except Exception:
                pass
        c3d = torch.from_numpy(c3).to(dev)
        T3 = torch.empty((1, m3, m3, m3), dtype=torch.float64, device=dev)
        ks, ts = [], []
        for _ in range(4):
            engine.solve3d(c3d, [g3], out=T3, nq=1, sync=False)
            s3d = engine.finish(dev)
            ks.append(s3d["solve_kernel_ms"])
            e0.record()
            o3, n3, st3 = engine.trace3d(T3, [s3], [g3], tau)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        k3 = float(np.mean(ks[1:]))
        vol = {"workload": f"{m3}^3 fp64 planner-like arm-workspace volume: full-field solve + 1 path",
               "solve_kernel_ms": k3, "trace_ms": float(np.mean(ts[1:])), "cells_per_s": m3 ** 3 / (k3 * 1e-3),
               "roofline_frac_solve_kernel": (2 * 8 * m3 ** 3 / (k3 * 1e-3) / 1e9) / peak,
               "evals_per_cell": s3d["evals"] / m3 ** 3, "path_rows": int(n3[0]), "path_status": int(st3[0])}
        # the exact arithmetic the drop-in uses (the reference's rounding of `**2`), timed on its own: from scratch / as a polish pass
        for mode, key in ((True, "solve_exact_ms"), ("polish", "solve_plus_exact_polish_ms")):
            engine.solve3d(c3d, [g3], out=T3, nq=1, sync=True, exact=mode)
            e0.record()
            engine.solve3d(c3d, [g3], out=T3, nq=1, sync=False, exact=mode)
            e1.record()
            torch.cuda.synchronize()
            vol[key] = e0.elapsed_time(e1)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            O.build()
            t0 = time.perf_counter()
            r3 = O.computeTmap3D(c3, g3)
            p3o, s3o = O.getPathGDM3D(r3, np.uint32(s3), np.uint32(g3), tau, return_status=True)
            dt3 = time.perf_counter() - t0
            engine.solve3d(c3d, [g3], out=T3, nq=1, sync=True)
            o3, n3, st3 = engine.trace3d(T3, [s3], [g3], tau)
            g3f = T3[0].cpu().numpy()
            f3 = np.isfinite(r3)
            p3g = o3[0, :int(n3[0])].cpu().numpy()
            vol["cpu_port_seconds"] = dt3
            vol["parity"] = {"field_max_rel_err": float(np.max(np.abs(g3f[f3] - r3[f3]) / np.maximum(r3[f3], 1e-300))),
                             "same_inf_pattern": bool(np.array_equal(np.isfinite(g3f), f3)),
                             "path_rows": [int(len(p3g)), int(len(p3o))], "path_status": [int(st3[0]), int(s3o)],
                             "path_max_abs_dev_cells": float(np.abs(p3g - p3o).max()) if p3g.shape == p3o.shape else None}
        del c3d, T3

    # ---- the cost-map construction that precedes the 2D solve (DEM -> cost map)
    cmap = None
    if not args.no_costmap and rank == 0:
        from planning_motion_planning_b200 import costmap, synth
        res_c = 0.05
        Zc = torch.from_numpy(synth.crater_dem(n, res_c, 1)).to(dev)
        for _ in range(2):
            cm = costmap.build_costmap_device(Zc, res_c, n * res_c)
        cts = []
        for _ in range(5):
            e0.record()
            cm = costmap.build_costmap_device(Zc, res_c, n * res_c, sync=False)
            e1.record()
            torch.cuda.synchronize()
            cts.append(e0.elapsed_time(e1))
        cms = float(np.median(cts))
        cmap = {"workload": f"{n}x{n} crater DEM (seed 1, resolution {res_c}) -> planner cost map, 31 kernel launches",
                "ms": cms, "cells_per_s": cells / (cms * 1e-3),
                "roofline_frac": (2 * 8 * cells / (cms * 1e-3) / 1e9) / peak,
                "algorithmic_bytes": "read DEM once + write cost once = 16 B per cell"}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import costmap_oracle as CO
            nc = 1024
            Zs = synth.crater_dem(nc, res_c, 1)
            t0 = time.perf_counter()
            cref = CO.costmap2d(Zs, res_c, nc * res_c)
            dtc = time.perf_counter() - t0
            got = costmap.build_costmap_device(torch.from_numpy(Zs).to(dev), res_c, nc * res_c).cpu().numpy().T
            finc = np.isfinite(cref)
            cmap["cpu_reference"] = {"sample": f"{nc}x{nc} DEM through the cv2/scipy restatement of Coupled_motion_planner.py:1144-1216",
                                     "seconds": dtc, "cells_per_s": nc * nc / dtc,
                                     "parity_max_rel_err": float(np.max(np.abs(got[finc] - cref[finc]) / cref[finc])),
                                     "same_inf_pattern": bool(np.array_equal(np.isfinite(got), finc))}
        del Zc, cm

    # ---- the 3D cost-volume construction that precedes the 3D solve
    cvol = None
    if not args.no_costmap and rank == 0:
        from planning_motion_planning_b200 import costvolume as CVP
        sv = args.size3d
        rxy, rz = 2.0 / sv, 0.02
        rngv = np.random.default_rng(5)
        Zv = 0.2 + 0.1 * rngv.random((sv, sv))
        mpose = 200
        sp = np.linspace(0, 1, mpose)
        pathv = np.stack([(0.2 + 0.6 * sp) * sv * rxy, (0.3 + 0.4 * sp ** 2) * sv * rxy, 0.4 * sv * rz + 0.3 * np.sin(3 * sp)], axis=1)
        headv = np.stack([0.2 * np.sin(5 * sp), 0.15 * np.cos(4 * sp), 0.3 + 1.2 * sp], axis=1)
        finv, iniv = np.uint32([int(0.75 * sv), int(0.6 * sv), int(0.4 * sv)]), np.uint32([int(0.25 * sv), int(0.3 * sv), int(0.4 * sv)])
        cv_args = (Zv, rxy, rxy, rz, sv, sv, sv, 0.3, 0.4, 0.527, 0.2673, 0.1105, pathv, headv, finv, iniv)
        t0 = time.perf_counter()
        tun = CVP._tables(0.527, 0.2673, 0.1105, pathv, rxy, rz, headv)
        host_ms = 1e3 * (time.perf_counter() - t0)
        for _ in range(2):
            volume = CVP.build_cost_volume_device(*cv_args)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        volume = CVP.build_cost_volume_device(*cv_args)
        torch.cuda.synchronize()
        call_ms = 1e3 * (time.perf_counter() - t0)
        events = 2 * mpose * tun["nX"] * tun["nZ"] + tun["nX"] * tun["nZ"] + 9000 * (len(tun["lr"]) + 1)
        cvol = {"workload": f"{sv}^3 volume, {mpose} base poses, {events} scatter events (GetObstMap + TunnelCost + product)",
                "call_ms_incl_host_tables_and_uploads": call_ms, "host_tables_ms": host_ms,
                "tunnel_voxels": int(((volume != 20) & torch.isfinite(volume)).sum())}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import costvol as CVO
            t0 = time.perf_counter()
            want = CVO.GetObstMap(Zv, rxy, rxy, rz, sv, sv, sv, np.zeros((sv, sv)), 0.3, 0.4) * \
                CVO.TunnelCost(0.527, 0.2673, 0.1105, pathv, sv, sv, sv, rxy, rxy, rz, headv, finv, iniv)
            cvol["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
            cvol["bit_exact_vs_cpu_port"] = bool(np.array_equal(volume.cpu().numpy(), want))
            cvol["reference_note"] = "the Python reference needs 2.4 s for 12 poses at 40^3 (measured, oracle/gen_golden.py)"
        del volume

    # ---- CPU baseline beside it (rank 0, N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        t0 = time.perf_counter()
        Tref = O.computeTmap(c, goal)
        pref, pst = O.getPathGDM(Tref, np.array(start), goal, tau, return_status=True)
        dt = time.perf_counter() - t0
        Tg = np.asarray(Tn)                      # the field the plugin call returned
        fin = np.isfinite(Tref)
        same_inf = bool(np.array_equal(np.isfinite(Tg), fin))
        rel = float(np.max(np.abs(Tg[fin] - Tref[fin]) / np.maximum(Tref[fin], 1e-300)))
        gp = np.asarray(pn)
        pdev = float(np.abs(gp - pref).max()) if gp.shape == pref.shape else None
        cpu = {"value": cells / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"the full {n}x{n} solve + path, once, single thread (C restatement of the reference FMM; "
                         f"the reference itself is single-threaded CPython at ~2e4 cells/s)",
               "seconds": dt, "host_cores_available": len(os.sched_getaffinity(0)),
               "parity": {"field_max_rel_err": rel, "same_inf_pattern": same_inf,
                          "field_cells_bitwise_equal": float(np.mean(Tg[fin] == Tref[fin])), "path_rows": [len(gp), len(pref)],
                          "path_max_abs_dev_cells": pdev}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": make_config(args),
            "run": {"path_rows": path_len, "path_status": path_status,
                    "device_resident_pipelining": f"{NF} independent queries in flight per GPU (own streams and buffers); a query's path "
                                                  "is traced while the next query of its lane solves"},
            "latency_ms_one_query": init_ms + solve_ms + trace_ms,
            "breakdown_ms": {"init_fill": init_ms, "solve_kernel": solve_ms, "trace_kernel": trace_ms},
            "resident_loop": {"inflight": NF, "streams_per_lane": 2,
                              "concurrent_solves": "fmb_options.concurrent_solves = min(inflight, queries left in the wave): each solve's persistent "
                                                   "grid takes that share of the resident CTA slots; reset to 0 (one solve alone: two CTAs per SM) "
                                                   "for solve_kernel_ms / latency_ms_one_query / the plugin-call e2e"},
            "solver_stats": {k: stats[k] for k in ("tile_visits", "steps", "evals", "pushes", "cells_written")},
            "evals_per_cell": stats["evals"] / cells,
            "roofline": roofline, "e2e": e2e, "e2e_pipelined": e2e_pipelined, "planner_call": bi, "cpu_baseline": cpu, "batch": batch, "batch_sharded": sharded, "volume3d": vol, "costmap2d": cmap, "costvolume3d": cvol,
            "gpu_launches": 4 * K, "clocks": clocks,
            "host_affinity": affinity, "solver_options": _capi.get_options(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        own_arm(args)


if __name__ == "__main__":
    main()
