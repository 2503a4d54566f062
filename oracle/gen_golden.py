"""Generates tests/golden/*.npz by RUNNING THE REFERENCE (build container only).

TEST INFRASTRUCTURE.  The reference ships no golden vectors for the Fast Marching path
(SURVEY.md section 4), so parity is pinned on the reference itself: this script imports the
unmodified ``/root/reference/src/FastMarching`` modules (and the unmodified planner
``Coupled_motion_planner.main``) with numpy 2.3.5, runs them on seeded inputs and freezes

  * the inputs (or the seed that regenerates them),
  * small outputs verbatim (join node, paths, sample values, sums),
  * a sha256 of every full output field,

and, before writing anything, asserts that the C restatement (oracle/fmm_oracle.c) gives
the same bits.  tests/test_oracle_golden.py re-checks the C oracle against these files
everywhere (no reference needed); the GPU tests compare CUDA results with the C oracle.

    python oracle/gen_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from oracle import ref_loader as R      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rand_map(shape, seed):
    """The survey's KAT maps: 1 + 4 U(0,1), one-cell inf border."""
    rng = np.random.default_rng(seed)
    c = 1.0 + rng.random(shape) * 4
    if len(shape) == 2:
        c[0, :] = c[-1, :] = c[:, 0] = c[:, -1] = np.inf
    else:
        c[0] = c[-1] = np.inf
        c[:, 0] = c[:, -1] = np.inf
        c[:, :, 0] = c[:, :, -1] = np.inf
    return c


def plateau_map(n, seed):
    """Blocks of cost 1 / 150.5 / 301 (heavy exact ties, planner-like contrast)."""
    rng = np.random.default_rng(seed)
    blocks = rng.choice([1.0, 1.0, 1.0, 150.5, 301.0], size=(n // 8 + 1, n // 8 + 1))
    c = np.kron(blocks, np.ones((8, 8)))[:n, :n].copy()
    c[0, :] = c[-1, :] = c[:, 0] = c[:, -1] = np.inf
    return c


def same(a, b, what):
    if not np.array_equal(a, b, equal_nan=True):
        raise SystemExit(f"C oracle differs from the reference on {what}")


def gen_2d():
    cases = {}
    # KAT-1
    c = np.pad(np.ones((7, 7)), 1, constant_values=np.inf)
    T = R.computeTmap2D(c, [4, 4])
    same(O.computeTmap(c, [4, 4]), T, "KAT-1")
    cases["kat1_T"] = T
    # KAT-3 / 3b and friends
    specs = [("kat3", "rand", 100, 0, [25, 25], [10, 10], [90, 90]),
             ("rand64", "rand", 64, 7, [40, 12], [5, 50], [57, 9]),
             ("plateau80", "plateau", 80, 3, [20, 60], [8, 8], [70, 71])]
    for name, kind, n, seed, goal, g2, s2 in specs:
        c = rand_map((n, n), seed) if kind == "rand" else plateau_map(n, seed)
        T = R.computeTmap2D(c, goal)
        same(O.computeTmap(c, goal), T, name + " full")
        Tt = R.computeTmap2D(c, goal, s2)
        same(O.computeTmap(c, goal, s2), Tt, name + " early-exit")
        TG, TS, j = R.biComputeTmap(c, g2, s2)
        oTG, oTS, oj = O.biComputeTmap(c, g2, s2)
        same(oTG, TG, name + " TG"); same(oTS, TS, name + " TS"); same(oj, j, name + " join")
        pG = R.getPathGDM2D(TG, j, g2, 0.5)
        pS = R.getPathGDM2D(TS, j, s2, 0.5)
        same(O.getPathGDM(TG, j, g2, 0.5), pG, name + " pathG")
        same(O.getPathGDM(TS, j, s2, 0.5), pS, name + " pathS")
        pF = R.getPathGDM2D(T, np.array(s2), goal, 0.5)
        same(O.getPathGDM(T, np.array(s2), goal, 0.5), pF, name + " path full")
        cases.update({
            f"{name}_kind": kind, f"{name}_n": n, f"{name}_seed": seed,
            f"{name}_goal": goal, f"{name}_g2": g2, f"{name}_s2": s2,
            f"{name}_full_sha": sha(T), f"{name}_full_sum": float(np.sum(T[np.isfinite(T)])),
            f"{name}_early_sha": sha(Tt), f"{name}_TG_sha": sha(TG), f"{name}_TS_sha": sha(TS),
            f"{name}_join": j, f"{name}_pathG": pG, f"{name}_pathS": pS, f"{name}_pathF": pF,
            f"{name}_nfinG": int(np.isfinite(TG).sum()), f"{name}_nfinS": int(np.isfinite(TS).sum()),
        })
    np.savez_compressed(os.path.join(OUT, "ref2d.npz"), **cases)
    print("ref2d.npz:", len(cases), "entries")


def gen_3d():
    cases = {}
    c = np.ones((9, 9, 9))
    c[0] = c[-1] = np.inf; c[:, 0] = c[:, -1] = np.inf; c[:, :, 0] = c[:, :, -1] = np.inf
    T = R.computeTmap3D(c, [4, 4, 4])
    same(O.computeTmap3D(c, [4, 4, 4]), T, "KAT-2")
    cases["kat2_T"] = T
    specs = [("kat4", (24, 24, 24), 0, [5, 6, 7], [18, 17, 16], None),
             ("slab20", (20, 20, 20), 1, [4, 4, 4], [15, 15, 15], (slice(8, 10), slice(3, 15), slice(3, 15))),
             ("box", (14, 22, 18), 5, [3, 10, 4], [18, 3, 14], None)]
    for name, shape, seed, goal, start, slab in specs:
        c = rand_map(shape, seed)
        if slab is not None:
            c[slab] = np.inf
        T = R.computeTmap3D(c, goal)
        same(O.computeTmap3D(c, goal), T, name + " full")
        Tt = R.computeTmap3D(c, goal, np.uint32(start))
        same(O.computeTmap3D(c, goal, start), Tt, name + " early-exit")
        out = {}
        for tag, F in (("full", T), ("trunc", Tt)):
            try:
                p = R.getPathGDM3D(F, np.uint32(start), np.uint32(goal), 0.5)
                exc = ""
            except Exception as e:          # the exception type is observable behaviour
                p = np.zeros((0, 3)); exc = type(e).__name__
            po, st = O.getPathGDM3D(F, np.uint32(start), np.uint32(goal), 0.5, return_status=True)
            oexc = {0: "", 2: "ValueError", 3: "IndexError", 4: "OverflowError"}[st]
            if exc != oexc:
                raise SystemExit(f"{name} {tag}: reference raised {exc!r}, oracle {oexc!r}")
            if not exc:
                same(po, p, f"{name} path {tag}")
            out[f"{name}_path_{tag}"] = p
            out[f"{name}_exc_{tag}"] = exc
        cases.update(out)
        cases.update({
            f"{name}_shape": shape, f"{name}_seed": seed, f"{name}_goal": goal, f"{name}_start": start,
            f"{name}_slab": np.array([[s.start, s.stop] for s in slab]) if slab is not None else np.zeros((0, 2), int),
            f"{name}_full_sha": sha(T), f"{name}_full_sum": float(np.sum(T[np.isfinite(T)])),
            f"{name}_trunc_sha": sha(Tt), f"{name}_nfin_trunc": int(np.isfinite(Tt).sum()),
        })
    np.savez_compressed(os.path.join(OUT, "ref3d.npz"), **cases)
    print("ref3d.npz:", len(cases), "entries")


def gen_planner(n=200, res=0.05):
    """Run the UNMODIFIED planner main() on a synthetic DEM and record what it passes to /
    receives from the five FastMarching calls (Coupled_motion_planner.py:1226-1230,1636-1639)."""
    import importlib
    sys.path.insert(0, R.REF_SRC)
    sys.dont_write_bytecode = True
    for k in [k for k in sys.modules if k == "FastMarching" or k.startswith("FastMarching.")]:
        del sys.modules[k]
    cmp_ = importlib.import_module("Coupled_motion_planner")
    FM, FM3D = cmp_.FM, cmp_.FM3D
    calls = []

    def wrap(mod, fname, tag):
        orig = getattr(mod, fname)

        def f(*a):
            r = orig(*a)
            calls.append((tag, a, r))
            return r
        setattr(mod, fname, f)
    wrap(FM, "biComputeTmap", "bi"); wrap(FM, "getPathGDM", "path2d")
    wrap(FM3D, "computeTmap", "tmap3d"); wrap(FM3D, "getPathGDM", "path3d")
    size = n * res
    ax = (np.arange(n) + 0.5) * res
    X, Y = np.meshgrid(ax, ax)
    Z = 0.03 * np.sin(2 * np.pi * X / (0.5 * size)) * np.cos(2 * np.pi * Y / (0.7 * size))
    Z += 0.5 * np.exp(-((X - 0.5 * size) ** 2 + (Y - 0.45 * size) ** 2) / (2 * (0.06 * size) ** 2))
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "PRL_DEM.txt"), "w") as f:
        for row in Z:
            f.write(",".join(repr(float(v)) for v in row) + "\n")
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cmp_.main(0.8 * size, 0.8 * size, 0.2 * size, 0.2 * size, 0.0, d + "/", res, size)
        status = "completed"
    except Exception as e:
        status = f"raised {type(e).__name__} after {len(calls)} FM calls: {e}"
    print("planner main():", status)
    tags = [c[0] for c in calls]
    print("captured calls:", tags)
    if tags[:3] != ["bi", "path2d", "path2d"]:
        raise SystemExit("planner did not reach the 2D FM calls")
    out = {"status": status, "n": n, "res": res}
    _, (cmapT, goal, start), (TG, TS, join) = calls[0]
    out["bi_cost"] = np.asarray(cmapT)         # F-ordered view in the planner; stored as values
    out["bi_cost_fortran"] = bool(cmapT.flags.f_contiguous and not cmapT.flags.c_contiguous)
    out["bi_goal"] = np.array(goal); out["bi_start"] = np.array(start)
    out["bi_join"] = join; out["bi_TG_sha"] = sha(TG); out["bi_TS_sha"] = sha(TS)
    oTG, oTS, oj = O.biComputeTmap(np.asarray(cmapT), goal, start)
    same(oTG, TG, "planner TG"); same(oTS, TS, "planner TS"); same(oj, join, "planner join")
    for k, idx in (("pathG", 1), ("pathS", 2)):
        _, (Tm, init, end, tau), p = calls[idx]
        out[f"{k}_init"] = np.asarray(init, dtype=np.float64); out[f"{k}_end"] = np.asarray(end, dtype=np.float64)
        out[f"{k}_tau"] = tau; out[k] = p
        same(O.getPathGDM(np.asarray(Tm), init, end, tau), p, "planner " + k)
    if "tmap3d" in tags:
        i3 = tags.index("tmap3d")
        _, (c3, g3, s3), T3 = calls[i3]
        out["c3"] = np.asarray(c3); out["g3"] = np.asarray(g3, dtype=np.int64); out["s3"] = np.asarray(s3, dtype=np.int64)
        out["T3_sha"] = sha(T3); out["T3_nfin"] = int(np.isfinite(T3).sum())
        same(O.computeTmap3D(np.asarray(c3), g3, s3), T3, "planner Tmap3D")
        if "path3d" in tags:
            _, (Tm, init, end, tau), p = calls[tags.index("path3d")]
            out["path3d_init"] = np.asarray(init, dtype=np.float64); out["path3d_end"] = np.asarray(end, dtype=np.float64)
            out["path3d_tau"] = tau; out["path3d"] = p
            same(O.getPathGDM3D(np.asarray(Tm), init, end, tau), p, "planner path3d")
    np.savez_compressed(os.path.join(OUT, "planner_calls.npz"), **out)
    print("planner_calls.npz: cost", out["bi_cost"].shape, "fortran", out["bi_cost_fortran"],
          "3D", out.get("c3", np.zeros(0)).shape, "size %.0f kB" % (os.path.getsize(os.path.join(OUT, "planner_calls.npz")) / 1e3))


def gen_costvolume(n=200, res=0.05):
    """3D cost-volume builder: arguments and results of the UNMODIFIED
    GetObstMap / TunnelCost (Coupled_motion_planner.py:319, :505) as called by the unmodified
    planner main() on the synthetic DEM of gen_planner, plus seeded direct calls of the same
    functions with rolled / pitched base frames.  The C restatement (oracle/costvol_oracle.c) must
    give the same bits before anything is written."""
    import importlib
    from oracle import costvol as CV
    sys.path.insert(0, R.REF_SRC)
    sys.dont_write_bytecode = True
    for k in [k for k in sys.modules if k == "FastMarching" or k.startswith("FastMarching.")]:
        del sys.modules[k]
    cmp_ = importlib.import_module("Coupled_motion_planner")
    rec = {}

    def wrap(fname):
        orig = getattr(cmp_, fname)

        def f(*a):
            r = orig(*a)
            rec[fname] = (a, r)
            return r
        setattr(cmp_, fname, f)
        return orig
    o1, o2 = wrap("GetObstMap"), wrap("TunnelCost")
    size = n * res
    ax = (np.arange(n) + 0.5) * res
    X, Y = np.meshgrid(ax, ax)
    Z = 0.03 * np.sin(2 * np.pi * X / (0.5 * size)) * np.cos(2 * np.pi * Y / (0.7 * size))
    Z += 0.5 * np.exp(-((X - 0.5 * size) ** 2 + (Y - 0.45 * size) ** 2) / (2 * (0.06 * size) ** 2))
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "PRL_DEM.txt"), "w") as f:
        for row in Z:
            f.write(",".join(repr(float(v)) for v in row) + "\n")
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cmp_.main(0.8 * size, 0.8 * size, 0.2 * size, 0.2 * size, 0.0, d + "/", res, size)
    except Exception as e:
        print("planner main() raised after the builders:", type(e).__name__)
    cmp_.GetObstMap, cmp_.TunnelCost = o1, o2
    out = {}
    (Zs, resX, resY, resZ, sX, sY, sZ, ob, xm, ym), (final, _, _) = rec["GetObstMap"]
    out.update(p_Zs=np.asarray(Zs), p_res=np.array([resX, resY, resZ]), p_shape=np.array([sX, sY, sZ]),
               p_obst=np.asarray(ob), p_xm_ym=np.array([xm, ym]), p_final=final)
    same(CV.GetObstMap(Zs, resX, resY, resZ, sX, sY, sZ, ob, xm, ym), final, "planner GetObstMap")
    (rlim, rO, rm, g2, sX, sY, sZ, resX, resY, resZ, head, fin, ini), cm2 = rec["TunnelCost"]
    out.update(p_radii=np.array([rlim, rO, rm]), p_path=np.asarray(g2), p_heading=np.asarray(head),
               p_fin=np.asarray(fin, dtype=np.int64), p_ini=np.asarray(ini, dtype=np.int64), p_tunnel=cm2)
    same(CV.TunnelCost(rlim, rO, rm, g2, sX, sY, sZ, resX, resY, resZ, head, fin, ini), cm2, "planner TunnelCost")
    print("planner volume", (sX, sY, sZ), "tunnel cells", int((cm2 != 10).sum()), "path points", len(g2))
    for s in range(2):                                  # direct calls, tilted frames
        rng = np.random.default_rng(40 + s)
        sX = sY = 36 + 8 * s
        sZ, m = 26 + 6 * s, 9 + 4 * s
        resX = resY = 0.02 * rng.uniform(0.9, 1.1)
        resZ = 0.02
        Zs = 0.1 + 0.05 * rng.random((sX, sY))
        ob = (rng.random((sX, sY)) < 0.1).astype(float)
        xm, ym = resX * 7, resY * 9
        final = cmp_.GetObstMap(Zs, resX, resY, resZ, sX, sY, sZ, ob, xm, ym)[0]
        same(CV.GetObstMap(Zs, resX, resY, resZ, sX, sY, sZ, ob, xm, ym), final, f"direct GetObstMap {s}")
        path = np.stack([np.linspace(0.2, 0.6, m) * sX * resX, np.linspace(0.3, 0.5, m) * sY * resY,
                         0.3 + 0.02 * rng.random(m)], axis=1)
        head = np.stack([0.15 * rng.normal(size=m), 0.15 * rng.normal(size=m),
                         rng.uniform(-1, 1) + 0.05 * np.arange(m)], axis=1)
        fin, ini = np.uint32([sX // 2, sY // 2, 5]), np.uint32([sX // 3, sY // 3, 12])
        rlim, rO, rm = 0.2, 0.12, 0.05
        cm2 = cmp_.TunnelCost(rlim, rO, rm, path, sX, sY, sZ, resX, resY, resZ, head, fin, ini)
        same(CV.TunnelCost(rlim, rO, rm, path, sX, sY, sZ, resX, resY, resZ, head, fin, ini), cm2, f"direct TunnelCost {s}")
        out.update({f"d{s}_Zs": Zs, f"d{s}_res": np.array([resX, resY, resZ]), f"d{s}_shape": np.array([sX, sY, sZ]),
                    f"d{s}_obst": ob, f"d{s}_xm_ym": np.array([xm, ym]), f"d{s}_final": final,
                    f"d{s}_radii": np.array([rlim, rO, rm]), f"d{s}_path": path, f"d{s}_heading": head,
                    f"d{s}_fin": fin.astype(np.int64), f"d{s}_ini": ini.astype(np.int64), f"d{s}_tunnel": cm2})
    np.savez_compressed(os.path.join(OUT, "costvolume.npz"), **out)
    print("costvolume.npz: %.0f kB" % (os.path.getsize(os.path.join(OUT, "costvolume.npz")) / 1e3))


def gen_pathpost(n=200, res=0.05):
    """Pins oracle/pathpost_oracle.py on the reference itself:
      * stitch (Coupled_motion_planner.py:1232-1234): the UNMODIFIED planner main() is run on the synthetic DEM of
        gen_planner and the local ``roverPath`` of main() is read at its return (profile hook);
      * smoothing + resampling (:1641-1671): on that DEM the arm path is shorter than the 11-tap window (scipy raises
        ValueError inside main(), frozen as the status-1 case), so the reference's OWN source lines 1641-1671 are read
        from the reference tree at run time and executed, unmodified, on longer seeded paths.
    Asserts the oracle reproduces both bit for bit, freezes inputs and outputs in tests/golden/pathpost.npz."""
    import importlib
    from scipy import interpolate, signal
    from oracle import pathpost_oracle as PO
    sys.path.insert(0, R.REF_SRC)
    sys.dont_write_bytecode = True
    for k in [k for k in sys.modules if k == "FastMarching" or k.startswith("FastMarching.")]:
        del sys.modules[k]
    cmp_ = importlib.import_module("Coupled_motion_planner")
    calls = []

    def wrap(mod, fname, tag):
        orig = getattr(mod, fname)

        def f(*a):
            r = orig(*a)
            calls.append((tag, a, r))
            return r
        setattr(mod, fname, f)
    wrap(cmp_.FM, "getPathGDM", "path2d"); wrap(cmp_.FM3D, "getPathGDM", "path3d")
    grabbed = {}
    main_code = cmp_.main.__code__

    def prof(frame, event, arg):
        if event == "return" and frame.f_code is main_code:
            grabbed.update(frame.f_locals)
    size = n * res
    ax = (np.arange(n) + 0.5) * res
    X, Y = np.meshgrid(ax, ax)
    Z = 0.03 * np.sin(2 * np.pi * X / (0.5 * size)) * np.cos(2 * np.pi * Y / (0.7 * size))
    Z += 0.5 * np.exp(-((X - 0.5 * size) ** 2 + (Y - 0.45 * size) ** 2) / (2 * (0.06 * size) ** 2))
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "PRL_DEM.txt"), "w") as f:
        for row in Z:
            f.write(",".join(repr(float(v)) for v in row) + "\n")
    sys.setprofile(prof)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cmp_.main(0.8 * size, 0.8 * size, 0.2 * size, 0.2 * size, 0.0, d + "/", res, size)
        status = "completed"
    except Exception as e:
        status = f"raised {type(e).__name__}: {e}"
    finally:
        sys.setprofile(None)
    print("planner main():", status)
    L = grabbed
    if "roverPath" not in L:
        raise SystemExit("planner did not reach the rover-path stitch")
    paths2d = [c for c in calls if c[0] == "path2d"]
    pathG, pathS = paths2d[0][2], paths2d[1][2]
    out = {"status": status, "resolution": float(L["resolution"]), "pathG": pathG, "pathS": pathS}
    # roverPath is trimmed in place after :1234 (waypoints near the sample / the rover are deleted, :1236-1247) and gets
    # a z column (:1250): every row the planner kept must be a row of the stitched array, bit for bit and in order
    stitched = PO.stitch_rover_path(pathS, pathG, out["resolution"])
    rp = np.asarray(L["roverPath"], dtype=np.float64)
    j = 0
    for row in rp:
        while j < len(stitched) and not np.array_equal(stitched[j], row[:2]):
            j += 1
        if j == len(stitched):
            raise SystemExit("stitched rover path does not contain the planner's roverPath rows")
        j += 1
    out["roverPath_trimmed"] = rp
    p3 = [c for c in calls if c[0] == "path3d"]
    if p3:
        out["planner_path3d"] = np.asarray(p3[0][2], dtype=np.float64)      # shorter than the window: ValueError in main()
    # the reference's own lines 1641-1671, executed unmodified
    with open(os.path.join(R.REF_SRC, "Coupled_motion_planner.py")) as f:
        lines = f.read().split("\n")
    import textwrap
    block = textwrap.dedent("\n".join(lines[1640:1671]))
    assert block.lstrip().startswith("gamma3D[:, 0] = gamma3D[:, 0]*resX") and "resizedGamma3D[:, 2] = fz(" in block, block[:200]
    code = compile(block, "Coupled_motion_planner.py:1641-1671", "exec")
    rng = np.random.default_rng(7)
    ncase = 0
    for (rows, m) in ((11, 25), (12, 7), (57, 57), (57, 200), (300, 41), (23, 1), (16, 2)):
        t = np.linspace(0.0, 1.0, rows)
        path = np.stack([5 + 40 * t + rng.normal(0, 0.4, rows), 30 - 12 * t ** 2 + rng.normal(0, 0.4, rows),
                         8 + 6 * np.sin(3 * t) + rng.normal(0, 0.2, rows)], axis=1)
        res3 = np.array([0.03 + 0.01 * rng.random(), 0.03 + 0.01 * rng.random(), 0.02 + 0.01 * rng.random()])
        off3 = rng.normal(0, 2.0, 3)
        last3 = rng.normal(0, 2.0, 3)
        ns = {"gamma3D": path.copy(), "resX": res3[0], "resY": res3[1], "resZ": res3[2], "signal": signal,
              "interpolate": interpolate, "np": np, "effectorBasePath": np.zeros((3, 3)), "Xmin": off3[0], "Ymin": off3[1],
              "Zmin": off3[2], "xm": last3[0], "ym": last3[1], "zm": last3[2], "finalBasePath": np.zeros((m + 4, 3)), "index": 5}
        exec(code, ns)
        ref = np.asarray(ns["resizedGamma3D"], dtype=np.float64)
        assert ref.shape == (m, 3), ref.shape
        same(PO.smooth_resample_arm(path, res3, off3, last3, m), ref, f"reference lines 1641-1671, case {ncase}")
        for k, v in (("path", path), ("res3", res3), ("off3", off3), ("last3", last3), ("resized", ref)):
            out[f"post{ncase}_{k}"] = v
        ncase += 1
    out["npost"] = ncase
    np.savez_compressed(os.path.join(OUT, "pathpost.npz"), **out)
    print("pathpost.npz: pathS", pathS.shape, "pathG", pathG.shape, "roverPath rows kept", len(rp), "of", len(stitched),
          "| post cases", ncase)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if not R.available():
        raise SystemExit("reference tree not found: run this in the build container")
    which = sys.argv[1:] or ["2d", "3d", "planner", "costvolume", "pathpost"]
    if "2d" in which:
        gen_2d()
    if "3d" in which:
        gen_3d()
    if "planner" in which:
        gen_planner()
    if "costvolume" in which:
        gen_costvolume()
    if "pathpost" in which:
        gen_pathpost()
