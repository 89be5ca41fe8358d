#!/usr/bin/env python
"""bench.py -- headline benchmark of the grid-diffusion solve path (BASELINE.json).

One "step" = one complete harmonic fill of the workload, solved to 1e-6 relative residual of the
reduced free-pixel system (true residual, recomputed in fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c5-strong|c5-weak]

  workload c3 (default; BASELINE configs[2], the configuration the metric is quoted on): a 16384^2 RGB image per
      GPU, 50 % of the pixels unknown (64x64-pixel tiles, Bernoulli(0.5), pixel (0,0) opaque).  With N > 1 ONE
      (16384 N) x 16384 grid is cut into N row slabs (weak scaling).
  workload c5-strong (configs[4]): ONE 32768^2 K=1 grid of the same generator cut into N row slabs (strong scaling).
  workload c5-weak: 4096 rows x 32768 columns per GPU (SURVEY 8(d) C5, weak form).

  value     M unknowns/s with the inputs resident in HBM (device-timed solve, CUDA events, max over ranks)
  e2e       the same through the public image API with HOST buffers (H2D + solve + D2H inside the timed region);
            e2e_one_shot adds the pattern upload and hierarchy set-up (what the reference redoes on every call)
  kernels   per-phase device time of one solve (CUDA events on the solver's stream) with algorithmic bytes and
            achieved GB/s; roofline = the single kernel with the largest share of the solve
  cpu_baseline  the UNMODIFIED reference (oracle/_ref through oracle/ref_shim.py) on a bounded sample, host cores

With N > 1 (torchrun, one rank per GPU) ghost bands travel over NCCL / NVLink and the CG scalars are all-reduced;
before the timed region a small slab solve is compared with the single-GPU solve of the same grid (slab_check).
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "harmonic_fill_throughput_at_1e-6_rel_residual"
UNIT = "Munknowns/s"
CPU_SAMPLE_N = 1024       # cpu_baseline leg of the default arm: one fill of this size by the unmodified reference
REF_STEP_N = 1024         # --impl reference: each step is one fill of this size


def workload(rows, cols=None, K=3, seed=0, tile=64):
    """C3 / C5 generator: uint8 image (rows, cols, K), keep-mask in `tile`-pixel tiles, Bernoulli(0.5)."""
    cols = rows if cols is None else cols
    rng = np.random.default_rng(seed)
    keep = np.kron(rng.random(((rows + tile - 1) // tile, (cols + tile - 1) // tile)) < 0.5, np.ones((tile, tile), bool))[:rows, :cols]
    keep = np.ascontiguousarray(keep)
    keep[0, 0] = True
    img8 = rng.integers(0, 256, size=(rows, cols, K), dtype=np.uint8)
    return img8, keep


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summary of the samples that arrived during [t0, t1] (the timed region); if nvidia-smi was too slow to
        deliver one inside the window, the samples nearest to it are used and the fact is noted."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = []
        for t, s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                rows.append((t, float(f[0]), float(f[1]), f[2:6]))
            except ValueError:
                continue
        note = None
        sel = rows
        if t0 is not None and rows:
            sel = [r for r in rows if t0 - 0.15 <= r[0] <= t1 + 0.15]
            if not sel:
                mid = 0.5 * (t0 + t1)
                sel = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
                note = "no sample inside the timed window (%.2f s); nearest samples used" % (t1 - t0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in sel:
            for name, v in zip(names, r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": float(np.median([r[1] for r in sel])) if sel else None, "sm_max_mhz": max(r[2] for r in sel) if sel else None,
               "reasons": sorted(reasons), "samples": len(sel)}
        if note:
            out["note"] = note
        return out


# ---- CPU legs (the only places that touch oracle/) ---------------------------------------------------------------
def cpu_fill(n):
    """One fill of an n x n K=3 sample of the workload on the host: the UNMODIFIED reference
    (fill_alpha_harmonic.fill_alpha_harmonic -> recovery.solve_grid_linear_simple, per-pixel Python assembly loops,
    direct solve) when oracle/_ref or /root/reference is present, else the vectorised port.  Returns
    (unknowns, seconds, kind, description)."""
    from oracle import ref_shim
    img8, keep = workload(n)
    img = img8 / 255.0
    if ref_shim.have_reference():
        mods = ref_shim.load_reference()
        t = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            mods["fill_alpha_harmonic"].fill_alpha_harmonic(img, keep)
        dt = time.perf_counter() - t
        kind = "reference"
        how = ("the unmodified reference (fill_alpha_harmonic.fill_alpha_harmonic: Python assembly loops + direct solve; "
               "cvxopt/CHOLMOD is not installed, SuperLU stands in through oracle/ref_shim.py)")
    else:
        from oracle import grid_oracle as go
        t = time.perf_counter()
        go.fill_alpha_harmonic(img, keep)
        dt = time.perf_counter() - t
        kind = "port"
        how = "the vectorised port oracle/grid_oracle.py (scipy.sparse assembly + SuperLU); oracle/_ref is absent"
    return float((~keep).sum()) * 3, dt, kind, how


def cpu_workers():
    """How many fills the CPU legs run side by side: the reference is single-threaded Python + a single-threaded direct solve,
    so all the host threads it can use = one independent fill per logical core (a batch of images), capped so that the
    ~1.3 GB each fill needs stay within a quarter of the free host memory."""
    P = min(os.cpu_count() or 1, 64)
    try:
        avail_kb = [int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        P = min(P, max(1, int(0.25 * avail_kb / 1024.0 / 1536.0)))
    except Exception:
        pass
    return max(1, P)


def _cpu_fill_worker(n):
    os.environ["OMP_NUM_THREADS"] = "1"
    return cpu_fill(n)


def cpu_fill_parallel(n, P, rounds=1):
    """`rounds` times: P independent fills side by side (one freshly spawned process each: the parent may hold a CUDA context).  Returns (unknowns per round, [wall seconds per
    round], kind, description)."""
    import multiprocessing as mp
    if P <= 1:
        out, times = None, []
        for _ in range(rounds):
            out = cpu_fill(n)
            times.append(out[1])
        return out[0], times, out[2], out[3]
    times, res = [], None
    with mp.get_context("spawn").Pool(P) as pool:
        for _ in range(rounds):
            t = time.perf_counter()
            res = pool.map(_cpu_fill_worker, [n] * P)
            times.append(time.perf_counter() - t)
    return sum(r[0] for r in res), times, res[0][2], res[0][3]


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = REF_STEP_N
    P = cpu_workers()
    unknowns, times, kind, how = cpu_fill_parallel(n, P, args.warmup + args.steps)
    times = times[args.warmup:]
    T = sum(times)
    v = unknowns * args.steps / T / 1e6
    sample = "per step %d independent %dx%d K=3 fills of the C3 generator side by side (one process per logical core; host has %d), each by %s" % (
        P, n, n, os.cpu_count(), how)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 harmonic alpha-fill, RGB, 50%% unknown in 64x64 tiles (CPU arm: %dx%d sample; the reference cannot hold 16384^2: "
                               "134 M Python tuples and a direct factorisation beyond host memory)" % (n, n)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": P, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- slab self-check (N > 1): a small grid through the slab mode against rank 0's own single-GPU solve ---------------
def slab_check(dist, dev, rank, world):
    from harmonic_interpolation_b200 import solver as S
    from harmonic_interpolation_b200.distributed import SlabSolver
    n, cols, K = 512, 1024, 3
    rows = n * world
    keeps, imgs = [], []
    for q in range(world):
        i8, kq = workload(n, cols, K, seed=500 + q)
        if q > 0:
            kq[0, 0] = False
        keeps.append(kq); imgs.append(i8)
    keep = np.concatenate(keeps); img = np.concatenate(imgs) / 255.0
    flags = S.encode_flags(rows, cols, hard=keep)
    s = SlabSolver(rows, cols, K, flags[rank * n:(rank + 1) * n], dist, "fp32", device=dev)
    x = s.solve(img[rank * n:(rank + 1) * n], rtol=1e-9)
    its = s.last_stats["iterations"]
    import torch
    mine = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = None
    if rank == 0:
        got = np.concatenate([p.cpu().numpy() for p in parts])
        s1 = S.GridSolver(rows, cols, K, False, "fp32", "multigrid")
        s1.set_pattern(flags)
        x1 = s1.solve(img, rtol=1e-9)
        out = {"grid": "%dx%d K=%d in %d slabs, fp32, rtol 1e-9" % (rows, cols, K, world), "max_abs_diff_vs_single_gpu": float(np.abs(got - x1).max()),
               "iterations": its, "iterations_single_gpu": s1.last_stats["iterations"], "hard_pixels_reproduced": bool((got[keep] == img[keep]).all())}
        s1.close()
    del s
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("HGRID_BENCH_WORKLOAD", "c3"), choices=["c3", "c5-strong", "c5-weak"])
    ap.add_argument("--size", type=int, default=int(os.environ.get("HGRID_BENCH_SIZE", 0)))
    ap.add_argument("--rtol", type=float, default=1e-6)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-one-shot", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # A multi-rank run that stops making progress must not sit on N GPUs until somebody else's limit kills it: the
    # library aborts its communicator after HGRID_COMM_TIMEOUT seconds without progress (hgrid.cu: sync_watchful), and
    # after HGRID_BENCH_WATCHDOG seconds every rank dumps its Python stacks to stderr and exits.
    # The collectives of this path are tiny (K doubles, a few MB of ghost rows): NVLink SHARP (NVLS) buys nothing.
    if world > 1:
        os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
        os.environ.setdefault("HGRID_COMM_TIMEOUT", "45")
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("HGRID_BENCH_WATCHDOG", 300)), exit=True)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from harmonic_interpolation_b200 import _capi as C
    from harmonic_interpolation_b200 import solver as S
    from harmonic_interpolation_b200.distributed import SlabSolver, partition_rows

    check = slab_check(dist, dev, rank, world) if world > 1 else None

    # ---- the workload: every rank generates the band it owns ---------------------------------------------------
    if args.workload == "c3":
        n = args.size or 16384
        K, cols, own_rows, grows = 3, n, n, n * world
        scaling = "weak"
        wname = "C3 harmonic alpha-fill %dx%d RGB, 50%% unknown in 64x64 tiles, rtol %g" % (grows, cols, args.rtol)
    elif args.workload == "c5-strong":
        n = args.size or 32768
        K, cols, grows = 1, n, n
        r0, r1 = partition_rows(grows, world)[rank]
        own_rows = r1 - r0
        scaling = "strong"
        wname = "C5 harmonic fill %dx%d K=1 (one fixed grid in %d row slabs), 50%% unknown in 64x64 tiles, rtol %g" % (grows, cols, world, args.rtol)
    else:
        cols = args.size or 32768
        K, own_rows, grows = 1, 4096, 4096 * world
        scaling = "weak"
        wname = "C5 (weak form) harmonic fill %dx%d K=1, 4096 rows per GPU, 50%% unknown in 64x64 tiles, rtol %g" % (grows, cols, args.rtol)
    img8, keep = workload(own_rows, cols, K, seed=rank)
    if rank > 0:
        keep[0, 0] = bool(np.random.default_rng(1000 + rank).random() < 0.5)
    free_px = float((~keep).sum())
    unknowns = free_px * K
    flags = S.encode_flags(own_rows, cols, hard=keep)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_solver():
        return SlabSolver(grows, cols, K, flags, dist if world > 1 else None, args.precision, device=dev if world > 1 else None)

    barrier()
    t0 = time.perf_counter()
    s = make_solver()
    info = s.info
    setup_s = time.perf_counter() - t0
    in_local, in_owned = s.pinned_u8()
    out_local, out_owned = s.pinned_u8()
    in_owned[...] = img8
    s.solve_u8(in_local, out_local, rtol=args.rtol, maxiter=500)      # uploads the values (and is the first warm-up)
    assert (out_owned[keep] == img8[keep]).all()

    # ---- value: device-resident solve; every step restarts from the zero guess ----------------
    sampler = ClockSampler(local)
    sampler.start()                      # nvidia-smi needs a second or so to deliver its first line: start before the warm-up
    for _ in range(args.warmup):
        s.reset_guess()
        st = s.solve_device(args.rtol, 500)
    barrier()
    t_win0 = time.perf_counter()
    dev_ms, launches, iters, relres = [], 0, 0, 0.0
    for _ in range(args.steps):
        s.reset_guess()
        st = s.solve_device(args.rtol, 500)
        dev_ms.append(st["solve_ms"])
        launches += st["launches"] + 1
        iters = st["iterations"]
        relres = st["relres"]
    barrier()
    clocks = sampler.stop(t_win0, time.perf_counter())
    t_dev = torch.tensor([sum(dev_ms)], dtype=torch.float64, device="cuda")
    tot_unknowns = torch.tensor([unknowns], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_unknowns, op=dist.ReduceOp.SUM)
    ms_per_step = float(t_dev.item()) / args.steps
    value = float(tot_unknowns.item()) / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host buffers in, host buffers out, through the public image API (hg_solve_u8) ----------
    # every step: H2D of the uint8 image from pinned host memory, the solve, D2H of the uint8 result
    e2e_t, up_ms, down_ms = [], 0.0, 0.0
    for i in range(1 + min(args.steps, 3)):
        barrier()
        t = time.perf_counter()
        s.solve_u8(in_local, out_local, rtol=args.rtol, maxiter=500)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if i > 0:
            e2e_t.append(dt)
            up_ms, down_ms = s.last_stats["upload_ms"], s.last_stats["download_ms"]
    t_e2e = torch.tensor([float(np.mean(e2e_t))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = float(tot_unknowns.item()) / float(t_e2e.item()) / 1e6
    assert (out_owned[keep] == img8[keep]).all()
    h2d_bytes, d2h_bytes = int(in_local.nbytes), int(out_local.nbytes)

    # ---- per-phase table of one solve (CUDA events between the launches, hg_set_phase_timing) -------------------
    W = 4 if args.precision == "fp32" else 8
    lib = C.lib()
    C.check(lib.hg_set_phase_timing(s.solver._h, 1))
    s.reset_guess()
    stp = s.solve_device(args.rtol, 500)
    ph_ms = np.zeros(C.N_PHASES); ph_n = np.zeros(C.N_PHASES, np.int32)
    import ctypes
    C.check(lib.hg_get_phase_ms(s.solver._h, ph_ms.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                ph_n.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), C.N_PHASES))
    C.check(lib.hg_set_phase_timing(s.solver._h, 0))
    # pixels of the 32-row x 64-column blocks that hold a free pixel: what the tile-skipping kernels touch
    act = np.add.reduceat(np.add.reduceat((~keep).astype(np.int64), np.arange(0, own_rows, 32), axis=0), np.arange(0, cols, 64), axis=1) > 0
    touched_px = float(act.sum()) * 32 * 64
    peak, peak_src = peaks()
    # algorithmic bytes per touched level-0 pixel and launch; level 1 has a quarter of the pixels
    alg = {"stencil+alpha": (K * 2 * W + 1, "sl_lap_kernel: q = A p, p.q"),
           "x/r update+check": (K * 6 * W, "cg_update_tile_kernel: x += alpha p, r -= alpha q, r.r"),
           "fine down": (K * W + 1 + 0.25 * K * W, "sl_down_kernel: red-black GS from zero + residual + restriction"),
           "level-1 down": (0.25 * (K * 2 * W + 1 + 0.25 * K * W), "level-1 downward leg: 4-colour GS from zero + residual + restriction"),
           "level-1 up": (0.25 * (K * 3 * W + 1 + 0.25 * K * W), "level-1 upward leg: prolongation + reversed 4-colour GS"),
           "fine up": (K * 2 * W + 1 + 0.25 * K * W, "sl_up_kernel: prolongation + black-red GS + r.z"),
           "beta+p update": (K * 3 * W, "cg_pupdate_tile_kernel: p = z + beta p")}
    table, tot_ms = [], float(ph_ms.sum())
    for i, name in enumerate(C.PHASE_NAMES):
        if ph_n[i] == 0:
            continue
        row = {"phase": name, "ms_total": float(ph_ms[i]), "launches": int(ph_n[i]), "share": float(ph_ms[i]) / tot_ms}
        if name in alg:
            b = alg[name][0] * touched_px
            row.update({"kernel": alg[name][1], "algorithmic_bytes_per_launch": b, "ms_per_launch": float(ph_ms[i]) / int(ph_n[i]),
                        "achieved_gbs": b / (float(ph_ms[i]) / int(ph_n[i]) * 1e-3) / 1e9})
            row["frac"] = row["achieved_gbs"] / peak
        table.append(row)
    single = [r for r in table if "kernel" in r]
    dom = max(single, key=lambda r: r["ms_total"]) if single else None
    # DRAM traffic of that kernel per launch from the committed `ncu --set full` capture of this workload (profiles/traffic.json)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if dom and os.path.isfile(tp) and args.workload == "c3" and cols == 16384 and world == 1 and args.precision == "fp32":
        ent = json.load(open(tp)).get(dom["phase"])
        if ent:
            traffic, traffic_src = ent["dram_read_bytes"] + ent["dram_write_bytes"], ent["source"]
    # the stencil kernel alone, back to back (the number round 1 reported as `roofline`)
    ms_apply = s.solver.bench_apply(20)
    stencil_alone = {"ms_per_launch": ms_apply, "achieved_gbs": touched_px * (K * 2 * W + 1) / (ms_apply * 1e-3) / 1e9}
    stencil_alone["frac"] = stencil_alone["achieved_gbs"] / peak

    # ---- one shot: pattern upload + hierarchy set-up + H2D + solve + D2H, as a caller without a cached solver pays ----
    one_shot = None
    if not args.no_one_shot:
        del s
        torch.cuda.synchronize()
        barrier()
        t = time.perf_counter()
        s2 = make_solver()
        t_setup2 = time.perf_counter() - t
        in2, in2o = s2.pinned_u8(); out2, out2o = s2.pinned_u8()     # (pinned allocation is the caller's, not timed as set-up)
        in2o[...] = img8
        barrier()
        t = time.perf_counter()
        s2.solve_u8(in2, out2, rtol=args.rtol, maxiter=500)
        torch.cuda.synchronize()
        t_solve2 = time.perf_counter() - t
        tt = torch.tensor([t_setup2, t_solve2], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        one_shot = {"setup_s": float(tt[0].item()), "solve_s": float(tt[1].item()), "seconds": float(tt.sum().item()),
                    "value": float(tot_unknowns.item()) / float(tt.sum().item()) / 1e6, "unit": UNIT,
                    "what": "second construction in the same process (CUDA context and NCCL warm): flags -> hierarchy, then hg_solve_u8"}
        del s2

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        faulthandler.cancel_dump_traceback_later()
        return
    alg_bytes_full = float(own_rows + (0 if world == 1 else 128)) * cols * (K * 2 * W + 1)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
        "config": {"workload": wname,
                   "parallelism": ("%d row slabs, 128-row ghost bands over NCCL, all-reduced CG scalars" % world) if world > 1 else "single GPU",
                   "l2": "working set (%.1f GB per vector pass and GPU) exceeds the 126 MB L2" % (alg_bytes_full / 1e9),
                   "solver": "MG-PCG, %d levels, %s storage, fp64 reductions and refinement" % (info.levels, args.precision),
                   "iterations": iters, "relres_true": relres, "setup_s": setup_s,
                   "time_to_solution_ms": ms_per_step, "free_pixels": free_px, "slab_check": check},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": d2h_bytes * world,
                "seconds_per_step": float(t_e2e.item()), "h2d_ms_rank0": up_ms, "d2h_ms_rank0": down_ms,
                "h2d_gbs_rank0": h2d_bytes / max(up_ms, 1e-9) / 1e6, "d2h_gbs_rank0": d2h_bytes / max(down_ms, 1e-9) / 1e6,
                "api": "hg_solve_u8 via SlabSolver.solve_u8: uint8 image in pinned host memory -> uint8 image (pattern and hierarchy "
                       "uploaded once, as with the reference's factor-once closure; e2e_one_shot includes them)"},
        "e2e_one_shot": one_shot,
        "gpu_launches": int(launches),
        "kernels": table,
        "stencil_alone": stencil_alone,
    }
    if dom:
        line["roofline"] = {"bound": "hbm", "kernel": dom["kernel"] + "  [the single kernel with the largest share of the solve: %.0f %%]" % (100 * dom["share"]),
                            "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": traffic,
                            "traffic_source": traffic_src, "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                            "ms_per_launch": dom["ms_per_launch"], "peak_source": peak_src, "frac_of_8TBs_nominal": dom["achieved_gbs"] / 8000.0,
                            "bytes_basis": "pixels of the 32x64 blocks that hold a free pixel (blocks without one are skipped by every kernel)",
                            "timing": "CUDA events on the solver's stream around the launch inside a real solve (hg_set_phase_timing); includes the one-block scalar stage that follows"}
    if single:
        # the kernel furthest below the roofline, and the time-weighted figure over every kernel of the table
        worst = min(single, key=lambda r: r["frac"])
        line["roofline_worst"] = {"kernel": worst["kernel"], "share": worst["share"], "frac": worst["frac"], "achieved": worst["achieved_gbs"],
                                  "unit": "GB/s", "ms_per_launch": worst["ms_per_launch"]}
        tb = sum(r["algorithmic_bytes_per_launch"] * r["launches"] for r in single)
        tm = sum(r["ms_total"] for r in single)
        line["roofline_time_weighted"] = {"achieved": tb / (tm * 1e-3) / 1e9, "unit": "GB/s", "frac": tb / (tm * 1e-3) / 1e9 / peak,
                                          "share_of_solve": tm / tot_ms,
                                          "what": "algorithmic bytes of the kernels in `kernels` that carry a byte model / their device time"}
    faulthandler.cancel_dump_traceback_later()
    if not args.no_cpu_baseline:
        P = cpu_workers()
        u, times, kind, how = cpu_fill_parallel(CPU_SAMPLE_N, P, 1)
        dt = times[0]
        line["cpu_baseline"] = {"value": u / dt / 1e6, "unit": UNIT, "cores": P, "kind": kind,
                                "sample": "%d independent %dx%d K=3 fills of the same generator side by side (one process per logical core; host has %d), "
                                          "%.1f s, each by %s" % (P, CPU_SAMPLE_N, CPU_SAMPLE_N, os.cpu_count(), dt, how)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
