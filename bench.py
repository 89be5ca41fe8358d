#!/usr/bin/env python
"""bench.py -- headline benchmark of the grid-diffusion solve path (BASELINE.json).

One "step" = one complete harmonic fill of the workload (config C3 of SURVEY 8(d)): a
16384^2 RGB image, 50 % of the pixels unknown (64x64-pixel tiles, Bernoulli(0.5), seed 0,
pixel (0,0) opaque), solved to 1e-6 relative residual of the reduced free-pixel system.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

  value   M unknowns/s with the inputs resident in HBM (device-timed solve, CUDA events)
  e2e     the same through the public array API with HOST buffers (H2D + solve + D2H)
  roofline  the 5-point stencil kernel against the measured HBM copy bandwidth
  cpu_baseline  the CPU restatement of the reference (oracle/grid_oracle.py) on a bounded sample

With N > 1 (torchrun, one rank per GPU) ONE (16384 N) x 16384 grid is cut into N row slabs
(weak scaling: every GPU owns 16384 x 16384 pixels); ghost bands travel over NCCL / NVLink and the
CG scalars are all-reduced; value is the whole-job aggregate over the max-over-ranks device time.
"""
import argparse
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
CPU_SAMPLE_N = 2048       # cpu_baseline leg of the default arm: one fill of this size
REF_STEP_N = 1024         # --impl reference: each step is one fill of this size


def workload(n, seed=0, tile=64):
    """C3 generator: uint8 RGB image in [0,1] as float64, keep-mask in `tile`-pixel tiles."""
    rng = np.random.default_rng(seed)
    keep = np.kron(rng.random((n // tile, n // tile)) < 0.5, np.ones((tile, tile), bool))
    keep[0, 0] = True
    img8 = rng.integers(0, 256, size=(n, n, 3), dtype=np.uint8)
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


def cpu_fill(n):
    """One fill of an n x n sample of the workload with the CPU restatement of the reference."""
    from oracle import grid_oracle as go
    img8, keep = workload(n)
    img = img8 / 255.0
    t = time.perf_counter()
    go.fill_alpha_harmonic(img, keep)
    dt = time.perf_counter() - t
    return float((~keep).sum()) * 3, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = REF_STEP_N
    for _ in range(args.warmup):
        cpu_fill(n)
    times = []
    unknowns = 0.0
    for _ in range(args.steps):
        unknowns, dt = cpu_fill(n)
        times.append(dt)
    T = sum(times)
    v = unknowns * args.steps / T / 1e6
    sample = "harmonic fill of a %dx%d K=3 sample of the workload per step (vectorised scipy.sparse assembly + SuperLU)" % (n, n)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 harmonic alpha-fill, RGB, 50%% unknown in 64x64 tiles (CPU arm: %dx%d sample)" % (n, n)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=int(os.environ.get("HGRID_BENCH_SIZE", 16384)))
    ap.add_argument("--rtol", type=float, default=1e-6)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # A multi-rank run that stops making progress (rendezvous, NCCL bootstrap, a collective one rank never enters)
    # must not sit on N GPUs until somebody else's limit kills it: after HGRID_BENCH_WATCHDOG seconds (default 420; a normal run takes one to two minutes)
    # every rank dumps its Python stacks to stderr and exits.
    # The collectives of this path are tiny (K doubles, a few MB of ghost rows): NVLink SHARP (NVLS) buys nothing and its
    # multicast set-up is one more thing that can go wrong when communicators of different sizes are created back to back.
    if world > 1:
        os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("HGRID_BENCH_WATCHDOG", 420)), exit=True)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from harmonic_interpolation_b200 import solver as S
    from harmonic_interpolation_b200.distributed import SlabSolver

    # Weak scaling: every GPU owns a size x size band of ONE (size * N) x size grid, cut into row
    # slabs (ghost bands over NVLink, all-reduced CG scalars).  Each rank generates its own band.
    n, K = args.size, 3
    img8, keep = workload(n, seed=rank)
    if rank > 0:
        keep[0, 0] = bool(np.random.default_rng(1000 + rank).random() < 0.5)
    img = img8 / 255.0                       # float64 (n, n, 3), what fill_alpha_harmonic receives
    free_px = float((~keep).sum())
    unknowns = free_px * K

    t0 = time.perf_counter()
    s = SlabSolver(n * world, n, K, S.encode_flags(n, n, hard=keep), dist if world > 1 else None, args.precision,
                   device=dev if world > 1 else None)
    info = s.info
    setup_s = time.perf_counter() - t0
    s.upload(hard_vals=img)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    in_local, in_owned = s.pinned_u8()
    out_local, out_owned = s.pinned_u8()
    in_owned[...] = img8
    e2e_t = []
    for i in range(1 + min(args.steps, 3)):
        barrier()
        t = time.perf_counter()
        s.solve_u8(in_local, out_local, rtol=args.rtol, maxiter=500)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if i > 0:
            e2e_t.append(dt)
    t_e2e = torch.tensor([float(np.mean(e2e_t))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = float(tot_unknowns.item()) / float(t_e2e.item()) / 1e6
    assert (out_owned[keep] == img8[keep]).all()
    h2d_bytes, d2h_bytes = int(in_local.nbytes), int(out_local.nbytes)

    # ---- roofline of the stencil kernel (CUDA events on the solver's stream) -------------------
    # algorithmic bytes per launch: K*2W + 1 per pixel over the pixels of the tiles
    # the kernel touches (tiles without a free pixel are skipped; the full-grid figure is given too)
    W = 4 if args.precision == "fp32" else 8
    ms_apply = s.solver.bench_apply(20)
    act = np.add.reduceat(np.add.reduceat((~keep).astype(np.int64), np.arange(0, n, 32), axis=0), np.arange(0, n, 64), axis=1) > 0
    touched_px = float(act.sum()) * 32 * 64
    alg_bytes = touched_px * (K * 2 * W + 1)
    alg_bytes_full = float(s.local_rows) * n * (K * 2 * W + 1)
    peak, peak_src = peaks()
    achieved = alg_bytes / (ms_apply * 1e-3) / 1e9
    # DRAM traffic of that kernel per launch from the committed `ncu --set full` capture of this workload
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    kname = "sl_lap_kernel<%s, %s, 0, 0>" % (("float", "float") if W == 4 else ("double", "double"))
    if os.path.isfile(tp) and n == 16384 and world == 1:
        ent = json.load(open(tp)).get(kname)
        if ent:
            traffic, traffic_src = ent["dram_read_bytes"] + ent["dram_write_bytes"], ent["source"]

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        faulthandler.cancel_dump_traceback_later()
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
        "config": {"workload": "C3 harmonic alpha-fill %dx%d RGB, 50%% unknown in 64x64 tiles, rtol %g" % (n * world, n, args.rtol),
                   "parallelism": ("%d row slabs of %d rows, 128-row ghost bands over NCCL, all-reduced CG scalars" % (world, n)) if world > 1 else "single GPU",
                   "l2": "working set (%.1f GB per GPU) exceeds the 126 MB L2" % (alg_bytes_full / 1e9),
                   "solver": "MG-PCG, %d levels, %s storage, fp64 reductions and refinement" % (info.levels, args.precision),
                   "iterations": iters, "relres_true": relres, "setup_s": setup_s,
                   "time_to_solution_ms": ms_per_step, "free_pixels": free_px},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": d2h_bytes * world,
                "seconds_per_step": float(t_e2e.item()),
                "api": "hg_solve_u8 via SlabSolver.solve_u8: uint8 RGB in pinned host memory -> uint8 RGB (pattern uploaded once, as with the reference's factor-once closure)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": kname + "  (q = A p and the partial sums of p.q: the stencil launch of every CG iteration; + its general-chunk twin sl_lap_kernel<.., 1, 0> on the 2 % of chunks that touch the grid border)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms_apply, "peak_source": peak_src,
                     "frac_of_8TBs_nominal": achieved / 8000.0,
                     "bytes_basis": "pixels of the tiles the kernel touches (tiles without a free pixel are skipped)",
                     "full_grid_bytes_per_launch": alg_bytes_full, "full_grid_gbs": alg_bytes_full / (ms_apply * 1e-3) / 1e9},
    }
    faulthandler.cancel_dump_traceback_later()
    if not args.no_cpu_baseline:
        u, dt = cpu_fill(CPU_SAMPLE_N)
        line["cpu_baseline"] = {"value": u / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "one %dx%d K=3 fill of the same generator, %.1f s (scipy.sparse assembly + SuperLU, 1 thread; host has %d logical cores)" % (
                                    CPU_SAMPLE_N, CPU_SAMPLE_N, dt, os.cpu_count())}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
