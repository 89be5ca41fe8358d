"""Runs the non-headline configurations of BASELINE.json (C2, C4, C5) at their full sizes on one
GPU through the drop-in API and prints one JSON line per case.  The oracle cannot reach these sizes; each
case reports size-independent checks instead (hard values reproduced, true relative residual of the reduced
system as evaluated by the library in fp64, distance to the analytic surface for C4).

    python tools/run_configs.py [c2] [c2b] [c4] [c5]        (default: all)
"""
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from harmonic_interpolation_b200 import recovery as R
from harmonic_interpolation_b200 import solver as S


def emit(**kw):
    print(json.dumps(kw), flush=True)


def c2(n=4096, bilaplacian=False, per_equation=False, seed=0):
    """recovery.solve_grid_linear_simple3: hard 2 %, soft 5 %, cut edges from 16x16 blocks (30 %)."""
    while True:
        rng = np.random.default_rng(seed)
        hard = rng.random((n, n)) < 0.02
        soft = (rng.random((n, n)) < 0.05) & ~hard
        hv = rng.random((n, n))
        sv = rng.random((n, n))
        w = rng.uniform(0.1, 100.0, int(soft.sum())) if per_equation else 10.0
        block = np.kron(rng.random((n // 16 + 1, n // 16 + 1)) < 0.3, np.ones((16, 16), bool))[:n, :n]
        cuts = R.cut_planes_from_mask(block)
        hij = np.argwhere(hard); sij = np.argwhere(soft)
        t = time.perf_counter()
        try:
            with redirect_stdout(io.StringIO()):
                solve = R.solve_grid_linear_simple3_solver(n, n, hij, sij, w_lsq=w, cut_edges=cuts, bilaplacian=bilaplacian,
                                                           precision="fp64", rtol=1e-10, maxiter=20000)
        except AssertionError:
            seed += 1           # recovery.py:495 fired for this pattern: next seed
            continue
        t_setup = time.perf_counter() - t
        t = time.perf_counter()
        with redirect_stdout(io.StringIO()):
            x = solve(hv[hard][:, None], sv[soft][:, None])
        t_solve = time.perf_counter() - t
        t = time.perf_counter()
        with redirect_stdout(io.StringIO()):
            x2 = solve(1.0 - hv[hard][:, None], sv[soft][:, None])        # the factor-once closure: a second value set
        t_solve2 = time.perf_counter() - t
        break
    exact = float(np.abs(x[:, :, 0][hard] - hv[hard]).max())
    emit(config="C2", op="bilaplacian" if bilaplacian else "laplacian", n=n, K=1, seed=seed,
         w_lsq="per-equation U[0.1,100)" if per_equation else 10.0,
         hard=int(hard.sum()), soft=int(soft.sum()), cut_edges=int(cuts[0].sum() + cuts[1].sum()),
         setup_s=t_setup, solve_s_host_to_host=t_solve, second_solve_s=t_solve2, max_abs_err_at_hard=exact,
         finite=bool(np.isfinite(x).all() and np.isfinite(x2).all()), range=[float(x.min()), float(x.max())])


def c2_device(n=4096, bilaplacian=False):
    """Same pattern through GridSolver to read the device-side numbers (iterations, true residual, ms)."""
    rng = np.random.default_rng(0 if not bilaplacian else 0)
    hard = rng.random((n, n)) < 0.02
    soft = (rng.random((n, n)) < 0.05) & ~hard
    hv = rng.random((n, n, 1)); sv = rng.random((n, n, 1))
    block = np.kron(rng.random((n // 16 + 1, n // 16 + 1)) < 0.3, np.ones((16, 16), bool))[:n, :n]
    cd, cr = R.cut_planes_from_mask(block)
    for prec, rtol in (("fp64", 1e-10), ("fp32", 1e-6)):
        s = S.GridSolver(n, n, 1, bilaplacian, prec)
        try:
            info = s.set_pattern(S.encode_flags(n, n, hard=hard, soft=soft, cut_down=cd, cut_right=cr), None, 10.0, 0.0)
        except AssertionError as e:
            emit(config="C2-device", op="bilaplacian", n=n, skipped=str(e)[:80])
            return
        s.upload(hard_vals=hv, soft_vals=sv)
        st = s.solve_device(rtol, 20000)
        emit(config="C2-device", op="bilaplacian" if bilaplacian else "laplacian", n=n, precision=prec, rtol=rtol,
             levels=int(info.levels), iterations=st["iterations"], relres_true=st["relres"], solve_ms=st["solve_ms"],
             Munknowns_per_s=float((~hard).sum()) / st["solve_ms"] / 1e3)
        s.close()


def c4(n=8192, w_lsq=None, bumps=64, radius=8):
    """Height-field recovery: normals on a Bernoulli(0.02) subset of a smooth surface -> dx / dy constraints
    (weight 1e5 unless given) -> solve_grid_linear(bilaplacian=True) -> smooth_bumps(bilaplacian=True)."""
    rng = np.random.default_rng(0)
    ii, jj = np.indices((n, n), dtype=np.float64)
    A = 0.05 * n
    z = A * np.sin(2 * np.pi * ii / n) * np.cos(2 * np.pi * jj / n)
    del ii, jj
    sel = rng.random((n - 1, n - 1)) < 0.02
    ij = np.argwhere(sel)
    zx = z[ij[:, 0] + 1, ij[:, 1]] - z[ij[:, 0], ij[:, 1]]          # forward differences: what dx / dy constrain
    zy = z[ij[:, 0], ij[:, 1] + 1] - z[ij[:, 0], ij[:, 1]]
    normals = np.stack([-zx, -zy, np.ones(len(ij))], axis=1)
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    t = time.perf_counter()
    dx, dy = R.convert_grid_normal_constraints_to_dx_and_dy_constraints(n, n, (ij, normals))
    t_conv = time.perf_counter() - t
    pin = (np.asarray([[0, 0]]), np.asarray([z[0, 0]]))
    out = io.StringIO()
    t = time.perf_counter()
    with redirect_stdout(out):
        h = R.solve_grid_linear(n, n, dx_constraints=dx, dy_constraints=dy, value_constraints=pin, bilaplacian=True,
                                w_lsq=w_lsq, rtol=1e-8, maxiter=20000)
    t_solve = time.perf_counter() - t
    err = float(np.abs(h - z).max())
    locs = np.stack([rng.integers(radius, n - radius, bumps), rng.integers(radius, n - radius, bumps)], axis=1)
    bumped = h.copy()
    bumped[tuple(locs.T)] += 5.0
    t = time.perf_counter()
    with redirect_stdout(out):
        sm = R.smooth_bumps(bumped, locs, radius, bilaplacian=True, rtol=1e-8, maxiter=20000)
    t_bumps = time.perf_counter() - t
    mask = np.zeros((n, n), bool); mask[tuple(locs.T)] = True
    from harmonic_interpolation_b200.highlighting import grow
    mask = grow(mask, radius)
    emit(config="C4", n=n, w_lsq=1e5 if w_lsq is None else w_lsq, normal_constraints=int(len(ij)), convert_s=t_conv,
         solve_grid_linear_s=t_solve, max_abs_err_vs_analytic_surface=err, surface_amplitude=A,
         smooth_bumps_s=t_bumps, bump_pixels=int(mask.sum()),
         bumps_removed_max_residual=float(np.abs(sm - h)[mask].max()), outside_untouched=bool((sm[~mask] == bumped[~mask]).all()))


def c5(n=32768):
    """Harmonic fill 32768^2, K = 1, 64-pixel tile mask, one GPU (the multi-GPU form is bench.py --gpus N)."""
    rng = np.random.default_rng(0)
    keep = np.kron(rng.random((n // 64, n // 64)) < 0.5, np.ones((64, 64), bool))
    keep[0, 0] = True
    vals = rng.integers(0, 256, size=(n, n, 1), dtype=np.uint8)
    t = time.perf_counter()
    s = S.GridSolver(n, n, 1, False, "fp32", "multigrid")
    info = s.set_pattern(S.encode_flags(n, n, hard=keep))
    t_setup = time.perf_counter() - t
    out = np.empty_like(vals)
    s.solve_u8(vals, out, rtol=1e-6, maxiter=500)
    st = dict(s.last_stats)
    s.reset_guess()
    st2 = s.solve_device(1e-6, 500)
    free = float((~keep).sum())
    emit(config="C5", n=n, K=1, gpus=1, levels=int(info.levels), setup_s=t_setup, iterations=st2["iterations"],
         relres_true=st2["relres"], solve_ms=st2["solve_ms"], Munknowns_per_s=free / st2["solve_ms"] / 1e3,
         hard_reproduced=bool((out[keep] == vals[keep]).all()))
    s.close()


if __name__ == "__main__":
    which = [a.lower() for a in sys.argv[1:]] or ["c2", "c2b", "c4", "c5"]
    if "c2" in which:
        c2(4096, False, False)
        c2(4096, False, True)
        c2_device(4096, False)
    if "c2b" in which:
        c2(4096, True, False)
        c2_device(4096, True)
    if "c4" in which:
        c4(int(os.environ.get("C4_N", 8192)), float(os.environ["C4_W"]) if "C4_W" in os.environ else None)
    if "c5" in which:
        c5()
