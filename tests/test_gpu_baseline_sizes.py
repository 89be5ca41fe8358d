"""GPU: parity at the sizes BASELINE.md section 4 promises, against the CPU oracle run on the box (scipy is there):

  C2  recovery.solve_grid_linear_simple3 on the SURVEY 8(d) generator -- 2 % hard, 5 % soft, cut edges around 16 x 16
      blocks, scalar and per-equation w_lsq, Laplacian and bilaplacian -- at 256^2, 512^2, 1024^2
  C3  the harmonic tile fill at 1024^2 and 2048^2, K = 3: parity settings, and the benchmark's own settings (fp32,
      rtol 1e-6) with the error that stopping rule buys reported
  C4  normals -> solve_grid_linear(bilaplacian, default w_lsq = 1e5) -> smooth_bumps at 256^2 and 512^2
  the reference's own 250^2 self-test (recovery.py:2020-2077) through the module's `python3 recovery.py` entry.

Tolerances are north_star's -- max abs error <= 1e-9 of the value range in fp64 mode, <= 1e-5 in fp32 mode -- except
where DESIGN.md section 8 ("Deviations") records why a looser bound is the attainable one; those cases say so here.
"""
import contextlib
import io
import os

import numpy as np
import pytest

from conftest import rel_err
from harmonic_interpolation_b200 import recovery as R
from harmonic_interpolation_b200 import solver as S
from oracle import grid_oracle as go

pytestmark = pytest.mark.gpu


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def c2_problem(n, bilaplacian, per_equation, seed=0):
    """SURVEY 8(d) C2 generator; for the bilaplacian the next seed is taken if recovery.py:495 fires."""
    for _ in range(8):
        rng = np.random.default_rng(seed)
        p = go.GridProblem(n, n, 1, bilaplacian)
        p.hard = rng.random((n, n)) < 0.02
        soft = (rng.random((n, n)) < 0.05) & ~p.hard
        p.hard_vals = rng.random((n, n, 1))
        p.soft_vals = rng.random((n, n, 1))
        p.soft_w = np.where(soft, rng.uniform(0.1, 100.0, (n, n)) if per_equation else 10.0, 0.0)
        block = np.kron(rng.random((n // 16 + 1, n // 16 + 1)) < 0.3, np.ones((16, 16), bool))[:n, :n]
        p.cut_down, p.cut_right = R.cut_planes_from_mask(block)
        p.have_cuts = True
        try:
            ref = go.solve_problem(p)
            return p, soft, ref, seed
        except AssertionError:
            seed += 1
    raise RuntimeError("no admissible C2 pattern within 8 seeds")


@pytest.mark.parametrize("n,bilaplacian,per_equation", [
    (256, False, False), (256, False, True), (512, False, True), (1024, False, False), (1024, False, True),
    (256, True, False), (256, True, True), (512, True, False), (1024, True, True)])
def test_c2_simple3_generator(n, bilaplacian, per_equation):
    p, soft, ref, seed = c2_problem(n, bilaplacian, per_equation)
    hij, sij = np.argwhere(p.hard), np.argwhere(soft)
    w = p.soft_w[soft] if per_equation else 10.0
    for prec, tol in (("fp64", 1e-9), ("fp32", 1e-5)):
        if bilaplacian and prec == "fp32" and n > 256:
            continue        # (minutes of fp32 refinement passes on a 1e8-conditioned system; the fp32 bound is checked at 256^2)
        solve = quiet(R.solve_grid_linear_simple3_solver, n, n, hij, sij, w_lsq=w, cut_edges=(p.cut_down, p.cut_right),
                      bilaplacian=bilaplacian, precision=prec, maxiter=200000)
        x = quiet(solve, p.hard_vals[p.hard], p.soft_vals[soft])
        err = rel_err(x, ref)
        print("C2 %s n=%d per_equation=%s %s seed %d: max|err|/range = %.2e" % ("bilaplacian" if bilaplacian else "laplacian", n, per_equation, prec, seed, err))
        assert (x[:, :, 0][p.hard] == p.hard_vals[:, :, 0][p.hard]).all()
        if bilaplacian and n >= 1024 and prec == "fp64":
            # cond(A_ff) ~ 1e9 here (thin-plate operator, h^-4, weights up to 100): two backward-stable fp64 solves agree to
            # cond * eps ~ 1e-7 at worst; measured 2.4e-9 between this solver and SuperLU.  DESIGN.md section 8.
            tol = 1e-8
        assert err <= tol, err


@pytest.mark.parametrize("n", [1024, 2048])
def test_c3_tile_fill(n):
    rng = np.random.default_rng(0)
    keep = np.kron(rng.random((n // 64, n // 64)) < 0.5, np.ones((64, 64), bool))
    keep[0, 0] = True
    img8 = rng.integers(0, 256, size=(n, n, 3), dtype=np.uint8)
    img = img8 / 255.0
    ref = go.fill_alpha_harmonic(img, keep)
    flags = S.encode_flags(n, n, hard=keep)
    for prec, tol in (("fp64", 1e-9), ("fp32", 1e-5)):
        s = S.GridSolver(n, n, 3, False, prec)
        s.set_pattern(flags)
        x = s.solve(hard_vals=img)
        err = rel_err(x, ref)
        print("C3 n=%d %s parity settings: max|err|/range = %.2e, %d iterations" % (n, prec, err, s.last_stats["iterations"]))
        assert err <= tol, err
        if prec == "fp32":
            # the benchmark's settings: rtol 1e-6 on the true residual.  What that stopping rule buys in error is REPORTED
            # (SURVEY section 7 "parity != residual"); the bound below only guards against a wrong answer.
            x6 = s.solve(hard_vals=img, rtol=1e-6)
            e6 = rel_err(x6, ref)
            out8 = s.solve_u8(img8, rtol=1e-6)
            want8 = np.clip(np.rint(255.0 * ref), 0, 255).astype(np.uint8)
            lsb = int(np.abs(out8.astype(int) - want8.astype(int)).max())
            print("C3 n=%d fp32 bench settings (rtol 1e-6): max|err|/range = %.2e, %d iterations, relres %.1e; uint8 image: max %d LSB" % (
                n, e6, s.last_stats["iterations"], s.last_stats["relres"], lsb))
            assert e6 <= 2e-3 and lsb <= 1
            assert (out8[keep] == img8[keep]).all()
        s.close()


@pytest.mark.parametrize("n", [256, 512])
def test_c4_height_field_pipeline(n):
    """The stiff default w_lsq = 1e5: cond(A) ~ 1e13 (256^2) .. 1e14 (512^2), so two backward-stable fp64 solvers agree to
    cond * eps at best; the ORACLE itself matches the unmodified reference to only 1e-7 on the small goldens
    (tests/test_oracle.py).  Bounds: 1e-6 at 256^2, 1e-5 at 512^2 (measured 3e-6) -- DESIGN.md section 8."""
    rng = np.random.default_rng(0)
    ii, jj = np.indices((n, n), dtype=np.float64)
    z = 0.05 * n * np.sin(2 * np.pi * ii / n) * np.cos(2 * np.pi * jj / n)
    ij = np.argwhere(rng.random((n - 1, n - 1)) < 0.02)
    zx = z[ij[:, 0] + 1, ij[:, 1]] - z[ij[:, 0], ij[:, 1]]
    zy = z[ij[:, 0], ij[:, 1] + 1] - z[ij[:, 0], ij[:, 1]]
    normals = np.stack([-zx, -zy, np.ones(len(ij))], axis=1)
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    nc = [(int(i), int(j), tuple(v)) for (i, j), v in zip(ij, normals)]
    dx, dy = R.convert_grid_normal_constraints_to_dx_and_dy_constraints(n, n, nc)
    pin = [(0, 0, float(z[0, 0]))]
    ref = go.solve_grid_linear(n, n, dx, dy, pin, bilaplacian=True)
    h = quiet(R.solve_grid_linear, n, n, dx_constraints=dx, dy_constraints=dy, value_constraints=pin, bilaplacian=True, maxiter=400000)
    err = rel_err(h, ref)
    print("C4 n=%d solve_grid_linear (w_lsq 1e5): max|err|/range = %.2e; vs the analytic surface %.2e" % (n, err, rel_err(h, z)))
    assert err <= (1e-6 if n <= 256 else 1e-5), err
    locs = [(int(a), int(b)) for a, b in zip(rng.integers(8, n - 8, 16), rng.integers(8, n - 8, 16))]
    locs = sorted(set(locs))
    bumped = ref.copy()
    for i, j in locs:
        bumped[i, j] += 5.0
    sm_ref = go.smooth_bumps(bumped, locs, 8, bilaplacian=True)
    sm = quiet(R.smooth_bumps, bumped, locs, 8, bilaplacian=True, maxiter=400000)
    e2 = rel_err(sm, sm_ref)
    print("C4 n=%d smooth_bumps: max|err|/range = %.2e" % (n, e2))
    assert e2 <= 1e-8, e2


def test_reference_selftest_entry(tmp_path, monkeypatch, golden_selftest):
    """`python3 recovery.py` (recovery.py:2096-2102, README.md:5-6): the 250^2 self-test through the drop-in module, its two
    1e-7 asserts, the three PNGs, and the solutions against those of the unmodified reference."""
    monkeypatch.chdir(tmp_path)
    hard, soft, mixed = quiet(R.test_solve_grid_linear_simpleN, R.solve_grid_linear_simple3)
    for name in ("sol_hard.png", "sol_soft.png", "sol_mixed.png"):
        assert os.path.isfile(tmp_path / name)
    # (1e-6: the per-equation / scalar soft systems have cond ~ 1e10; the oracle itself matches the reference to 1e-6 here, tests/test_oracle.py)
    for got, key in ((hard, "sol_hard"), (soft, "sol_soft"), (mixed, "sol_mixed")):
        err = rel_err(got, golden_selftest[key])
        print("self-test %s: max|err|/range = %.2e" % (key, err))
        assert err <= 1e-6, (key, err)
    with pytest.raises(SystemExit) as e:
        quiet(R.main)
    assert e.value.code == 0
