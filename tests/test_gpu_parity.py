"""GPU: the CUDA path, called through the C ABI (ctypes -> libhgrid.so), against the CPU oracle
and against the golden outputs of the unmodified reference.

Tolerances (BASELINE.json north_star): max abs error <= 1e-5 of the value range in
fp32-with-fp64-reductions mode, <= 1e-9 in fp64 mode, <= 1 LSB on the golden PNG.
Where the system itself is ill-conditioned (w_lsq = 1e5 on the bilaplacian, cond >~ 1e9) two
exact float64 solvers already differ by cond*eps, and the bound is stated in the test.
"""
import os

import numpy as np
import pytest
from PIL import Image

import golden_cases as gc
from conftest import GOLDEN, rel_err
from harmonic_interpolation_b200 import _capi as C
from harmonic_interpolation_b200 import fill_alpha_harmonic as F
from harmonic_interpolation_b200 import recovery as R
from harmonic_interpolation_b200 import solver as S
from oracle import grid_oracle as go

pytestmark = pytest.mark.gpu

TOL32 = 1e-5
TOL64 = 1e-9


def random_problem(rng, rows, cols, K, bilap, cuts=True, pen=True, soft_plane=True):
    p = go.GridProblem(rows, cols, K, bilap)
    p.hard = rng.random((rows, cols)) < 0.06
    p.hard[0, 0] = True
    p.hard_vals = rng.random((rows, cols, K))
    soft = (rng.random((rows, cols)) < 0.1) & ~p.hard
    p.soft_w = np.where(soft, rng.uniform(0.5, 20.0, (rows, cols)) if soft_plane else 3.0, 0.0)
    p.soft_vals = rng.random((rows, cols, K))
    if cuts:
        block = np.zeros((rows, cols), bool)
        block[rows // 4: rows // 2, cols // 3: 2 * cols // 3] = True
        p.cut_down, p.cut_right = R.cut_planes_from_mask(block)
        p.have_cuts = True
    if pen:
        p.pen_down = rng.random((rows, cols)) < 0.05
        p.pen_down[-1, :] = False
        p.pen_right = rng.random((rows, cols)) < 0.05
        p.pen_right[:, -1] = False
        p.d_down = rng.normal(size=(rows, cols, K))
        p.d_right = rng.normal(size=(rows, cols, K))
        p.w_g = 7.0
    return p


def make_solver(p, precision, soft_plane=True):
    s = S.GridSolver(p.rows, p.cols, p.K, p.bilaplacian, precision)
    flags = S.encode_flags(p.rows, p.cols, hard=p.hard, soft=p.soft_w > 0, cut_down=p.cut_down, cut_right=p.cut_right,
                           pen_down=p.pen_down, pen_right=p.pen_right)
    if soft_plane:
        s.set_pattern(flags, p.soft_w, 0.0, p.w_g)
    else:
        s.set_pattern(flags, None, float(p.soft_w.max()), p.w_g)
    return s


@pytest.mark.parametrize("bilap", [False, True])
@pytest.mark.parametrize("shape", [(1, 7), (7, 1), (2, 9), (3, 3), (17, 33), (64, 64), (70, 131), (257, 300)])
def test_operator_matches_oracle_elementwise(shape, bilap):
    """hg_apply against the matrix-free / assembled oracle operator, fp64 (<= 1e-13 of the
    largest entry) and fp32 (<= 4 ulp-ish: 2e-6), on ragged shapes with cuts, soft weights and
    gradient-penalty edges; output must be 0 at hard pixels."""
    rows, cols = shape
    rng = np.random.default_rng(rows * 1000 + cols + bilap)
    p = random_problem(rng, rows, cols, 3, bilap, cuts=min(rows, cols) >= 8)
    x = rng.normal(size=(rows, cols, 3))
    want = go.apply_full_operator(p, x)
    want[p.hard] = 0.0
    scale = np.abs(want).max() + 1e-30
    for prec, tol in (("fp64", 1e-13), ("fp32", 3e-6)):
        s = make_solver_unchecked(p, prec)
        if s is None:
            continue
        got = s.apply(x)
        assert np.abs(got - want).max() / scale < tol, (prec, shape)
        s.close()


def make_solver_unchecked(p, prec):
    """Bilaplacian domains that the reference rejects (recovery.py:495) are rejected here too."""
    try:
        go.energy_matrix(p)
        ok = True
    except AssertionError:
        ok = False
    if ok:
        return make_solver(p, prec)
    with pytest.raises(AssertionError):
        make_solver(p, prec)
    return None


@pytest.mark.parametrize("bilap", [False, True])
def test_random_problem_solve_parity(bilap):
    rng = np.random.default_rng(11 + bilap)
    p = random_problem(rng, 60, 52, 2, bilap)
    ref = go.solve_problem(p)
    s = make_solver(p, "fp64")
    out = s.solve(p.hard_vals, p.soft_vals, p.d_down, p.d_right, rtol=1e-13, maxiter=200000)
    assert rel_err(out, ref) < TOL64
    # hard pixels are returned exactly
    assert (out[p.hard] == p.hard_vals[p.hard]).all()
    # the reported residual is the reduced-system residual the oracle computes
    rr = go.residual_norms(p, out).max()
    assert rr < 1e-11 and abs(np.log10(max(s.last_stats["relres"], 1e-300)) - np.log10(max(rr, 1e-300))) < 1.5
    s.close()
    s = make_solver(p, "fp32")
    out = s.solve(p.hard_vals, p.soft_vals, p.d_down, p.d_right, rtol=1e-7, maxiter=200000, allow_unconverged=True)
    assert rel_err(out, ref) < (TOL32 if not bilap else 1e-4)
    s.close()


def test_fill_golden_png(tmp_path):
    """config 1: fill_alpha_harmonic-test/test.png -> expected.png through the CLI, <= 1 LSB."""
    outp = tmp_path / "out.png"
    F.main(["fill_alpha_harmonic.py", os.path.join(GOLDEN, "fill_alpha_harmonic_test.png"), str(outp)])
    out = np.asarray(Image.open(outp))
    exp = np.asarray(Image.open(os.path.join(GOLDEN, "fill_alpha_harmonic_expected.png")))
    assert out.shape == exp.shape and out.dtype == exp.dtype
    assert np.abs(out.astype(int) - exp.astype(int)).max() <= 1
    # refuses to clobber (fill_alpha_harmonic.py:67-69)
    with pytest.raises(SystemExit):
        F.main(["fill_alpha_harmonic.py", os.path.join(GOLDEN, "fill_alpha_harmonic_test.png"), str(outp)])


def test_fill_golden_png_fp32():
    rgba = np.asarray(Image.open(os.path.join(GOLDEN, "fill_alpha_harmonic_test.png")).convert("RGBA"))
    exp = np.asarray(Image.open(os.path.join(GOLDEN, "fill_alpha_harmonic_expected.png")))
    arr = rgba / 255.0
    res = F.fill_alpha_harmonic(arr[:, :, :3], (255 * arr[:, :, 3]).round(0).astype(int) == 255, precision="fp32", rtol=1e-7)
    out = np.asarray((res * 255).round(0).clip(0, 255), dtype=np.uint8)
    assert np.abs(out.astype(int) - exp.astype(int)).max() <= 1


def test_fill_rgba8_golden_png():
    """uint8 in / uint8 out entry point (hg_solve_u8) on the golden image, fp64 and fp32: <= 1 LSB; kept pixels exact."""
    rgba = np.asarray(Image.open(os.path.join(GOLDEN, "fill_alpha_harmonic_test.png")).convert("RGBA"))
    exp = np.asarray(Image.open(os.path.join(GOLDEN, "fill_alpha_harmonic_expected.png")))
    for prec in ("fp64", "fp32"):
        out = F.fill_rgba8(rgba, precision=prec)
        assert out.dtype == np.uint8 and out.shape == exp.shape
        assert np.abs(out.astype(int) - exp.astype(int)).max() <= 1
        keep = rgba[:, :, 3] == 255
        assert (out[keep] == rgba[:, :, :3][keep]).all()
    assert (go.fill_rgba8(rgba) == exp).all()


def test_ktest_all_ones():
    """fill_alpha_harmonic.py:85-99: a hole in an all-ones image fills with ones."""
    for sl in (np.s_[:3, :3], np.s_[3:6, 3:6]):
        img = np.full((9, 9, 4), 255, np.uint8)
        img[sl + (3,)] = 0
        arr = img / 255.0
        res = F.fill_alpha_harmonic(arr[:, :, :3], (255 * arr[:, :, 3]).round(0).astype(int) == 255)
        assert np.allclose(res, arr[:, :, :3])


def test_simple3_golden(golden):
    for key, kw in gc.s3_cases(golden):
        out = R.solve_grid_linear_simple3(maxiter=400000, **kw)
        assert out.shape == golden[key].shape
        # w_lsq = 1e3 against O(1) stencil weights on the bilaplacian: cond ~ 1e9, the reference's
        # own float64 answer is only good to ~1e-8 of the range
        assert rel_err(out, golden[key]) < (1e-8 if ("bilap" in key and "w1e3" in key) else TOL64), key


def test_simple3_array_form_equals_tuple_form(golden):
    rows, cols, hard, soft, cuts = gc.s3_args(golden)
    a = R.solve_grid_linear_simple3(rows, cols, hard, soft, 10.0, cuts, False)
    b = R.solve_grid_linear_simple3(rows, cols, (golden["s3_hard_ij"], golden["s3_hard_v"]),
                                    (golden["s3_soft_ij"], golden["s3_soft_v"]), 10.0, golden["s3_cuts"], False)
    assert np.abs(a - b).max() == 0.0


def test_simple3_solver_reuse(golden):
    """recovery.py:1307-1334: one pattern, several value sets, K differing between calls."""
    rows, cols, hard, soft, cuts = gc.s3_args(golden)
    solve = R.solve_grid_linear_simple3_solver(rows, cols, [(i, j) for i, j, _ in hard], [(i, j) for i, j, _ in soft],
                                               w_lsq=3.0, cut_edges=cuts, bilaplacian=False)
    osolve = go.solve_grid_linear_simple3_solver(rows, cols, [(i, j) for i, j, _ in hard], [(i, j) for i, j, _ in soft],
                                                 3.0, cuts, False)
    rng = np.random.default_rng(4)
    for K in (1, 3, 2):
        hv = rng.random((len(hard), K))
        sv = rng.random((len(soft), K))
        assert rel_err(solve(list(hv), list(sv)), osolve(list(hv), list(sv))) < TOL64
    with pytest.raises(AssertionError):
        solve(list(hv)[:-1], list(sv))   # recovery.py:1310


def test_simple_golden(golden):
    for key, kw in gc.s1_cases(golden):
        out = R.solve_grid_linear_simple(maxiter=400000, **kw)
        assert rel_err(out, golden[key]) < TOL64, key


def test_solve_grid_linear_golden(golden):
    for key, kw in gc.g_cases(golden):
        out = R.solve_grid_linear(maxiter=2000000, **kw)
        assert out.shape == golden[key].shape
        # w_lsq = 1e5 (the default) makes cond ~ 1e8..1e12: bound by what two direct solvers agree to
        tol = 1e-6 if "wdef" in key else TOL64
        assert rel_err(out, golden[key]) < tol, key


def test_smooth_bumps_golden(golden):
    for key, fn, args, kw in gc.sb_cases(golden):
        out = getattr(R, fn)(*args, maxiter=400000, **kw)
        assert rel_err(out, golden[key]) < 1e-8, key
    g = golden["sb_grid"]
    assert (R.smooth_bumps2(g, [(1, 1)], 1, 0) == g).all()                 # recovery.py:1489-1490
    with pytest.raises(AssertionError):
        R.smooth_bumps2(g, [(g.shape[0] - 1, 0)], 1, 1)                    # recovery.py:1465


def test_reference_selftest_asserts():
    """recovery.py:2020-2069 at the 'small' size (9x9) and at 64x64: scalar vs per-equation w_lsq
    agree to 1e-7, hard / soft / mixed, bilaplacian with cut edges."""
    for rows, br0, br1 in ((9, 3, 6), (64, 24, 40)):
        cols = rows
        vcs = [(0, 0, [br0 * 2.0]), (rows - 1, 0, [0.0]), (0, cols - 1, [0.0]), (rows - 1, cols - 1, [0.0]),
               (br0, br0, [br0]), (br1 - 1, br0, [br0]), (br0, br1 - 1, [br0]), (br1 - 1, br1 - 1, [br0])]
        mask = np.zeros((rows, cols), bool)
        mask[br0:br1, br0:br1] = True
        cuts = R.cut_edges_from_mask(mask)
        kw = dict(cut_edges=cuts, bilaplacian=True, maxiter=2000000)
        soft = R.solve_grid_linear_simple3(rows, cols, value_constraints_soft=vcs, w_lsq=1e3, **kw)
        soft2 = R.solve_grid_linear_simple3(rows, cols, value_constraints_soft=vcs, w_lsq=[1e3] * len(vcs), **kw)
        assert np.abs(soft2 - soft).max() < 1e-7
        mid = [((br0 + br1) // 2, (br0 + br1) // 2, [0.5 * br0])]
        mixed = R.solve_grid_linear_simple3(rows, cols, value_constraints_hard=vcs, value_constraints_soft=mid, w_lsq=1e3, **kw)
        mixed2 = R.solve_grid_linear_simple3(rows, cols, value_constraints_hard=vcs, value_constraints_soft=mid, w_lsq=[1e3], **kw)
        assert np.abs(mixed2 - mixed).max() < 1e-7
        assert rel_err(mixed, go.solve_grid_linear_simple3(rows, cols, vcs, mid, 1e3, cuts, True)) < 1e-7


def test_documented_equivalences(golden):
    """The reference's disabled cross-checks: simple == per-channel solve_grid_linear
    (recovery.py:1080-1086); simple2 == simple3 (:1332)."""
    keep, img = golden["s1_keep"], golden["s1_img"]
    rows, cols = keep.shape
    a = R.solve_grid_linear_simple(rows, cols, (keep, img))
    for c in range(img.shape[2]):
        b = R.solve_grid_linear(rows, cols, value_constraints=(keep, img[:, :, c]))
        assert np.abs(a[:, :, c] - b).max() < 1e-9
    rows, cols, hard, soft, _ = gc.s3_args(golden)
    s2 = R.solve_grid_linear_simple2(rows, cols, hard, soft, 10.0, False)
    s3 = R.solve_grid_linear_simple3(rows, cols, hard, soft, 10.0, None, False)
    assert np.abs(s2 - s3).max() < 1e-9


def test_error_behaviour():
    with pytest.raises(AssertionError):
        R.solve_grid_linear_simple(5, 5, [])                                  # recovery.py:1050
    with pytest.raises(AssertionError):
        R.solve_grid_linear_simple3(5, 5, [], [])                             # recovery.py:1258
    with pytest.raises(AssertionError):
        R.solve_grid_linear_simple3(5, 5, [(0, 0, [1.0])], [(1, 1, [1.0])], w_lsq=-1.0)   # :1262
    with pytest.raises(AssertionError):
        R.solve_grid_linear_simple(2, 2, [(0, 0, [1.0])], bilaplacian=True)   # recovery.py:495
    with pytest.raises(AssertionError):
        R.solve_grid_linear(5, 5, [(4, 0, 1.0)])                              # recovery.py:523


def test_analytic_large_grid():
    """Beyond the oracle's reach (2048^2): constants are reproduced exactly-ish by the Laplacian
    fill and the residual reported is honest (size-independent property)."""
    rows = cols = 2048
    rng = np.random.default_rng(0)
    keep = np.kron(rng.random((rows // 64, cols // 64)) < 0.5, np.ones((64, 64), bool))
    keep[0, 0] = True
    img = np.full((rows, cols, 1), 0.625)
    s = S.GridSolver(rows, cols, 1, False, "fp32")
    s.set_pattern(S.encode_flags(rows, cols, hard=keep))
    out = s.solve(hard_vals=img, rtol=1e-6, maxiter=20000)
    assert s.last_stats["relres"] <= 1e-6
    assert np.abs(out - 0.625).max() < 1e-5
    s.close()


def test_unconstrained_component_raises(monkeypatch):
    """cvxopt.cholmod raises ArithmeticError on the singular system of a connected component that no value
    constraint touches (recovery.py:993).  hg_setup finds such components by spreading an `anchored` mark over uncut edges."""
    rows, cols = 90, 120
    block = np.zeros((rows, cols), bool)
    block[30:50, 40:70] = True
    cd, cr = R.cut_planes_from_mask(block)
    hard = np.zeros((rows, cols), bool)
    hard[0, 0] = True
    for bilap in (False, True):
        s = S.GridSolver(rows, cols, 1, bilap, "fp64")
        with pytest.raises(ArithmeticError):
            s.set_pattern(S.encode_flags(rows, cols, hard=hard, cut_down=cd, cut_right=cr))
        assert s.info.unconstrained == 20 * 30
        # a soft pixel inside the cut-off block, or a hard one, makes the system regular
        soft = np.zeros((rows, cols), bool); soft[40, 50] = True
        s.set_pattern(S.encode_flags(rows, cols, hard=hard, soft=soft, cut_down=cd, cut_right=cr), None, 2.0)
        h2 = hard.copy(); h2[35, 45] = True
        s.set_pattern(S.encode_flags(rows, cols, hard=h2, cut_down=cd, cut_right=cr))
        assert s.info.unconstrained == 0
        s.close()
    # no constraint at all
    with pytest.raises(ArithmeticError):
        S.GridSolver(20, 30, 1, False, "fp64").set_pattern(np.zeros((20, 30), np.uint8))
    # a serpentine corridor one pixel wide, anchored only at its far end: many rounds of spreading, no false alarm
    n = 130
    free = np.zeros((n, n), bool)
    for k in range(0, n, 4):
        free[k, :] = True
        free[k:k + 4, (n - 1) if (k // 4) % 2 == 0 else 0] = True
    free = free[:n, :n]
    hard = ~free
    hard[0, 0] = False
    cdp = np.zeros((n, n), bool); crp = np.zeros((n, n), bool)
    # hard pixels would anchor the corridor everywhere: cut every edge between free and hard pixels, keep one hard pixel joined at the end
    cdp[:-1] = free[:-1] != free[1:]; crp[:, :-1] = free[:, :-1] != free[:, 1:]
    last = np.argwhere(free)[-1]
    s = S.GridSolver(n, n, 1, False, "fp64")
    with pytest.raises(ArithmeticError):
        s.set_pattern(S.encode_flags(n, n, hard=~free, cut_down=cdp, cut_right=crp))
    soft = np.zeros((n, n), bool); soft[last[0], last[1]] = True
    s.set_pattern(S.encode_flags(n, n, hard=~free, soft=soft, cut_down=cdp, cut_right=crp), None, 1.0)
    assert s.info.unconstrained == 0
    x = s.solve(soft_vals=np.full((n, n, 1), 3.0))
    assert np.abs(x[free] - 3.0).max() < 1e-8            # the harmonic function with one soft pin is the constant
    s.close()
    monkeypatch.setenv("HGRID_ALLOW_SINGULAR", "1")
    s = S.GridSolver(20, 30, 1, False, "fp64")
    s.set_pattern(np.zeros((20, 30), np.uint8))
    s.close()
