"""CPU: host-side logic of the drop-in modules that needs no GPU (array forms of the constraint encoders)."""
import numpy as np

from harmonic_interpolation_b200 import recovery as R


def test_normals_conversion_array_form_matches_tuple_form():
    """recovery.py:1532-1620 (tuple form) against the vectorised array form, including the global shift,
    constraints in the last row / column, and duplicates."""
    rng = np.random.default_rng(0)
    for rows, cols in ((9, 7), (12, 12), (2, 30)):
        for trial in range(24):
            m = int(rng.integers(3, 30))
            ij = np.stack([rng.integers(trial % 2, rows, m), rng.integers((trial // 2) % 2, cols, m)], 1)
            nrm = rng.normal(size=(m, 3))
            nrm[:, 2] = np.abs(nrm[:, 2]) + 0.1
            tup = [(int(i), int(j), tuple(v)) for (i, j), v in zip(ij, nrm)]
            dx, dy = R.convert_grid_normal_constraints_to_dx_and_dy_constraints(rows, cols, tup)
            (aij, adx), (bij, ady) = R.convert_grid_normal_constraints_to_dx_and_dy_constraints(rows, cols, (ij, nrm))
            assert sorted((i, j, round(v, 12)) for i, j, v in dx) == sorted((int(i), int(j), round(float(v), 12)) for (i, j), v in zip(aij, adx))
            assert sorted((i, j, round(v, 12)) for i, j, v in dy) == sorted((int(i), int(j), round(float(v), 12)) for (i, j), v in zip(bij, ady))


def test_cut_planes_match_cut_edges_from_mask():
    rng = np.random.default_rng(1)
    mask = rng.random((13, 17)) < 0.4
    cd, cr = R.cut_planes_from_mask(mask)
    edges = R.cut_edges_from_mask(mask)
    cd2, cr2, n = R._cut_planes(13, 17, edges)
    assert n == len(edges) == int(cd.sum() + cr.sum())
    assert (cd == cd2).all() and (cr == cr2).all()


def test_poisson_gradient_rhs_matches_the_gradient_operator():
    """harmonic_interpolation_b200.poisson.gradient_rhs (vectorised G.T u with mask borders cut) against the oracle's
    assembled gradient operator (poisson.py:204-251)."""
    from harmonic_interpolation_b200 import poisson as P
    from oracle import grid_oracle as go
    rng = np.random.default_rng(4)
    for rows, cols, K, with_mask in ((7, 9, 1, False), (12, 5, 3, True), (1, 6, 2, False), (6, 1, 1, True)):
        u = rng.normal(size=(rows, cols, 2, K))
        mask = None
        if with_mask:
            mask = rng.random((rows, cols)) < 0.7
        got, down, right = P.gradient_rhs(u, mask)
        G = go.gradient_operator(rows, cols, mask)
        want = (G.T @ u.reshape(rows * cols * 2, K)).reshape(rows, cols, K)
        assert np.abs(got - want).max() < 1e-13
        # the edges that exist are exactly the rows of G that are not empty
        nz = np.asarray(abs(G).sum(axis=1)).ravel().reshape(rows, cols, 2) > 0
        assert (nz[:, :, 0] == down).all() and (nz[:, :, 1] == right).all()


def test_cli_argument_errors_exit_minus_one(tmp_path, capsys):
    """fill_alpha_harmonic.py:47-69: wrong argument count, missing input, existing output -> usage text and exit(-1),
    before any GPU work."""
    import pytest
    from harmonic_interpolation_b200 import fill_alpha_harmonic as F
    from conftest import GOLDEN
    import os
    inp = os.path.join(GOLDEN, "fill_alpha_harmonic_test.png")
    existing = tmp_path / "out.png"
    existing.write_bytes(b"x")
    for argv in (["prog"], ["prog", inp], ["prog", inp, "a", "b"], ["prog", str(tmp_path / "missing.png"), str(tmp_path / "o.png")],
                 ["prog", inp, str(existing)]):
        with pytest.raises(SystemExit) as e:
            F.main(argv)
        assert e.value.code == -1
        assert "Usage:" in capsys.readouterr().err


def test_constraint_array_forms_match_tuple_forms():
    """The array forms of the constraint arguments (extension) encode exactly what the tuple forms do."""
    rng = np.random.default_rng(9)
    rows, cols = 11, 14
    ij = np.unique(np.stack([rng.integers(0, rows - 1, 30), rng.integers(0, cols - 1, 30)], 1), axis=0)
    vals = rng.normal(size=len(ij))
    tup = [(int(i), int(j), float(v)) for (i, j), v in zip(ij, vals)]
    i1, v1 = R._scalar_constraints(tup, rows, cols, rows - 1, cols)
    i2, v2 = R._scalar_constraints((ij, vals), rows, cols, rows - 1, cols)
    mask = np.zeros((rows, cols), bool); mask[tuple(ij.T)] = True
    dense = np.zeros((rows, cols)); dense[tuple(ij.T)] = vals
    i3, v3 = R._scalar_constraints((mask, dense), rows, cols, rows - 1, cols)
    o1, o2, o3 = np.argsort(i1), np.argsort(i2), np.argsort(i3)
    assert (i1[o1] == i2[o2]).all() and (i1[o1] == i3[o3]).all()
    assert np.allclose(v1[o1], v2[o2]) and np.allclose(v1[o1], v3[o3])
    # cut edges: list of pairs, (n,4) array, pair of planes
    m = rng.random((rows, cols)) < 0.4
    edges = R.cut_edges_from_mask(m)
    arr = np.asarray([(a[0], a[1], b[0], b[1]) for a, b in edges])
    planes = R.cut_planes_from_mask(m)
    a = R._cut_planes(rows, cols, edges); b = R._cut_planes(rows, cols, arr); c = R._cut_planes(rows, cols, planes)
    for x in (b, c):
        assert (a[0] == x[0]).all() and (a[1] == x[1]).all() and a[2] == x[2]


def test_tint_functions_host_logic_against_the_reference(monkeypatch):
    """highlighting.py:139-213, 361-464: crop / buffer / grow / rim / constraint sets / blending of the drop-in, with the CPU
    oracle standing in for the GPU solve, against outputs of the unmodified reference (tests/golden/reference_highlighting.npz)."""
    import contextlib
    import io

    import highlighting_cases as hc
    from harmonic_interpolation_b200 import highlighting as H
    from oracle import grid_oracle as go

    def oracle_solve(rows, cols, value_constraints=None, bilaplacian=False, **kw):
        mask, values = value_constraints
        vc = [(int(i), int(j), float(values[i, j])) for i, j in zip(*np.nonzero(mask))]
        return go.solve_grid_linear(rows, cols, value_constraints=vc, bilaplacian=bilaplacian)

    monkeypatch.setattr(R, "solve_grid_linear", oracle_solve)
    for name in hc.NAMES:
        case = hc.load(name)
        with contextlib.redirect_stdout(io.StringIO()):
            got = hc.run(H, case)
        diff = np.abs(got.astype(int) - case["out"].astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3, (name, diff.max())


def test_tictoc_dec_and_nesting(capsys):
    """tictoc.py:5-35: '+' * depth on entry, '-' * depth on exit, the decorator names the function."""
    from harmonic_interpolation_b200 import tictoc as T
    del T.starts[:]            # a tic() whose solve raised in an earlier test leaves its entry behind (as in the reference)

    @T.tictoc_dec
    def work(a, b=2):
        with T.tictoc("inner"):
            pass
        return a * b

    assert work(3, b=5) == 15
    lines = capsys.readouterr().out.splitlines()
    assert lines[0] == "+ work" and lines[1] == "++ inner"
    assert lines[2].startswith("-- inner ") and lines[3].startswith("- work ")
    assert T.starts == []


class _ScipyComposedSolver:
    """TEST INFRASTRUCTURE: what libhgrid's hg_solve_composed is specified to compute (include/hgrid.h), assembled with scipy,
    standing in for GridSolver so that the HOST side of poisson.solve_poisson_simple's composed path -- scaling, the
    preconditioner diagonal, masked-out pixels that a constraint touches -- can be held to the reference's goldens on CPU."""
    def __init__(self, rows, cols, channels=1, bilaplacian=False, precision=None, preconditioner=None, uniform_edges=False):
        assert uniform_edges and not bilaplacian and precision == "fp64"
        self.rows, self.cols, self.K = rows, cols, channels

    def set_pattern(self, flags, soft_weights=None, soft_scalar=0.0, w_g=0.0):
        self.flags, self.D = np.asarray(flags), soft_weights
        assert ((self.D > 0) == ((self.flags & 32) != 0)).all()

    def solve_composed(self, constraints, weight, squared=False, hard_vals=None, rhs=None, x0=None, rtol=None, maxiter=None, accept=None):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        rows, cols, K, f = self.rows, self.cols, self.K, self.flags
        N = rows * cols
        idx = np.arange(N).reshape(rows, cols)
        I, J = [], []
        dn = (f[:-1] & 2) == 0
        rt = (f[:, :-1] & 4) == 0
        I += [idx[:-1][dn]]; J += [idx[1:][dn]]
        I += [idx[:, :-1][rt]]; J += [idx[:, 1:][rt]]
        I, J = np.concatenate(I), np.concatenate(J)
        Wm = sp.coo_matrix((np.full(len(I), 0.25), (I, J)), shape=(N, N))
        Wm = (Wm + Wm.T).tocsr()
        L = (sp.diags(np.asarray(Wm.sum(1)).ravel()) - Wm).tocsr()
        hard = (f & 1).astype(bool).ravel()
        free = np.flatnonzero(~hard)
        assert abs(L[free][:, np.flatnonzero(hard)]).sum() == 0        # hard pixels are cut off (needed for m = 2)
        g = np.asarray(rhs, float).reshape(N, K)
        A = sp.lil_matrix((len(constraints), N)); c = np.zeros((len(constraints), K))
        for n, (pix, coef, val) in enumerate(constraints):
            for p, cf in zip(pix, coef):
                assert not hard[p]
                A[n, p] += cf
            c[n] = val
        A = A.tocsr()
        Lf = L[free][:, free]
        S = (Lf @ Lf if squared else Lf) + weight * (A.T @ A)[free][:, free]
        b = (Lf @ g[free] if squared else g[free]) + weight * (A.T @ c)[free]
        x = np.zeros((N, K))
        if hard_vals is not None:
            x[hard] = np.asarray(hard_vals, float).reshape(N, K)[hard]
        x[free] = spla.splu(S.tocsc()).solve(b)
        return x.reshape(rows, cols, K)

    def close(self):
        pass


def test_poisson_composed_host_logic_against_reference_goldens(monkeypatch):
    """poisson.solve_poisson_simple's composed path (bilaplacian=True, rows over several pixels, constraints on masked-out
    pixels) with libhgrid replaced by its scipy specification: the host-side transformation reproduces the unmodified
    reference (tests/golden/reference_poisson.npz, cases of oracle/make_golden_poisson.py: cases2)."""
    import contextlib
    import io
    import poisson_cases as pc
    from harmonic_interpolation_b200 import poisson as P
    monkeypatch.setattr(P, "GridSolver", _ScipyComposedSolver)
    for name in pc.NAMES2:
        kw, want = pc.load2(name)
        with contextlib.redirect_stdout(io.StringIO()):
            got = P.solve_poisson_simple(**kw)
        tol = 1e-7 if kw["bilaplacian"] else 1e-9
        assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max()), name


def test_poisson_components_and_rank_check():
    """poisson.components against scipy's labelling of the same graph; check_constrained raises exactly when the reference's
    system matrix is singular."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    from harmonic_interpolation_b200 import poisson as P
    rng = np.random.default_rng(2)
    for trial in range(20):
        rows, cols = int(rng.integers(1, 14)), int(rng.integers(1, 14))
        free = rng.random((rows, cols)) < 0.8
        down = rng.random((rows, cols)) < 0.7; down[-1] = False
        right = rng.random((rows, cols)) < 0.7; right[:, -1] = False
        labels, nc = P.components(free, down, right)
        idx = np.arange(rows * cols).reshape(rows, cols)
        ed = free[:-1] & free[1:] & down[:-1]
        er = free[:, :-1] & free[:, 1:] & right[:, :-1]
        I = np.concatenate([idx[:-1][ed], idx[:, :-1][er]]); J = np.concatenate([idx[1:][ed], idx[:, 1:][er]])
        g = sp.coo_matrix((np.ones(len(I)), (I, J)), shape=(rows * cols, rows * cols))
        _, want = connected_components(g, directed=False)
        f = free.ravel()
        assert nc == len(np.unique(want[f]))
        assert (labels.ravel()[~f] == -1).all()
        # the same partition
        pairs = set(zip(labels.ravel()[f].tolist(), want[f].tolist()))
        assert len(pairs) == nc
    free = np.ones((4, 6), bool); down = np.ones((4, 6), bool); right = np.ones((4, 6), bool)
    right[:, 2] = False                                   # two components: columns 0..2 and 3..5
    P.check_constrained(free, down, right, [([0], [1.0], None), ([5], [2.0], None)])
    P.check_constrained(free, down, right, [([0, 5], [1.0, 1.0], None), ([1, 4], [1.0, -1.0], None)])
    import pytest
    with pytest.raises(ArithmeticError):
        P.check_constrained(free, down, right, [([0], [1.0], None)])                      # the right half floats
    with pytest.raises(ArithmeticError):
        P.check_constrained(free, down, right, [([0, 1], [1.0, -1.0], None), ([5], [1.0], None)])   # a difference pins nothing
    with pytest.raises(ArithmeticError):
        P.check_constrained(free, down, right, [([0, 5], [1.0, 1.0], None)])               # one row, two constants
