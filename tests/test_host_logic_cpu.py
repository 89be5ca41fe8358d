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
