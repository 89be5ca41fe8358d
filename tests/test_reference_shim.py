"""CPU, build container only: the oracle against the UNMODIFIED reference run through
oracle/ref_shim.py.  Skipped where /root/reference does not exist (the GPU box)."""
import contextlib
import io

import numpy as np
import pytest

from oracle import grid_oracle as go
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.have_reference(), reason="reference tree not present")


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def test_matrices_identical():
    R = ref_shim.load_reference()["recovery"]
    rng = np.random.default_rng(1)
    for rows, cols in [(1, 5), (5, 1), (2, 7), (7, 2), (3, 3), (6, 9)]:
        L = R.gen_symmetric_grid_laplacian(rows, cols)
        assert abs(L - go.symmetric_laplacian_matrix(rows, cols)).max() == 0
        Lr = R.gen_grid_laplacian2_with_boundary_reflection(rows, cols)
        assert abs(Lr - go.reflected_laplacian_matrix(rows, cols)).max() == 0
    rows, cols = 12, 15
    block = np.zeros((rows, cols), bool)
    block[3:8, 4:11] = True
    cuts = R.cut_edges_from_mask(block)
    assert cuts == go.cut_edges_from_mask(block)
    cd, cr = go.cut_planes_from_edges(rows, cols, cuts)
    assert abs(R.gen_symmetric_grid_laplacian(rows, cols, cuts) - go.symmetric_laplacian_matrix(rows, cols, cd, cr)).max() == 0
    assert abs(R.gen_grid_laplacian2_with_boundary_reflection(rows, cols, cuts) -
               go.reflected_laplacian_matrix(rows, cols, cd, cr)).max() == 0
    # random cuts: both must agree on whether the domain is valid (recovery.py:495)
    for seed in range(6):
        m = np.random.default_rng(seed).random((8, 9)) < 0.35
        cu = R.cut_edges_from_mask(m)
        cd, cr = go.cut_planes_from_edges(8, 9, cu)
        try:
            quiet(R.gen_grid_laplacian2_with_boundary_reflection, 8, 9, cu)
            ok_ref = True
        except AssertionError:
            ok_ref = False
        try:
            go.reflected_laplacian_matrix(8, 9, cd, cr, have_cuts=len(cu) > 0)
            ok_o = True
        except AssertionError:
            ok_o = False
        assert ok_ref == ok_o


def test_reference_selftest_passes_under_shim(tmp_path, monkeypatch):
    """python3 recovery.py == test_solve_grid_linear_simpleN(solve_grid_linear_simple3) at the 'small' size."""
    R = ref_shim.load_reference()["recovery"]
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(3)
    rows, cols = 30, 26
    hard = [(int(i), int(j), list(rng.random(2))) for i, j in zip(*np.where(rng.random((rows, cols)) < 0.05))]
    soft = [(int(i), int(j), list(rng.random(2))) for i, j in zip(*np.where(rng.random((rows, cols)) < 0.1))
            if (int(i), int(j)) not in set((a, b) for a, b, _ in hard)]
    for bil in (False, True):
        a = quiet(R.solve_grid_linear_simple3, rows, cols, hard, soft, 3.0, None, bil)
        b = go.solve_grid_linear_simple3(rows, cols, hard, soft, 3.0, None, bil)
        assert np.abs(a - b).max() < 1e-10


def test_host_helpers_match_reference():
    """The product's pure-Python helpers against the reference's (same names)."""
    ref = ref_shim.load_reference()
    R = ref["recovery"]
    from harmonic_interpolation_b200 import recovery as P
    from harmonic_interpolation_b200.highlighting import grow
    rng = np.random.default_rng(7)
    mask = rng.random((9, 11)) < 0.4
    assert P.cut_edges_from_mask(mask) == R.cut_edges_from_mask(mask)
    edges = R.cut_edges_from_mask(mask)
    # R.zero_dx_dy_constraints_from_cut_edges cannot be compared: under `from numpy import *` its
    # `min(i, k)` is numpy.min(a, axis) and raises AxisError for every vertical edge (recovery.py:87)
    for it in (0, 1, 3):
        assert (grow(mask, it) == ref["highlighting"].grow(mask, it)).all()
    for rows, cols, locs in [(6, 7, [(1, 1), (5, 6), (5, 2), (0, 6)]), (5, 5, [(4, 4), (2, 2)]), (5, 5, [(0, 0), (4, 4), (4, 3), (3, 4)])]:
        normals = [(i, j, tuple(rng.normal(size=2)) + (abs(rng.normal()) + 0.1,)) for i, j in locs]
        a = R.convert_grid_normal_constraints_to_dx_and_dy_constraints(rows, cols, normals)
        b = P.convert_grid_normal_constraints_to_dx_and_dy_constraints(rows, cols, normals)
        assert a == b
    img = rng.normal(size=(5, 6))
    assert (P.normalize_to_char_img(img) == R.normalize_to_char_img(img)).all()
