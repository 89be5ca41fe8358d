"""GPU: the poisson.solve_poisson_simple drop-in (libhgrid: uniform-edge Laplacian, cut edges at mask borders, explicit
right-hand side) against golden outputs of the unmodified reference and against the CPU oracle."""
import contextlib
import io

import numpy as np
import pytest

import poisson_cases as pc
from conftest import rel_err
from harmonic_interpolation_b200 import poisson as P
from oracle import grid_oracle as go

pytestmark = pytest.mark.gpu


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


@pytest.mark.parametrize("name", pc.NAMES)
def test_poisson_matches_reference_golden(name):
    kw, want = pc.load(name)
    got = quiet(P.solve_poisson_simple, **kw, precision="fp64")
    assert got.shape == want.shape
    assert rel_err(got, want) < 1e-9
    got32 = quiet(P.solve_poisson_simple, **kw, precision="fp32")
    assert rel_err(got32, want) < 1e-5


def test_poisson_integrates_a_gradient_field_exactly():
    """The forward differences of a field, plus one pin: the field comes back (G x = u is consistent)."""
    rng = np.random.default_rng(3)
    rows, cols = 150, 211
    f = rng.random((rows, cols, 2))
    u = np.zeros((rows, cols, 2, 2))
    u[:-1, :, 0] = f[1:] - f[:-1]
    u[:, :-1, 1] = f[:, 1:] - f[:, :-1]
    lc = [([(5, 7, 1.0)], f[5, 7])]
    x = quiet(P.solve_poisson_simple, u, lc, 10.0, precision="fp64")
    assert np.abs(x - f).max() < 1e-8
    ref = go.solve_poisson_simple(u, lc, 10.0)
    assert rel_err(x, ref) < 1e-9


def test_poisson_argument_checks():
    u = np.zeros((4, 5, 2, 1))
    with pytest.raises(AssertionError):
        quiet(P.solve_poisson_simple, u[:, :, :1], [([(0, 0, 1.0)], [0.0])], 1.0)      # poisson.py:37
    with pytest.raises(AssertionError):
        quiet(P.solve_poisson_simple, u, [], 1.0)                                       # poisson.py:49
    with pytest.raises(ArithmeticError):
        # a difference constraint pins no value: the system is singular (CHOLMOD fails in the reference, poisson.py:95)
        quiet(P.solve_poisson_simple, u + 1.0, [([(0, 0, 1.0), (0, 1, -1.0)], [0.5])], 1.0, rtol=1e-10, maxiter=300)


@pytest.mark.parametrize("name", pc.NAMES2)
def test_poisson_composed_matches_reference_golden(name):
    """bilaplacian=True ((G.T G)^2, poisson.py:65-67), constraint rows over several pixels (poisson.py:329-366), constraints on
    masked-out pixels (poisson.py:419-423): hg_solve_composed against outputs of the unmodified reference.  The squared
    operator's condition number (1e6 .. 1e9 at these sizes, times the weight 1e5 of the self-test pattern) bounds the agreement
    of two backward-stable solvers by cond * eps: 1e-7 of the range there, 1e-9 for the Laplacian form."""
    kw, want = pc.load2(name)
    got = quiet(P.solve_poisson_simple, **kw)
    assert got.shape == want.shape
    tol = 1e-7 if kw["bilaplacian"] else 1e-9
    assert rel_err(got, want) < tol, rel_err(got, want)


@pytest.mark.parametrize("squared", [False, True])
def test_poisson_composed_larger_grid_against_oracle(squared):
    """250 x 250 (the size of the reference's own test, poisson.py:405-406) with a mask, K = 2, one-pixel and several-pixel rows."""
    rng = np.random.default_rng(5)
    rows = cols = 250
    f = np.stack([np.add.outer(np.sin(np.arange(rows) / 17.0), np.cos(np.arange(cols) / 23.0)), rng.random((rows, cols)) * 0.1], -1)
    u = np.zeros((rows, cols, 2, 2))
    u[:-1, :, 0] = f[1:] - f[:-1]
    u[:, :-1, 1] = f[:, 1:] - f[:, :-1]
    mask = np.ones((rows, cols), bool); mask[100:150, 60:120] = False
    lc = [([(5, 7, 1.0)], f[5, 7]), ([(240, 200, 2.0)], 2.0 * f[240, 200]), ([(30, 220, 1.0), (200, 20, -1.0)], f[30, 220] - f[200, 20]),
          ([(120, 10, 0.5), (121, 10, 0.5), (10, 120, 1.0)], rng.random(2)), ([(110, 70, 1.0)], [0.3, 0.4])]
    got = quiet(P.solve_poisson_simple, u, lc, 25.0, mask=mask, mask_value=[0.25, 0.5], bilaplacian=squared)
    ref = go.solve_poisson_simple(u, lc, 25.0, mask, [0.25, 0.5], squared)
    assert rel_err(got, ref) < (1e-7 if squared else 1e-9), rel_err(got, ref)
