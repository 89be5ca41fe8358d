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
    with pytest.raises(NotImplementedError):
        quiet(P.solve_poisson_simple, u, [([(0, 0, 1.0), (0, 1, -1.0)], [0.0])], 1.0)
    with pytest.raises(NotImplementedError):
        quiet(P.solve_poisson_simple, u, [([(0, 0, 1.0)], [0.0])], 1.0, bilaplacian=True)
