"""CPU: the oracle (oracle/grid_oracle.py) against the reference's golden data."""
import os

import numpy as np
import pytest
from PIL import Image

import golden_cases as gc
from conftest import GOLDEN, rel_err
from oracle import grid_oracle as go


def test_oracle_reproduces_expected_png():
    """fill_alpha_harmonic-test/test.png -> expected.png, the reference's only known-answer test: 0 LSB."""
    rgba = np.asarray(Image.open(os.path.join(GOLDEN, "fill_alpha_harmonic_test.png")).convert("RGBA"))
    exp = np.asarray(Image.open(os.path.join(GOLDEN, "fill_alpha_harmonic_expected.png")))
    out = go.fill_rgba8(rgba)
    assert out.shape == exp.shape
    assert np.abs(out.astype(int) - exp.astype(int)).max() == 0


def test_oracle_ktest_all_ones():
    """The reference's disabled inline test (fill_alpha_harmonic.py:85-99)."""
    for sl in (np.s_[:3, :3], np.s_[3:6, 3:6]):
        img = np.full((9, 9, 4), 255, np.uint8)
        img[sl + (3,)] = 0
        arr = img / 255.0
        res = go.fill_alpha_harmonic(arr[:, :, :3], (255 * arr[:, :, 3]).round(0).astype(int) == 255)
        assert np.allclose(res, arr[:, :, :3])


def test_oracle_simple3_golden(golden):
    for key, kw in gc.s3_cases(golden):
        out = go.solve_grid_linear_simple3(**kw)
        assert rel_err(out, golden[key]) < 1e-10, key


def test_oracle_simple_golden(golden):
    for key, kw in gc.s1_cases(golden):
        out = go.solve_grid_linear_simple(**kw)
        assert rel_err(out, golden[key]) < 1e-10, key


def test_oracle_solve_grid_linear_golden(golden):
    for key, kw in gc.g_cases(golden):
        out = go.solve_grid_linear(**kw)
        # w_lsq = 1e5 on the bilaplacian is ill-conditioned: the reference's own answer carries cond*eps
        assert rel_err(out, golden[key]) < 1e-7, key


def test_oracle_smooth_bumps_golden(golden):
    for key, fn, args, kw in gc.sb_cases(golden):
        out = getattr(go, fn)(*args, **kw)
        assert rel_err(out, golden[key]) < 1e-8, key


def test_oracle_selftest_consistency(golden_selftest):
    """recovery.py:2064, 2069: scalar w_lsq and per-equation w_lsq agree to 1e-7, and the 250x250 outputs of the reference's self-test are reproduced."""
    rows = cols = 250
    br0, br1 = 100, 150
    vcs = [(0, 0, [br0 * 2.0]), (rows - 1, 0, [0.0]), (0, cols - 1, [0.0]), (rows - 1, cols - 1, [0.0]),
           (br0, br0, [br0]), (br1 - 1, br0, [br0]), (br0, br1 - 1, [br0]), (br1 - 1, br1 - 1, [br0])]
    mask = np.zeros((rows, cols), bool)
    mask[br0:br1, br0:br1] = True
    cuts = go.cut_edges_from_mask(mask)
    hard = go.solve_grid_linear_simple3(rows, cols, vcs, [], 1.0, cuts, True)
    # 250^2 bilaplacian pinned at 8 pixels: cond ~ 1e9, two backward-stable direct solves of the
    # same system differ by cond * eps ~ 3e-8 of the range
    assert rel_err(hard, golden_selftest["sol_hard"]) < 1e-6
    soft = go.solve_grid_linear_simple3(rows, cols, [], vcs, 1e3, cuts, True)
    soft2 = go.solve_grid_linear_simple3(rows, cols, [], vcs, [1e3] * len(vcs), cuts, True)
    assert np.abs(soft2 - soft).max() < 1e-7
    assert rel_err(soft, golden_selftest["sol_soft"]) < 1e-6


def test_operator_matrix_free_equals_assembled():
    rng = np.random.default_rng(5)
    for rows, cols in [(1, 6), (6, 1), (2, 7), (3, 3), (9, 12)]:
        x = rng.random((rows, cols, 2))
        L = go.symmetric_laplacian_matrix(rows, cols)
        y = go.apply_symmetric_laplacian(x)
        for k in range(2):
            assert np.abs((L @ x[:, :, k].ravel()).reshape(rows, cols) - y[:, :, k]).max() < 1e-14
    rows, cols = 14, 17
    block = np.zeros((rows, cols), bool)
    block[4:9, 5:12] = True
    cd, cr = go.cut_planes_from_edges(rows, cols, go.cut_edges_from_mask(block))
    x = rng.random((rows, cols, 2))
    Lr = go.reflected_laplacian_matrix(rows, cols, cd, cr)
    B = Lr.T @ Lr
    y = go.apply_bilaplacian(x, cd, cr)
    for k in range(2):
        assert np.abs((B @ x[:, :, k].ravel()).reshape(rows, cols) - y[:, :, k]).max() < 1e-14


def test_reflected_domain_asserts():
    """recovery.py:495: a 2x2 grid has no referenced pixel -> AssertionError."""
    with pytest.raises(AssertionError):
        go.reflected_laplacian_matrix(2, 2)
    go.reflected_laplacian_matrix(3, 3)


@pytest.mark.parametrize("name", ["plain", "rgb", "masked", "island"])
def test_poisson_oracle_matches_reference_golden(name):
    """oracle restatement of poisson.solve_poisson_simple (poisson.py:6-99) against outputs of the unmodified reference"""
    import poisson_cases as pc
    kw, want = pc.load(name)
    got = go.solve_poisson_simple(kw["target_gradients"], kw["linear_constraints"], kw["linear_constraints_weight"],
                                  kw.get("mask"), kw.get("mask_value"))
    assert np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("name", ["bilap_plain", "bilap_masked", "multi_lap", "multi_bilap", "multi_free_mask", "selftest_small", "selftest_bilap", "row_bilap", "tiny_multi"])
def test_poisson_oracle_matches_reference_golden_round2(name):
    """the (G.T G)^2 form (poisson.py:65-67), constraint rows over several pixels (poisson.py:329-366) and constraints on
    masked-out pixels (poisson.py:419-423), against outputs of the unmodified reference"""
    import poisson_cases as pc
    kw, want = pc.load2(name)
    got = go.solve_poisson_simple(kw["target_gradients"], kw["linear_constraints"], kw["linear_constraints_weight"],
                                  kw.get("mask"), kw.get("mask_value"), kw["bilaplacian"])
    assert np.abs(got - want).max() <= (1e-7 if kw["bilaplacian"] else 1e-9) * max(1.0, np.abs(want).max())
