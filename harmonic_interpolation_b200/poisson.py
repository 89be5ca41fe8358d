"""
Drop-in for `poisson.solve_poisson_simple` of the reference (poisson.py:6-99): gradient-domain
integration on a pixel grid with a mask and soft linear constraints.  The reference assembles the
gradient operator G in Python loops (poisson.py:204-251), forms L = G.T*G and calls a sparse direct
solver; here L/4 is the library's Laplacian with uniform 1/4 edge weights (`hg_set_edge_weights`),
mask borders are cut edges, masked-out pixels with a `mask_value` are hard pixels, one-pixel linear
constraints are soft value constraints, and G.T*u is handed to the solver as an explicit right-hand
side (`hg_values.rhs`).

Supported: the Laplacian form (`bilaplacian=False`) with linear constraints that involve ONE pixel each
(`([(row, col, coeff)], rhs)` -- what every caller in the reference passes).  Constraints coupling
several pixels and the `bilaplacian=True` variant ((G.T G)^2, a different 13-point operator than
recovery.py's) raise NotImplementedError.
"""
import numpy as np

from .highlighting import grow, shrink  # noqa: F401  (poisson.py:154-202 re-exports them)
from .solver import GridSolver, encode_flags


def gradient_rhs(target_gradients, mask=None):
    """G.T * target_gradients for the gradient operator of poisson.py:204-251, vectorised:
    (rows, cols, 2, K) -> (rows, cols, K).  Edges across a mask border do not exist."""
    u = np.asarray(target_gradients, float)
    rows, cols = u.shape[:2]
    down = np.ones((rows, cols), bool); down[-1, :] = False
    right = np.ones((rows, cols), bool); right[:, -1] = False
    if mask is not None:
        down[:-1] &= mask[:-1] == mask[1:]
        right[:, :-1] &= mask[:, :-1] == mask[:, 1:]
    ud = u[:, :, 0, :] * down[:, :, None]
    ur = u[:, :, 1, :] * right[:, :, None]
    out = -ud - ur
    out[1:] += ud[:-1]
    out[:, 1:] += ur[:, :-1]
    return out, down, right


def solve_poisson_simple(target_gradients, linear_constraints=None, linear_constraints_weight=None, mask=None,
                         mask_value=None, bilaplacian=False, *, precision=None, rtol=None, maxiter=None):
    """poisson.py:6-99.  Returns (rows, cols, K) float64."""
    target_gradients = np.asarray(target_gradients, float)
    print('solve_poisson_simple( target_gradients.shape = %s, |constraints| = %s, linear_constraints_weight = %s, |mask == False| = %s )' % (
        target_gradients.shape, len(linear_constraints), linear_constraints_weight,
        None if mask is None else (mask == False).sum()))  # noqa: E712

    assert len(target_gradients.shape) == 4
    rows, cols = target_gradients.shape[:2]
    assert rows > 0 and cols > 0
    assert target_gradients.shape[2] == 2
    K = target_gradients.shape[3]
    assert mask is None or mask.shape == target_gradients.shape[:2]
    assert mask_value is None or len(mask_value) == K
    assert (linear_constraints is None) or (linear_constraints_weight is not None)
    assert len(linear_constraints) > 0
    if bilaplacian:
        raise NotImplementedError("solve_poisson_simple(bilaplacian=True): (G.T G)^2 is not among the library's operators")
    if mask is not None:
        mask = np.asarray(mask, bool)

    rhs, down, right = gradient_rhs(target_gradients, mask)

    # one-pixel linear constraints coeff * x[row, col] = value, weight w: + w coeff^2 on the diagonal, + w coeff value on the rhs
    w = float(linear_constraints_weight)
    S = np.zeros((rows, cols))
    Sv = np.zeros((rows, cols, K))
    for coeffs, value in linear_constraints:
        if len(coeffs) != 1:
            raise NotImplementedError("linear constraints coupling several pixels are not supported")
        (i, j, c), = coeffs
        assert 0 <= i < rows and 0 <= j < cols
        S[i, j] += w * c * c
        Sv[i, j] += w * c * np.asarray(value, float).reshape(K)

    hard = None
    hard_vals = None
    if mask is not None and mask_value is not None:
        hard = ~mask
        hard_vals = np.zeros((rows, cols, K))
        hard_vals[hard] = np.asarray(mask_value, float)
    soft = S > 0
    if hard is not None:
        soft &= ~hard          # the value rows replace whatever else touches a masked pixel (poisson.py:76-82 before :85-89
                               # would add the constraint to the identity row; a constrained masked pixel is not a meaningful input)
    soft_vals = np.zeros((rows, cols, K))
    soft_vals[soft] = Sv[soft] / S[soft][:, None]

    # the library solves (G.T G / 4 + S / 4) x = G.T u / 4 + S v / 4
    flags = encode_flags(rows, cols, hard=hard, soft=soft, cut_down=~down, cut_right=~right)
    s = GridSolver(rows, cols, K, False, precision, uniform_edges=True)
    s.set_pattern(flags, 0.25 * S, 0.0, 0.0)
    x = s.solve(hard_vals=hard_vals, soft_vals=soft_vals, rhs=0.25 * rhs, rtol=rtol, maxiter=maxiter)
    s.close()
    return x.reshape(rows, cols, -1)
