"""
Drop-in for `poisson.solve_poisson_simple` of the reference (poisson.py:6-99): gradient-domain
integration on a pixel grid with a mask and soft linear constraints.  The reference assembles the
gradient operator G in Python loops (poisson.py:204-251), forms L = G.T*G (or its square for
`bilaplacian=True`, poisson.py:65-67) and calls a sparse direct solver; here L/4 is the library's
Laplacian with uniform 1/4 edge weights (`hg_set_edge_weights`), mask borders are cut edges,
masked-out pixels with a `mask_value` are hard pixels and G.T*u is handed to the solver as an
explicit right-hand side (`hg_values.rhs`).

Two device paths:
  * the Laplacian form with constraints of ONE pixel each (`([(row, col, coeff)], rhs)` -- what every
    caller in the reference passes): the constraints are soft value constraints of the standard MG-PCG
    solve (`hg_solve`);
  * everything else -- `bilaplacian=True`, constraint rows over several pixels (poisson.py:329-366),
    constraints on masked-out pixels (the reference's own test pins the four corners of the masked-out
    region, poisson.py:419-423) --: `hg_solve_composed`, PCG on (L^m + w A.T A) with the V-cycle of
    L + diag as the preconditioner (csrc/solver_composed.cu).
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
    if mask is not None:
        mask = np.asarray(mask, bool)

    rhs, down, right = gradient_rhs(target_gradients, mask)
    w = float(linear_constraints_weight)
    hard = None
    if mask is not None and mask_value is not None:
        hard = ~mask
    rows_of = []
    for coeffs, value in linear_constraints:
        ent = [(int(i), int(j), float(c)) for i, j, c in coeffs]
        for i, j, _ in ent:
            assert 0 <= i < rows and 0 <= j < cols
        rows_of.append((ent, np.asarray(value, float).reshape(K)))
    general = bool(bilaplacian) or any(len(ent) != 1 for ent, _ in rows_of) or \
        (hard is not None and any(hard[i, j] for ent, _ in rows_of for i, j, _ in ent))
    if general:
        return _solve_composed(rows, cols, K, rhs, down, right, rows_of, w, hard, mask_value, bool(bilaplacian), rtol, maxiter)

    # one-pixel linear constraints coeff * x[row, col] = value, weight w: + w coeff^2 on the diagonal, + w coeff value on the rhs
    S = np.zeros((rows, cols))
    Sv = np.zeros((rows, cols, K))
    for ((i, j, c),), value in rows_of:
        S[i, j] += w * c * c
        Sv[i, j] += w * c * value
    hard_vals = None
    if hard is not None:
        hard_vals = np.zeros((rows, cols, K))
        hard_vals[hard] = np.asarray(mask_value, float)
    soft = S > 0
    soft_vals = np.zeros((rows, cols, K))
    soft_vals[soft] = Sv[soft] / S[soft][:, None]

    # the library solves (G.T G / 4 + S / 4) x = G.T u / 4 + S v / 4
    flags = encode_flags(rows, cols, hard=hard, soft=soft, cut_down=~down, cut_right=~right)
    s = GridSolver(rows, cols, K, False, precision, uniform_edges=True)
    s.set_pattern(flags, 0.25 * S, 0.0, 0.0)
    x = s.solve(hard_vals=hard_vals, soft_vals=soft_vals, rhs=0.25 * rhs, rtol=rtol, maxiter=maxiter)
    s.close()
    return x.reshape(rows, cols, -1)


def components(free, down, right):
    """Connected components of the free pixels over the edges that exist (down[i, j]: (i, j)-(i+1, j), right[i, j]:
    (i, j)-(i, j+1)).  Returns (labels, count), labels (rows, cols) int, -1 at pixels that are not free.  Row runs are found
    with array operations, the runs of adjacent rows are merged by a union-find over the (few) distinct run pairs."""
    free = np.asarray(free, bool)
    linked = np.zeros_like(free)
    linked[:, 1:] = free[:, 1:] & free[:, :-1] & right[:, :-1]
    start = free & ~linked
    run = np.cumsum(start.ravel()).reshape(free.shape) - 1      # the run a free pixel belongs to = the latest run start
    nr = int(start.sum())
    vert = free[:-1] & free[1:] & down[:-1]
    pairs = np.unique(np.stack([run[:-1][vert], run[1:][vert]], 1), axis=0) if vert.any() else np.zeros((0, 2), int)
    parent = list(range(nr))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    for a, b in pairs:
        ra, rb = find(int(a)), find(int(b))
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    root = np.array([find(r) for r in range(nr)], int)
    uniq, inv = np.unique(root, return_inverse=True)
    labels = np.where(free, inv[np.clip(run, 0, max(nr - 1, 0))] if nr else -1, -1)
    return labels, len(uniq)


def check_constrained(free, down, right, cons):
    """The null space of L^m + w A.T A is spanned by the per-component constants v with A v = 0: the system is singular
    unless the (rows x components) matrix of per-component coefficient sums has full column rank.  The reference's direct
    solver fails on such a system (cvxopt.cholmod raises ArithmeticError, poisson.py:95); an iterative solver would quietly
    return one of the solutions."""
    labels, nc = components(free, down, right)
    B = np.zeros((len(cons), nc))
    flat = labels.ravel()
    for n, (pix, coef, _) in enumerate(cons):
        np.add.at(B[n], flat[np.asarray(pix, int)], np.asarray(coef, float))
    if nc and np.linalg.matrix_rank(B) < nc:
        raise ArithmeticError("singular system: the linear constraints do not fix the constant of every connected component "
                              "of the domain (%d components, rank %d)" % (nc, np.linalg.matrix_rank(B)))


def _solve_composed(rows, cols, K, rhs, down, right, rows_of, w, hard, mask_value, squared, rtol, maxiter):
    """(L^m + w A.T A) x = L^(m-1) G.T u + w A.T c through hg_solve_composed (fp64), m = 2 for `bilaplacian`.

    Masked-out pixels that a constraint touches stay unknowns, as in the reference: there the value constraints turn
    their rows into identity rows first (poisson.py:76-82) and `system += w*C` then adds the constraint terms on top
    (poisson.py:85-89).  Such a pixel is handed to the library as a free pixel without edges whose identity row is one
    more constraint row (coefficient 1/sqrt(w), so that w c^2 = 1)."""
    assert w > 0, "linear_constraints_weight must be positive"
    down, right, rhs = down.copy(), right.copy(), rhs.copy()
    cons = [([i * cols + j for i, j, _ in ent], [c for _, _, c in ent], value) for ent, value in rows_of]
    hard_vals = None
    if hard is not None:
        hard = hard.copy()
        hard_vals = np.zeros((rows, cols, K))
        hard_vals[hard] = np.asarray(mask_value, float)
        for i, j in sorted({(i, j) for ent, _ in rows_of for i, j, _ in ent if hard[i, j]}):
            hard[i, j] = False
            down[i, j] = False; right[i, j] = False
            if i > 0: down[i - 1, j] = False
            if j > 0: right[i, j - 1] = False
            rhs[i, j] = 0.0                        # (the reference overwrites the right-hand side with mask_value, poisson.py:289)
            cons.append(([i * cols + j], [1.0 / np.sqrt(w)], np.asarray(mask_value, float) / np.sqrt(w)))
    check_constrained(np.ones((rows, cols), bool) if hard is None else ~hard, down, right, cons)
    # the library's L is G.T G / 4: the system is divided by 4 (m = 1) or 16 (m = 2); the preconditioner's diagonal D is
    # diag(w' A.T A), or its square root for m = 2 ((L + D)^2 = L^2 + D^2 + a term of low rank)
    wl = w / 16.0 if squared else w / 4.0
    D = np.zeros(rows * cols)
    for pix, coef, _ in cons:
        np.add.at(D, pix, wl * np.square(coef))
    if squared:
        D = np.sqrt(D)
    D = D.reshape(rows, cols)
    flags = encode_flags(rows, cols, hard=hard, soft=D > 0, cut_down=~down, cut_right=~right)
    s = GridSolver(rows, cols, K, False, "fp64", uniform_edges=True)
    s.set_pattern(flags, D, 0.0, 0.0)
    x = s.solve_composed(cons, wl, squared, hard_vals=hard_vals, rhs=0.25 * rhs, rtol=rtol, maxiter=maxiter,
                         accept=1e-8 if squared else None)
    s.close()
    return x.reshape(rows, cols, -1)
