"""
TEST INFRASTRUCTURE ONLY.  CPU restatement (numpy / scipy.sparse, float64) of the
grid-diffusion solve path of yig/harmonic_interpolation.  It is the checker the
CUDA path is compared against and the `cpu_baseline` leg of bench.py; the product
package never imports it and has no CPU fallback.

Parity status: PINNED.  The restatement is checked (tests/test_oracle.py) against
  * the reference's golden image  fill_alpha_harmonic-test/expected.png (0 LSB),
  * the reference's self-test asserts (recovery.py:2064, 2069),
  * the disabled `kTest` (fill_alpha_harmonic.py:85-99),
  * outputs of the UNMODIFIED reference run in the build container through
    oracle/ref_shim.py, stored in tests/golden/*.npz by oracle/make_golden.py.

All citations are file:line in the reference tree (yig/harmonic_interpolation).

The solver back end of the reference is third-party and un-vendored (cvxopt ->
SuiteSparse CHOLMOD / UMFPACK, versions unpinned, README.md:30-35).  Both are exact
sparse direct solvers, so any backward-stable direct solver of the same system is
an equivalent oracle; this file uses SuperLU (scipy.sparse.linalg.splu).

Conventions: N = rows*cols, row-major index p = i*cols + j (recovery.py:254-256).
A "down" edge of (i,j) joins (i+1,j); a "right" edge joins (i,j+1).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


# --------------------------------------------------------------------------------------
# cut edges
# --------------------------------------------------------------------------------------

def cut_planes_from_edges(rows, cols, cut_edges):
    """((i,j),(k,l)) pairs -> boolean planes cut_down[i,j], cut_right[i,j].

    Validation follows assert_valid_cut_edges (recovery.py:120-130): in-grid,
    4-adjacent, unique as ordered pairs.  The same edge given in both orientations
    is undefined behaviour in the reference (negative weight); it is
    canonicalised here (cut once)."""
    assert rows > 0 and cols > 0
    cd = np.zeros((rows, cols), dtype=bool)
    cr = np.zeros((rows, cols), dtype=bool)
    if cut_edges is None or len(cut_edges) == 0:
        return cd, cr
    e = np.asarray([(a[0], a[1], b[0], b[1]) for a, b in cut_edges], dtype=np.int64)
    i, j, k, l = e.T
    assert ((i >= 0) & (i < rows) & (j >= 0) & (j < cols)).all()
    assert ((k >= 0) & (k < rows) & (l >= 0) & (l < cols)).all()
    assert (np.abs(i - k) + np.abs(j - l) == 1).all()
    assert len(np.unique(e, axis=0)) == len(e)
    vert = i != k
    cd[np.minimum(i, k)[vert], j[vert]] = True
    cr[i[~vert], np.minimum(j, l)[~vert]] = True
    return cd, cr


def cut_edges_from_mask(mask):
    """recovery.py:94-118: transitions between True and False become cut edges."""
    mask = np.asarray(mask, dtype=bool)
    out = []
    for i, j in zip(*np.where(mask[:-1, :] != mask[1:, :])):
        out.append(((int(i), int(j)), (int(i) + 1, int(j))))
    for i, j in zip(*np.where(mask[:, :-1] != mask[:, 1:])):
        out.append(((int(i), int(j)), (int(i), int(j) + 1)))
    return out


# --------------------------------------------------------------------------------------
# operators, matrix-free and assembled
# --------------------------------------------------------------------------------------

def symmetric_edge_weights(rows, cols, cut_down=None, cut_right=None):
    """Edge weights of gen_symmetric_grid_laplacian2 (recovery.py:232-347).

    wd[i,j] is the weight of edge (i,j)-(i+1,j) (last row: 0), wr[i,j] of (i,j)-(i,j+1).
    Down edges weigh .25, but .125 in the first and last column (recovery.py:263-279);
    right edges .25, but .125 in the first and last row (:283-299).  With cols == 1 the
    loop `for j in (0, cols-1)` visits column 0 twice and the COO constructor sums the
    duplicates, giving .25 (same for rows == 1).  A cut edge has weight 0 (:309-319)."""
    wd = np.full((rows, cols), 0.25)
    wr = np.full((rows, cols), 0.25)
    if cols > 1:
        wd[:, 0] = 0.125
        wd[:, -1] = 0.125
    if rows > 1:
        wr[0, :] = 0.125
        wr[-1, :] = 0.125
    wd[-1, :] = 0.0
    wr[:, -1] = 0.0
    if cut_down is not None:
        wd[cut_down] = 0.0
    if cut_right is not None:
        wr[cut_right] = 0.0
    return wd, wr


def _edge_laplacian_matrix(rows, cols, wd, wr):
    """L = diag(rowsum(Adj)) - Adj for symmetric edge weights (recovery.py:337-338)."""
    N = rows * cols
    idx = np.arange(N).reshape(rows, cols)
    I = []
    J = []
    V = []
    m = wd[:-1, :] != 0
    a = idx[:-1, :][m]
    b = idx[1:, :][m]
    w = wd[:-1, :][m]
    I += [a, b, a, b]
    J += [b, a, a, b]
    V += [-w, -w, w, w]
    m = wr[:, :-1] != 0
    a = idx[:, :-1][m]
    b = idx[:, 1:][m]
    w = wr[:, :-1][m]
    I += [a, b, a, b]
    J += [b, a, a, b]
    V += [-w, -w, w, w]
    return sp.coo_matrix((np.concatenate(V), (np.concatenate(I), np.concatenate(J))), shape=(N, N)).tocsr()


def symmetric_laplacian_matrix(rows, cols, cut_down=None, cut_right=None):
    wd, wr = symmetric_edge_weights(rows, cols, cut_down, cut_right)
    return _edge_laplacian_matrix(rows, cols, wd, wr)


def apply_symmetric_laplacian(x, cut_down=None, cut_right=None):
    """(Lx)_p = sum_{q in N4(p)} w_pq (x_p - x_q); x has shape (rows, cols[, K])."""
    rows, cols = x.shape[:2]
    wd, wr = symmetric_edge_weights(rows, cols, cut_down, cut_right)
    if x.ndim == 3:
        wd = wd[..., None]
        wr = wr[..., None]
    y = np.zeros_like(x, dtype=float)
    d = wd[:-1] * (x[:-1] - x[1:])
    y[:-1] += d
    y[1:] -= d
    d = wr[:, :-1] * (x[:, :-1] - x[:, 1:])
    y[:, :-1] += d
    y[:, 1:] -= d
    return y


def reflected_axis_masks(rows, cols, cut_down=None, cut_right=None):
    """Axis masks of gen_grid_laplacian2_with_boundary_reflection (recovery.py:349-500).

    All directed weights are .25 (:385-400).  A cut (ij,kl) removes ij->kl and
    ij->(2 ij - kl), and the same from kl's side (:422-439); grid borders are treated as
    cuts (:441-449).  Net effect: pixel p keeps its two neighbours along an axis iff both
    exist and both edges are uncut, otherwise it loses both.  av/ah are those masks."""
    cd = np.zeros((rows, cols), bool) if cut_down is None else cut_down
    cr = np.zeros((rows, cols), bool) if cut_right is None else cut_right
    av = np.zeros((rows, cols), dtype=bool)
    ah = np.zeros((rows, cols), dtype=bool)
    if rows > 2:
        av[1:-1, :] = ~cd[:-2, :] & ~cd[1:-1, :]
    if cols > 2:
        ah[:, 1:-1] = ~cr[:, :-2] & ~cr[:, 1:-1]
    return av, ah


def reflected_laplacian_matrix(rows, cols, cut_down=None, cut_right=None, check=True, have_cuts=None):
    av, ah = reflected_axis_masks(rows, cols, cut_down, cut_right)
    N = rows * cols
    idx = np.arange(N).reshape(rows, cols)
    I = []
    J = []
    V = []
    p = idx[av]
    if len(p):
        I += [p, p, p]
        J += [p, p - cols, p + cols]
        V += [np.full(len(p), 0.5), np.full(len(p), -0.25), np.full(len(p), -0.25)]
    p = idx[ah]
    if len(p):
        I += [p, p, p]
        J += [p, p - 1, p + 1]
        V += [np.full(len(p), 0.5), np.full(len(p), -0.25), np.full(len(p), -0.25)]
    if I:
        L = sp.coo_matrix((np.concatenate(V), (np.concatenate(I), np.concatenate(J))), shape=(N, N)).tocsr()
    else:
        L = sp.csr_matrix((N, N))
    if check:
        if have_cuts is None:
            have_cuts = bool((cut_down is not None and cut_down.any()) or (cut_right is not None and cut_right.any()))
        check_reflected_domain(rows, cols, av, ah, have_cuts)
    return L


def check_reflected_domain(rows, cols, av, ah, have_cuts):
    """The reference's structural asserts on the reflected operator:
    recovery.py:493 (number of all-zero rows is 4, or 2 for a 1-D grid, unless cuts were
    given) and :495 (every column is non-empty, i.e. every pixel is referenced by a row)."""
    empty_rows = int((~(av | ah)).sum())
    assert empty_rows == (2 if rows == 1 or cols == 1 else 4) or have_cuts
    ref = av | ah
    ref[:-1, :] |= av[1:, :]
    ref[1:, :] |= av[:-1, :]
    ref[:, :-1] |= ah[:, 1:]
    ref[:, 1:] |= ah[:, :-1]
    assert ref.all()


def apply_reflected_laplacian(x, cut_down=None, cut_right=None):
    """(L_refl x)_p = 1/4 sum_axis a_axis(p) (2 x_p - x_{p-e} - x_{p+e})."""
    rows, cols = x.shape[:2]
    av, ah = reflected_axis_masks(rows, cols, cut_down, cut_right)
    if x.ndim == 3:
        av = av[..., None]
        ah = ah[..., None]
    u = np.zeros_like(x, dtype=float)
    if rows > 2:
        u[1:-1] += 0.25 * av[1:-1] * (2 * x[1:-1] - x[:-2] - x[2:])
    if cols > 2:
        u[:, 1:-1] += 0.25 * ah[:, 1:-1] * (2 * x[:, 1:-1] - x[:, :-2] - x[:, 2:])
    return u


def apply_bilaplacian(x, cut_down=None, cut_right=None):
    """B x = L_refl^T (L_refl x)  (recovery.py:911, 1039, 1119, 1247), matrix-free:
    (L^T u)_p = 1/4 sum_axis [ 2 a(p) u_p - a(p-e) u_{p-e} - a(p+e) u_{p+e} ] with zero extension."""
    rows, cols = x.shape[:2]
    av, ah = reflected_axis_masks(rows, cols, cut_down, cut_right)
    if x.ndim == 3:
        av = av[..., None]
        ah = ah[..., None]
    u = apply_reflected_laplacian(x, cut_down, cut_right)
    y = np.zeros_like(u)
    t = av * u
    y += 0.5 * t
    y[:-1] -= 0.25 * t[1:]
    y[1:] -= 0.25 * t[:-1]
    t = ah * u
    y += 0.5 * t
    y[:, :-1] -= 0.25 * t[:, 1:]
    y[:, 1:] -= 0.25 * t[:, :-1]
    return y


# --------------------------------------------------------------------------------------
# the problem in array form, its assembled system, and the direct solve
# --------------------------------------------------------------------------------------

class GridProblem:
    """Everything the reference's entry points can express, as arrays.

    hard       bool (rows, cols)      hard value constraints (A6/A7)
    hard_vals  float (rows, cols, K)  read where hard
    soft_w     float (rows, cols)     per-pixel least-squares weight, 0 where absent (A5)
    soft_vals  float (rows, cols, K)
    cut_down / cut_right   bool (rows, cols)
    pen_down / pen_right   bool (rows, cols)   edges carrying a dx / dy constraint (A4)
    d_down / d_right       float (rows, cols, K)   the constrained differences
    w_g        float          weight of the gradient constraints (recovery.py:879)
    bilaplacian bool
    """

    def __init__(self, rows, cols, K=1, bilaplacian=False):
        self.rows, self.cols, self.K = rows, cols, K
        self.bilaplacian = bool(bilaplacian)
        self.hard = np.zeros((rows, cols), bool)
        self.hard_vals = np.zeros((rows, cols, K))
        self.soft_w = np.zeros((rows, cols))
        self.soft_vals = np.zeros((rows, cols, K))
        self.cut_down = np.zeros((rows, cols), bool)
        self.cut_right = np.zeros((rows, cols), bool)
        self.pen_down = np.zeros((rows, cols), bool)
        self.pen_right = np.zeros((rows, cols), bool)
        self.d_down = np.zeros((rows, cols, K))
        self.d_right = np.zeros((rows, cols, K))
        self.w_g = 0.0
        self.have_cuts = False


def energy_matrix(prob, check=True):
    if prob.bilaplacian:
        L = reflected_laplacian_matrix(prob.rows, prob.cols, prob.cut_down, prob.cut_right,
                                       check=check, have_cuts=prob.have_cuts)
        return (L.T @ L).tocsr()
    return symmetric_laplacian_matrix(prob.rows, prob.cols, prob.cut_down, prob.cut_right)


def full_system(prob, check=True):
    """system and rhs BEFORE hard-constraint elimination:
    system = L + C^T W C, rhs = C^T W Crhs (recovery.py:926-927, 1280-1283, 1319-1321)."""
    rows, cols, K = prob.rows, prob.cols, prob.K
    N = rows * cols
    A = energy_matrix(prob, check)
    b = np.zeros((N, K))
    sw = prob.soft_w.ravel()
    if sw.any():
        A = A + sp.diags(sw)
        b += sw[:, None] * prob.soft_vals.reshape(N, K)
    if prob.w_g != 0 and (prob.pen_down.any() or prob.pen_right.any()):
        pd = prob.pen_down.copy()
        pd[-1, :] = False
        pr = prob.pen_right.copy()
        pr[:, -1] = False
        G = _edge_laplacian_matrix(rows, cols, prob.w_g * pd, prob.w_g * pr)
        A = A + G
        # C^T W Crhs: +w d at the far endpoint, -w d at (i,j)  (recovery.py:549-557)
        dd = prob.w_g * pd[..., None] * prob.d_down
        dr = prob.w_g * pr[..., None] * prob.d_right
        bb = np.zeros((rows, cols, K))
        bb -= dd
        bb[1:] += dd[:-1]
        bb -= dr
        bb[:, 1:] += dr[:, :-1]
        b += bb.reshape(N, K)
    return A.tocsr(), b


def apply_full_operator(prob, x):
    """Matrix-free y = system @ x (no hard elimination), x of shape (rows, cols, K)."""
    if prob.bilaplacian:
        y = apply_bilaplacian(x, prob.cut_down, prob.cut_right)
    else:
        y = apply_symmetric_laplacian(x, prob.cut_down, prob.cut_right)
    y = y + prob.soft_w[..., None] * x
    if prob.w_g != 0:
        pd = prob.pen_down.copy()
        pd[-1, :] = False
        pr = prob.pen_right.copy()
        pr[:, -1] = False
        d = prob.w_g * pd[:-1, :, None] * (x[:-1] - x[1:])
        y[:-1] += d
        y[1:] -= d
        d = prob.w_g * pr[:, :-1, None] * (x[:, :-1] - x[:, 1:])
        y[:, :-1] += d
        y[:, 1:] -= d
    return y


def reduced_system(prob, check=True):
    """Hard-constraint elimination (recovery.py:647-777): x_c = val_c and
    A_ff x_f = b_f - A_fc x_c.  Returns (A_ff, b_f, free_index)."""
    A, b = full_system(prob, check)
    N = prob.rows * prob.cols
    hard = prob.hard.ravel()
    free = np.flatnonzero(~hard)
    cons = np.flatnonzero(hard)
    xc = prob.hard_vals.reshape(N, prob.K)[cons]
    Af = A[free]
    bf = b[free] - Af[:, cons] @ xc
    return Af[:, free].tocsc(), bf, free


def solve_problem(prob, check=True):
    """Direct solve; returns float64 (rows, cols, K)."""
    N = prob.rows * prob.cols
    Aff, bf, free = reduced_system(prob, check)
    x = np.where(prob.hard[..., None], prob.hard_vals, 0.0).reshape(N, prob.K).copy()
    if len(free):
        x[free] = spla.splu(Aff).solve(bf)
    return x.reshape(prob.rows, prob.cols, prob.K)


def residual_norms(prob, x):
    """||b_f - A_ff x_f||_2 / ||b_f||_2 per channel on the reduced free-dof system."""
    Aff, bf, free = reduced_system(prob, check=False)
    xf = x.reshape(-1, prob.K)[free]
    r = bf - Aff @ xf
    return np.linalg.norm(r, axis=0) / np.maximum(np.linalg.norm(bf, axis=0), 1e-300)


# --------------------------------------------------------------------------------------
# the reference's entry points, same signatures (tuple lists in, float64 arrays out)
# --------------------------------------------------------------------------------------

def _set_cuts(prob, cut_edges):
    prob.cut_down, prob.cut_right = cut_planes_from_edges(prob.rows, prob.cols, cut_edges)
    prob.have_cuts = cut_edges is not None and len(cut_edges) > 0


def solve_grid_linear(rows, cols, dx_constraints=None, dy_constraints=None, value_constraints=None,
                      bilaplacian=False, iterative=False, cut_edges=None, w_lsq=None):
    """recovery.py:838-1009.  w_smooth = 1, w_gradients = 1e5 unless w_lsq is given (:875-884);
    a SOFT pin (0,0,0.) of weight w_gradients is added iff gradient constraints exist and no
    value constraints do (:920).  `iterative` selects the solver in the reference, not the
    system; the oracle always solves directly."""
    dx_constraints = [] if dx_constraints is None else dx_constraints
    dy_constraints = [] if dy_constraints is None else dy_constraints
    value_constraints = [] if value_constraints is None else value_constraints
    if w_lsq is None:
        w_g = 1e5
    else:
        assert float(w_lsq) == w_lsq
        w_g = float(w_lsq)
    prob = GridProblem(rows, cols, 1, bilaplacian)
    _set_cuts(prob, cut_edges)
    prob.w_g = w_g
    # gen_constraint_system's index and duplicate asserts (recovery.py:523-537)
    assert all(0 <= i < rows - 1 and 0 <= j < cols for i, j, _ in dx_constraints)
    assert all(0 <= i < rows and 0 <= j < cols - 1 for i, j, _ in dy_constraints)
    assert all(0 <= i < rows and 0 <= j < cols for i, j, _ in value_constraints)
    assert len(set((i, j) for i, j, _ in dx_constraints)) == len(dx_constraints)
    assert len(set((i, j) for i, j, _ in dy_constraints)) == len(dy_constraints)
    assert len(set((i, j) for i, j, _ in value_constraints)) == len(value_constraints)
    for i, j, d in dx_constraints:
        prob.pen_down[i, j] = True
        prob.d_down[i, j, 0] = d
    for i, j, d in dy_constraints:
        prob.pen_right[i, j] = True
        prob.d_right[i, j, 0] = d
    if len(dx_constraints) + len(dy_constraints) > 0 and len(value_constraints) == 0:
        prob.soft_w[0, 0] = w_g
        prob.soft_vals[0, 0, 0] = 0.0
    for i, j, v in value_constraints:
        prob.hard[i, j] = True
        prob.hard_vals[i, j, 0] = v
    return solve_problem(prob)[:, :, 0]


def _vector_values(values):
    vals = [np.asarray(v, dtype=float) for v in values]
    assert all(v.ndim == 1 for v in vals)
    dims = set(len(v) for v in vals)
    assert len(dims) == 1
    return vals, dims.pop()


def solve_grid_linear_simple(rows, cols, value_constraints, bilaplacian=False):
    """recovery.py:1011-1088: hard K-vector value constraints only, no cuts."""
    assert len(value_constraints) > 0
    vals, K = _vector_values([v for _, _, v in value_constraints])
    assert len(set((i, j) for i, j, _ in value_constraints)) == len(value_constraints)
    prob = GridProblem(rows, cols, K, bilaplacian)
    for (i, j, _), v in zip(value_constraints, vals):
        prob.hard[i, j] = True
        prob.hard_vals[i, j] = v
    return solve_problem(prob)


def solve_grid_linear_simple3_solver(rows, cols, value_constraints_hard=(), value_constraints_soft=(),
                                     w_lsq=1.0, cut_edges=None, bilaplacian=False):
    """recovery.py:1202-1336.  Hard wins over soft at the same pixel (the hard row replaces
    the row after the soft diagonal was added, :1286-1290)."""
    hard = [tuple(c) for c in value_constraints_hard]
    soft = [tuple(c) for c in value_constraints_soft]
    assert len(hard) + len(soft) > 0
    w = np.asarray(w_lsq, dtype=float)
    assert (w >= 0).all()
    if w.ndim > 0:
        assert len(w) == len(soft) or len(soft) == 0
    else:
        w = np.full(len(soft), float(w))
    assert len(set(hard)) == len(hard)
    assert len(set(soft)) == len(soft)
    assert all(0 <= i < rows and 0 <= j < cols for i, j in hard + soft)

    def solve(hard_values=(), soft_values=()):
        assert len(hard) == len(hard_values)
        assert len(soft) == len(soft_values)
        vals, K = _vector_values(list(hard_values) + list(soft_values))
        prob = GridProblem(rows, cols, K, bilaplacian)
        _set_cuts(prob, cut_edges)
        for (i, j), wk, v in zip(soft, w, vals[len(hard):]):
            prob.soft_w[i, j] = wk
            prob.soft_vals[i, j] = v
        for (i, j), v in zip(hard, vals[:len(hard)]):
            prob.hard[i, j] = True
            prob.hard_vals[i, j] = v
        return solve_problem(prob)

    return solve


def solve_grid_linear_simple3(rows, cols, value_constraints_hard=(), value_constraints_soft=(),
                              w_lsq=1.0, cut_edges=None, bilaplacian=False):
    """recovery.py:1338-1344."""
    solve = solve_grid_linear_simple3_solver(
        rows, cols, [(i, j) for i, j, _ in value_constraints_hard],
        [(i, j) for i, j, _ in value_constraints_soft], w_lsq, cut_edges, bilaplacian)
    return solve([v for _, _, v in value_constraints_hard], [v for _, _, v in value_constraints_soft])


def grow(mask, pixel_iterations):
    """highlighting.py:28-74: diamond dilation by repeated 4-neighbour OR."""
    assert int(pixel_iterations) == pixel_iterations and pixel_iterations >= 0
    mask = np.array(mask, dtype=bool)
    for _ in range(int(pixel_iterations)):
        m = mask.copy()
        mask[:-1, :] |= m[1:, :]
        mask[1:, :] |= m[:-1, :]
        mask[:, :-1] |= m[:, 1:]
        mask[:, 1:] |= m[:, :-1]
    return mask


def smooth_bumps(grid, bump_locations, bump_radius, **kw):
    """recovery.py:1346-1418: re-solve the dilated bump neighbourhood with every other
    pixel hard-constrained at its current value."""
    grid = np.asarray(grid, dtype=float)
    assert grid.ndim == 2 and grid.shape[0] > 0 and grid.shape[1] > 0
    loc = np.asarray(bump_locations, dtype=int)
    assert loc.ndim == 2 and loc.shape[1] == 2
    assert loc[:, 0].min() >= 0 and loc[:, 1].min() >= 0
    assert loc[:, 0].max() < grid.shape[0] and loc[:, 1].max() < grid.shape[1]
    assert bump_radius >= 0 and int(bump_radius) == bump_radius
    mask = np.zeros(grid.shape, bool)
    mask[tuple(loc.T)] = True
    mask = ~grow(mask, int(bump_radius))
    vc = [(int(i), int(j), grid[i, j]) for i, j in zip(*np.where(mask))]
    kw.pop("w_lsq", None)
    return solve_grid_linear(grid.shape[0], grid.shape[1], value_constraints=vc, **kw)


def smooth_bumps2(grid, bump_locations, bump_radius, smooth_bumps_iterations, smooth_func=None, **kw):
    """recovery.py:1420-1530: repeatedly replace the gradient at each bump by the
    median / average gradient of its neighbourhood and re-solve."""
    grid = np.asarray(grid, dtype=float)
    assert grid.ndim == 2 and grid.shape[0] > 0 and grid.shape[1] > 0
    loc = np.asarray(bump_locations, dtype=int)
    assert loc.ndim == 2 and loc.shape[1] == 2
    assert loc[:, 0].min() >= 0 and loc[:, 1].min() >= 0
    assert loc[:, 0].max() < grid.shape[0] - 1 and loc[:, 1].max() < grid.shape[1] - 1
    assert bump_radius >= 0 and int(bump_radius) == bump_radius
    assert smooth_bumps_iterations >= 0 and int(smooth_bumps_iterations) == smooth_bumps_iterations
    smooth_func = "median" if smooth_func is None else smooth_func
    assert smooth_func in ("median", "average")
    if int(smooth_bumps_iterations) == 0:
        return np.array(grid)
    f = {"median": np.median, "average": np.average}[smooth_func]
    nbrs = []
    for l in loc:
        m = np.zeros(grid[:-1, :-1].shape, bool)
        m[tuple(l)] = True
        nbrs.append(np.where(grow(m, int(bump_radius))))
    for _ in range(int(smooth_bumps_iterations)):
        dxs = [(int(l[0]), int(l[1]), f((grid[1:, :] - grid[:-1, :])[ix])) for l, ix in zip(loc, nbrs)]
        dys = [(int(l[0]), int(l[1]), f((grid[:, 1:] - grid[:, :-1])[ix])) for l, ix in zip(loc, nbrs)]
        grid = solve_grid_linear(grid.shape[0], grid.shape[1], dx_constraints=dxs, dy_constraints=dys, **kw)
    return grid


def fill_alpha_harmonic(arr, mask):
    """fill_alpha_harmonic.py:6-25: keep pixels where mask is True, harmonic elsewhere."""
    arr = np.asarray(arr, dtype=float)
    if arr.ndim == 2:
        arr = arr[..., None]
    mask = np.asarray(mask, dtype=bool)
    assert mask.any()
    prob = GridProblem(arr.shape[0], arr.shape[1], arr.shape[2], False)
    prob.hard = mask.copy()
    prob.hard_vals = arr.copy()
    return solve_problem(prob)


def fill_rgba8(rgba):
    """The CLI's array pipeline (fill_alpha_harmonic.py:27-45, 72-76): uint8 RGBA in,
    uint8 RGB out; keep-mask round(255*alpha) == 255; output round-half-even, clipped."""
    rgba = np.asarray(rgba, dtype=np.uint8)
    arr = rgba / 255.0
    mask = (255 * arr[:, :, 3]).round(0).astype(int) == 255
    res = fill_alpha_harmonic(arr[:, :, :3], mask)
    return np.asarray((res * 255).round(0).clip(0, 255), dtype=np.uint8)


# --------------------------------------------------------------------------------------
# poisson.py: gradient-domain integration
# --------------------------------------------------------------------------------------
def gradient_operator(rows, cols, mask=None):
    """poisson.py:204-251, vectorised: G (2 N x N); row (i*cols + j)*2 + k holds the forward difference of
    pixel (i,j) along axis k (0: i, 1: j); edges across a mask border are left out."""
    idx = np.arange(rows * cols).reshape(rows, cols)
    down = np.ones((rows, cols), bool); down[-1, :] = False
    right = np.ones((rows, cols), bool); right[:, -1] = False
    if mask is not None:
        mask = np.asarray(mask, bool)
        down[:-1] &= mask[:-1] == mask[1:]
        right[:, :-1] &= mask[:, :-1] == mask[:, 1:]
    I, J, V = [], [], []
    p = idx[down]
    I += [2 * p, 2 * p]; J += [p, p + cols]; V += [-np.ones(len(p)), np.ones(len(p))]
    p = idx[right]
    I += [2 * p + 1, 2 * p + 1]; J += [p, p + 1]; V += [-np.ones(len(p)), np.ones(len(p))]
    return sp.coo_matrix((np.concatenate(V), (np.concatenate(I), np.concatenate(J))), shape=(2 * rows * cols, rows * cols)).tocsr()


def solve_poisson_simple(target_gradients, linear_constraints, linear_constraints_weight, mask=None, mask_value=None, bilaplacian=False):
    """poisson.py:6-99: system = G.T G -- or (G.T G)^2 with G <- G (G.T G) for `bilaplacian`, poisson.py:65-67 --
    (+ Dirichlet rows for masked pixels, poisson.py:76-82) + w A.T A, rhs = G.T u + w A.T b (poisson.py:85-89, 329-366);
    sparse direct solve."""
    u = np.asarray(target_gradients, float)
    rows, cols, _, K = u.shape
    N = rows * cols
    G = gradient_operator(rows, cols, mask)
    system = (G.T @ G).tocsr()
    if bilaplacian:
        G = (G @ system).tocsr()
        system = (system.T @ system).tocsr()
    rhs = G.T @ u.reshape(N * 2, K)
    if mask is not None and mask_value is not None:
        c = np.flatnonzero(~np.asarray(mask, bool).ravel())
        vals = np.tile(np.asarray(mask_value, float), (len(c), 1))
        rhs = rhs - system[:, c] @ vals            # symmetric elimination (poisson.py:253-327)
        keep = np.ones(N); keep[c] = 0.0
        D = sp.diags(keep)
        system = (D @ system @ D + sp.diags(1.0 - keep)).tocsr()
        rhs[c] = vals
    I, J, V = [], [], []
    b = np.zeros((len(linear_constraints), K))
    for n, (coeffs, value) in enumerate(linear_constraints):
        for i, j, cf in coeffs:
            I.append(n); J.append(i * cols + j); V.append(cf)
        b[n] = value
    A = sp.coo_matrix((V, (I, J)), shape=(len(linear_constraints), N)).tocsr()
    system = system + linear_constraints_weight * (A.T @ A)
    rhs = rhs + linear_constraints_weight * (A.T @ b)
    x = spla.splu(system.tocsc()).solve(rhs)
    return x.reshape(rows, cols, K)
