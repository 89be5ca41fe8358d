"""
Drop-in for the grid-diffusion entry points of the reference's `recovery` module: same
names, argument meaning, return shapes, asserts and printed header lines -- but the
operator assembly (scipy.sparse, recovery.py:232-500), the constraint algebra
(recovery.py:504-836) and the direct solve (cvxopt, recovery.py:993, 1075, 1304-1327) are
replaced by one flag byte per pixel and CUDA kernels in libhgrid.so (include/hgrid.h).

Extensions (new surface, not reference behaviour), needed where lists of Python tuples are
infeasible:
  * constraints may be given as arrays: `(ij, values)` with ij an (n,2) integer array, or
    `(mask, values)` with a boolean (rows, cols) mask and a dense (rows, cols[, K]) array;
  * `cut_edges` may be an (n,4) integer array or a `(cut_down, cut_right)` pair of masks;
  * keyword-only `precision=`, `rtol=`, `maxiter=` (defaults: HGRID_PRECISION / HGRID_RTOL /
    HGRID_MAXITER, else fp64 and 1e-12, chosen to agree with the reference's float64
    direct solve to ~1e-9 of the value range).
"""
import numpy as np

from . import _capi as C
from .highlighting import grow
from .solver import GridSolver, encode_flags
from .tictoc import tic, toc


# --------------------------------------------------------------------------------------
# helpers with the reference's names
# --------------------------------------------------------------------------------------

def zero_dx_dy_constraints_from_cut_edges(cut_edges):
    """recovery.py:68-92: dx / dy constraints of value 0 across every cut edge."""
    dx_constraints = []
    dy_constraints = []
    for (i, j), (k, l) in cut_edges:
        assert i == k or j == l
        assert abs(i - k) + abs(j - l) == 1
        if i != k:
            dx_constraints.append((min(i, k), j, 0.0))
        else:
            dy_constraints.append((i, min(j, l), 0.0))
    return dx_constraints, dy_constraints


def cut_edges_from_mask(mask):
    """recovery.py:94-118: True/False transitions of `mask` become cut edges, vertical
    transitions (row i | row i+1) first, then horizontal ones."""
    mask = np.asarray(mask, dtype=bool)
    ii, jj = np.nonzero(mask[:-1, :] != mask[1:, :])
    edges = [((i, j), (i + 1, j)) for i, j in zip(ii.tolist(), jj.tolist())]
    ii, jj = np.nonzero(mask[:, :-1] != mask[:, 1:])
    edges += [((i, j), (i, j + 1)) for i, j in zip(ii.tolist(), jj.tolist())]
    return edges


def cut_planes_from_mask(mask):
    """Array-native form of cut_edges_from_mask: (cut_down, cut_right) boolean planes."""
    mask = np.asarray(mask, dtype=bool)
    cd = np.zeros(mask.shape, bool)
    cr = np.zeros(mask.shape, bool)
    cd[:-1, :] = mask[:-1, :] != mask[1:, :]
    cr[:, :-1] = mask[:, :-1] != mask[:, 1:]
    return cd, cr


def _cut_edge_array(cut_edges):
    if isinstance(cut_edges, np.ndarray):
        e = cut_edges.reshape(-1, 4).astype(np.int64)
    else:
        e = np.asarray([(a[0], a[1], b[0], b[1]) for a, b in cut_edges], dtype=np.int64).reshape(-1, 4)
    return e


def assert_valid_cut_edges(rows, cols, cut_edges):
    """recovery.py:120-130: in the grid, 4-adjacent, unique."""
    assert rows > 0 and cols > 0
    if len(cut_edges) > 0:
        e = _cut_edge_array(cut_edges)
        i, j, k, l = e.T
        assert ((i >= 0) & (i < rows) & (j >= 0) & (j < cols) & (k >= 0) & (k < rows) & (l >= 0) & (l < cols)).all()
        assert (np.abs(i - k) + np.abs(j - l) == 1).all()
        assert len(np.unique(e, axis=0)) == len(e)


def _cut_planes(rows, cols, cut_edges):
    """cut_edges in any accepted form -> (cut_down, cut_right, count or None)."""
    if cut_edges is None:
        return None, None, None
    if isinstance(cut_edges, tuple) and len(cut_edges) == 2 and isinstance(cut_edges[0], np.ndarray) and cut_edges[0].dtype == bool:
        cd = np.asarray(cut_edges[0], bool).copy()
        cr = np.asarray(cut_edges[1], bool).copy()
        assert cd.shape == (rows, cols) and cr.shape == (rows, cols)
        cd[-1, :] = False
        cr[:, -1] = False
        return cd, cr, int(cd.sum() + cr.sum())
    assert_valid_cut_edges(rows, cols, cut_edges)
    cd = np.zeros((rows, cols), bool)
    cr = np.zeros((rows, cols), bool)
    if len(cut_edges) > 0:
        i, j, k, l = _cut_edge_array(cut_edges).T
        v = i != k
        cd[np.minimum(i, k)[v], j[v]] = True
        cr[i[~v], np.minimum(j, l)[~v]] = True
    return cd, cr, len(cut_edges)


def _is_array_form(c):
    return isinstance(c, tuple) and len(c) == 2 and isinstance(c[0], np.ndarray)


def _scalar_constraints(cons, rows, cols, i_hi, j_hi):
    """[(i, j, value)] or (ij, values) / (mask, values) -> (linear index, float values)."""
    if _is_array_form(cons):
        a, v = cons
        if a.dtype == bool:
            assert a.shape == (rows, cols)
            idx = np.flatnonzero(a)
            v = np.asarray(v, float)
            vals = v.reshape(rows * cols)[idx] if v.size == rows * cols else v.reshape(-1)
            assert len(vals) == len(idx)
            return idx, vals
        ij = np.asarray(a, np.int64).reshape(-1, 2)
        vals = np.asarray(v, float).reshape(-1)
    else:
        if len(cons) == 0:
            return np.zeros(0, np.int64), np.zeros(0)
        arr = np.asarray([(c[0], c[1], c[2]) for c in cons], dtype=float)
        ij = arr[:, :2].astype(np.int64)
        vals = arr[:, 2]
    assert len(ij) == len(vals)
    # index ranges and duplicates: recovery.py:523-537, 660
    assert ((ij[:, 0] >= 0) & (ij[:, 0] < i_hi)).all()
    assert ((ij[:, 1] >= 0) & (ij[:, 1] < j_hi)).all()
    idx = ij[:, 0] * cols + ij[:, 1]
    # no pixel twice (a mark per pixel instead of sorting the indices)
    seen = np.zeros(rows * cols, bool)
    seen[idx] = True
    assert int(np.count_nonzero(seen)) == len(idx)
    return idx, vals


def _vector_values(values):
    """sequence of 1-D equal-length vectors (or an (n,K) array) -> float (n, K).
    recovery.py:1052-1056, 1313-1317."""
    if isinstance(values, np.ndarray):
        vals = np.asarray(values, float)
    else:
        vs = [np.asarray(v, dtype=float) for v in values]
        assert all(v.ndim == 1 for v in vs)
        assert len(set(len(v) for v in vs)) <= 1
        vals = np.asarray(vs, float).reshape(len(vs), -1)
    assert vals.ndim == 2
    return vals


def _vector_constraints(cons, rows, cols):
    """[(i, j, [v1..vK])] or (ij, values) / (mask, values) -> (linear index, (n,K) values)."""
    if _is_array_form(cons):
        a, v = cons
        v = np.asarray(v, float)
        if a.dtype == bool:
            assert a.shape == (rows, cols)
            idx = np.flatnonzero(a)
            if v.ndim == 2 and v.shape == (rows, cols):
                v = v[:, :, None]
            vals = v.reshape(rows * cols, -1)[idx] if v.shape[:2] == (rows, cols) else v.reshape(len(idx), -1)
            return idx, vals
        ij = np.asarray(a, np.int64).reshape(-1, 2)
        vals = _vector_values(v)
    else:
        if len(cons) == 0:
            return np.zeros(0, np.int64), np.zeros((0, 0))
        ij = np.asarray([(c[0], c[1]) for c in cons], dtype=np.int64).reshape(-1, 2)
        vals = _vector_values([c[2] for c in cons])
    assert len(ij) == len(vals)
    assert ((ij[:, 0] >= 0) & (ij[:, 0] < rows)).all()
    assert ((ij[:, 1] >= 0) & (ij[:, 1] < cols)).all()
    idx = ij[:, 0] * cols + ij[:, 1]
    # no pixel twice (a mark per pixel instead of sorting the indices)
    seen = np.zeros(rows * cols, bool)
    seen[idx] = True
    assert int(np.count_nonzero(seen)) == len(idx)
    return idx, vals


def _index_constraints(cons, rows, cols):
    """[(i, j)] or an (n,2) array or a boolean mask -> linear index (duplicates asserted)."""
    if isinstance(cons, np.ndarray) and cons.dtype == bool:
        assert cons.shape == (rows, cols)
        return np.flatnonzero(cons)
    if len(cons) == 0:
        return np.zeros(0, np.int64)
    ij = np.asarray(cons, dtype=np.int64).reshape(-1, 2)
    assert ((ij[:, 0] >= 0) & (ij[:, 0] < rows)).all()
    assert ((ij[:, 1] >= 0) & (ij[:, 1] < cols)).all()
    idx = ij[:, 0] * cols + ij[:, 1]
    # no pixel twice (a mark per pixel instead of sorting the indices)
    seen = np.zeros(rows * cols, bool)
    seen[idx] = True
    assert int(np.count_nonzero(seen)) == len(idx)
    return idx


def _count(c):
    if _is_array_form(c):
        return int(c[0].sum()) if c[0].dtype == bool else len(c[0])
    if isinstance(c, np.ndarray) and c.dtype == bool:
        return int(c.sum())
    return len(c)


def _plane(rows, cols, idx, dtype=bool, value=True):
    p = np.zeros(rows * cols, dtype)
    p[idx] = value
    return p.reshape(rows, cols)


def _scatter(rows, cols, K, idx, vals):
    a = np.zeros((rows * cols, K))
    if len(idx):
        a[idx] = vals
    return a.reshape(rows, cols, K)


# --------------------------------------------------------------------------------------
# solve entry points
# --------------------------------------------------------------------------------------

def solve_grid_linear(rows, cols, dx_constraints=None, dy_constraints=None, value_constraints=None,
                      bilaplacian=False, iterative=False, cut_edges=None, w_lsq=None,
                      *, precision=None, rtol=None, maxiter=None):
    """recovery.py:838-1009.  Minimises the (bi)laplacian energy subject to soft dx / dy
    constraints (weight w_lsq, default 1e5) and hard value constraints; returns (rows, cols).
    `iterative` selected the solver in the reference; here every solve is the GPU PCG."""
    if dx_constraints is None: dx_constraints = []
    if dy_constraints is None: dy_constraints = []
    if value_constraints is None: value_constraints = []

    if w_lsq is None:
        w_gradients = 1e5
    else:
        assert float(w_lsq) == w_lsq
        w_gradients = float(w_lsq)

    cd, cr, ncut = _cut_planes(rows, cols, cut_edges)
    print('solve_grid_linear( rows = %s, cols = %s, |constraints| = %s, iterative = %s, bilaplacian = %s, |cut_edges| = %s, w_lsq = %s )' % (
        rows, cols,
        _count(dx_constraints) + _count(dy_constraints) + _count(value_constraints),
        iterative, bilaplacian, ncut, w_gradients))

    tic('build energy:')
    assert rows > 0 and cols > 0
    dx_idx, dx_v = _scalar_constraints(dx_constraints, rows, cols, rows - 1, cols)
    dy_idx, dy_v = _scalar_constraints(dy_constraints, rows, cols, rows, cols - 1)
    vc_idx, vc_v = _scalar_constraints(value_constraints, rows, cols, rows, cols)
    toc()

    tic('add constraints:')
    N = rows * cols
    hard = _plane(rows, cols, vc_idx)
    soft = None
    have_grad = len(dx_idx) + len(dy_idx) > 0
    if have_grad and len(vc_idx) == 0:
        # the default SOFT pin (0,0,0.) of weight w_gradients (recovery.py:920)
        soft = np.zeros((rows, cols), bool)
        soft[0, 0] = True
    flags = encode_flags(rows, cols, hard=hard, soft=soft, cut_down=cd, cut_right=cr,
                         pen_down=_plane(rows, cols, dx_idx) if len(dx_idx) else None,
                         pen_right=_plane(rows, cols, dy_idx) if len(dy_idx) else None)
    toc()

    tic('solve system of equations on the GPU:')
    s = GridSolver(rows, cols, 1, bilaplacian, precision)
    s.set_pattern(flags, None, w_gradients, w_gradients)
    result = s.solve(hard_vals=_scatter(rows, cols, 1, vc_idx, vc_v[:, None]) if len(vc_idx) else None,
                     d_down=_scatter(rows, cols, 1, dx_idx, dx_v[:, None]) if len(dx_idx) else None,
                     d_right=_scatter(rows, cols, 1, dy_idx, dy_v[:, None]) if len(dy_idx) else None,
                     rtol=rtol, maxiter=maxiter)
    s.close()
    toc()
    return result.reshape((rows, cols))


def solve_grid_linear_simple(rows, cols, value_constraints, bilaplacian=False,
                             *, precision=None, rtol=None, maxiter=None):
    """recovery.py:1011-1088: hard K-vector value constraints only; returns (rows, cols, K)."""
    print('solve_grid_linear_simple( rows = %s, cols = %s, |constraints| = %s, bilaplacian = %s )' % (
        rows, cols, _count(value_constraints), bilaplacian))

    tic('build energy:')
    assert rows > 0 and cols > 0
    toc()
    tic('add constraints:')
    ## We must have at least one value constraint or our system is under-constrained.
    assert _count(value_constraints) > 0
    idx, vals = _vector_constraints(value_constraints, rows, cols)
    K = vals.shape[1]
    flags = encode_flags(rows, cols, hard=_plane(rows, cols, idx))
    toc()

    tic('solve system of equations on the GPU:')
    result = _solve_channels(rows, cols, K, bilaplacian, precision, flags, None, 0.0, 0.0,
                             dict(hard_vals=_scatter(rows, cols, K, idx, vals)), rtol, maxiter)
    toc()
    return result


def _solve_channels(rows, cols, K, bilaplacian, precision, flags, soft_weights, soft_scalar, w_g, value_planes,
                    rtol, maxiter, solvers=None):
    """Solves K channels in chunks of at most 16 (HG_MAXK), reusing one solver per chunk width."""
    out = np.empty((rows, cols, K))
    MAXK = 16
    for k0 in range(0, K, MAXK):
        k1 = min(K, k0 + MAXK)
        key = k1 - k0
        s = solvers.get(key) if solvers is not None else None
        if s is None:
            s = GridSolver(rows, cols, key, bilaplacian, precision)
            s.set_pattern(flags, soft_weights, soft_scalar, w_g)
            if solvers is not None:
                solvers[key] = s
        chunk = {n: (None if a is None else a[:, :, k0:k1]) for n, a in value_planes.items()}
        out[:, :, k0:k1] = s.solve(rtol=rtol, maxiter=maxiter, **chunk)
        if solvers is None:
            s.close()
    return out


def _simple23_pattern(rows, cols, hard_idx, soft_idx, w_lsq, cut_edges):
    ## Least squares weight should be non-negative (recovery.py:1142, 1262)
    w = np.asarray(w_lsq, dtype=float)
    assert (w >= 0.0).all()
    cd, cr, _ = _cut_planes(rows, cols, cut_edges)
    soft_weights = None
    soft_scalar = 0.0
    if len(soft_idx) > 0:
        if w.ndim > 0:
            assert len(w) == len(soft_idx)
            soft_weights = _plane(rows, cols, soft_idx, float, w)
        else:
            soft_scalar = float(w)
    flags = encode_flags(rows, cols, hard=_plane(rows, cols, hard_idx), soft=_plane(rows, cols, soft_idx),
                         cut_down=cd, cut_right=cr)
    return flags, soft_weights, soft_scalar


def _w_repr(w_lsq):
    return w_lsq if not hasattr(w_lsq, '__iter__') else '%s ...' % w_lsq[:3]


def solve_grid_linear_simple2(rows, cols, value_constraints_hard=[], value_constraints_soft=[], w_lsq=1., bilaplacian=False,
                              *, precision=None, rtol=None, maxiter=None):
    """recovery.py:1090-1200: hard and soft K-vector value constraints, no cuts."""
    print('solve_grid_linear_simple2( rows = %s, cols = %s, |hard constraints| = %s, |soft constraints| = %s, w_lsq = %s, bilaplacian = %s )' % (
        rows, cols, _count(value_constraints_hard), _count(value_constraints_soft), _w_repr(w_lsq), bilaplacian))
    assert _count(value_constraints_hard) + _count(value_constraints_soft) > 0
    hidx, hvals = _vector_constraints(value_constraints_hard, rows, cols)
    sidx, svals = _vector_constraints(value_constraints_soft, rows, cols)
    dims = set(v.shape[1] for v, i in ((hvals, hidx), (svals, sidx)) if len(i))
    assert len(dims) == 1
    K = dims.pop()
    flags, sw, ss = _simple23_pattern(rows, cols, hidx, sidx, w_lsq, None)
    return _solve_channels(rows, cols, K, bilaplacian, precision, flags, sw, ss, 0.0,
                           dict(hard_vals=_scatter(rows, cols, K, hidx, hvals) if len(hidx) else None,
                                soft_vals=_scatter(rows, cols, K, sidx, svals) if len(sidx) else None), rtol, maxiter)


def solve_grid_linear_simple3_solver(rows, cols, value_constraints_hard=[], value_constraints_soft=[], w_lsq=1.,
                                     cut_edges=None, bilaplacian=False, *, precision=None, rtol=None, maxiter=None):
    """recovery.py:1202-1336.  Fixes the constraint pattern, uploads it once, and returns
    solve(hard_values, soft_values) -> (rows, cols, K), the GPU analogue of the reference's
    factor-once / solve-many closure."""
    print('solve_grid_linear_simple3( rows = %s, cols = %s, |hard constraints| = %s, |soft constraints| = %s, w_lsq = %s, bilaplacian = %s )' % (
        rows, cols, _count(value_constraints_hard), _count(value_constraints_soft), _w_repr(w_lsq), bilaplacian))

    tic('build energy:')
    assert rows > 0 and cols > 0
    toc()
    tic('add constraints:')
    ## We must have at least one value constraint or our system is under-constrained.
    assert _count(value_constraints_hard) + _count(value_constraints_soft) > 0
    hidx = _index_constraints(value_constraints_hard, rows, cols)
    sidx = _index_constraints(value_constraints_soft, rows, cols)
    flags, sw, ss = _simple23_pattern(rows, cols, hidx, sidx, w_lsq, cut_edges)
    toc()

    tic('upload pattern to the GPU:')
    solvers = {}
    # validate the domain now (the reference asserts while building L, before solve() exists)
    probe = GridSolver(rows, cols, 1, bilaplacian, precision)
    probe.set_pattern(flags, sw, ss, 0.0)
    solvers[1] = probe
    toc()

    def solve(hard_values=[], soft_values=[]):
        assert len(hidx) == len(hard_values)
        assert len(sidx) == len(soft_values)
        hv = _vector_values(hard_values) if len(hidx) else np.zeros((0, 0))
        sv = _vector_values(soft_values) if len(sidx) else np.zeros((0, 0))
        dims = set(v.shape[1] for v, i in ((hv, hidx), (sv, sidx)) if len(i))
        assert len(dims) == 1
        K = dims.pop()
        if K <= 16:
            # the values go to the device as lists and are scattered there (hg_upload_values_sparse): no dense planes on the host
            s = solvers.get(K)
            if s is None:
                s = GridSolver(rows, cols, K, bilaplacian, precision)
                s.set_pattern(flags, sw, ss, 0.0)
                solvers[K] = s
            return s.solve_sparse(hidx, hv if len(hidx) else None, sidx, sv if len(sidx) else None, rtol=rtol, maxiter=maxiter)
        return _solve_channels(rows, cols, K, bilaplacian, precision, flags, sw, ss, 0.0,
                               dict(hard_vals=_scatter(rows, cols, K, hidx, hv) if len(hidx) else None,
                                    soft_vals=_scatter(rows, cols, K, sidx, sv) if len(sidx) else None),
                               rtol, maxiter, solvers)

    return solve


def solve_grid_linear_simple3(rows, cols, value_constraints_hard=[], value_constraints_soft=[], w_lsq=1.,
                              cut_edges=None, bilaplacian=False, *, precision=None, rtol=None, maxiter=None):
    """recovery.py:1338-1344."""
    if _is_array_form(value_constraints_hard) or _is_array_form(value_constraints_soft):
        h_ij, h_v = value_constraints_hard if _is_array_form(value_constraints_hard) else (np.zeros((0, 2), np.int64), [])
        s_ij, s_v = value_constraints_soft if _is_array_form(value_constraints_soft) else (np.zeros((0, 2), np.int64), [])
    else:
        h_ij = [(i, j) for i, j, val in value_constraints_hard]
        h_v = [val for i, j, val in value_constraints_hard]
        s_ij = [(i, j) for i, j, val in value_constraints_soft]
        s_v = [val for i, j, val in value_constraints_soft]
    solve = solve_grid_linear_simple3_solver(rows, cols, h_ij, s_ij, w_lsq=w_lsq, cut_edges=cut_edges,
                                             bilaplacian=bilaplacian, precision=precision, rtol=rtol, maxiter=maxiter)
    return solve(hard_values=h_v, soft_values=s_v)


def smooth_bumps(grid, bump_locations, bump_radius, **solve_grid_linear_kwargs):
    """recovery.py:1346-1418: re-solves the L1 ball of radius `bump_radius` around each bump
    with every other pixel held at its current value."""
    if 'w_lsq' in solve_grid_linear_kwargs:
        print("WARNING: parameter 'w_lsq' has no effect smooth_bumps()")

    grid = np.asarray(grid)
    assert len(grid.shape) == 2
    assert grid.shape[0] > 0
    assert grid.shape[1] > 0

    bump_locations = np.asarray(bump_locations, dtype=int)
    assert len(bump_locations.shape) == 2 and bump_locations.shape[1] == 2
    assert bump_locations[:, 0].min() >= 0
    assert bump_locations[:, 1].min() >= 0
    assert bump_locations[:, 0].max() < grid.shape[0]
    assert bump_locations[:, 1].max() < grid.shape[1]

    assert bump_radius >= 0
    assert int(bump_radius) == bump_radius
    bump_radius = int(bump_radius)

    mask = np.zeros(grid.shape, dtype=bool)
    mask[tuple(bump_locations.T)] = True
    mask = grow(mask, bump_radius)
    # everything outside the mask is a hard constraint at its current value; passed in array
    # form (mask, values) instead of ~rows*cols tuples (recovery.py:1409)
    keep = np.logical_not(mask)
    result = solve_grid_linear(grid.shape[0], grid.shape[1],
                               value_constraints=(keep, np.asarray(grid, float)), **solve_grid_linear_kwargs)
    assert list(result.shape) == list(grid.shape)
    return result


def smooth_bumps2(grid, bump_locations, bump_radius, smooth_bumps_iterations, smooth_func=None, **solve_grid_linear_kwargs):
    """recovery.py:1420-1530: iteratively replaces the gradient at each bump by the median /
    average gradient of its neighbourhood and re-solves the whole grid."""
    grid = np.asarray(grid)
    assert len(grid.shape) == 2
    assert grid.shape[0] > 0
    assert grid.shape[1] > 0

    bump_locations = np.asarray(bump_locations, dtype=int)
    assert len(bump_locations.shape) == 2 and bump_locations.shape[1] == 2
    assert bump_locations[:, 0].min() >= 0
    assert bump_locations[:, 1].min() >= 0
    ## bumps are derivative constraints: not in the last row or column (recovery.py:1465-1466)
    assert bump_locations[:, 0].max() < grid.shape[0] - 1
    assert bump_locations[:, 1].max() < grid.shape[1] - 1

    assert bump_radius >= 0
    assert int(bump_radius) == bump_radius
    bump_radius = int(bump_radius)

    assert smooth_bumps_iterations >= 0
    assert int(smooth_bumps_iterations) == smooth_bumps_iterations
    smooth_bumps_iterations = int(smooth_bumps_iterations)

    if smooth_func is None: smooth_func = 'median'
    assert smooth_func in ('median', 'average')

    if smooth_bumps_iterations == 0:
        return np.array(grid)

    bump_neighbors = []
    for loc in bump_locations:
        mask = np.zeros(grid[:-1, :-1].shape, dtype=bool)
        mask[tuple(loc)] = True
        bump_neighbors.append(np.where(grow(mask, bump_radius)))

    smooth_func = {'median': np.median, 'average': np.average}[smooth_func]
    for _ in range(smooth_bumps_iterations):
        ddown = grid[1:, :] - grid[:-1, :]
        dright = grid[:, 1:] - grid[:, :-1]
        dx_constraints = [(loc[0], loc[1], smooth_func(ddown[ix])) for loc, ix in zip(bump_locations, bump_neighbors)]
        dy_constraints = [(loc[0], loc[1], smooth_func(dright[ix])) for loc, ix in zip(bump_locations, bump_neighbors)]
        grid = solve_grid_linear(grid.shape[0], grid.shape[1], dx_constraints=dx_constraints,
                                 dy_constraints=dy_constraints, **solve_grid_linear_kwargs)
    return grid


def convert_grid_normal_constraints_to_dx_and_dy_constraints(rows, cols, normal_constraints):
    """recovery.py:1532-1620: (i, j, 3-D normal) -> dx = -n.x/n.z on edge (i,j)-(i+1,j) and
    dy = -n.y/n.z on edge (i,j)-(i,j+1).  Constraints in the last row / column are shifted
    globally by one if the first row / column is empty, else locally if the slot is free."""
    assert rows > 0 and cols > 0
    if _is_array_form(normal_constraints):
        return _convert_normals_arrays(rows, cols, *normal_constraints)
    all_is = set(i for i, j, n in normal_constraints)
    all_js = set(j for i, j, n in normal_constraints)
    if (rows - 1) in all_is and 0 not in all_is:
        normal_constraints = [(i - 1, j, n) for i, j, n in normal_constraints]
    if (cols - 1) in all_js and 0 not in all_js:
        normal_constraints = [(i, j - 1, n) for i, j, n in normal_constraints]

    ij2normal = {}
    unsafe = []
    for i, j, n in normal_constraints:
        if i < rows - 1 and j < cols - 1:
            ij2normal[(i, j)] = n
        else:
            unsafe.append((i, j, n))
    for i, j, n in unsafe:
        tgt = (i - (i == rows - 1), j - (j == cols - 1))
        if tgt not in ij2normal:
            ij2normal[tgt] = n

    dx_constraints = []
    dy_constraints = []
    for (i, j), n in ij2normal.items():
        if n[2] < 1e-5:
            print('Silhouette normal!  TODO: Make it a height map discontinuity.  Skipping for now...')
            continue
        dx_constraints.append((i, j, -n[0] / n[2]))
        dy_constraints.append((i, j, -n[1] / n[2]))
    return dx_constraints, dy_constraints


def _convert_normals_arrays(rows, cols, ij, normals):
    """Array form of convert_grid_normal_constraints_to_dx_and_dy_constraints (extension): ij (n,2) integer
    pixels, normals (n,3); returns ((ij, dx values), (ij, dy values)) in the array form solve_grid_linear accepts.
    Same rules as the tuple form: global shift by one if the last row / column is used and the first is
    not; a constraint in the last row / column moves one pixel inwards unless that slot is taken."""
    ij = np.asarray(ij, np.int64).reshape(-1, 2).copy()
    n = np.asarray(normals, float).reshape(-1, 3)
    assert len(ij) == len(n)
    if len(ij) and (ij[:, 0] == rows - 1).any() and not (ij[:, 0] == 0).any():
        ij[:, 0] -= 1
    if len(ij) and (ij[:, 1] == cols - 1).any() and not (ij[:, 1] == 0).any():
        ij[:, 1] -= 1
    safe = (ij[:, 0] < rows - 1) & (ij[:, 1] < cols - 1)
    key = ij[:, 0] * cols + ij[:, 1]
    # later duplicates of a safe pixel overwrite earlier ones (dict semantics)
    ks, first_of_rev = np.unique(key[safe][::-1], return_index=True)
    idx_safe = np.flatnonzero(safe)[::-1][first_of_rev]
    taken = set(ks.tolist())
    extra = []
    for q in np.flatnonzero(~safe):          # few: only the last row / column
        i, j = ij[q]
        t = (i - (i == rows - 1)) * cols + (j - (j == cols - 1))
        if t not in taken:
            taken.add(int(t))
            extra.append((int(t), q))
    keys = np.concatenate([ks, np.asarray([t for t, q in extra], np.int64)]) if extra else ks
    src = np.concatenate([idx_safe, np.asarray([q for t, q in extra], np.int64)]) if extra else idx_safe
    nn = n[src]
    ok = nn[:, 2] >= 1e-5
    if not ok.all():
        print('Silhouette normal!  TODO: Make it a height map discontinuity.  Skipping for now...')
    keys, nn = keys[ok], nn[ok]
    out_ij = np.stack([keys // cols, keys % cols], axis=1)
    return (out_ij, -nn[:, 0] / nn[:, 2]), (out_ij, -nn[:, 1] / nn[:, 2])


def normalize_to_char_img(img):
    """recovery.py:1901-1942: linear map of [min, max] to [0, 255], uint8."""
    img = np.asarray(img)
    if img.dtype.kind != 'f':
        img = np.asarray(img, dtype=float)
    img = img.squeeze()
    assert len(img.shape) in (2, 3)
    lo, hi = img.min(), img.max()
    return np.asarray((255. * ((img - lo) / (hi - lo))).clip(0, 255).round(), dtype=np.uint8)


def test_solve_grid_linear_simpleN(solve_grid_linear_simpleN):
    """The reference's self-test (recovery.py:2020-2077, run by `python3 recovery.py`, README.md:5-6): a 250 x 250 thin-plate
    solve with a cut-out square, hard / soft / per-equation-soft / mixed constraints, the two 1e-7 self-consistency asserts,
    three PNGs written into the working directory."""
    from PIL import Image
    rows = cols = 250
    br0, br1 = 100, 150
    bc0, bc1 = br0, br1
    K = 1
    value_constraints = [
        (0, 0, [br0 * 2.] * K), (rows - 1, 0, [0.] * K), (0, cols - 1, [0.] * K), (rows - 1, cols - 1, [0.] * K),
        (br0, bc0, [br0] * K), (br1 - 1, bc0, [br0] * K), (br0, bc1 - 1, [br0] * K), (br1 - 1, bc1 - 1, [br0] * K)]
    mask = np.zeros((rows, cols), dtype=bool)
    mask[br0:br1, bc0:bc1] = True
    cut_edges = cut_edges_from_mask(mask)

    sol_hard = solve_grid_linear_simpleN(rows, cols, value_constraints_hard=value_constraints, cut_edges=cut_edges, bilaplacian=True)
    sol_soft = solve_grid_linear_simpleN(rows, cols, value_constraints_soft=value_constraints, w_lsq=1e3, cut_edges=cut_edges, bilaplacian=True)
    sol_soft2 = solve_grid_linear_simpleN(rows, cols, value_constraints_soft=value_constraints, w_lsq=[1e3] * len(value_constraints),
                                          cut_edges=cut_edges, bilaplacian=True)
    assert abs(sol_soft2 - sol_soft).max() < 1e-7
    mid = [((br0 + br1) // 2, (bc0 + bc1) // 2, [.5 * br0] * K)]
    sol_mixed = solve_grid_linear_simpleN(rows, cols, value_constraints_hard=value_constraints, value_constraints_soft=mid, w_lsq=1e3,
                                          cut_edges=cut_edges, bilaplacian=True)
    sol_mixed2 = solve_grid_linear_simpleN(rows, cols, value_constraints_hard=value_constraints, value_constraints_soft=mid, w_lsq=[1e3] * 1,
                                           cut_edges=cut_edges, bilaplacian=True)
    assert abs(sol_mixed2 - sol_mixed).max() < 1e-7

    Image.fromarray(normalize_to_char_img(sol_hard)).save('sol_hard.png')
    Image.fromarray(normalize_to_char_img(sol_hard)).save('sol_soft.png')       # (sic: recovery.py:2072 saves sol_hard twice)
    Image.fromarray(normalize_to_char_img(sol_mixed)).save('sol_mixed.png')
    print('[Saved "sol_soft.png".]')
    print('[Saved "sol_hard.png".]')
    print('[Saved "sol_mixed.png".]')
    return sol_hard, sol_soft, sol_mixed


def main():
    """recovery.py:2096-2102: `python3 recovery.py` runs the self-test and exits."""
    import sys
    test_solve_grid_linear_simpleN(solve_grid_linear_simple3)
    sys.exit(0)


if __name__ == '__main__':
    main()
