"""Numpy / scipy prototype for the stiff gradient-constraint problems of solve_grid_linear (default w_lsq = 1e5, BASELINE
config C4: normals on a sparse subset of the pixels -> dx / dy rows of weight 1e5 on top of the bilaplacian).  Design tool,
not product code, not shipped in the package: it measures PCG iteration counts of candidate multigrid preconditioners on
explicitly assembled matrices (the oracle's), before anything is written as a CUDA kernel.

Finding (round 2).  A stiff edge pins the DIFFERENCE of its two pixels, so the error of any reasonable
iterate is constant on every connected cluster of stiff edges.  Bilinear interpolation cannot represent "smooth, but flat on
a few hundred random three-pixel clusters", so the Galerkin coarse operator sees a penalty of w (difference of interpolation
weights)^2 at every cluster and the coarse correction degenerates: 343 / 435 iterations at 128^2 / 256^2, growing.  With the
level-0 interpolation rows AVERAGED over each cluster (P' = C P, C the cluster-averaging projector; Galerkin A_1 = P'^T A P',
every coarser level unchanged) the stiff term drops out of the coarse operators altogether: 19 / 26 iterations at 128^2 /
256^2 with point Gauss-Seidel, 15-17 with exact cluster solves added to the smoother; 26 at 10 % constraint density, 83 at
30 % (clusters of 66 pixels), and worse than plain at 60 % (one giant cluster: the structure would have to be carried through
every level -- an aggregation hierarchy); dense constraints (100 %) need nothing (8 iterations with the plain hierarchy).
What does NOT work: smoothing the interpolation with a few damped Jacobi steps of the stiff part (251-800 iterations: it
has to be flat exactly), flattening only the clusters up to a size cap (141 iterations at 128^2 with 6 % of the clusters
left out, 440 at 256^2), and preconditioning in the collapsed space with the V-cycle of the bilaplacian alone plus cluster
block solves (91 / 113 iterations, still growing).  The price: the level-1 stencil grows to 7 x 7 and beyond (its radius is
2 + the largest cluster's extent), i.e. level 1 has to be stored as a general sparse matrix.

usage: python tools/stiff_proto.py [w]
"""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from scipy.sparse.csgraph import connected_components
sys.path.insert(0, '/root/repo')
from oracle import grid_oracle as go

def c4_problem(n, w=1e5, dens=0.02, seed=0):
    rng = np.random.default_rng(seed)
    ii, jj = np.indices((n, n), dtype=float)
    A = 0.05 * n
    z = A * np.sin(2*np.pi*ii/n) * np.cos(2*np.pi*jj/n)
    sel = rng.random((n-1, n-1)) < dens
    prob = go.GridProblem(n, n, 1, True)
    prob.w_g = w
    prob.pen_down[:-1, :-1] = sel; prob.pen_right[:-1, :-1] = sel
    prob.d_down[:-1, :-1, 0] = np.where(sel, z[1:, :-1] - z[:-1, :-1], 0)
    prob.d_right[:-1, :-1, 0] = np.where(sel, z[:-1, 1:] - z[:-1, :-1], 0)
    prob.hard[0, 0] = True; prob.hard_vals[0, 0, 0] = z[0, 0]
    return prob, z

def bilinear_P(rows, cols):
    """vertex-centred: coarse (I,J) at fine (2I,2J); coarse dims ceil(n/2); returns (N_f x N_c)"""
    def P1(n):
        nc = (n + 1) // 2
        I, J, V = [], [], []
        for i in range(n):
            if i % 2 == 0:
                I.append(i); J.append(i // 2); V.append(1.0)
            else:
                a, b = i // 2, i // 2 + 1
                if b < nc:
                    I += [i, i]; J += [a, b]; V += [0.5, 0.5]
                else:
                    I.append(i); J.append(a); V.append(1.0)
        return sp.coo_matrix((V, (I, J)), shape=(n, nc)).tocsr()
    return sp.kron(P1(rows), P1(cols)).tocsr(), (rows + 1) // 2, (cols + 1) // 2

class MG:
    def __init__(self, A, free_mask, rows, cols, clusters=None, cluster_P=False, cluster_smooth=False, min_n=16):
        # A on free dofs of the fine grid; free_mask (rows*cols) bool
        self.levels = []
        N = rows * cols
        E = sp.identity(N, format='csr')[:, np.flatnonzero(free_mask)]   # embed free -> full
        Afull = (E @ A @ E.T).tocsr()    # zero rows/cols at hard
        self.E = E
        cur = Afull; r, c = rows, cols
        first = True
        while True:
            act = cur.diagonal() != 0
            lev = dict(A=cur, act=act, r=r, c=c)
            idx = np.flatnonzero(act)
            As = cur[idx][:, idx].tocsr()
            lev['idx'] = idx; lev['As'] = As
            lev['L'] = sp.tril(As, format='csr'); lev['U'] = sp.triu(As, format='csr')
            if first and clusters is not None:
                lev['clusters'] = clusters if cluster_smooth else None
            self.levels.append(lev)
            if r * c <= min_n * min_n:
                lev['lu'] = spla.splu(As.tocsc())
                break
            P, rc, cc = bilinear_P(r, c)
            # mask: rows of inactive fine pixels zero
            P = sp.diags(act.astype(float)) @ P
            if first and cluster_P and clusters is not None:
                # average the interpolation rows over each cluster
                lab = clusters['label_full']     # (N,) cluster id or -1
                members = lab >= 0
                ncl = lab.max() + 1
                M = sp.coo_matrix((np.ones(members.sum()), (lab[members], np.flatnonzero(members))), shape=(ncl, N)).tocsr()
                cnt = np.asarray(M.sum(1)).ravel()
                avg = sp.diags(1.0 / cnt) @ M @ P          # (ncl x Nc)
                P = P.tolil()
                Pavg = (M.T @ avg).tocsr()                  # rows for members
                keep = sp.diags((~members).astype(float))
                P = (keep @ P.tocsr() + Pavg).tocsr()
            lev['P'] = P
            cur = (P.T @ cur @ P).tocsr()
            cur.eliminate_zeros()
            r, c = rc, cc
            first = False

    def smooth(self, lev, x, b, forward):
        As, idx = lev['As'], lev['idx']
        xs, bs = x[idx], b[idx]
        cl = lev.get('clusters')
        def gs(xs, fwd):
            if fwd:
                return xs + spla.spsolve_triangular(lev['L'], bs - As @ xs, lower=True)
            return xs + spla.spsolve_triangular(lev['U'], bs - As @ xs, lower=False)
        def cluster_solves(xs):
            if cl is None: return xs
            # exact solve on each cluster (block Jacobi with damping 1 on non-overlapping clusters, residual recomputed once)
            r = bs - As @ xs
            xs = xs.copy()
            for blk, inv in zip(cl['blocks'], cl['invs']):
                xs[blk] += inv @ r[blk]
            return xs
        if forward:
            xs = gs(xs, True); xs = cluster_solves(xs)
        else:
            xs = cluster_solves(xs); xs = gs(xs, False)
        x = x.copy(); x[idx] = xs
        return x

    def vcycle(self, l, b):
        lev = self.levels[l]
        x = np.zeros_like(b)
        if 'lu' in lev:
            x[lev['idx']] = lev['lu'].solve(b[lev['idx']])
            return x
        x = self.smooth(lev, x, b, True)
        r = b - lev['A'] @ x
        ec = self.vcycle(l + 1, lev['P'].T @ r)
        x = x + lev['P'] @ ec
        x = self.smooth(lev, x, b, False)
        return x

    def precond(self, rf):
        return self.E.T @ self.vcycle(0, self.E @ rf)

def pcg(A, b, M, tol=1e-8, maxit=3000):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r @ z; bn = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        q = A @ p; a = rz / (p @ q); x += a * p; r -= a * q
        if np.linalg.norm(r) <= tol * bn: return x, it
        z = M(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit

def clusters_of(prob, free_idx):
    rows, cols = prob.rows, prob.cols
    N = rows * cols
    idx = np.arange(N).reshape(rows, cols)
    pd = prob.pen_down.copy(); pd[-1] = False
    pr = prob.pen_right.copy(); pr[:, -1] = False
    I = np.concatenate([idx[:-1][pd[:-1]], idx[:, :-1][pr[:, :-1]]]); J = np.concatenate([idx[1:][pd[:-1]], idx[:, 1:][pr[:, :-1]]])
    g = sp.coo_matrix((np.ones(len(I)), (I, J)), shape=(N, N))
    ncomp, lab = connected_components(g, directed=False)
    size = np.bincount(lab, minlength=ncomp)
    big = size[lab] > 1
    hard = prob.hard.ravel()
    label_full = np.where(big & ~hard, lab, -1)
    uniq, inv = np.unique(label_full[label_full >= 0], return_inverse=True)
    lf = np.full(N, -1); lf[label_full >= 0] = inv
    return lf, len(uniq), size

if __name__ == '__main__':
    w = float(sys.argv[1]) if len(sys.argv) > 1 else 1e5
    for n in (32, 64, 128, 192):
        prob, z = c4_problem(n, w)
        A, b, free = go.reduced_system(prob, check=False)
        A = A.tocsr(); b = b[:, 0]
        free_mask = np.zeros(n * n, bool); free_mask[free] = True
        lf, ncl, size = clusters_of(prob, free)
        # blocks in free-dof numbering (level 0 idx == free since diag != 0 everywhere free)
        pos = np.full(n * n, -1); pos[free] = np.arange(len(free))
        order = np.argsort(lf[free], kind='stable'); labs = lf[free][order]
        blocks = []
        start = np.searchsorted(labs, np.arange(ncl)); end = np.searchsorted(labs, np.arange(ncl), side='right')
        for s, e in zip(start, end): blocks.append(order[s:e])
        Ad = A.tocsr()
        invs = [np.linalg.inv(Ad[blk][:, blk].toarray()) for blk in blocks]
        cl = dict(label_full=lf, blocks=blocks, invs=invs)
        xe = spla.splu(A.tocsc()).solve(b)
        out = []
        for name, kw in (('plain', {}), ('clsmooth', dict(clusters=cl, cluster_smooth=True)), ('clP', dict(clusters=cl, cluster_P=True)),
                         ('clP+clsmooth', dict(clusters=cl, cluster_P=True, cluster_smooth=True))):
            t = time.time()
            mg = MG(A, free_mask, n, n, **kw)
            x, it = pcg(A, b, mg.precond, 1e-8, 1500)
            out.append('%s %d (err %.1e, %.0fs)' % (name, it, np.abs(x - xe).max() / np.abs(xe).max(), time.time() - t))
        print('n=%d w=%g clusters=%d maxsize=%d | ' % (n, w, ncl, size.max()) + ' | '.join(out), flush=True)
