"""Numpy prototype of the multigrid preconditioner (design tool, not product code, not shipped
in the package): measures PCG iteration counts for candidate coarse-operator / transfer
choices on the mask families of SURVEY 8(d) before they are written as CUDA kernels."""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class Level:
    def __init__(self, hard, wd, wr, s):
        self.hard, self.wd, self.wr, self.s = hard, wd, wr, s
        self.rows, self.cols = hard.shape
        self.build()

    def build(self):
        r, c = self.rows, self.cols
        N = r * c
        idx = np.arange(N).reshape(r, c)
        free = ~self.hard
        I, J, V = [], [], []
        diag = self.s.copy()
        # down edges
        w = self.wd[:-1]
        diag[:-1] += w
        diag[1:] += w
        m = (w != 0) & free[:-1] & free[1:]
        I += [idx[:-1][m], idx[1:][m]]; J += [idx[1:][m], idx[:-1][m]]; V += [-w[m], -w[m]]
        w = self.wr[:, :-1]
        diag[:, :-1] += w
        diag[:, 1:] += w
        m = (w != 0) & free[:, :-1] & free[:, 1:]
        I += [idx[:, :-1][m], idx[:, 1:][m]]; J += [idx[:, 1:][m], idx[:, :-1][m]]; V += [-w[m], -w[m]]
        self.diag = np.where(free, diag, 1.0)
        self.off = sp.coo_matrix((np.concatenate(V), (np.concatenate(I), np.concatenate(J))), shape=(N, N)).tocsr()
        self.free = free
        ii, jj = np.indices((r, c))
        self.red = (((ii + jj) & 1) == 0) & free
        self.black = (((ii + jj) & 1) == 1) & free
        self.dinv = np.where((self.diag > 0) & free, 1.0 / np.where(self.diag > 0, self.diag, 1), 0.0)

    def A(self, x):
        y = self.diag * x + (self.off @ x.ravel()).reshape(x.shape)
        return np.where(self.free, y, 0.0)

    def gs(self, x, b, order):
        for col in order:
            m = self.red if col == 0 else self.black
            t = (b - (self.off @ x.ravel()).reshape(x.shape)) * self.dinv
            x = np.where(m, t, x)
        return x

    def matrix(self):
        N = self.rows * self.cols
        f = self.free.ravel()
        return (sp.diags(self.diag.ravel()) + self.off)[f][:, f].tocsc()


def fine_level(rows, cols, hard, cut_down=None, cut_right=None, soft=None):
    wd = np.full((rows, cols), 0.25); wr = np.full((rows, cols), 0.25)
    if cols > 1: wd[:, 0] = wd[:, -1] = 0.125
    if rows > 1: wr[0, :] = wr[-1, :] = 0.125
    wd[-1, :] = 0; wr[:, -1] = 0
    if cut_down is not None: wd[cut_down] = 0
    if cut_right is not None: wr[cut_right] = 0
    s = np.zeros((rows, cols)) if soft is None else soft.copy()
    return Level(hard, wd, wr, s)


def pad_to(a, r, c, val=0):
    out = np.full((r, c), val, a.dtype)
    out[:a.shape[0], :a.shape[1]] = a
    return out


def coarsen(L, variant):
    """vertex-centred: coarse (I,J) <-> fine (2I,2J); coarse dims ceil(n/2)."""
    r, c = L.rows, L.cols
    cr, cc = (r + 1) // 2, (c + 1) // 2
    R, Cc = 2 * cr + 1, 2 * cc + 1   # padded fine dims so that indices 2I+2 exist
    hard = pad_to(L.hard, R, Cc, True)
    wd = pad_to(L.wd, R, Cc); wr = pad_to(L.wr, R, Cc); s = pad_to(L.s, R, Cc)
    # zero weights into hard pixels but remember the leak to ground
    free = ~hard
    # leak g_p = s_p + sum of conductances to hard neighbours (for free p)
    g = s.copy()
    g[:-1] += wd[:-1] * hard[1:]
    g[1:] += wd[:-1] * hard[:-1]
    g[:, :-1] += wr[:, :-1] * hard[:, 1:]
    g[:, 1:] += wr[:, :-1] * hard[:, :-1]
    # conductances only between free pixels
    wdf = wd.copy(); wdf[:-1] *= free[:-1] & free[1:]
    wrf = wr.copy(); wrf[:, :-1] *= free[:, :-1] & free[:, 1:]
    chard = hard[0:2 * cr:2, 0:2 * cc:2].copy()

    def series(w1, w2, gm):
        den = w1 + w2 + gm
        den = np.where(den > 0, den, 1)
        return w1 * w2 / den, w1 * gm / den, w2 * gm / den   # through, leak at first end, leak at second end

    if variant == 'redisc':
        # plain rediscretisation: constant weights, cut if either fine edge on the path is cut, hard by injection
        cwd = np.full((cr, cc), 0.25); cwr = np.full((cr, cc), 0.25)
        if cc > 1: cwd[:, 0] = cwd[:, -1] = 0.125
        if cr > 1: cwr[0, :] = cwr[-1, :] = 0.125
        cwd[-1, :] = 0; cwr[:, -1] = 0
        w1 = wd[0:2 * cr:2, 0:2 * cc:2]; w2 = wd[1:2 * cr:2, 0:2 * cc:2]
        cwd[(w1 == 0) | (w2 == 0)] = 0
        w1 = wr[0:2 * cr:2, 0:2 * cc:2]; w2 = wr[0:2 * cr:2, 1:2 * cc:2]
        cwr[(w1 == 0) | (w2 == 0)] = 0
        cs = restrict_fw(np.where(free, g, 0.0)[:r, :c], L, None, cr, cc, masked=False)
        cs = np.where(chard, 0, cs)
        return Level(chard, cwd, cwr, cs)

    # 'network': star-mesh elimination of the mid-edge nodes, parallel half-paths through the
    # neighbouring fine lines, leaks to ground accumulate on the coarse diagonal
    cwd = np.zeros((cr, cc)); cwr = np.zeros((cr, cc)); cs = np.zeros((cr, cc))
    # vertical paths along fine column j: nodes (2I,j) -m(2I+1,j)- (2I+2,j)
    def vpath(j_sl):
        w1 = wdf[0:2 * cr:2, j_sl]; w2 = wdf[1:2 * cr:2, j_sl]; gm = g[1:2 * cr:2, j_sl] * free[1:2 * cr:2, j_sl]
        # mid node hard -> w1, w2 already zeroed by wdf, its effect is in g of the end nodes
        return series(w1, w2, gm)
    def hpath(i_sl):
        w1 = wrf[i_sl, 0:2 * cc:2]; w2 = wrf[i_sl, 1:2 * cc:2]; gm = g[i_sl, 1:2 * cc:2] * free[i_sl, 1:2 * cc:2]
        return series(w1, w2, gm)

    # centre lines
    t, l1, l2 = vpath(slice(0, 2 * cc, 2))
    cwd += t; cs += l1; cs[1:] += l2[:-1]
    t, l1, l2 = hpath(slice(0, 2 * cr, 2))
    cwr += t; cs += l1; cs[:, 1:] += l2[:, :-1]
    # neighbouring lines (odd fine columns / rows) contribute half to each adjacent coarse line
    # vertical paths in odd columns j = 2J+1: ends are (2I, 2J+1) which are mid-edge nodes of horizontal coarse edges;
    # heuristic: give half of the through-conductance to coarse columns J and J+1
    t, l1, l2 = vpath(slice(1, 2 * cc, 2))          # shape (cr, cc) for odd columns 1,3,..
    cwd += 0.5 * t
    cwd[:, 1:] += 0.5 * t[:, :-1]
    t, l1, l2 = hpath(slice(1, 2 * cr, 2))
    cwr += 0.5 * t
    cwr[1:, :] += 0.5 * t[:-1, :]
    # own leak of the coarse nodes + leaks of the other fine nodes, distributed by full weighting
    gg = np.where(free, g, 0.0)
    cs = restrict_fw(gg[:r, :c], L, None, cr, cc, masked=False)
    cwd[-1, :] = 0; cwr[:, -1] = 0
    cs = np.where(chard, 0, cs)
    return Level(chard, cwd, cwr, cs)


def restrict_fw(rf, L, Lc, cr, cc, masked=True):
    """R = P^T with bilinear P (weights 1, 1/2, 1/4); for even fine sizes the last fine row / column
    hangs beyond the last coarse one and P extrapolates by a constant, so its weight folds back."""
    r, c = rf.shape
    er, ec_ = cr + 1, cc + 1                      # extended coarse grid
    R, Cc = 2 * er + 1, 2 * ec_ + 1
    f = np.zeros((R + 1, Cc + 1))
    f[1:r + 1, 1:c + 1] = rf
    ctr = f[1:2 * er:2, 1:2 * ec_:2]
    up = f[0:2 * er:2, 1:2 * ec_:2]; dn = f[2:2 * er + 1:2, 1:2 * ec_:2]
    lf = f[1:2 * er:2, 0:2 * ec_:2]; rt = f[1:2 * er:2, 2:2 * ec_ + 1:2]
    ul = f[0:2 * er:2, 0:2 * ec_:2]; ur = f[0:2 * er:2, 2:2 * ec_ + 1:2]
    dl = f[2:2 * er + 1:2, 0:2 * ec_:2]; dr = f[2:2 * er + 1:2, 2:2 * ec_ + 1:2]
    out = ctr + 0.5 * (up + dn + lf + rt) + 0.25 * (ul + ur + dl + dr)
    out[cr - 1, :] += out[cr, :]
    out[:, cc - 1] += out[:, cc]
    return out[:cr, :cc].copy()


def prolong_bl(ec, r, c):
    cr, cc = ec.shape
    e = np.zeros((2 * cr + 1, 2 * cc + 1))
    ecp = np.pad(ec, ((0, 1), (0, 1)), mode='edge')
    e[0:2 * cr:2, 0:2 * cc:2] = ec
    e[1:2 * cr:2, 0:2 * cc:2] = 0.5 * (ecp[:cr, :cc] + ecp[1:, :cc])
    e[0:2 * cr:2, 1:2 * cc:2] = 0.5 * (ecp[:cr, :cc] + ecp[:cr, 1:])
    e[1:2 * cr:2, 1:2 * cc:2] = 0.25 * (ecp[:cr, :cc] + ecp[1:, :cc] + ecp[:cr, 1:] + ecp[1:, 1:])
    return e[:r, :c]


class MG:
    def __init__(self, L0, variant='network', min_size=8, nu=1, coarse_sweeps=30, galerkin=False):
        self.levels = [L0]
        self.nu = nu
        self.coarse_sweeps = coarse_sweeps
        while min(self.levels[-1].rows, self.levels[-1].cols) > min_size:
            self.levels.append(coarsen(self.levels[-1], variant))
        self.lu = None
        Lc = self.levels[-1]
        if Lc.free.any():
            self.lu = spla.splu(Lc.matrix())

    def vcycle(self, l, b):
        L = self.levels[l]
        if l == len(self.levels) - 1:
            x = np.zeros_like(b)
            if self.lu is not None:
                x[L.free] = self.lu.solve(b[L.free])
            return x
        x = np.zeros_like(b)
        for _ in range(self.nu):
            x = L.gs(x, b, (0, 1))
        res = np.where(L.free, b - L.A(x), 0.0)
        Lc = self.levels[l + 1]
        rc = restrict_fw(res, L, Lc, Lc.rows, Lc.cols)
        rc = np.where(Lc.free, rc, 0.0)
        ec = self.vcycle(l + 1, rc)
        x = x + np.where(L.free, prolong_bl(ec, L.rows, L.cols), 0.0)
        for _ in range(self.nu):
            x = L.gs(x, b, (1, 0))
        return x


def pcg(L, b, M, tol=1e-6, maxit=500):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = (r * z).sum()
    bb = np.sqrt((b * b).sum())
    hist = []
    for it in range(1, maxit + 1):
        q = L.A(p)
        a = rz / (p * q).sum()
        x += a * p
        r -= a * q
        rn = np.sqrt((r * r).sum()) / bb
        hist.append(rn)
        if rn < tol:
            break
        z = M(r)
        rz2 = (r * z).sum()
        p = z + (rz2 / rz) * p
        rz = rz2
    return x, it, hist


def make_case(name, n, rng):
    hard = np.zeros((n, n), bool); cd = cr = soft = None
    if name == 'tiles64':
        hard = np.kron(rng.random((n // 64, n // 64)) < 0.5, np.ones((64, 64), bool)); hard[0, 0] = True
    elif name == 'tiles16':
        hard = np.kron(rng.random((n // 16, n // 16)) < 0.5, np.ones((16, 16), bool)); hard[0, 0] = True
    elif name == 'iid':
        hard = rng.random((n, n)) < 0.5
    elif name == 'disk':
        ii, jj = np.indices((n, n))
        hard = (ii - n / 2) ** 2 + (jj - n / 2) ** 2 > n * n / (2 * np.pi)
    elif name == 'pin':
        hard[0, 0] = True
    elif name == 'c2':
        hard = rng.random((n, n)) < 0.02
        softm = (rng.random((n, n)) < 0.05) & ~hard
        soft = np.where(softm, 10.0, 0.0)
        m = np.kron(rng.random((n // 16 + 1, n // 16 + 1)) < 0.3, np.ones((16, 16), bool))[:n, :n]
        cd = np.zeros((n, n), bool); cr = np.zeros((n, n), bool)
        cd[:-1] = m[:-1] != m[1:]; cr[:, :-1] = m[:, :-1] != m[:, 1:]
    elif name == 'c2w':
        hard = rng.random((n, n)) < 0.02
        softm = (rng.random((n, n)) < 0.05) & ~hard
        soft = np.where(softm, rng.uniform(0.1, 100, (n, n)), 0.0)
    elif name == 'sparse':
        hard = rng.random((n, n)) < 0.001
    return fine_level(n, n, hard, cd, cr, soft)


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    cases = sys.argv[2].split(',') if len(sys.argv) > 2 else ['tiles64', 'iid', 'disk', 'pin', 'c2', 'c2w', 'sparse']
    variants = sys.argv[3].split(',') if len(sys.argv) > 3 else ['redisc', 'network']
    for name in cases:
        rng = np.random.default_rng(0)
        L = make_case(name, n, rng)
        # rhs: random hard values -> b = -A_fc x_c, plus random rhs where soft
        xc = np.where(L.hard, rng.random((n, n)), 0.0)
        full = fine_level(n, n, np.zeros((n, n), bool))
        full.wd, full.wr, full.s = L.wd, L.wr, L.s
        full.build()
        b = np.where(L.free, -full.A(xc) + L.s * rng.random((n, n)), 0.0)
        if not b.any():
            b = np.where(L.free, rng.normal(size=(n, n)), 0.0)
        for v in variants:
            t = time.time()
            mg = MG(L, v)
            x, it, hist = pcg(L, b, lambda r: mg.vcycle(0, r), 1e-6)
            x2, it2, _ = pcg(L, b, lambda r: mg.vcycle(0, r), 1e-10)
            print('%-8s n=%d %-8s levels=%d  its(1e-6)=%3d its(1e-10)=%3d  rate=%.3f  %.1fs' % (
                name, n, v, len(mg.levels), it, it2, (hist[-1] / hist[0]) ** (1.0 / max(len(hist) - 1, 1)), time.time() - t))
        x, it, hist = pcg(L, b, lambda r: r * L.dinv, 1e-6, 20000)
        print('%-8s n=%d jacobi   its(1e-6)=%d' % (name, n, it))


# ---------------------------------------------------------------------------------------------
# Galerkin variant: explicit sparse operators, masked bilinear P, 4-colour GS on coarse levels
# ---------------------------------------------------------------------------------------------
def bilinear_P(r, c, free_f, chard):
    cr, cc = (r + 1) // 2, (c + 1) // 2
    rows_, cols_, vals_ = [], [], []
    ii, jj = np.indices((r, c))
    fidx = (ii * c + jj)
    for di in (0, 1):
        for dj in (0, 1):
            # fine pixel (i,j) with i = 2I + a: contributions from coarse rows I and I+1
            pass
    I0 = ii // 2; J0 = jj // 2
    ai = ii % 2; aj = jj % 2
    for ci, wi in ((I0, np.where(ai == 0, 1.0, 0.5)), (I0 + 1, np.where(ai == 0, 0.0, 0.5))):
        for cj, wj in ((J0, np.where(aj == 0, 1.0, 0.5)), (J0 + 1, np.where(aj == 0, 0.0, 0.5))):
            cI = np.minimum(ci, cr - 1); cJ = np.minimum(cj, cc - 1)   # replicate at the far edge
            w = wi * wj
            m = (w > 0) & free_f
            rows_.append(fidx[m]); cols_.append((cI * cc + cJ)[m]); vals_.append(w[m])
    P = sp.coo_matrix((np.concatenate(vals_), (np.concatenate(rows_), np.concatenate(cols_))), shape=(r * c, cr * cc)).tocsr()
    return P, cr, cc


class GLevel:
    def __init__(self, A, r, c, free):
        self.A_, self.rows, self.cols, self.free = A.tocsr(), r, c, free
        d = A.diagonal()
        self.dinv = np.where(d > 0, 1.0 / np.where(d > 0, d, 1), 0.0)
        ii, jj = np.indices((r, c))
        self.colors = [(((ii % 2) == a) & ((jj % 2) == b)).ravel() for a in (0, 1) for b in (0, 1)]

    def gs(self, x, b, order):
        for col in order:
            m = self.colors[col]
            res = b - self.A_ @ x
            x = np.where(m, x + res * self.dinv, x)
        return x


class GMG:
    def __init__(self, L0, min_size=8, nu=1):
        r, c = L0.rows, L0.cols
        N = r * c
        f = L0.free.ravel()
        D = sp.diags(f.astype(float))
        A = D @ (sp.diags(L0.diag.ravel()) + L0.off) @ D
        self.levels = [GLevel(A, r, c, L0.free)]
        self.P = []
        self.nu = nu
        free = L0.free
        while min(r, c) > min_size:
            P, cr, cc = bilinear_P(r, c, free, None)
            Ac = (P.T @ A @ P).tocsr()
            cfree = (np.asarray(abs(P).sum(0)).ravel() > 0).reshape(cr, cc)
            self.P.append(P)
            A, r, c, free = Ac, cr, cc, cfree
            self.levels.append(GLevel(A, r, c, free))
        Lc = self.levels[-1]
        ff = Lc.free.ravel()
        self.cf = ff
        self.lu = spla.splu(Lc.A_[ff][:, ff].tocsc()) if ff.any() else None
        self.L0 = L0

    def vcycle(self, l, b):
        L = self.levels[l]
        if l == len(self.levels) - 1:
            x = np.zeros_like(b)
            if self.lu is not None:
                x[self.cf] = self.lu.solve(b[self.cf])
            return x
        x = np.zeros_like(b)
        if l == 0:
            x2 = x.reshape(L.rows, L.cols); b2 = b.reshape(L.rows, L.cols)
            for _ in range(self.nu):
                x2 = self.L0.gs(x2, b2, (0, 1))
            x = x2.ravel()
        else:
            for _ in range(self.nu):
                x = L.gs(x, b, (0, 1, 2, 3))
        res = b - L.A_ @ x
        ec = self.vcycle(l + 1, self.P[l].T @ res)
        x = x + self.P[l] @ ec
        if l == 0:
            x2 = x.reshape(L.rows, L.cols)
            for _ in range(self.nu):
                x2 = self.L0.gs(x2, b2, (1, 0))
            x = x2.ravel()
        else:
            for _ in range(self.nu):
                x = L.gs(x, b, (3, 2, 1, 0))
        return x


def run_galerkin(n, cases):
    for name in cases:
        rng = np.random.default_rng(0)
        L = make_case(name, n, rng)
        xc = np.where(L.hard, rng.random((n, n)), 0.0)
        full = fine_level(n, n, np.zeros((n, n), bool))
        full.wd, full.wr, full.s = L.wd, L.wr, L.s
        full.build()
        b = np.where(L.free, -full.A(xc) + L.s * rng.random((n, n)), 0.0)
        if not b.any():
            b = np.where(L.free, rng.normal(size=(n, n)), 0.0)
        t = time.time()
        mg = GMG(L)
        M = lambda r: mg.vcycle(0, r.ravel()).reshape(r.shape)
        x, it, hist = pcg(L, b, M, 1e-6)
        x2, it2, _ = pcg(L, b, M, 1e-10)
        print('%-8s n=%d galerkin levels=%d  its(1e-6)=%3d its(1e-10)=%3d  rate=%.3f  %.1fs' % (
            name, n, len(mg.levels), it, it2, (hist[-1] / hist[0]) ** (1.0 / max(len(hist) - 1, 1)), time.time() - t))
