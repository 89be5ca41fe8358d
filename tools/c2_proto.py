"""Numpy / scipy prototype for the bilaplacian hierarchy on the C2 pattern (2 % hard, 5 % soft of weight 10, cut edges around
16 x 16 blocks): which ingredient costs the iterations, and what a cut-aware interpolation buys.  Design tool, not product
code; uses the classes of tools/stiff_proto.py.

Findings (round 2), PCG iterations to 1e-10 with a V(1,1) cycle, lexicographic Gauss-Seidel:
  no cuts, hard pins only     86 /  90 /  96   at 64^2 / 128^2 / 256^2   (flat, but high: point pins of a thin plate)
  no cuts, + soft             52 /  54 /  48
  cuts, hard pins only       150 / 472 / 989   (growing: bilinear interpolation averages across edges that decouple pixels)
  cuts + soft (= C2)          76 / 135 / 144
  cut-aware interpolation (a parent that can only be reached across a cut is dropped, the weights renormalised; coarse cuts
  inherited; Galerkin operators of the modified P):   cuts only 99 / 177, C2 59 / 74 at 64^2 / 128^2 -- about half.
On the device this needs transfer kernels with per-pixel parent masks and Galerkin products for them on every level (the
probing of mg25_clusters.cuh would do); not written in round 2.

usage: python tools/c2_proto.py
"""
import sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, time
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stiff_proto import MG, pcg, bilinear_P
from oracle import grid_oracle as go

def c2_problem(n, seed=0, cuts=True, soft=True):
    rng = np.random.default_rng(seed)
    hard = rng.random((n, n)) < 0.02
    softm = (rng.random((n, n)) < 0.05) & ~hard
    hv = rng.random((n, n)); sv = rng.random((n, n))
    block = np.kron(rng.random((n // 16 + 1, n // 16 + 1)) < 0.3, np.ones((16, 16), bool))[:n, :n]
    prob = go.GridProblem(n, n, 1, True)
    prob.hard = hard; prob.hard_vals[:, :, 0] = hv
    if soft:
        prob.soft_w = np.where(softm, 10.0, 0.0); prob.soft_vals[:, :, 0] = sv
    if cuts:
        cd = np.zeros((n, n), bool); cr = np.zeros((n, n), bool)
        cd[:-1] = block[:-1] != block[1:]; cr[:, :-1] = block[:, :-1] != block[:, 1:]
        prob.cut_down, prob.cut_right = cd, cr; prob.have_cuts = True
    return prob

def ingredients():
    for n in (64, 128, 256):
        for cuts, soft in ((False, False), (False, True), (True, False), (True, True)):
            prob = c2_problem(n, 0, cuts, soft)
            try:
                A, b, free = go.reduced_system(prob, check=False)
            except AssertionError as e:
                print('assert', e); continue
            A = A.tocsr(); b = b[:, 0]
            fm = np.zeros(n * n, bool); fm[free] = True
            mg = MG(A, fm, n, n)
            x, it = pcg(A, b, mg.precond, 1e-10, 1500)
            print('n=%d cuts=%d soft=%d: %d iterations' % (n, cuts, soft, it), flush=True)


def cut_aware_P(rows, cols, cd, cr):
    """bilinear, but a parent that can only be reached across a cut edge is dropped and the remaining weights renormalised.
    cd[i,j]: edge (i,j)-(i+1,j) cut; cr[i,j]: edge (i,j)-(i,j+1) cut.  Returns P (N_f x N_c), coarse dims, coarse cuts."""
    rc, cc = (rows + 1) // 2, (cols + 1) // 2
    I, J, V = [], [], []
    def ok_v(i0, i1, j):   # vertical path i0 -> i1 (adjacent rows) at column j uncut
        a = min(i0, i1); return not cd[a, j]
    def ok_h(i, j0, j1):
        a = min(j0, j1); return not cr[i, a]
    for i in range(rows):
        for j in range(cols):
            p = i * cols + j
            pi = [i // 2] if i % 2 == 0 else [i // 2, i // 2 + 1]
            pj = [j // 2] if j % 2 == 0 else [j // 2, j // 2 + 1]
            cand = []
            for a in pi:
                for b in pj:
                    if a >= rc or b >= cc: continue
                    fi, fj = 2 * a, 2 * b
                    if fi >= rows or fj >= cols: continue
                    # reachable without crossing a cut: direct, or one of the two L paths
                    if fi == i and fj == j: r = True
                    elif fi == i: r = ok_h(i, j, fj)
                    elif fj == j: r = ok_v(i, fi, j)
                    else: r = (ok_v(i, fi, j) and ok_h(fi, j, fj)) or (ok_h(i, j, fj) and ok_v(i, fi, fj))
                    w = (1.0 if i % 2 == 0 else 0.5) * (1.0 if j % 2 == 0 else 0.5)
                    if r: cand.append((a * cc + b, w))
            if not cand: continue
            tot = sum(w for _, w in cand)
            full = 1.0
            for q, w in cand:
                I.append(p); J.append(q); V.append(w / tot * full)
    P = sp.coo_matrix((V, (I, J)), shape=(rows * cols, rc * cc)).tocsr()
    ccd = np.zeros((rc, cc), bool); ccr = np.zeros((rc, cc), bool)
    for a in range(rc):
        for b in range(cc):
            fi, fj = 2 * a, 2 * b
            if fi + 2 < rows + 0 and fi + 1 < rows: ccd[a, b] = cd[fi, fj] or (fi + 1 < rows and cd[fi + 1, fj])
            if fj + 1 < cols: ccr[a, b] = cr[fi, fj] or (fj + 1 < cols and cr[fi, fj + 1])
    return P, rc, cc, ccd, ccr

class MGc:
    def __init__(self, A, free_mask, rows, cols, cd, cr, aware, min_n=16):
        N = rows * cols
        E = sp.identity(N, format='csr')[:, np.flatnonzero(free_mask)]
        self.E = E; self.levels = []
        cur = (E @ A @ E.T).tocsr(); r, c = rows, cols
        while True:
            act = cur.diagonal() != 0
            idx = np.flatnonzero(act); As = cur[idx][:, idx].tocsr()
            lev = dict(A=cur, idx=idx, As=As, L=sp.tril(As, format='csr'), U=sp.triu(As, format='csr'))
            self.levels.append(lev)
            if r * c <= min_n * min_n:
                lev['lu'] = spla.splu(As.tocsc()); break
            if aware: P, rc, cc, cd, cr = cut_aware_P(r, c, cd, cr)
            else: P, rc, cc, _, _ = cut_aware_P(r, c, np.zeros_like(cd), np.zeros_like(cr)); cd = np.zeros((rc, cc), bool); cr = np.zeros((rc, cc), bool)
            P = sp.diags(act.astype(float)) @ P
            lev['P'] = P
            cur = (P.T @ cur @ P).tocsr(); cur.eliminate_zeros(); r, c = rc, cc
    def smooth(self, lev, x, b, fwd):
        As, idx = lev['As'], lev['idx']; xs, bs = x[idx], b[idx]
        xs = xs + spla.spsolve_triangular(lev['L'] if fwd else lev['U'], bs - As @ xs, lower=fwd)
        x = x.copy(); x[idx] = xs; return x
    def vcycle(self, l, b):
        lev = self.levels[l]; x = np.zeros_like(b)
        if 'lu' in lev: x[lev['idx']] = lev['lu'].solve(b[lev['idx']]); return x
        x = self.smooth(lev, x, b, True)
        x = x + lev['P'] @ self.vcycle(l + 1, lev['P'].T @ (b - lev['A'] @ x))
        return self.smooth(lev, x, b, False)
    def precond(self, rf): return self.E.T @ self.vcycle(0, self.E @ rf)

def cut_aware():
  for n in (64, 128):
    for soft in (False, True):
        prob = c2_problem(n, 0, True, soft)
        A, b, free = go.reduced_system(prob, check=False); A = A.tocsr(); b = b[:, 0]
        fm = np.zeros(n * n, bool); fm[free] = True
        out = []
        for aware in (False, True):
            mg = MGc(A, fm, n, n, prob.cut_down.copy(), prob.cut_right.copy(), aware)
            x, it = pcg(A, b, mg.precond, 1e-10, 1500)
            out.append('%s %d' % ('cut-aware P' if aware else 'bilinear P', it))
        print('n=%d soft=%d: %s' % (n, soft, ' | '.join(out)), flush=True)


if __name__ == '__main__':
    ingredients()
    cut_aware()
