"""scipy restatement of the multigrid transfer and Galerkin coarse operator (test-side only):
masked bilinear P with constant extrapolation at the far edge of even-sized grids, A_c = P^T A P."""
import numpy as np
import scipy.sparse as sp


def bilinear_P(r, c, active):
    cr, cc = (r + 1) // 2, (c + 1) // 2
    ii, jj = np.indices((r, c))
    fidx = ii * c + jj
    I0, J0 = ii // 2, jj // 2
    ai, aj = ii % 2, jj % 2
    rows_, cols_, vals_ = [], [], []
    for ci, wi in ((I0, np.where(ai == 0, 1.0, 0.5)), (I0 + 1, np.where(ai == 0, 0.0, 0.5))):
        for cj, wj in ((J0, np.where(aj == 0, 1.0, 0.5)), (J0 + 1, np.where(aj == 0, 0.0, 0.5))):
            cI = np.minimum(ci, cr - 1)
            cJ = np.minimum(cj, cc - 1)
            w = wi * wj
            m = (w > 0) & active
            rows_.append(fidx[m]); cols_.append((cI * cc + cJ)[m]); vals_.append(w[m])
    P = sp.coo_matrix((np.concatenate(vals_), (np.concatenate(rows_), np.concatenate(cols_))), shape=(r * c, cr * cc)).tocsr()
    return P, cr, cc


def stencil_from_matrix(A, r, c):
    """(r, c, 9) array of the entries A[p, p+d], d in {-1,0,1}^2."""
    A = A.tocsr()
    out = np.zeros((r, c, 9))
    ii, jj = np.indices((r, c))
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            qi, qj = ii + di, jj + dj
            ok = (qi >= 0) & (qi < r) & (qj >= 0) & (qj < c)
            p = (ii * c + jj)[ok]
            q = (qi * c + qj)[ok]
            out[ii[ok], jj[ok], (di + 1) * 3 + (dj + 1)] = np.asarray(A[p, q]).ravel()
    return out
