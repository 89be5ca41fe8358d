"""Array-native interface to libhgrid.so.

A `GridSolver` is the device-resident counterpart of the reference's factor-once closure
(recovery.py:1304-1336): the constraint *pattern* (which pixels are hard / soft, which
edges are cut or carry a gradient constraint, the weights) is fixed at construction and
uploaded once; `solve()` may then be called with any number of value sets.
"""
import ctypes
import os
import weakref

import numpy as np

from . import _capi as C


def default_precision():
    """'fp64' unless HGRID_PRECISION=fp32.  The drop-in entry points default to fp64 so that
    results agree with the reference's float64 direct solve to ~1e-9 of the value range."""
    p = os.environ.get("HGRID_PRECISION", "fp64").lower()
    assert p in ("fp32", "fp64")
    return p


def default_rtol(precision):
    v = os.environ.get("HGRID_RTOL")
    if v is not None:
        return float(v)
    return 1e-13 if precision == "fp64" else 1e-9


def accept_relres(precision):
    """A solve that stops short of `rtol` (stagnation at the attainable accuracy, or maxiter) is
    still returned if its true relative residual is at most this; otherwise ArithmeticError."""
    v = os.environ.get("HGRID_ACCEPT_RELRES")
    if v is not None:
        return float(v)
    return 1e-10 if precision == "fp64" else 1e-7


def encode_flags(rows, cols, hard=None, soft=None, cut_down=None, cut_right=None, pen_down=None, pen_right=None):
    """Boolean (rows, cols) planes -> one flag byte per pixel (include/hgrid.h).  A few vectorised passes over the
    planes that are given (a 16384^2 hard-only mask costs one copy)."""
    f = None
    tmp = None
    for plane, bit in ((hard, C.FLAG_HARD), (cut_down, C.FLAG_CUT_DOWN), (cut_right, C.FLAG_CUT_RIGHT),
                       (pen_down, C.FLAG_PEN_DOWN), (pen_right, C.FLAG_PEN_RIGHT), (soft, C.FLAG_SOFT)):
        if plane is None:
            continue
        plane = np.ascontiguousarray(plane, dtype=bool)
        assert plane.shape == (rows, cols)
        v = plane.view(np.uint8)                      # 0 / 1 bytes
        shift = bit.bit_length() - 1
        if f is None:
            f = np.left_shift(v, shift) if shift else v.copy()
        else:
            tmp = np.left_shift(v, shift, out=tmp)
            np.bitwise_or(f, tmp, out=f)
    if f is None:
        return np.zeros((rows, cols), np.uint8)
    if hard is not None and soft is not None:
        # hard wins over soft at the same pixel (recovery.py:1286-1290)
        tmp = np.left_shift(np.bitwise_and(f, C.FLAG_HARD, out=tmp), C.FLAG_SOFT.bit_length() - 1, out=tmp)
        np.bitwise_and(f, np.invert(tmp, out=tmp), out=f)
    return f


class GridSolver:
    def __init__(self, rows, cols, channels=1, bilaplacian=False, precision=None, preconditioner=None, uniform_edges=False):
        self.rows, self.cols, self.K = int(rows), int(cols), int(channels)
        self.bilaplacian = bool(bilaplacian)
        self.precision = precision or default_precision()
        assert self.precision in ("fp32", "fp64")
        self._lib = C.lib()
        h = ctypes.c_void_p()
        C.check(self._lib.hg_create(self.rows, self.cols, self.K,
                                    C.OP_BILAPLACIAN if self.bilaplacian else C.OP_LAPLACIAN,
                                    C.PREC_FP64 if self.precision == "fp64" else C.PREC_FP32, ctypes.byref(h)))
        self._h = h
        self._finalizer = weakref.finalize(self, self._lib.hg_destroy, h)
        if preconditioner is None:
            # multigrid where it exists (the Laplacian); the library uses Jacobi for the bilaplacian
            preconditioner = os.environ.get("HGRID_PRECOND", "multigrid")
        C.check(self._lib.hg_set_preconditioner(self._h, {"jacobi": C.PRECOND_JACOBI, "multigrid": C.PRECOND_MULTIGRID}[preconditioner]))
        if uniform_edges:
            # every edge weighs 1/4: the operator is G.T*G / 4 of poisson.py (hg_set_edge_weights)
            C.check(self._lib.hg_set_edge_weights(self._h, C.EDGES_UNIFORM))
        self.flags = None
        self.info = None
        self.last_stats = None

    def close(self):
        self._finalizer()

    # ---- pattern ------------------------------------------------------------------------
    def set_pattern(self, flags, soft_weights=None, soft_scalar=0.0, w_g=0.0):
        """flags: uint8 (rows, cols); soft_weights: float64 (rows, cols) per-pixel weights or None
        (then `soft_scalar` applies to every FLAG_SOFT pixel); w_g: weight of FLAG_PEN_* edges."""
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        assert flags.shape == (self.rows, self.cols)
        self.flags = flags
        C.check(self._lib.hg_set_flags(self._h, flags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
        if soft_weights is not None:
            soft_weights = np.ascontiguousarray(soft_weights, dtype=np.float64)
            assert soft_weights.shape == (self.rows, self.cols)
        C.check(self._lib.hg_set_soft_weights(self._h, C.dptr(soft_weights), float(soft_scalar)))
        C.check(self._lib.hg_set_grad_weight(self._h, float(w_g)))
        info = C.SetupInfo()
        code = self._lib.hg_setup(self._h, ctypes.byref(info))
        self.info = info
        if code in (C.ERR_UNREFERENCED, C.ERR_EMPTY_ROWS):
            # the reference raises AssertionError here (recovery.py:493, 495)
            raise AssertionError(C.last_error())
        if code == C.ERR_SINGULAR:
            # cvxopt.cholmod raises ArithmeticError on the singular system (recovery.py:993)
            raise ArithmeticError(C.last_error())
        C.check(code)
        return info

    # ---- values ---------------------------------------------------------------------------
    def _vals(self, a):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.ndim == 2:
            a = a[:, :, None]
        assert a.shape == (self.rows, self.cols, self.K), (a.shape, (self.rows, self.cols, self.K))
        return np.ascontiguousarray(a)

    def apply(self, x):
        """y = system * x at free pixels, 0 at hard pixels (hg_apply)."""
        x = self._vals(x)
        y = np.empty_like(x)
        C.check(self._lib.hg_apply(self._h, C.dptr(x), C.dptr(y)))
        return y

    def solve(self, hard_vals=None, soft_vals=None, d_down=None, d_right=None, x0=None, rtol=None, maxiter=None,
              allow_unconverged=False, rhs=None):
        """Returns float64 (rows, cols, K).  Raises ArithmeticError if the iteration breaks down
        (singular / under-constrained system; cvxopt raises ArithmeticError there too) or does
        not converge (the reference asserts `0 == info`, recovery.py:972)."""
        keep = [self._vals(a) for a in (hard_vals, soft_vals, d_down, d_right, x0, rhs)]
        vals = C.Values(*[C.dptr(a) for a in keep])
        out = np.empty((self.rows, self.cols, self.K), np.float64)
        stats = C.Stats()
        if rtol is None:
            rtol = default_rtol(self.precision)
        if maxiter is None:
            maxiter = int(os.environ.get("HGRID_MAXITER", 0)) or max(2000, 40 * (self.rows + self.cols))
        code = self._lib.hg_solve(self._h, ctypes.byref(vals), C.dptr(out), float(rtol), int(maxiter), ctypes.byref(stats))
        self.last_stats = stats.as_dict()
        if code == C.ERR_BREAKDOWN:
            raise ArithmeticError(C.last_error())
        if code == C.ERR_NOT_CONVERGED:
            ok = allow_unconverged or os.environ.get("HGRID_ALLOW_UNCONVERGED") or stats.relres <= accept_relres(self.precision)
            if not ok:
                raise ArithmeticError("%s (relative residual %.3e after %d iterations)" % (
                    C.last_error(), stats.relres, stats.iterations))
        else:
            C.check(code)
        return out

    def _finish(self, code, stats, allow_unconverged=False):
        """Error mapping shared by the solve entry points (see `solve`)."""
        if code == C.ERR_BREAKDOWN:
            raise ArithmeticError(C.last_error())
        if code == C.ERR_NOT_CONVERGED:
            ok = allow_unconverged or os.environ.get("HGRID_ALLOW_UNCONVERGED") or stats.relres <= accept_relres(self.precision)
            if not ok:
                raise ArithmeticError("%s (relative residual %.3e after %d iterations)" % (
                    C.last_error(), stats.relres, stats.iterations))
        else:
            C.check(code)

    def solve_sparse(self, hard_idx=None, hard_vals=None, soft_idx=None, soft_vals=None, rtol=None, maxiter=None,
                     allow_unconverged=False):
        """`solve` for value constraints given as lists -- row-major pixel indices (unique) and (n, K) values, the form the
        factor-once closure of recovery.py:1304-1336 is called with: the values are scattered on the device
        (hg_upload_values_sparse) instead of the caller filling dense (rows, cols, K) planes."""
        def pack(idx, vals):
            if idx is None or len(idx) == 0:
                return 0, None, None
            idx = np.ascontiguousarray(idx, dtype=np.int64).ravel()
            vals = np.ascontiguousarray(vals, dtype=np.float64).reshape(len(idx), self.K)
            return len(idx), idx, vals
        nh, hi, hv = pack(hard_idx, hard_vals)
        ns, si, sv = pack(soft_idx, soft_vals)
        lp = ctypes.POINTER(ctypes.c_longlong)
        ip = lambda a: ctypes.cast(None, lp) if a is None else a.ctypes.data_as(lp)  # noqa: E731
        stats = C.Stats()
        C.check(self._lib.hg_upload_values_sparse(self._h, nh, ip(hi), C.dptr(hv), ns, ip(si), C.dptr(sv), ctypes.byref(stats)))
        if rtol is None:
            rtol = default_rtol(self.precision)
        if maxiter is None:
            maxiter = int(os.environ.get("HGRID_MAXITER", 0)) or max(2000, 40 * (self.rows + self.cols))
        code = self._lib.hg_solve_device(self._h, float(rtol), int(maxiter), ctypes.byref(stats))
        self.last_stats = stats.as_dict()
        if code != C.ERR_NOT_CONVERGED:
            self._finish(code, stats)
        out = np.empty((self.rows, self.cols, self.K), np.float64)
        C.check(self._lib.hg_download(self._h, C.dptr(out), ctypes.byref(stats)))
        self.last_stats = stats.as_dict()
        self._finish(code, stats, allow_unconverged)
        return out

    def solve_composed(self, constraints, weight, squared=False, hard_vals=None, rhs=None, x0=None, rtol=None, maxiter=None,
                       accept=None):
        """(L^m + weight A^T A) x = L^(m-1) rhs + weight A^T c (hg_solve_composed; poisson.py:55-89), m = 2 if `squared`.
        L is this solver's Laplacian without its soft weights (they only shape the preconditioner).  constraints: a
        sequence of (pixels, coeffs, c) rows -- row-major pixel indices, their coefficients, the (K,) right-hand side.
        fp64 solvers only.  Raises ArithmeticError like `solve`; a solve that stagnates short of `rtol` is returned if its
        true relative residual is at most `accept`."""
        assert self.precision == "fp64"
        ptr = np.zeros(len(constraints) + 1, np.int32)
        pix, coef = [], []
        c = np.zeros((max(len(constraints), 1), self.K), np.float64)
        for n, (pixels, coeffs, value) in enumerate(constraints):
            pixels = np.asarray(pixels, np.int64).ravel()
            coeffs = np.asarray(coeffs, np.float64).ravel()
            assert pixels.shape == coeffs.shape
            pix.append(pixels); coef.append(coeffs)
            ptr[n + 1] = ptr[n] + len(pixels)
            c[n] = np.asarray(value, np.float64).reshape(self.K)
        pix = np.ascontiguousarray(np.concatenate(pix) if pix else np.zeros(0), dtype=np.int32)
        coef = np.ascontiguousarray(np.concatenate(coef) if coef else np.zeros(0), dtype=np.float64)
        ip = ctypes.POINTER(ctypes.c_int)
        lc = C.LinearConstraints(int(bool(squared)), len(constraints), ptr.ctypes.data_as(ip), pix.ctypes.data_as(ip), C.dptr(coef),
                                 C.dptr(c), float(weight))
        keep = [self._vals(a) for a in (hard_vals, None, None, None, x0, rhs)]
        vals = C.Values(*[C.dptr(a) for a in keep])
        out = np.empty((self.rows, self.cols, self.K), np.float64)
        stats = C.Stats()
        if rtol is None:
            rtol = default_rtol(self.precision)
        if maxiter is None:
            maxiter = int(os.environ.get("HGRID_MAXITER", 0)) or max(5000, 40 * (self.rows + self.cols))
        code = self._lib.hg_solve_composed(self._h, ctypes.byref(vals), ctypes.byref(lc), C.dptr(out), float(rtol), int(maxiter),
                                           ctypes.byref(stats))
        self.last_stats = stats.as_dict()
        if code == C.ERR_BREAKDOWN:
            raise ArithmeticError(C.last_error())
        if code == C.ERR_NOT_CONVERGED:
            if accept is None:
                accept = accept_relres(self.precision)
            if not (os.environ.get("HGRID_ALLOW_UNCONVERGED") or stats.relres <= accept):
                raise ArithmeticError("%s (relative residual %.3e after %d iterations)" % (
                    C.last_error(), stats.relres, stats.iterations))
        else:
            C.check(code)
        return out

    def solve_u8(self, hard_vals_u8, out=None, rtol=None, maxiter=None):
        """Image form (hg_solve_u8): uint8 (rows, cols, K) in -- meaning u8 / 255 at hard pixels -- and
        uint8 out = clip(round(255 x)).  `out` may be a preallocated (e.g. pinned) uint8 array."""
        v = np.ascontiguousarray(hard_vals_u8, dtype=np.uint8).reshape(self.rows, self.cols, self.K)
        if out is None:
            out = np.empty((self.rows, self.cols, self.K), np.uint8)
        assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"] and out.size == v.size
        stats = C.Stats()
        if rtol is None:
            rtol = default_rtol(self.precision)
        if maxiter is None:
            maxiter = max(2000, 40 * (self.rows + self.cols))
        p8 = ctypes.POINTER(ctypes.c_uint8)
        code = self._lib.hg_solve_u8(self._h, v.ctypes.data_as(p8), out.ctypes.data_as(p8), float(rtol), int(maxiter), ctypes.byref(stats))
        self.last_stats = stats.as_dict()
        if code == C.ERR_BREAKDOWN:
            raise ArithmeticError(C.last_error())
        if code == C.ERR_NOT_CONVERGED:
            if stats.relres > accept_relres(self.precision):
                raise ArithmeticError("%s (relative residual %.3e after %d iterations)" % (C.last_error(), stats.relres, stats.iterations))
        else:
            C.check(code)
        return out

    # ---- three-step form used by bench.py ---------------------------------------------------
    def upload(self, hard_vals=None, soft_vals=None, d_down=None, d_right=None, x0=None, rhs=None):
        keep = [self._vals(a) for a in (hard_vals, soft_vals, d_down, d_right, x0, rhs)]
        vals = C.Values(*[C.dptr(a) for a in keep])
        stats = C.Stats()
        C.check(self._lib.hg_upload_values(self._h, ctypes.byref(vals), ctypes.byref(stats)))
        return stats.as_dict()

    def solve_device(self, rtol, maxiter):
        stats = C.Stats()
        code = self._lib.hg_solve_device(self._h, float(rtol), int(maxiter), ctypes.byref(stats))
        self.last_stats = stats.as_dict()
        if code not in (C.OK, C.ERR_NOT_CONVERGED):
            C.check(code)
        return self.last_stats

    def reset_guess(self):
        C.check(self._lib.hg_reset_guess(self._h))

    def download(self):
        out = np.empty((self.rows, self.cols, self.K), np.float64)
        stats = C.Stats()
        C.check(self._lib.hg_download(self._h, C.dptr(out), ctypes.byref(stats)))
        return out, stats.as_dict()

    def precond_apply(self, r):
        """z = M^-1 r (one Jacobi scaling or one multigrid V-cycle), r masked to the free pixels."""
        r = self._vals(r)
        z = np.empty_like(r)
        C.check(self._lib.hg_precond_apply(self._h, C.dptr(r), C.dptr(z)))
        return z

    def mg_levels(self):
        return self._lib.hg_mg_levels(self._h)

    def mg_level_stencil(self, level):
        """(rows, cols, 9) float64 Galerkin stencil of hierarchy level >= 1."""
        r, c = ctypes.c_int(), ctypes.c_int()
        C.check(self._lib.hg_mg_get_level(self._h, level, ctypes.byref(r), ctypes.byref(c), C.dptr(None)))
        out = np.empty((r.value, c.value, 9), np.float64)
        C.check(self._lib.hg_mg_get_level(self._h, level, ctypes.byref(r), ctypes.byref(c), C.dptr(out)))
        return out

    def mg_level_vectors(self, level):
        """(x, b) of hierarchy level >= 1 as the last V-cycle left them, each (rows, cols, K) float64."""
        r, c = ctypes.c_int(), ctypes.c_int()
        C.check(self._lib.hg_mg_get_level(self._h, level, ctypes.byref(r), ctypes.byref(c), C.dptr(None)))
        x = np.empty((r.value, c.value, self.K)); b = np.empty((r.value, c.value, self.K))
        C.check(self._lib.hg_mg_get_level_vectors(self._h, level, C.dptr(x), C.dptr(b)))
        return x, b

    def bench_apply(self, iters=20):
        ms = ctypes.c_double()
        C.check(self._lib.hg_bench_apply(self._h, int(iters), ctypes.byref(ms)))
        return ms.value
