"""
TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Import shim that lets the UNMODIFIED reference (`/root/reference/recovery.py`,
`fill_alpha_harmonic.py`) run in this container, so that it can (a) validate the
numpy restatement in `oracle/grid_oracle.py` and (b) generate the golden vectors
committed under `tests/golden/` (see `oracle/make_golden.py`).

`/root/reference` does not exist on the GPU box; there the verbatim copy under
oracle/_ref/ (oracle/build_ref.py) is used by bench.py's CPU legs only.
`have_reference()` says whether either is present.

Why a shim is needed:
  * `tictoc.py:9,12` calls `time.clock`, removed in Python 3.8.
  * `recovery.py:975,1064,1181,1295` import `cvxopt` (CHOLMOD / UMFPACK), which is
    not installed here and cannot be (no network).  Both are exact sparse direct
    solvers of a well-posed system; SuperLU (`scipy.sparse.linalg.splu`) is an
    equivalent backward-stable stand-in.
  * `recovery.py:1154,1262,1274` use `asfarray`, removed in numpy 2.0.
"""
import os
import sys
import time
import types

import numpy as np

# the reference tree where it lies (build container), else the verbatim copy made by oracle/build_ref.py
# (oracle/_ref/, git-ignored, travels to the GPU box)
_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_DIR = os.environ.get("HGRID_REFERENCE_DIR") or (
    "/root/reference" if os.path.isfile("/root/reference/recovery.py") else os.path.join(_HERE, "_ref"))


def have_reference():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "recovery.py"))


class _Matrix:
    """Stand-in for cvxopt.matrix: dense float64, column vector if 1-D."""

    def __init__(self, a):
        a = np.array(a, dtype=float)
        if a.ndim == 1:
            a = a[:, None]
        self.a = a

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)


class _SpMatrix:
    """Stand-in for cvxopt.spmatrix(x, I, J): square, duplicates summed."""

    def __init__(self, x, I, J):
        import scipy.sparse as sp
        I = np.asarray(I)
        J = np.asarray(J)
        n = int(max(I.max(), J.max())) + 1
        self.m = sp.coo_matrix((np.asarray(x, dtype=float), (I, J)), shape=(n, n)).tocsc()


def _install_fake_cvxopt():
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    cvxopt = types.ModuleType("cvxopt")
    cholmod = types.ModuleType("cvxopt.cholmod")
    umfpack = types.ModuleType("cvxopt.umfpack")
    cvxopt.matrix = _Matrix
    cvxopt.spmatrix = _SpMatrix

    def linsolve(A, B):
        # CHOLMOD reads only the lower triangle of a symmetric matrix.
        lower = sp.tril(A.m)
        sym = lower + sp.tril(A.m, -1).T
        B.a[...] = spla.splu(sym.tocsc()).solve(B.a)

    cholmod.linsolve = linsolve
    umfpack.symbolic = lambda A: None
    umfpack.numeric = lambda A, sym: spla.splu(A.m)

    def solve(A, F, B):
        B.a[...] = F.solve(B.a)

    umfpack.solve = solve
    cvxopt.cholmod = cholmod
    cvxopt.umfpack = umfpack
    sys.modules["cvxopt"] = cvxopt
    sys.modules["cvxopt.cholmod"] = cholmod
    sys.modules["cvxopt.umfpack"] = umfpack


_loaded = {}


def load_reference():
    """Returns the reference's modules as a dict: recovery, fill_alpha_harmonic, highlighting, poisson."""
    if _loaded:
        return _loaded
    if not have_reference():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    if not hasattr(time, "clock"):
        time.clock = time.perf_counter
    _install_fake_cvxopt()
    # The reference modules are flat top-level modules; import them under their own
    # names from REFERENCE_DIR, then restore sys.path and sys.modules so that the
    # product's same-named drop-in modules are never shadowed.
    saved_path = list(sys.path)
    names = ["tictoc", "recovery", "fill_alpha_harmonic", "highlighting", "poisson"]
    saved_mods = {n: sys.modules.pop(n, None) for n in names}
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import importlib
        for n in names:
            _loaded[n] = importlib.import_module(n)
        _loaded["recovery"].asfarray = lambda a: np.asarray(a, dtype=float)
    finally:
        sys.path[:] = saved_path
        for n in names:
            sys.modules.pop(n, None)
            if saved_mods[n] is not None:
                sys.modules[n] = saved_mods[n]
    # the reference's functions do `from tictoc import tic, toc` at call time
    # and smooth_bumps* do `from highlighting import grow`
    sys.modules.setdefault("tictoc", _loaded["tictoc"])
    sys.modules.setdefault("highlighting", _loaded["highlighting"])
    return _loaded
