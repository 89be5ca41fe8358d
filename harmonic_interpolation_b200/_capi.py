"""ctypes binding of libhgrid.so (include/hgrid.h).  Nothing here computes: it loads the
library that holds the CUDA kernels and fails loudly if it is missing."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HGRID_LIB", os.path.join(_HERE, "libhgrid.so"))

FLAG_HARD, FLAG_CUT_DOWN, FLAG_CUT_RIGHT, FLAG_PEN_DOWN, FLAG_PEN_RIGHT, FLAG_SOFT = 1, 2, 4, 8, 16, 32
OP_LAPLACIAN, OP_BILAPLACIAN = 0, 1
PREC_FP32, PREC_FP64 = 0, 1
PRECOND_JACOBI, PRECOND_MULTIGRID = 0, 1
EDGES_REFERENCE, EDGES_UNIFORM = 0, 1

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_UNREFERENCED, ERR_EMPTY_ROWS = -1, -2, -3, -4, -5
ERR_NOT_CONVERGED, ERR_BREAKDOWN, ERR_NOMEM, ERR_COMM, ERR_SINGULAR = -6, -7, -8, -9, -10

_dp = ctypes.POINTER(ctypes.c_double)


class SetupInfo(ctypes.Structure):
    _fields_ = [("levels", ctypes.c_int), ("empty_rows", ctypes.c_int), ("unreferenced", ctypes.c_int),
                ("n_free", ctypes.c_longlong), ("n_hard", ctypes.c_longlong), ("setup_ms", ctypes.c_double),
                ("unconstrained", ctypes.c_longlong)]


class Values(ctypes.Structure):
    _fields_ = [("hard_vals", _dp), ("soft_vals", _dp), ("d_down", _dp), ("d_right", _dp), ("x0", _dp), ("rhs", _dp)]


class LinearConstraints(ctypes.Structure):
    _fields_ = [("squared", ctypes.c_int), ("n_constraints", ctypes.c_int), ("row_ptr", ctypes.POINTER(ctypes.c_int)),
                ("pixel", ctypes.POINTER(ctypes.c_int)), ("coeff", _dp), ("rhs", _dp), ("weight", ctypes.c_double)]


class Stats(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int), ("status", ctypes.c_int), ("relres", ctypes.c_double),
                ("relres_cg", ctypes.c_double), ("solve_ms", ctypes.c_double), ("upload_ms", ctypes.c_double),
                ("download_ms", ctypes.c_double), ("bytes_moved", ctypes.c_double), ("launches", ctypes.c_longlong)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/hgrid.h declares: name -> (restype, argtypes)
_H = ctypes.c_void_p
SYMBOLS = {
    "hg_version": (ctypes.c_int, []),
    "hg_last_error": (ctypes.c_char_p, []),
    "hg_device_count": (ctypes.c_int, []),
    "hg_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_H)]),
    "hg_destroy": (ctypes.c_int, [_H]),
    "hg_trim_memory": (ctypes.c_int, []),
    "hg_set_flags": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_uint8)]),
    "hg_set_soft_weights": (ctypes.c_int, [_H, _dp, ctypes.c_double]),
    "hg_set_grad_weight": (ctypes.c_int, [_H, ctypes.c_double]),
    "hg_set_preconditioner": (ctypes.c_int, [_H, ctypes.c_int]),
    "hg_set_edge_weights": (ctypes.c_int, [_H, ctypes.c_int]),
    "hg_setup": (ctypes.c_int, [_H, ctypes.POINTER(SetupInfo)]),
    "hg_apply": (ctypes.c_int, [_H, _dp, _dp]),
    "hg_solve": (ctypes.c_int, [_H, ctypes.POINTER(Values), _dp, ctypes.c_double, ctypes.c_int, ctypes.POINTER(Stats)]),
    "hg_solve_composed": (ctypes.c_int, [_H, ctypes.POINTER(Values), ctypes.POINTER(LinearConstraints), _dp, ctypes.c_double, ctypes.c_int,
                                         ctypes.POINTER(Stats)]),
    "hg_solve_u8": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint8), ctypes.c_double, ctypes.c_int, ctypes.POINTER(Stats)]),
    "hg_upload_values": (ctypes.c_int, [_H, ctypes.POINTER(Values), ctypes.POINTER(Stats)]),
    "hg_upload_values_sparse": (ctypes.c_int, [_H, ctypes.c_longlong, ctypes.POINTER(ctypes.c_longlong), _dp, ctypes.c_longlong,
                                               ctypes.POINTER(ctypes.c_longlong), _dp, ctypes.POINTER(Stats)]),
    "hg_solve_device": (ctypes.c_int, [_H, ctypes.c_double, ctypes.c_int, ctypes.POINTER(Stats)]),
    "hg_download": (ctypes.c_int, [_H, _dp, ctypes.POINTER(Stats)]),
    "hg_reset_guess": (ctypes.c_int, [_H]),
    "hg_comm_unique_id": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint8)]),
    "hg_comm_init": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "hg_mg_get_level_vectors": (ctypes.c_int, [_H, ctypes.c_int, _dp, _dp]),
    "hg_set_phase_timing": (ctypes.c_int, [_H, ctypes.c_int]),
    "hg_get_phase_ms": (ctypes.c_int, [_H, _dp, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "hg_bench_apply": (ctypes.c_int, [_H, ctypes.c_int, _dp]),
    "hg_precond_apply": (ctypes.c_int, [_H, _dp, _dp]),
    "hg_mg_levels": (ctypes.c_int, [_H]),
    "hg_mg_get_level": (ctypes.c_int, [_H, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), _dp]),
}

N_PHASES = 10
PHASE_NAMES = ["other", "stencil+alpha", "x/r update+check", "fine down", "level-1 down", "levels>=2", "level-1 up", "fine up",
               "beta+p update", "outer residual"]

_lib = None


class HgridError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libhgrid error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Loads libhgrid.so once.  Raises if it has not been built: there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "libhgrid.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This package has no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def last_error():
    return lib().hg_last_error().decode("utf-8", "replace")


def check(code):
    if code != OK:
        raise HgridError(code, last_error())


def dptr(a):
    """float64 C-contiguous ndarray (or None) -> double*"""
    if a is None:
        return ctypes.cast(None, _dp)
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)
