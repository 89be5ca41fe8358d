"""CPU: libhgrid.so builds, loads and exports every symbol include/hgrid.h declares; the
host-side encoders behave; the product fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from harmonic_interpolation_b200 import _capi as C
from harmonic_interpolation_b200 import recovery as R
from harmonic_interpolation_b200 import solver as S
from harmonic_interpolation_b200.highlighting import grow, shrink


def header_symbols():
    src = open(os.path.join(ROOT, "include", "hgrid.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(C.LIB_PATH)
    names = header_symbols()
    assert "hg_solve" in names and "hg_create" in names
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(C.SYMBOLS) == names
    assert C.lib().hg_version() == 100


def test_header_flag_values_match_binding():
    src = open(os.path.join(ROOT, "include", "hgrid.h")).read()
    for name, val in (("HG_FLAG_HARD", C.FLAG_HARD), ("HG_FLAG_CUT_DOWN", C.FLAG_CUT_DOWN), ("HG_FLAG_CUT_RIGHT", C.FLAG_CUT_RIGHT),
                      ("HG_FLAG_PEN_DOWN", C.FLAG_PEN_DOWN), ("HG_FLAG_PEN_RIGHT", C.FLAG_PEN_RIGHT), ("HG_FLAG_SOFT", C.FLAG_SOFT)):
        m = re.search(name + r"\s*=\s*(\d+)", src)
        assert m and int(m.group(1)) == val


def test_no_gpu_means_error_not_fallback():
    if C.lib().hg_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(C.HgridError) as e:
        S.GridSolver(8, 8, 1)
    assert e.value.code == C.ERR_NO_DEVICE
    with pytest.raises(C.HgridError):
        R.solve_grid_linear_simple(4, 4, [(0, 0, [1.0])])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "harmonic_interpolation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # no import of the oracle or of a CPU solver anywhere in the product
                assert not re.search(r"^\s*(import|from)\s+(oracle|scipy)", text, flags=re.M), f
                assert "grid_oracle" not in text and "ref_shim" not in text, f


def test_encode_flags_hard_wins_over_soft():
    hard = np.zeros((3, 4), bool); hard[1, 1] = True
    soft = np.zeros((3, 4), bool); soft[1, 1] = True; soft[2, 2] = True
    f = S.encode_flags(3, 4, hard=hard, soft=soft)
    assert f[1, 1] == C.FLAG_HARD and f[2, 2] == C.FLAG_SOFT


def test_cut_edges_helpers():
    mask = np.zeros((5, 6), bool); mask[1:3, 2:5] = True
    edges = R.cut_edges_from_mask(mask)
    R.assert_valid_cut_edges(5, 6, edges)
    cd, cr = R.cut_planes_from_mask(mask)
    cd2, cr2, n = R._cut_planes(5, 6, edges)
    assert n == len(edges) and (cd == cd2).all() and (cr == cr2).all()
    dx, dy = R.zero_dx_dy_constraints_from_cut_edges(edges)
    assert len(dx) + len(dy) == len(edges)
    with pytest.raises(AssertionError):
        R.assert_valid_cut_edges(5, 6, edges + [edges[0]])       # duplicate
    with pytest.raises(AssertionError):
        R.assert_valid_cut_edges(5, 6, [((0, 0), (1, 1))])        # not adjacent
    with pytest.raises(AssertionError):
        R.assert_valid_cut_edges(5, 6, [((4, 0), (5, 0))])        # outside


def test_constraint_validation_asserts():
    with pytest.raises(AssertionError):
        R._scalar_constraints([(0, 0, 1.0), (0, 0, 2.0)], 4, 4, 4, 4)      # duplicate (recovery.py:537)
    with pytest.raises(AssertionError):
        R._scalar_constraints([(3, 0, 1.0)], 4, 4, 3, 4)                   # dx in the last row (recovery.py:523)
    with pytest.raises(AssertionError):
        R._vector_constraints([(0, 0, [1.0, 2.0]), (1, 1, [1.0])], 4, 4)    # ragged (recovery.py:1054)
    with pytest.raises(AssertionError):
        R._vector_constraints([(0, 0, 1.0)], 4, 4)                         # scalar, not a vector (:1052)


def test_grow_matches_diamond():
    m = np.zeros((9, 9), bool); m[4, 4] = True
    g2 = grow(m, 2)
    ii, jj = np.indices(m.shape)
    assert (g2 == (np.abs(ii - 4) + np.abs(jj - 4) <= 2)).all()
    assert (shrink(g2, 2) == m).all()
    assert (grow(m, 0) == m).all()


def test_normals_to_gradients():
    dx, dy = R.convert_grid_normal_constraints_to_dx_and_dy_constraints(4, 4, [(1, 1, (0.6, 0.0, 0.8)), (3, 3, (0.0, 0.6, 0.8))])
    d = {(i, j): v for i, j, v in dx}
    # a constraint in the last row / column and none in the first: everything shifts by one (recovery.py:1564-1565)
    assert abs(d[(0, 0)] + 0.75) < 1e-15 and (2, 2) in d
    e = {(i, j): v for i, j, v in dy}
    assert abs(e[(2, 2)] + 0.75) < 1e-15


def test_compat_modules_export_reference_names():
    """`from recovery import solve_grid_linear, solve_grid_linear_simple` (fill_alpha_harmonic.py:4) keeps working."""
    import subprocess
    import sys
    code = ("import recovery, fill_alpha_harmonic, highlighting, tictoc;"
            "from recovery import solve_grid_linear, solve_grid_linear_simple, solve_grid_linear_simple3, "
            "solve_grid_linear_simple3_solver, smooth_bumps, smooth_bumps2, cut_edges_from_mask, "
            "zero_dx_dy_constraints_from_cut_edges, convert_grid_normal_constraints_to_dx_and_dy_constraints, normalize_to_char_img;"
            "from highlighting import grow, shrink; from tictoc import tic, toc;"
            "from poisson import solve_poisson_simple; print('ok')")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "harmonic_interpolation_b200", "compat")]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr
