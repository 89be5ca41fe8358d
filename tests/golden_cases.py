"""Rebuilds the argument lists of the golden cases in tests/golden/reference_outputs.npz
(written by oracle/make_golden.py) for any implementation with the reference's signatures."""
from conftest import cuts_list


def s3_args(g):
    rows, cols = int(g["s3_rows"]), int(g["s3_cols"])
    hard = [(int(i), int(j), list(v)) for (i, j), v in zip(g["s3_hard_ij"], g["s3_hard_v"])]
    soft = [(int(i), int(j), list(v)) for (i, j), v in zip(g["s3_soft_ij"], g["s3_soft_v"])]
    return rows, cols, hard, soft, cuts_list(g["s3_cuts"])


def s3_cases(g):
    """yields (key, kwargs) for solve_grid_linear_simple3"""
    rows, cols, hard, soft, cuts = s3_args(g)
    for bil in (False, True):
        for cname, cu in (("nocut", None), ("cut", cuts)):
            for wname, w in (("w1", 1.0), ("w1e3", 1e3), ("wlist", list(g["s3_w_list"]))):
                key = "s3_out_%s_%s_%s" % ("bilap" if bil else "lap", cname, wname)
                yield key, dict(rows=rows, cols=cols, value_constraints_hard=hard, value_constraints_soft=soft,
                                w_lsq=w, cut_edges=cu, bilaplacian=bil)
    yield "s3_out_lap_hardonly", dict(rows=rows, cols=cols, value_constraints_hard=hard, value_constraints_soft=[],
                                      w_lsq=1.0, cut_edges=cuts, bilaplacian=False)
    yield "s3_out_bilap_softonly", dict(rows=rows, cols=cols, value_constraints_hard=[], value_constraints_soft=soft,
                                        w_lsq=10.0, cut_edges=cuts, bilaplacian=True)


def s1_cases(g):
    import numpy as np
    keep, img = g["s1_keep"], g["s1_img"]
    rows, cols = keep.shape
    vc = [(int(i), int(j), img[i, j]) for i, j in zip(*np.where(keep))]
    yield "s1_out_lap", dict(rows=rows, cols=cols, value_constraints=vc, bilaplacian=False)
    yield "s1_out_bilap", dict(rows=rows, cols=cols, value_constraints=vc, bilaplacian=True)


def g_cases(g):
    rows, cols = int(g["g_rows"]), int(g["g_cols"])
    dx = [(int(i), int(j), float(v)) for (i, j), v in zip(g["g_dx_ij"], g["g_dx_v"])]
    dy = [(int(i), int(j), float(v)) for (i, j), v in zip(g["g_dy_ij"], g["g_dy_v"])]
    vc = [(int(i), int(j), float(v)) for (i, j), v in zip(g["g_vc_ij"], g["g_vc_v"])]
    cuts = cuts_list(g["g_cuts"])
    for bil in (False, True):
        for cname, cu in (("nocut", None), ("cut", cuts)):
            for vname, v in (("pin", None), ("hard", vc)):
                for wname, w in (("wdef", None), ("w10", 10.0)):
                    key = "g_out_%s_%s_%s_%s" % ("bilap" if bil else "lap", cname, vname, wname)
                    yield key, dict(rows=rows, cols=cols, dx_constraints=dx, dy_constraints=dy, value_constraints=v,
                                    bilaplacian=bil, cut_edges=cu, w_lsq=w)
    yield "g_out_lap_valuesonly", dict(rows=rows, cols=cols, value_constraints=vc, bilaplacian=False, cut_edges=cuts)
    vc4 = [(int(i), int(j), float(v)) for (i, j), v in zip(g["g_vc4_ij"], g["g_vc4_v"])]
    yield "g_out_bilap_valuesonly", dict(rows=rows, cols=cols, value_constraints=vc4, bilaplacian=True, cut_edges=None)


def sb_cases(g):
    grid = g["sb_grid"]
    locs = [tuple(int(v) for v in l) for l in g["sb_locs"]]
    for bil in (False, True):
        b = "bilap" if bil else "lap"
        yield "sb_out_%s_r0" % b, "smooth_bumps", (grid, locs, 0), dict(bilaplacian=bil)
        yield "sb_out_%s_r2" % b, "smooth_bumps", (grid, locs, 2), dict(bilaplacian=bil)
        for fn in ("median", "average"):
            yield "sb2_out_%s_%s" % (b, fn), "smooth_bumps2", (grid, locs, 2, 2, fn), dict(bilaplacian=bil, w_lsq=100.0)
