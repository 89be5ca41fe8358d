"""
TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference, through oracle/ref_shim.py) on seeded inputs.  Run in the
build container only:

    python oracle/make_golden.py

Each case stores its inputs as plain arrays (so the tests never depend on a RNG
stream) and the reference's float64 output.  The two PNGs of the reference's only
known-answer test (fill_alpha_harmonic-test/test.png -> expected.png) are copied
beside them as data fixtures.
"""
import contextlib
import io
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def cut_edges_array(cut_edges):
    if not cut_edges:
        return np.zeros((0, 4), np.int32)
    return np.asarray([(a[0], a[1], b[0], b[1]) for a, b in cut_edges], np.int32)


def pick(rng, shape, p, exclude=None):
    m = rng.random(shape) < p
    if exclude is not None:
        m[: exclude.shape[0], : exclude.shape[1]] &= ~exclude[: shape[0], : shape[1]]
    return np.argwhere(m).astype(np.int32)


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_shim.load_reference()
    R = ref["recovery"]
    cwd = os.getcwd()
    os.chdir("/tmp")  # the reference's self-test writes PNGs into the CWD

    for n in ("test.png", "expected.png"):
        shutil.copyfile(os.path.join(ref_shim.REFERENCE_DIR, "fill_alpha_harmonic-test", n),
                        os.path.join(GOLD, "fill_alpha_harmonic_" + n))

    rng = np.random.default_rng(20261019)
    out = {}

    # ---- solve_grid_linear_simple3: hard + soft (+ overlap), cuts, both operators ----
    rows, cols, K = 48, 40, 2
    block = np.zeros((rows, cols), bool)
    block[12:30, 9:25] = True
    cuts = R.cut_edges_from_mask(block)
    hard_ij = pick(rng, (rows, cols), 0.03)
    hmask = np.zeros((rows, cols), bool)
    hmask[tuple(hard_ij.T)] = True
    soft_ij = pick(rng, (rows, cols), 0.08, exclude=hmask)
    soft_ij = np.concatenate([soft_ij, hard_ij[:1]])  # one overlapping pixel: hard wins
    hard_v = rng.random((len(hard_ij), K))
    soft_v = rng.random((len(soft_ij), K))
    w_list = rng.uniform(0.1, 100.0, len(soft_ij))
    out.update(s3_rows=rows, s3_cols=cols, s3_hard_ij=hard_ij, s3_soft_ij=soft_ij, s3_hard_v=hard_v,
               s3_soft_v=soft_v, s3_w_list=w_list, s3_cuts=cut_edges_array(cuts))
    hard = [(int(i), int(j), list(v)) for (i, j), v in zip(hard_ij, hard_v)]
    soft = [(int(i), int(j), list(v)) for (i, j), v in zip(soft_ij, soft_v)]
    for bil in (False, True):
        for cname, cu in (("nocut", None), ("cut", cuts)):
            for wname, w in (("w1", 1.0), ("w1e3", 1e3), ("wlist", list(w_list))):
                key = "s3_out_%s_%s_%s" % ("bilap" if bil else "lap", cname, wname)
                out[key] = quiet(R.solve_grid_linear_simple3, rows, cols, hard, soft, w, cu, bil)
    # hard only / soft only
    out["s3_out_lap_hardonly"] = quiet(R.solve_grid_linear_simple3, rows, cols, hard, [], 1.0, cuts, False)
    out["s3_out_bilap_softonly"] = quiet(R.solve_grid_linear_simple3, rows, cols, [], soft, 10.0, cuts, True)

    # ---- solve_grid_linear_simple: K = 3 fill with a tile mask -----------------------
    rows, cols, K = 64, 56, 3
    keep = np.kron(rng.random((8, 7)) < 0.5, np.ones((8, 8), bool))
    keep[0, 0] = True
    img = rng.random((rows, cols, K))
    out.update(s1_keep=keep, s1_img=img)
    vc = [(int(i), int(j), img[i, j]) for i, j in zip(*np.where(keep))]
    out["s1_out_lap"] = quiet(R.solve_grid_linear_simple, rows, cols, vc, False)
    out["s1_out_bilap"] = quiet(R.solve_grid_linear_simple, rows, cols, vc, True)

    # ---- solve_grid_linear: gradient constraints, soft pin, hard values, cuts ----------
    rows, cols = 36, 44
    block = np.zeros((rows, cols), bool)
    block[10:22, 14:30] = True
    cuts = R.cut_edges_from_mask(block)
    dx_ij = pick(rng, (rows - 1, cols), 0.04)
    dy_ij = pick(rng, (rows, cols - 1), 0.04)
    dx_v = rng.normal(size=len(dx_ij))
    dy_v = rng.normal(size=len(dy_ij))
    vc_ij = np.asarray([(3, 4), (20, 20), (35, 43)], np.int32)
    vc_v = np.asarray([1.0, -2.0, 0.5])
    out.update(g_rows=rows, g_cols=cols, g_dx_ij=dx_ij, g_dy_ij=dy_ij, g_dx_v=dx_v, g_dy_v=dy_v,
               g_vc_ij=vc_ij, g_vc_v=vc_v, g_cuts=cut_edges_array(cuts))
    dx = [(int(i), int(j), float(v)) for (i, j), v in zip(dx_ij, dx_v)]
    dy = [(int(i), int(j), float(v)) for (i, j), v in zip(dy_ij, dy_v)]
    vc = [(int(i), int(j), float(v)) for (i, j), v in zip(vc_ij, vc_v)]
    for bil in (False, True):
        for cname, cu in (("nocut", None), ("cut", cuts)):
            for vname, v in (("pin", None), ("hard", vc)):
                for wname, w in (("wdef", None), ("w10", 10.0)):
                    key = "g_out_%s_%s_%s_%s" % ("bilap" if bil else "lap", cname, vname, wname)
                    out[key] = quiet(R.solve_grid_linear, rows, cols, dx, dy, v, bil, False, cu, w)
    # value constraints only (no gradients): plain Dirichlet problem
    out["g_out_lap_valuesonly"] = quiet(R.solve_grid_linear, rows, cols, None, None, vc, False, False, cuts, None)
    # The null space of L_refl^T L_refl is spanned by {1, i, j, i*j} (the bilinear saddle is discrete
    # harmonic and affine along every side), per cut-off component: with the 3 pins above the
    # bilaplacian system is singular and SuperLU's answer arbitrary -- not a valid golden.  Use 4 pins
    # in general position and no cuts.
    vc4 = vc + [(30, 5, 0.25)]
    out["g_vc4_ij"] = np.asarray([(i, j) for i, j, _ in vc4], np.int32)
    out["g_vc4_v"] = np.asarray([v for _, _, v in vc4])
    out["g_out_bilap_valuesonly"] = quiet(R.solve_grid_linear, rows, cols, None, None, vc4, True, False, None, None)

    # ---- smooth_bumps / smooth_bumps2 (doctest inputs, recovery.py:1360-1365, 1437-1444) -
    g = 5 * np.outer(np.sin(np.linspace(0, np.pi, 20)), np.cos(np.linspace(0, np.pi, 24)))
    g += 0.3 * rng.normal(size=g.shape)
    out["sb_grid"] = g
    locs = [(3, 3), (0, 0), (0, 4), (10, 12)]
    out["sb_locs"] = np.asarray(locs, np.int32)
    for bil in (False, True):
        b = "bilap" if bil else "lap"
        out["sb_out_%s_r0" % b] = quiet(R.smooth_bumps, g, locs, 0, bilaplacian=bil)
        out["sb_out_%s_r2" % b] = quiet(R.smooth_bumps, g, locs, 2, bilaplacian=bil)
        for fn in ("median", "average"):
            out["sb2_out_%s_%s" % (b, fn)] = quiet(R.smooth_bumps2, g, locs, 2, 2, fn, bilaplacian=bil, w_lsq=100.0)
    np.savez_compressed(os.path.join(GOLD, "reference_outputs.npz"), **out)

    # ---- the reference's own self-test (recovery.py:2020-2077) at its 'large' size ------
    rows = cols = 250
    br0, br1 = 100, 150
    vcs = [(0, 0, [br0 * 2.0]), (rows - 1, 0, [0.0]), (0, cols - 1, [0.0]), (rows - 1, cols - 1, [0.0]),
           (br0, br0, [br0]), (br1 - 1, br0, [br0]), (br0, br1 - 1, [br0]), (br1 - 1, br1 - 1, [br0])]
    mask = np.zeros((rows, cols), bool)
    mask[br0:br1, br0:br1] = True
    cuts = R.cut_edges_from_mask(mask)
    f = R.solve_grid_linear_simple3
    st = {}
    st["sol_hard"] = quiet(f, rows, cols, value_constraints_hard=vcs, cut_edges=cuts, bilaplacian=True)
    st["sol_soft"] = quiet(f, rows, cols, value_constraints_soft=vcs, w_lsq=1e3, cut_edges=cuts, bilaplacian=True)
    st["sol_mixed"] = quiet(f, rows, cols, value_constraints_hard=vcs,
                            value_constraints_soft=[((br0 + br1) // 2, (br0 + br1) // 2, [0.5 * br0])],
                            w_lsq=1e3, cut_edges=cuts, bilaplacian=True)
    np.savez_compressed(os.path.join(GOLD, "reference_selftest_250.npz"),
                        **{k: v.astype(np.float64) for k, v in st.items()})
    os.chdir(cwd)
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)))


if __name__ == "__main__":
    main()
