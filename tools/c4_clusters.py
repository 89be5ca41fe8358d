"""A/B of the cluster-flat level-0 transfers (HGRID_MG25_CLUSTERS, mg25_clusters.cuh) on the C4 pattern: normals on a
Bernoulli(0.02) subset -> dx / dy constraints of weight 1e5 on top of the bilaplacian, one hard pin.  Prints iterations, time
and the agreement of the two answers.  usage: python tools/c4_clusters.py [n ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harmonic_interpolation_b200 import solver as S  # noqa: E402


def problem(n, dens=0.02, seed=0):
    rng = np.random.default_rng(seed)
    ii, jj = np.indices((n, n), dtype=np.float64)
    z = 0.05 * n * np.sin(2 * np.pi * ii / n) * np.cos(2 * np.pi * jj / n)
    sel = np.zeros((n, n), bool)
    sel[:-1, :-1] = rng.random((n - 1, n - 1)) < dens
    dd = np.zeros((n, n, 1)); dr = np.zeros((n, n, 1))
    dd[:-1, :, 0] = z[1:] - z[:-1]; dr[:, :-1, 0] = z[:, 1:] - z[:, :-1]
    dd[~sel] = 0; dr[~sel] = 0
    hard = np.zeros((n, n), bool); hard[0, 0] = True
    hv = np.zeros((n, n, 1)); hv[0, 0, 0] = z[0, 0]
    return S.encode_flags(n, n, hard=hard, pen_down=sel, pen_right=sel), hv, dd, dr, z


def run(n, clusters, dens=0.02, rtol=1e-8):
    os.environ["HGRID_MG25_CLUSTERS"] = "1" if clusters else "0"
    flags, hv, dd, dr, z = problem(n, dens)
    s = S.GridSolver(n, n, 1, True, "fp64")
    t = time.perf_counter()
    info = s.set_pattern(flags, None, 1e5, 1e5)
    t_setup = time.perf_counter() - t
    s.upload(hard_vals=hv, d_down=dd, d_right=dr)
    st = s.solve_device(rtol, 20000)
    x, _ = s.download()
    s.close()
    print("n=%d dens=%.2f clusters=%d: %d iterations, relres %.2e, solve %.1f ms, set-up %.2f s, |x - z|max %.3e" % (
        n, dens, int(clusters), st["iterations"], st["relres"], st["solve_ms"], t_setup, np.abs(x[:, :, 0] - z).max()), flush=True)
    return x


if __name__ == "__main__":
    only = "--only-clusters" in sys.argv
    sizes = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [256, 512, 1024]
    for n in sizes:
        b = run(n, True)
        if not only:
            a = run(n, False)
            print("   max |difference| / range = %.2e" % (np.abs(a - b).max() / np.ptp(a)), flush=True)
