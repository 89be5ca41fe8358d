"""Iteration counts of the bilaplacian solve: multigrid (mg25.cuh) against Jacobi, C4-like problems."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from harmonic_interpolation_b200 import solver as S

def problem(n, seed=0, w_g=1e5, dens=0.02):
    rng = np.random.default_rng(seed)
    ii, jj = np.indices((n, n))
    z = 10.0 * np.sin(2 * np.pi * ii / n) * np.cos(2 * np.pi * jj / n)
    pen = rng.random((n, n)) < dens
    pd = pen.copy(); pd[-1, :] = False
    pr = pen.copy(); pr[:, -1] = False
    dd = np.zeros((n, n, 1)); dr = np.zeros((n, n, 1))
    dd[:-1, :, 0] = z[1:] - z[:-1]; dr[:, :-1, 0] = z[:, 1:] - z[:, :-1]
    hard = np.zeros((n, n), bool); hard[0, 0] = True
    hv = np.zeros((n, n, 1)); hv[0, 0, 0] = z[0, 0]
    return hard, hv, pd, pr, dd, dr, z

WG = float(os.environ.get("WG", 1e5))
for n in [int(a) for a in sys.argv[1:]] or [128, 256, 512]:
    hard, hv, pd, pr, dd, dr, z = problem(n)
    for pre in ("multigrid", "jacobi"):
        if pre == "jacobi" and (n > 512 or os.environ.get("NOJAC")): continue
        s = S.GridSolver(n, n, 1, True, "fp64", pre)
        t = time.perf_counter()
        s.set_pattern(S.encode_flags(n, n, hard=hard, pen_down=pd, pen_right=pr), None, 0.0, WG)
        ts = time.perf_counter() - t
        try:
            x = s.solve(hv, None, dd, dr, rtol=1e-10, maxiter=200000, allow_unconverged=True)
            st = s.last_stats
            print("n=%d %-9s setup %.2fs  iterations %6d  relres %.2e  solve %.1f ms  max|x - z| = %.3e" % (
                n, pre, ts, st["iterations"], st["relres"], st["solve_ms"], np.abs(x[:, :, 0] - z).max()), flush=True)
        except ArithmeticError as e:
            print("n=%d %s: %s" % (n, pre, e), s.last_stats, flush=True)
        s.close()
