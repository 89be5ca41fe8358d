"""Debug aid: V-cycle output of the streaming legs against the shared-memory tile legs on the same input."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from harmonic_interpolation_b200 import solver as S

def run(n, m, K, hard, prec, r, env):
    for k in ("HGRID_TILE_LEGS", "HGRID_UNFUSED_COARSE"):
        os.environ.pop(k, None)
    for k in env:
        os.environ[k] = "1"
    s = S.GridSolver(n, m, K, False, prec, "multigrid")
    s.set_pattern(S.encode_flags(n, m, hard=hard))
    z = s.precond_apply(r)
    s.close()
    return z

rng = np.random.default_rng(1)
for (n, m, case) in [(512, 512, "pin"), (200, 300, "iid"), (512, 512, "tiles")]:
    hard = np.zeros((n, m), bool)
    if case == "pin": hard[0, 0] = True
    elif case == "iid": hard = rng.random((n, m)) < 0.3
    else: hard = np.kron(rng.random((n // 64, m // 64)) < 0.5, np.ones((64, 64), bool)); hard[0, 0] = True
    K = 2
    r = rng.normal(size=(n, m, K)); r[hard] = 0
    for prec in ("fp64",):
        a = run(n, m, K, hard, prec, r, [])
        b = run(n, m, K, hard, prec, r, ["HGRID_TILE_LEGS"])
        d = np.abs(a - b).max(axis=2)
        print(case, prec, "max diff %.3e (scale %.3e)" % (d.max(), np.abs(b).max()))
        bad = np.argwhere(d > 1e-9 * np.abs(b).max())
        if len(bad):
            print("  bad pixels:", len(bad), "rows", bad[:, 0].min(), bad[:, 0].max(), "cols", bad[:, 1].min(), bad[:, 1].max())
            rows_hist = np.bincount(bad[:, 0], minlength=n); cols_hist = np.bincount(bad[:, 1], minlength=m)
            print("  worst rows:", np.argsort(-rows_hist)[:12], " worst cols:", np.argsort(-cols_hist)[:12])
            i, j = np.unravel_index(d.argmax(), d.shape)
            print("  argmax", i, j, a[i, j], b[i, j])
