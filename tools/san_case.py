"""Small solves for `compute-sanitizer --tool memcheck python tools/san_case.py` (one GPU): the fp32 fill (streaming legs,
streaming stencil and the fp64 true-residual kernel) with free pixels in the first and last rows and columns, a soft /
cut case (general chunks, tile stencil), a bilaplacian case."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harmonic_interpolation_b200 import solver as S  # noqa: E402

rng = np.random.default_rng(3)
for rows, cols, K in ((300, 260, 3), (517, 1030, 1)):
    keep = np.kron(rng.random((rows // 32 + 1, cols // 32 + 1)) < 0.5, np.ones((32, 32), bool))[:rows, :cols]
    keep[0, :] = False; keep[-1, :] = False; keep[:, 0] = False; keep[:, -1] = False
    keep[5, 5] = True
    img = rng.random((rows, cols, K))
    for prec in ("fp32", "fp64"):
        s = S.GridSolver(rows, cols, K, False, prec, "multigrid")
        s.set_pattern(S.encode_flags(rows, cols, hard=keep))
        out = s.solve(hard_vals=img)
        print("fill", rows, cols, K, prec, s.last_stats["iterations"], s.last_stats["relres"], np.isfinite(out).all(), flush=True)
        s.close()
rows, cols, K = 200, 333, 2
hard = rng.random((rows, cols)) < 0.03
soft = (rng.random((rows, cols)) < 0.1) & ~hard
cd = rng.random((rows, cols)) < 0.05
cr = rng.random((rows, cols)) < 0.05
for bil in (False, True):
    s = S.GridSolver(rows, cols, K, bil, "fp64", "multigrid")
    s.set_pattern(S.encode_flags(rows, cols, hard=hard, soft=soft, cut_down=cd, cut_right=cr), np.where(soft, 3.0, 0.0))
    out = s.solve(hard_vals=rng.random((rows, cols, K)), soft_vals=rng.random((rows, cols, K)), maxiter=3000, allow_unconverged=True)
    print("soft/cuts bilap=%s" % bil, s.last_stats["iterations"], s.last_stats["relres"], np.isfinite(out).all(), flush=True)
    s.close()
print("SAN_CASE_DONE")
