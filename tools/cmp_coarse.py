"""Debug aid: one V-cycle with the tile legs, the streaming legs and the unfused reference kernels of the 9-point levels;
prints, level by level, the largest difference of b (made by the downward leg of the level above) and x (after the upward leg)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harmonic_interpolation_b200 import solver as S  # noqa: E402

rows, cols, K = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (70, 90, 2)
prec = sys.argv[4] if len(sys.argv) > 4 else "fp64"
rng = np.random.default_rng(5)
hard = rng.random((rows, cols)) < 0.04
if len(sys.argv) > 5:
    hard = np.kron(rng.random((rows // 64 + 1, cols // 64 + 1)) < 0.5, np.ones((64, 64), bool))[:rows, :cols]
hard[0, 0] = True
a = rng.normal(size=(rows, cols, K)); a[hard] = 0
res = {}
for name, env in (("tiles", "HGRID_COARSE_TILES"), ("unfused", "HGRID_UNFUSED_COARSE"), ("stream", None)):
    for e in ("HGRID_COARSE_TILES", "HGRID_UNFUSED_COARSE"):
        os.environ.pop(e, None)
    if env:
        os.environ[env] = "1"
    s = S.GridSolver(rows, cols, K, False, prec, "multigrid")
    s.set_pattern(S.encode_flags(rows, cols, hard=hard))
    z = s.precond_apply(a)
    res[name] = (z, [s.mg_level_vectors(l) for l in range(1, s.mg_levels())])
    s.close()
for other in ("unfused", "stream"):
    print("== %s vs tiles: z max diff %.3e (scale %.3e)" % (other, np.abs(res[other][0] - res["tiles"][0]).max(), np.abs(res["tiles"][0]).max()))
    for l, ((x0, b0), (x1, b1)) in enumerate(zip(res["tiles"][1], res[other][1]), start=1):
        db, dx = np.abs(b1 - b0), np.abs(x1 - x0)
        print("  level %d %s: b diff %.3e at %s (scale %.2e) | x diff %.3e at %s (scale %.2e)" % (
            l, x0.shape, db.max(), np.unravel_index(db.argmax(), db.shape), np.abs(b0).max(), dx.max(), np.unravel_index(dx.argmax(), dx.shape), np.abs(x0).max()))
