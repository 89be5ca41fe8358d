"""One device-resident bilaplacian solve of the C2 pattern (hard 2 %, soft 5 %, cut cells) for timing and ncu.
Usage: python tools/profile_bilap.py [size] [precision] [rtol]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harmonic_interpolation_b200 import recovery as R  # noqa: E402
from harmonic_interpolation_b200 import solver as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "fp64"
rtol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
rng = np.random.default_rng(0)
hard = rng.random((n, n)) < 0.02
soft = (rng.random((n, n)) < 0.05) & ~hard
hv = rng.random((n, n, 1)); sv = rng.random((n, n, 1))
block = np.kron(rng.random((n // 16 + 1, n // 16 + 1)) < 0.3, np.ones((16, 16), bool))[:n, :n]
cd, cr = R.cut_planes_from_mask(block)
s = S.GridSolver(n, n, 1, True, prec)
info = s.set_pattern(S.encode_flags(n, n, hard=hard, soft=soft, cut_down=cd, cut_right=cr), None, 10.0, 0.0)
s.upload(hard_vals=hv, soft_vals=sv)
st = s.solve_device(rtol, 20000)
print("bilaplacian C2 %d^2 %s rtol %g: %d levels, %d iterations, relres %.2e, %.1f ms (%.2f ms / iteration), %d launches" % (
    n, prec, rtol, info.levels, st["iterations"], st["relres"], st["solve_ms"], st["solve_ms"] / max(st["iterations"], 1), st["launches"]))
print("stencil apply alone: %.3f ms" % s.bench_apply(10))
