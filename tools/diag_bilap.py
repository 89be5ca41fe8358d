import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, io, contextlib
import golden_cases as gc
from harmonic_interpolation_b200 import recovery as R, solver as S
g = dict(np.load("tests/golden/reference_outputs.npz"))
for key, kw in gc.g_cases(g):
    if "valuesonly" not in key and "wdef" not in key: continue
    for rtol in (1e-13, 1e-15):
        with contextlib.redirect_stdout(io.StringIO()):
            try:
                out = R.solve_grid_linear(rtol=rtol, maxiter=3000000, **kw)
                err = np.abs(out - g[key]).max() / np.ptp(g[key])
            except ArithmeticError as e:
                err = str(e)[:80]
        print(key, rtol, err)
