"""One warm + one measured device-resident solve of the headline workload; the command that is
run plain and then under ncu (profiles/).  Usage: python tools/profile_step.py [size] [solves]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bench import workload
from harmonic_interpolation_b200 import solver as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
solves = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
img8, keep = workload(n)
s = S.GridSolver(n, n, 3, False, prec, "multigrid")
s.set_pattern(S.encode_flags(n, n, hard=keep))
s.upload(hard_vals=img8 / 255.0)
for i in range(solves):
    s.reset_guess()
    st = s.solve_device(1e-6, 500)
    print("solve %d: %.2f ms, %d iterations, relres %.2e, %d launches" % (i, st["solve_ms"], st["iterations"], st["relres"], st["launches"]))
print("apply: %.3f ms" % s.bench_apply(10))
