"""hg_set_flags + hg_setup of the headline workload, nothing else (for an ncu launch list of the set-up kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from harmonic_interpolation_b200 import solver as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
img8, keep = workload(n)
s = S.GridSolver(n, n, 3, False, "fp32", "multigrid")
info = s.set_pattern(S.encode_flags(n, n, hard=keep))
print("setup_ms", info.setup_ms, "levels", info.levels)
