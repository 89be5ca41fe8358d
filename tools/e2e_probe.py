import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from bench import workload
from harmonic_interpolation_b200 import solver as S
n = 16384
img8, keep = workload(n)
s = S.GridSolver(n, n, 3, False, "fp32", "multigrid")
s.set_pattern(S.encode_flags(n, n, hard=keep))
tin = torch.empty((n, n, 3), dtype=torch.uint8).pin_memory(); tout = torch.empty((n, n, 3), dtype=torch.uint8).pin_memory()
a = tin.numpy(); a[...] = img8; o = tout.numpy()
for i in range(4):
    t = time.perf_counter(); s.solve_u8(a, o, rtol=1e-6, maxiter=500); dt = time.perf_counter() - t
    st = s.last_stats
    print("e2e %.1f ms: upload %.1f solve %.1f download %.1f  (sum %.1f)" % (dt * 1e3, st["upload_ms"], st["solve_ms"], st["download_ms"], st["upload_ms"] + st["solve_ms"] + st["download_ms"]))
