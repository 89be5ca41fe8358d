"""Where the one-shot time of a harmonic fill goes (host wall clock around every call of the C ABI).
Usage: python tools/setup_probe.py [size]"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload  # noqa: E402
from harmonic_interpolation_b200 import _capi as C  # noqa: E402
from harmonic_interpolation_b200 import solver as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
img8, keep = workload(n)
lib = C.lib()
import torch  # noqa: E402  (context warm-up, pinned buffers)
torch.zeros(1, device="cuda"); torch.cuda.synchronize()
inp = torch.empty((n, n, 3), dtype=torch.uint8).pin_memory().numpy(); inp[...] = img8
out = torch.empty((n, n, 3), dtype=torch.uint8).pin_memory().numpy()


def T(label, f):
    t = time.perf_counter(); r = f(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("%-34s %8.1f ms" % (label, 1e3 * dt), flush=True)
    return r


for rep in range(2):
    print("--- pass %d" % rep)
    flags = T("encode_flags (numpy)", lambda: S.encode_flags(n, n, hard=keep))
    s = T("GridSolver() = hg_create", lambda: S.GridSolver(n, n, 3, False, "fp32", "multigrid"))
    p8 = ctypes.POINTER(ctypes.c_uint8)
    T("hg_set_flags", lambda: C.check(lib.hg_set_flags(s._h, flags.ctypes.data_as(p8))))
    T("hg_set_soft_weights(NULL)", lambda: C.check(lib.hg_set_soft_weights(s._h, C.dptr(None), 0.0)))
    info = C.SetupInfo()
    T("hg_setup", lambda: C.check(lib.hg_setup(s._h, ctypes.byref(info))))
    print("   (device part of hg_setup: %.1f ms)" % info.setup_ms)
    for k in range(2):
        st = C.Stats()
        T("hg_solve_u8 #%d" % k, lambda: lib.hg_solve_u8(s._h, inp.ctypes.data_as(p8), out.ctypes.data_as(p8), 1e-6, 500, ctypes.byref(st)))
        print("   (upload %.1f ms, solve %.1f ms, download %.1f ms, %d iterations)" % (st.upload_ms, st.solve_ms, st.download_ms, st.iterations))
    T("close", s.close)
