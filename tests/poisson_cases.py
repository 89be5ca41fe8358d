"""The stored inputs of tests/golden/reference_poisson.npz (written by oracle/make_golden_poisson.py from the
unmodified reference) as keyword arguments of solve_poisson_simple."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_poisson.npz")
NAMES = ("plain", "rgb", "masked", "island")
# round 2: (G.T G)^2, constraint rows over several pixels, constraints on masked-out pixels (oracle/make_golden_poisson.py: cases2)
NAMES2 = ("bilap_plain", "bilap_masked", "multi_lap", "multi_bilap", "multi_free_mask", "selftest_small", "selftest_bilap",
          "row_bilap", "tiny_multi")


def load(name):
    z = np.load(GOLD)
    lc = [([(int(i), int(j), float(c))], v) for (i, j), c, v in zip(z[name + "_ij"], z[name + "_coeff"], z[name + "_vals"])]
    kw = dict(target_gradients=z[name + "_u"], linear_constraints=lc, linear_constraints_weight=float(z[name + "_w"]))
    if name + "_mask" in z.files:
        kw["mask"] = z[name + "_mask"]
        kw["mask_value"] = z[name + "_mask_value"]
    return kw, z[name + "_out"]


def load2(name):
    z = np.load(GOLD)
    ent, vals = z[name + "_ent"], z[name + "_vals"]
    lc = [([(int(i), int(j), float(c)) for r, i, j, c in ent if int(r) == n], vals[n]) for n in range(len(vals))]
    kw = dict(target_gradients=z[name + "_u"], linear_constraints=lc, linear_constraints_weight=float(z[name + "_w"]),
              bilaplacian=bool(z[name + "_bilap"]))
    if name + "_mask" in z.files:
        kw["mask"] = z[name + "_mask"]
    if name + "_mask_value" in z.files:
        kw["mask_value"] = z[name + "_mask_value"]
    return kw, z[name + "_out"]
