"""The seeded cases of tests/golden/reference_highlighting.npz (outputs of the unmodified reference, oracle/make_golden_highlighting.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_highlighting.npz")
NAMES = ["center_blob", "center_edge", "center_thin", "border_pair", "border_three"]


def load(name):
    g = np.load(GOLDEN)
    masks = g[name + "_masks"]
    return dict(kind=str(g[name + "_kind"]), img=g[name + "_img"], regions=[np.nonzero(m) for m in masks],
                colors=[tuple(int(c) for c in col) for col in g[name + "_colors"]], alpha=float(g[name + "_alpha"]), out=g[name + "_out"])


def run(H, case):
    if case["kind"] == "center":
        return H.array_tinting_region_with_falloff_from_center(case["img"], case["regions"][0], case["colors"][0], case["alpha"])
    return H.array_tinting_region_border_with_falloff(case["img"], case["regions"], case["colors"], case["alpha"])
