"""
TEST INFRASTRUCTURE ONLY.  tests/golden/reference_highlighting.npz: outputs of the UNMODIFIED reference's two tint functions
(highlighting.py:139-213, 361-464; each solves a small cropped (bi)laplacian problem through recovery.solve_grid_linear) on
seeded inputs.  Run in the build container only:

    python oracle/make_golden_highlighting.py

Two extra shims on top of oracle/ref_shim.py.  (1) `from numpy import *` (highlighting.py:1) shadows the builtins `max` / `min`
under numpy 2 (they joined numpy.__all__), so `max( 0., min( 1., alpha ) )` at :160 / :384 calls numpy.min( 1., axis = alpha );
the builtins are put back into the module's namespace.  (2) highlighting.py:184 computes `rim = r - rinner` on boolean arrays, which numpy
removed.  The reference's own `grow` is wrapped to return a bool ndarray SUBCLASS whose reflected subtraction is the set
difference (a AND NOT b) -- what boolean subtraction meant when the reference was written -- so that the function runs
unmodified.  `array_tinting_region_border_with_falloff` needs no such help.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class BoolDiff(np.ndarray):
    def __sub__(self, other):
        return np.logical_and(np.asarray(self), np.logical_not(np.asarray(other)))

    def __rsub__(self, other):
        return np.logical_and(np.asarray(other), np.logical_not(np.asarray(self)))


def cases():
    rng = np.random.default_rng(20261020)
    out = []
    # (name, image shape, list of boolean region masks)
    ii, jj = np.mgrid[0:48, 0:60]
    disk = (ii - 20) ** 2 + (jj - 25) ** 2 <= 11 ** 2
    blob = disk | ((ii >= 18) & (ii < 30) & (jj >= 30) & (jj < 50))
    edge = np.zeros((48, 60), bool); edge[0:14, 40:60] = True; edge[10:20, 52:60] = True          # reaches the image border
    thin = np.zeros((48, 60), bool); thin[30:32, 5:40] = True                                        # two pixels wide
    left = np.zeros((40, 44), bool); left[6:34, 4:22] = True
    right = np.zeros((40, 44), bool); right[10:30, 22:40] = True                                     # abuts `left` along a column
    tri_a = np.zeros((50, 50), bool); tri_b = np.zeros((50, 50), bool); tri_c = np.zeros((50, 50), bool)
    tri_a[5:25, 5:45] = True; tri_b[25:45, 5:25] = True; tri_c[25:45, 25:45] = True                  # three mutually abutting regions
    for name, masks, kind in (("center_blob", [blob], "center"), ("center_edge", [edge], "center"), ("center_thin", [thin], "center"),
                              ("border_pair", [left, right], "border"), ("border_three", [tri_a, tri_b, tri_c], "border")):
        img = rng.integers(0, 256, size=masks[0].shape + (3,), dtype=np.uint8)
        colors = rng.integers(0, 256, size=(len(masks), 3)).astype(np.int64)
        alpha = float(rng.uniform(0.4, 1.0))
        out.append((name, kind, img, masks, colors, alpha))
    return out


def main():
    ref = ref_shim.load_reference()
    H = ref["highlighting"]
    sys.modules["recovery"] = ref["recovery"]          # the tint functions do `from recovery import solve_grid_linear` at call time
    import builtins
    H.max, H.min = builtins.max, builtins.min
    orig_grow = H.grow
    H.grow = lambda mask, n: orig_grow(mask, n).view(BoolDiff)
    store = {}
    try:
        for name, kind, img, masks, colors, alpha in cases():
            regions = [np.nonzero(m) for m in masks]
            with contextlib.redirect_stdout(io.StringIO()):
                if kind == "center":
                    res = H.array_tinting_region_with_falloff_from_center(img, regions[0], tuple(int(c) for c in colors[0]), alpha)
                else:
                    res = H.array_tinting_region_border_with_falloff(img, regions, [tuple(int(c) for c in col) for col in colors], alpha)
            store[name + "_kind"] = np.array(kind)
            store[name + "_img"] = img
            store[name + "_masks"] = np.stack(masks)
            store[name + "_colors"] = colors
            store[name + "_alpha"] = np.array(alpha)
            store[name + "_out"] = np.asarray(res)
            print(name, kind, img.shape, "changed pixels:", int((np.asarray(res) != img).any(axis=2).sum()))
    finally:
        H.grow = orig_grow
        sys.modules.pop("recovery", None)
    np.savez_compressed(os.path.join(GOLD, "reference_highlighting.npz"), **store)


if __name__ == "__main__":
    main()
