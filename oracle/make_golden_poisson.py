"""
TEST INFRASTRUCTURE ONLY.  Golden vectors for `poisson.solve_poisson_simple`: runs the UNMODIFIED reference
(/root/reference/poisson.py through oracle/ref_shim.py) on seeded inputs and stores inputs + outputs in
tests/golden/reference_poisson.npz.  Build container only:   python oracle/make_golden_poisson.py
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


def cases():
    rng = np.random.default_rng(7)
    out = []
    for name, rows, cols, K, with_mask, with_value in (("plain", 24, 31, 1, False, False), ("rgb", 17, 20, 3, False, False),
                                                       ("masked", 30, 26, 2, True, True), ("island", 22, 22, 1, True, True)):
        u = rng.normal(size=(rows, cols, 2, K))
        mask = None
        mask_value = None
        if with_mask:
            mask = np.ones((rows, cols), bool)
            mask[rows // 3: rows // 2, cols // 4: cols // 2] = False
            if name == "island":
                mask[2:5, 2:5] = False
            mask_value = rng.random(K) if with_value else None
        n = 6
        ij = np.stack([rng.integers(0, rows, n), rng.integers(0, cols, n)], axis=1)
        if mask is not None:
            ij = ij[mask[ij[:, 0], ij[:, 1]]]
        ij = np.unique(ij, axis=0)
        coeff = rng.uniform(0.5, 2.0, len(ij))
        vals = rng.random((len(ij), K))
        w = float(rng.uniform(1.0, 50.0))
        out.append(dict(name=name, u=u, mask=mask, mask_value=mask_value, ij=ij, coeff=coeff, vals=vals, w=w))
    return out


def cases2():
    """Round 2: the (G.T G)^2 form (poisson.py:65-67), constraint rows over several pixels (poisson.py:329-366), constraints
    on masked-out pixels (the pattern of the reference's own test, poisson.py:404-447).  A constraint row is stored as
    (row index, i, j, coefficient) entries + one right-hand side per row."""
    rng = np.random.default_rng(11)
    out = []
    for name, rows, cols, K, bilap, multi, kind in (("bilap_plain", 26, 33, 1, True, False, None), ("bilap_masked", 30, 26, 2, True, False, "value"),
                                                    ("multi_lap", 21, 28, 2, False, True, None), ("multi_bilap", 24, 24, 1, True, True, "value"),
                                                    ("multi_free_mask", 20, 23, 1, False, True, "free"),
                                                    ("selftest_small", 36, 36, 1, False, False, "selftest"), ("selftest_bilap", 36, 36, 1, True, False, "selftest"),
                                                    ("row_bilap", 1, 9, 1, True, False, None), ("tiny_multi", 3, 2, 2, True, True, None)):
        u = rng.normal(size=(rows, cols, 2, K))
        mask = mask_value = None
        ent, vals = [], []
        if kind == "selftest":
            # poisson.py:404-447 at 36 x 36: the box is the domain, the outside is pinned to 0, constraints at the four outer
            # corners (masked-out pixels) and the four corners of the box
            br0, br1 = 14, 22
            mask = np.zeros((rows, cols), bool); mask[br0:br1, br0:br1] = True
            mask_value = np.zeros(K)
            u = np.zeros((rows, cols, 2, K))
            pts = [(0, 0, 2.0 * br0), (rows - 1, 0, 0.0), (0, cols - 1, 0.0), (rows - 1, cols - 1, 0.0),
                   (br0, br0, br0), (br1 - 1, br0, br0), (br0, br1 - 1, br0), (br1 - 1, br1 - 1, br0)]
            for n, (i, j, v) in enumerate(pts):
                ent.append((n, i, j, 1.0)); vals.append([float(v)] * K)
            w = 1e5
        else:
            if kind is not None:
                mask = np.ones((rows, cols), bool)
                mask[rows // 3: rows // 2, cols // 4: cols // 2] = False
                mask_value = rng.random(K) if kind == "value" else None
            ok = np.ones((rows, cols), bool) if (mask is None or kind == "free") else mask
            cand = np.argwhere(ok)
            n = 0
            for _ in range(5):          # one-pixel rows
                i, j = cand[rng.integers(len(cand))]
                ent.append((n, int(i), int(j), float(rng.uniform(0.5, 2.0)))); vals.append(rng.random(K)); n += 1
            if kind == "free":           # the masked-out block is a component of its own and needs a value
                ent.append((n, rows // 3 + 1, cols // 4 + 1, 1.0)); vals.append(rng.random(K)); n += 1
            if multi:
                for m in (2, 3, 2):     # rows over several pixels: differences and weighted means
                    pts = cand[rng.choice(len(cand), m, replace=False)]
                    cf = rng.uniform(-1.5, 1.5, m)
                    for (i, j), c in zip(pts, cf):
                        ent.append((n, int(i), int(j), float(c)))
                    vals.append(rng.normal(size=K)); n += 1
            w = float(rng.uniform(1.0, 50.0))
        out.append(dict(name=name, u=u, mask=mask, mask_value=mask_value, ent=np.array(ent, float), vals=np.array(vals, float), w=w, bilap=bilap))
    return out


def rows_of(ent, vals):
    lc = []
    for n in range(len(vals)):
        lc.append(([(int(i), int(j), float(c)) for r, i, j, c in ent if int(r) == n], vals[n]))
    return lc


def main():
    P = ref_shim.load_reference()["poisson"]
    store = {}
    for c in cases2():
        with contextlib.redirect_stdout(io.StringIO()):
            x = P.solve_poisson_simple(c["u"], rows_of(c["ent"], c["vals"]), c["w"], mask=c["mask"], mask_value=c["mask_value"], bilaplacian=c["bilap"])
        n = c["name"]
        store[n + "_u"] = c["u"]; store[n + "_ent"] = c["ent"]; store[n + "_vals"] = c["vals"]; store[n + "_w"] = np.float64(c["w"])
        store[n + "_bilap"] = np.bool_(c["bilap"])
        if c["mask"] is not None:
            store[n + "_mask"] = c["mask"]
        if c["mask_value"] is not None:
            store[n + "_mask_value"] = c["mask_value"]
        store[n + "_out"] = np.asarray(x, np.float64).reshape(c["u"].shape[0], c["u"].shape[1], -1)
    for c in cases():
        lc = [([(int(i), int(j), float(cf))], v) for (i, j), cf, v in zip(c["ij"], c["coeff"], c["vals"])]
        with contextlib.redirect_stdout(io.StringIO()):
            x = P.solve_poisson_simple(c["u"], lc, c["w"], mask=c["mask"], mask_value=c["mask_value"])
        n = c["name"]
        store[n + "_u"] = c["u"]; store[n + "_ij"] = c["ij"]; store[n + "_coeff"] = c["coeff"]; store[n + "_vals"] = c["vals"]
        store[n + "_w"] = np.float64(c["w"])
        if c["mask"] is not None:
            store[n + "_mask"] = c["mask"]
            store[n + "_mask_value"] = c["mask_value"]
        store[n + "_out"] = np.asarray(x, np.float64)
    path = os.path.join(ROOT, "tests", "golden", "reference_poisson.npz")
    np.savez_compressed(path, **store)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
