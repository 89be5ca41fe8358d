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


def main():
    P = ref_shim.load_reference()["poisson"]
    store = {}
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
