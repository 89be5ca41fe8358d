"""CPU: the design claim behind mg25_clusters.cuh, on explicit matrices (tools/stiff_proto.py, scipy): with the level-0
interpolation averaged over the clusters of stiff gradient constraints, the multigrid-preconditioned CG no longer feels the
weight 1e5 -- and without it, it does."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_cluster_flat_interpolation_removes_the_stiffness():
    import stiff_proto as sp_
    from oracle import grid_oracle as go
    n = 64
    prob, _ = sp_.c4_problem(n, 1e5)
    A, b, free = go.reduced_system(prob, check=False)
    A = A.tocsr(); b = b[:, 0]
    fm = np.zeros(n * n, bool); fm[free] = True
    lf, ncl, size = sp_.clusters_of(prob, free)
    assert ncl > 50 and size.max() <= 8
    cl = dict(label_full=lf, blocks=[], invs=[])
    _, plain = sp_.pcg(A, b, sp_.MG(A, fm, n, n).precond, 1e-8, 400)
    _, flat = sp_.pcg(A, b, sp_.MG(A, fm, n, n, clusters=cl, cluster_P=True).precond, 1e-8, 400)
    assert flat <= 30 and plain >= 5 * flat, (plain, flat)
