"""Row-slab decomposition of one grid over the GPUs of a box, one process per GPU
(torch.distributed is the plumbing: rendezvous, the NCCL id broadcast, the exchange of the
boundary flag rows at setup; the solve itself talks NCCL from inside libhgrid.so).

Every rank owns a contiguous band of rows.  Its local solver is built on the band plus one GHOST
row towards each neighbouring slab.  Ghost rows are flagged hard, so for the local solver they are
just more Dirichlet pixels: the operator rows of the owned pixels are exactly the rows of the
global operator (the library refreshes the ghost values from the neighbour before every operator
application), the preconditioner is the slab's own V-cycle (block-Jacobi over slabs), and the
K conjugate-gradient scalars are all-reduced.  Laplacian only.
"""
import ctypes

import numpy as np

from . import _capi as C
from .solver import GridSolver


def partition_rows(rows, world):
    """Balanced contiguous bands: list of (row0, row1) per rank."""
    assert rows >= world >= 1
    base, extra = divmod(rows, world)
    out, r = [], 0
    for k in range(world):
        n = base + (1 if k < extra else 0)
        out.append((r, r + n))
        r += n
    return out


def local_flags_with_ghosts(flags_owned, row_above, row_below):
    """Stacks [ghost row above | owned rows | ghost row below].  A ghost row keeps the neighbour's
    edge bits (its CUT_DOWN / PEN_DOWN bits describe the edge into our first row) and is made hard."""
    parts = []
    if row_above is not None:
        parts.append((np.asarray(row_above, np.uint8) | np.uint8(C.FLAG_HARD))[None, :] & np.uint8(~C.FLAG_SOFT & 0xFF))
    parts.append(np.asarray(flags_owned, np.uint8))
    if row_below is not None:
        parts.append((np.asarray(row_below, np.uint8) | np.uint8(C.FLAG_HARD))[None, :] & np.uint8(~C.FLAG_SOFT & 0xFF))
    return np.ascontiguousarray(np.concatenate(parts, axis=0))


def exchange_boundary_rows(first_row, last_row, rank, world, dist, device=None):
    """All-gathers every rank's first and last owned flag row; returns (row_above, row_below) for
    this rank (None at the ends).  Works on the gloo (CPU) and nccl (GPU) back ends."""
    import torch
    cols = len(first_row)
    mine = torch.from_numpy(np.stack([first_row, last_row]).astype(np.uint8))
    if device is not None:
        mine = mine.to(device)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine)
    above = bufs[rank - 1][1].cpu().numpy() if rank > 0 else None
    below = bufs[rank + 1][0].cpu().numpy() if rank < world - 1 else None
    assert above is None or len(above) == cols
    return above, below


class SlabSolver:
    """The rank-local part of a distributed harmonic solve.

        s = SlabSolver(global_rows, cols, K, flags_owned, dist, precision="fp32")
        x_owned = s.solve(hard_vals_owned, ...)       # (owned rows, cols, K) float64
    """

    def __init__(self, global_rows, cols, channels, flags_owned, dist=None, precision=None, soft_weights_owned=None,
                 soft_scalar=0.0, w_g=0.0, device=None):
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.row0, self.row1 = partition_rows(global_rows, self.world)[self.rank]
        flags_owned = np.ascontiguousarray(flags_owned, np.uint8)
        assert flags_owned.shape == (self.row1 - self.row0, cols)
        above = below = None
        if self.world > 1:
            above, below = exchange_boundary_rows(flags_owned[0], flags_owned[-1], self.rank, self.world, dist, device)
        self.gt, self.gb = int(above is not None), int(below is not None)
        flags = local_flags_with_ghosts(flags_owned, above, below)
        self.local_rows = flags.shape[0]
        self.cols, self.K = cols, channels
        self.solver = GridSolver(self.local_rows, cols, channels, False, precision, "multigrid")
        if self.world > 1:
            import torch
            ident = np.zeros(128, np.uint8)
            if self.rank == 0:
                C.check(C.lib().hg_comm_unique_id(ident.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
            t = torch.from_numpy(ident)
            if device is not None:
                t = t.to(device)
            dist.broadcast(t, 0)
            ident = np.ascontiguousarray(t.cpu().numpy())
            C.check(C.lib().hg_comm_init(self.solver._h, ident.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                         self.rank, self.world, self.gt, self.gb))
        sw = None
        if soft_weights_owned is not None:
            sw = self._pad(np.asarray(soft_weights_owned, float)[:, :, None])[:, :, 0]
        self.info = self.solver.set_pattern(flags, sw, soft_scalar, w_g)

    def _pad(self, a):
        """(owned, cols, K) -> (local rows, cols, K) with zero ghost rows."""
        if a is None:
            return None
        a = np.asarray(a, np.float64)
        if a.ndim == 2:
            a = a[:, :, None]
        out = np.zeros((self.local_rows, self.cols, a.shape[2]))
        out[self.gt:self.gt + a.shape[0]] = a
        return out

    def _owned(self, a):
        return a[self.gt:self.local_rows - self.gb]

    def upload(self, hard_vals=None, soft_vals=None, d_down=None, d_right=None):
        return self.solver.upload(self._pad(hard_vals), self._pad(soft_vals), self._pad(d_down), self._pad(d_right))

    def solve_device(self, rtol, maxiter):
        return self.solver.solve_device(rtol, maxiter)

    def reset_guess(self):
        self.solver.reset_guess()

    def download(self):
        out, st = self.solver.download()
        return self._owned(out), st

    def solve(self, hard_vals=None, soft_vals=None, d_down=None, d_right=None, rtol=None, maxiter=None):
        out = self.solver.solve(self._pad(hard_vals), self._pad(soft_vals), self._pad(d_down), self._pad(d_right),
                                rtol=rtol, maxiter=maxiter)
        self.last_stats = self.solver.last_stats
        return self._owned(out)
