"""Row-slab decomposition of one grid over the GPUs of a box, one process per GPU
(torch.distributed is the plumbing: rendezvous, the NCCL id broadcast, the exchange of the
boundary flag bands at setup; the solve itself talks NCCL from inside libhgrid.so).

The global rows are cut at multiples of GHOST_ROWS (128).  Every rank builds its local solver on
the band it owns plus a GHOST band of 128 rows towards each neighbouring slab.  Ghost rows mirror
the neighbour: same flags and soft weights, values refreshed by the library during the solve.
All kernels run unchanged on the local grid; what they compute on ghost rows is redundant with
the neighbour, so a whole smoothing sweep needs no communication and the V-cycle is the
single-GPU V-cycle (iteration counts do not depend on the number of slabs).  Dot products run
over owned rows and are all-reduced.  Laplacian only.
"""
import ctypes

import numpy as np

from . import _capi as C
from .solver import GridSolver

GHOST_ROWS = 128


def partition_rows(rows, world, align=GHOST_ROWS):
    """Contiguous bands cut at multiples of `align`: list of (row0, row1) per rank.  Every band has
    at least 2*align rows; the last one takes the remainder."""
    assert world >= 1
    if world == 1:
        return [(0, rows)]
    nblocks = rows // align
    assert nblocks >= 2 * world, "grid too small for %d slabs (needs >= %d rows)" % (world, 2 * align * world)
    base, extra = divmod(nblocks, world)
    out, r = [], 0
    for k in range(world):
        n = (base + (1 if k < extra else 0)) * align
        r1 = rows if k == world - 1 else r + n
        out.append((r, r1))
        r = r1
    return out


def exchange_bands(first_band, last_band, rank, world, dist, device=None):
    """All-gathers every rank's first and last owned band (GHOST_ROWS rows each); returns
    (band_above, band_below) for this rank (None at the ends).  gloo (CPU) or nccl (GPU) back end."""
    import torch
    mine = torch.from_numpy(np.ascontiguousarray(np.stack([first_band, last_band])))
    if device is not None:
        mine = mine.to(device)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine)
    above = bufs[rank - 1][1].cpu().numpy() if rank > 0 else None
    below = bufs[rank + 1][0].cpu().numpy() if rank < world - 1 else None
    return above, below


def with_ghost_bands(owned, above, below):
    parts = [p for p in (above, owned, below) if p is not None]
    if len(parts) == 1:
        return np.ascontiguousarray(owned)           # (no copy: a 16384^2 flag plane is a quarter of a GB)
    return np.ascontiguousarray(np.concatenate(parts, axis=0))


class SlabSolver:
    """The rank-local part of a distributed harmonic solve.

        s = SlabSolver(global_rows, cols, K, flags_owned, dist, precision="fp32")
        x_owned = s.solve(hard_vals_owned, ...)       # (owned rows, cols, K) float64
    """

    def __init__(self, global_rows, cols, channels, flags_owned, dist=None, precision=None, soft_weights_owned=None,
                 soft_scalar=0.0, w_g=0.0, device=None):
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        parts = partition_rows(global_rows, self.world)
        self.row0, self.row1 = parts[self.rank]
        flags_owned = np.ascontiguousarray(flags_owned, np.uint8)
        assert flags_owned.shape == (self.row1 - self.row0, cols)
        G = GHOST_ROWS
        above = below = sw_above = sw_below = None
        if self.world > 1:
            above, below = exchange_bands(flags_owned[:G], flags_owned[-G:], self.rank, self.world, dist, device)
            if soft_weights_owned is not None:
                sw = np.ascontiguousarray(soft_weights_owned, np.float64)
                sw_above, sw_below = exchange_bands(sw[:G], sw[-G:], self.rank, self.world, dist, device)
        self.gt = 0 if above is None else G
        self.gb = 0 if below is None else G
        flags = with_ghost_bands(flags_owned, above, below)
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
            starts = np.ascontiguousarray([p[0] for p in parts], np.int32)
            C.check(C.lib().hg_comm_init(self.solver._h, ident.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                         self.rank, self.world, self.gt, self.gb,
                                         starts.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), int(global_rows)))
        sw_local = None
        if soft_weights_owned is not None:
            sw_local = with_ghost_bands(np.asarray(soft_weights_owned, np.float64), sw_above, sw_below)
        self.info = self.solver.set_pattern(flags, sw_local, soft_scalar, w_g)

    def _pad(self, a):
        """(owned, cols, K) -> (local rows, cols, K); ghost rows are left 0 (the library fills them)."""
        if a is None:
            return None
        a = np.asarray(a, np.float64)
        if a.ndim == 2:
            a = a[:, :, None]
        if self.gt == 0 and self.gb == 0:
            return a
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

    # ---- image form: bytes both ways, caller-provided (pinned) local buffers -----------------------
    def pinned_u8(self):
        """A page-locked uint8 (local rows, cols, K) buffer and its view of the owned rows."""
        import torch
        t = torch.empty((self.local_rows, self.cols, self.K), dtype=torch.uint8).pin_memory()
        a = t.numpy()
        a[...] = 0
        self._keep = getattr(self, "_keep", []) + [t]
        return a, self._owned(a)

    def solve_u8(self, inp_local, out_local, rtol=None, maxiter=None):
        """hg_solve_u8 on local-sized buffers (ghost rows of the input are ignored: the library fills
        them); returns the owned rows of `out_local`."""
        self.solver.solve_u8(inp_local, out_local, rtol=rtol, maxiter=maxiter)
        self.last_stats = self.solver.last_stats
        return self._owned(out_local)
