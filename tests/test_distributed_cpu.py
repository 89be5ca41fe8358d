"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU slab mode -- row partition,
ghost-row flag assembly through torch.distributed -- without touching a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from harmonic_interpolation_b200 import _capi as C
from harmonic_interpolation_b200.distributed import exchange_boundary_rows, local_flags_with_ghosts, partition_rows


def test_partition_rows():
    for rows, world in [(10, 1), (10, 3), (16384, 8), (7, 7), (33, 4)]:
        parts = partition_rows(rows, world)
        assert parts[0][0] == 0 and parts[-1][1] == rows
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, rows, cols, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    flags = rng.integers(0, 64, size=(rows, cols), dtype=np.uint8)
    r0, r1 = partition_rows(rows, world)[rank]
    above, below = exchange_boundary_rows(flags[r0], flags[r1 - 1], rank, world, dist)
    local = local_flags_with_ghosts(flags[r0:r1], above, below)
    want = flags[max(r0 - 1, 0):min(r1 + 1, rows)].copy()
    ghost = np.uint8(C.FLAG_HARD)
    if rank > 0:
        want[0] = (want[0] | ghost) & np.uint8(~C.FLAG_SOFT & 0xFF)
    if rank < world - 1:
        want[-1] = (want[-1] | ghost) & np.uint8(~C.FLAG_SOFT & 0xFF)
    ret[rank] = bool(local.shape == want.shape and (local == want).all())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ghost_rows_over_gloo(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 41, 23, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(world))
