"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU slab mode -- row partition,
ghost-row flag assembly through torch.distributed -- without touching a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from harmonic_interpolation_b200.distributed import GHOST_ROWS, exchange_bands, partition_rows, with_ghost_bands


def test_partition_rows():
    assert partition_rows(100, 1) == [(0, 100)]
    for rows, world in [(1024, 2), (16384, 8), (32768, 8), (1000, 3), (2049, 4)]:
        parts = partition_rows(rows, world)
        assert parts[0][0] == 0 and parts[-1][1] == rows
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(a[0] % GHOST_ROWS == 0 for a in parts)
        assert all(b - a >= 2 * GHOST_ROWS for a, b in parts)
    with pytest.raises(AssertionError):
        partition_rows(500, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, rows, cols, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    flags = rng.integers(0, 64, size=(rows, cols), dtype=np.uint8)
    soft = rng.random((rows, cols))
    r0, r1 = partition_rows(rows, world)[rank]
    G = GHOST_ROWS
    above, below = exchange_bands(flags[r0:r0 + G], flags[r1 - G:r1], rank, world, dist)
    local = with_ghost_bands(flags[r0:r1], above, below)
    sa, sb = exchange_bands(soft[r0:r0 + G], soft[r1 - G:r1], rank, world, dist)
    lsoft = with_ghost_bands(soft[r0:r1], sa, sb)
    lo, hi = max(r0 - G, 0), min(r1 + G, rows)
    want = flags[lo:hi]
    ok = local.shape == want.shape and (local == want).all() and (lsoft == soft[lo:hi]).all()
    ok = ok and (above is None) == (rank == 0) and (below is None) == (rank == world - 1)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_ghost_rows_over_gloo(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 2 * GHOST_ROWS * world + 37, 23, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(world))
