"""Multi-rank parity check, launched by torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank solves its row slab of one grid; rank 0 gathers the result and compares it with the
CPU oracle and prints DIST_CHECK_OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from harmonic_interpolation_b200 import solver as S  # noqa: E402
from harmonic_interpolation_b200.distributed import SlabSolver, partition_rows  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for case, (rows, cols, K) in enumerate([(300 * world + 17, 300, 2), (1024 * max(world // 2, 1), 1024, 3)]):
        rng = np.random.default_rng(case)          # same stream on every rank
        if case == 0:
            hard = rng.random((rows, cols)) < 0.05
            soft = (rng.random((rows, cols)) < 0.08) & ~hard
        else:
            hard = np.kron(rng.random((rows // 64, cols // 64)) < 0.5, np.ones((64, 64), bool))
            soft = np.zeros((rows, cols), bool)
        hard[0, 0] = True
        sw = np.where(soft, rng.uniform(0.5, 20.0, (rows, cols)), 0.0)
        block = np.kron(rng.random((rows // 32 + 1, cols // 32 + 1)) < 0.3, np.ones((32, 32), bool))[:rows, :cols]
        cd = np.zeros((rows, cols), bool); cr = np.zeros((rows, cols), bool)
        if case == 0:
            cd[:-1] = block[:-1] != block[1:]; cr[:, :-1] = block[:, :-1] != block[:, 1:]
            hard[::32, ::32] = True
        hv = rng.random((rows, cols, K)); sv = rng.random((rows, cols, K))
        flags = S.encode_flags(rows, cols, hard=hard, soft=soft, cut_down=cd, cut_right=cr)
        r0, r1 = partition_rows(rows, world)[rank]
        for prec, tol in (("fp64", 1e-9), ("fp32", 1e-5)):
            s = SlabSolver(rows, cols, K, flags[r0:r1], dist, prec, soft_weights_owned=sw[r0:r1] if soft.any() else None, device=dev)
            x = s.solve(hv[r0:r1], sv[r0:r1] if soft.any() else None)
            its = s.last_stats["iterations"]
            parts = [None] * world
            dist.all_gather_object(parts, x)
            if rank == 0:
                from oracle import grid_oracle as go
                p = go.GridProblem(rows, cols, K, False)
                p.hard, p.hard_vals, p.soft_w, p.soft_vals, p.cut_down, p.cut_right = hard, hv, sw, sv, cd, cr
                ref = go.solve_problem(p)
                got = np.concatenate(parts, axis=0)
                err = np.abs(got - ref).max() / np.ptp(ref)
                # the distributed V-cycle is the single-GPU V-cycle: same iteration count
                s1 = S.GridSolver(rows, cols, K, False, prec, "multigrid")
                s1.set_pattern(flags, sw if soft.any() else None, 0.0, 0.0)
                s1.solve(hv, sv if soft.any() else None)
                its1 = s1.last_stats["iterations"]
                s1.close()
                print("case %d %s world %d: max|err|/range = %.3e, iterations %d (single GPU: %d)" % (case, prec, world, err, its, its1), flush=True)
                ok = ok and err < tol and abs(its - its1) <= 1
    if rank == 0:
        print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
