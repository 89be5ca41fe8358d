"""Slab mode on real GPUs: tests/dist_check.py under torch.distributed.run at every world size the box
offers (2, 3, 4, 8) -- oracle parity, agreement with the single-GPU solve, identical iteration counts,
the benchmark's own call sequence.  Skipped on a one-GPU box (there is nothing to decompose over)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_slab_mode_parity(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_check.py")]
    env = dict(os.environ, HGRID_COMM_TIMEOUT="60", HGRID_DIST_WATCHDOG="240")
    p = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=420)
    sys.stdout.write(p.stdout[-6000:])
    assert p.returncode == 0, p.stdout[-3000:]
    assert "DIST_CHECK_OK" in p.stdout, p.stdout[-3000:]
