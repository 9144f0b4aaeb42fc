"""N-rank run (one process per GPU, NCCL ghost exchange + reductions inside libquokka_b200) is bit-identical to the
single-process oracle.  Needs >= 2 GPUs; skipped on a 1-GPU box (tests/test_gloo_exchange.py covers the host logic)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("world,n,b,steps", [(2, 64, 32, 6), (4, 64, 32, 4), (8, 64, 32, 4)])
def test_multirank_sedov_bit_exact(world, n, b, steps):
    if ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           str(free_port()), os.path.join(ROOT, "tests", "multirank_worker.py"), str(n), str(b), str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTIRANK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
