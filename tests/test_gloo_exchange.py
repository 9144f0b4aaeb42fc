"""world_size-2 (and 4) run of the multi-rank host logic on CPU over gloo: each rank builds its own level plan through
the C ABI (no GPU needed for planning), exchanges ghost messages with its peers, and checks every ghost cell."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n,b,periodic", [(2, 32, 16, 0), (2, 32, 16, 1), (4, 32, 8, 1), (8, 32, 8, 0)])
def test_gloo_ghost_exchange(world, n, b, periodic):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           str(free_port()), os.path.join(ROOT, "tests", "gloo_worker.py"), str(n), str(b), str(periodic)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "GLOO_EXCHANGE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
