"""N-rank GPU test: the multi-GPU ingest output is compared with the CPU oracle, not with
another GPU path.  Needs >= 2 GPUs on the box; skipped otherwise (the world-size-2 gloo tests in test_sharding.py cover
the host logic without a GPU)."""

import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_ingest_matches_oracle(torch_cuda):
    n = torch_cuda.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (found %d)" % n)
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mp", "ingest_vs_oracle.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(res.stdout[-4000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count(" OK") == 4 and "MISMATCH" not in res.stdout
