"""-m gpu, needs >= 2 GPUs on the box (skipped on a single-GPU box): sharded answers are
bit-identical to the single-GPU answer for both partitions and both exchanges (tests/mgpu_worker.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_modes_and_exchanges_are_bit_identical_to_single_gpu():
    world = 2 if torch.cuda.device_count() < 4 else 4
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and f"MGPU_OK world={world}" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
