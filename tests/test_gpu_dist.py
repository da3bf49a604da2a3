"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the row-partitioned SpMV is
bit-identical to the single-GPU one, the distributed V-cycle agrees to 1e-12 and PCG iteration
counts agree within +-1."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("halo", ["p2p", "nccl"])
@pytest.mark.parametrize("dims,rep", [((32, 32, 32), 500), ((24, 20, 16), 100)])
def test_two_rank_parity(dims, rep, halo):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    # a port of its own per case: back-to-back rendezvous on one port can find it still in TIME_WAIT
    port = 29517 + 2 * ["p2p", "nccl"].index(halo) + (0 if rep == 500 else 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "dist_check.py"), *map(str, dims), str(rep)]
    env = dict(os.environ, FAMG_HALO=halo)  # peer-memory (CUDA IPC) stores vs NCCL send/recv
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert "DIST_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
