"""-m gpu, needs >= 2 GPUs (skipped otherwise): the diagnostics exchange across ranks, one process per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("comm", ["p2p", "nccl"])
@pytest.mark.parametrize("S", [1, 2])
def test_two_rank_global_diagnostics(fcmod, comm, S):
    if fcmod.lib.fc_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + (os.getpid() + S * 7 + (comm == "nccl")) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "multi_gpu_worker.py"), "--comm", comm, "--S", str(S)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-gpu diagnostics ok" in r.stdout
