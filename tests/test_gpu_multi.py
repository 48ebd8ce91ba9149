"""GPU multi-rank parity (needs >= 2 GPUs on the box; skipped otherwise): N ranks vs one GPU vs oracle, over both transports."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("transport", ["peer_memory", "nccl"])
def test_two_rank_run_matches_single_gpu_and_oracle(transport):
    """Both transports of the halo exchange / allreduce (csrc/comm.cu): peer-memory windows over NVLink (default) and
    NCCL send/recv + allreduce (GLIMS_NO_P2P=1)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611" if transport == "peer_memory" else "29617",
           os.path.join(ROOT, "tests", "dist_check.py")]
    env = dict(os.environ)
    if transport == "nccl":
        env["GLIMS_NO_P2P"] = "1"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_CHECK_OK world=2" in out.stdout
