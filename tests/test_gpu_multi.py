"""GPU multi-rank parity (needs >= 2 GPUs on the box; skipped otherwise): N ranks vs one GPU vs oracle, over both transports."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("transport", ["peer_memory", "nccl", "peer_memory_no_graph", "peer_memory_plain_graph",
                                       "peer_memory_fused_coarse_only", "nccl_no_graph"])
def test_two_rank_run_matches_single_gpu_and_oracle(transport):
    """Both transports of the halo exchange / allreduce / level-1 all-gather (csrc/comm.cu): peer-memory windows over
    NVLink (default) and NCCL (GLIMS_NO_P2P=1); and the PCG driver without CUDA graphs / with plain graphs, where the
    stop decision is taken on the host and has to be identical on every rank; and the V-cycle with only
    levels >= 2 fused (GLIMS_AMG_FUSED=1) instead of the whole cycle below level 0."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29611 + 3 * ["peer_memory", "nccl", "peer_memory_no_graph", "peer_memory_plain_graph",
                                                          "peer_memory_fused_coarse_only", "nccl_no_graph"].index(transport)),
           os.path.join(ROOT, "tests", "dist_check.py")]
    env = dict(os.environ)
    if transport.startswith("nccl"):
        env["GLIMS_NO_P2P"] = "1"
    if transport.endswith("no_graph"):
        env["GLIMS_NO_GRAPH"] = "1"
    if transport.endswith("plain_graph"):
        env["GLIMS_NO_COND_GRAPH"] = "1"
    if transport.endswith("fused_coarse_only"):
        env["GLIMS_AMG_FUSED"] = "1"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_CHECK_OK world=2" in out.stdout
