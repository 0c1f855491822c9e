"""Multi-GPU: z-sharded ESDF over NCCL equals the single-GPU result (needs >= 2 GPUs; the
1-GPU round-end run skips it -- the N>1 host logic is covered on CPU by tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("p2p,n", [("1", ("128", "96", "64")), ("0", ("128", "96", "64")), ("1", ("1024", "64", "64"))])
def test_sharded_esdf_matches_single_gpu(p2p, n):
    """Both forms of the partial's exchange -- peer-memory stores of the tile kernels (CUDA IPC, default) and the
    ncclSend/ncclRecv rounds (FUELGPU_SHARDED_P2P=0) -- and a map whose x lines are 2-CTA cluster tiles read as pieces."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "shard_esdf.py"),
           *n, "--check"]
    env = dict(os.environ, FUELGPU_SHARDED_P2P=p2p)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout
    assert ("peer-memory stores" if p2p == "1" else "ncclSend/Recv") in out.stdout


def test_sharded_frontier_search_matches_whole_search():
    """SURVEY 8e row 2 on 2 GPUs: every rank knows only its own z planes (+ one exchanged halo plane per side), sweeps them,
    and clusters the gathered candidate list: clusters / cells / average_ / filtered_cells_ / flags == the whole search."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tools", "shard_frontier.py"),
           "160", "128", "48"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout
