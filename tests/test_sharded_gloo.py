"""N > 1 path on CPU: world_size 2 and 4 over gloo (one process per shard).  The GPU twin is
tests/test_gpu_parity.py::test_two_gpu_sharded (NCCL and NVLink peer stores)."""
import os
import subprocess
import sys

import pytest


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_gloo(hb, world):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "GLOO_SHARDED_OK" in out.stdout
