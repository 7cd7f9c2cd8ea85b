"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): the block-sharded solver with the NCCL
all-reduce of partial A x must reproduce the single-GPU trajectory and iterates."""
import os
import subprocess
import sys

import pytest

import cuadmm_b200 as cu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2])
def test_sharded_solver_matches_single_gpu(world):
    if cu.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, NBLK="120", CON="20000", ITERS="40")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29541",
                          os.path.join(ROOT, "scripts", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=280, env=env)
    assert "MULTI_GPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
