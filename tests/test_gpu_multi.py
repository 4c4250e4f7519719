"""Multi-GPU parity: P z-slabs (stencil halos over NCCL, SOR halos and residuals
through peer-mapped memory) == one GPU, bit for bit.  Needs >= 2 GPUs
on the box (`gpurun --gpus 2`); on a 1-GPU box the test is skipped with that reason (two NCCL
ranks cannot share one device)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_decomposition_is_bitwise_equal_to_one_gpu(gpu, world):
    if gpu.device_count() < world:
        pytest.skip("needs %d GPUs on one box, found %d" % (world, gpu.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
           str(world), "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:]
    assert r.stdout.count("BITWISE-EQUAL") >= 18
