"""Real multi-GPU path: x-slabs over N GPUs of one box with the library's NCCL halo exchange.
Skipped on boxes with a single GPU (the single-GPU emulation in test_gpu_parity.py and the gloo
test in test_decomp_gloo.py cover the same logic there)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,axis,mp", [(2, "x", 2), (2, "y", 2), (4, "y", 2), (2, "x", 1), (2, "y", 1)])
def test_nccl_slabs_equal_single_image(world, axis, mp):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(ROOT, "tests", "multi_gpu_worker.py"), "96", "12", "96", "4", "3", axis, str(mp)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
