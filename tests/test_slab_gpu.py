"""Two-GPU run of the slab decomposition (skipped on boxes with fewer than 2 GPUs): one world in two x-slabs
(NCCL halo exchange + cross-GPU dataflow solve over NVLink peer memory, csrc/slab.cu) must be bit-identical,
every checked step, to the same world on one GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_two_gpus_bit_exact():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tools", "slab_check.py"), "16", "12", "16", "20", "10", "5"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout[-3000:])
    assert out.returncode == 0, out.stderr[-3000:]
    assert "SLAB CHECK PASSED" in out.stdout
