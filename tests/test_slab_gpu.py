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


def test_cpp_host_drives_slab_mode_through_the_c_abi():
    """nans_projekat_b200/host/nans_slab_host: one process per GPU, C ABI only (no Python / torch in the world's
    path), the set-up blobs over pipes.  Its per-slab state hashes after 20 steps of a squeezed pile (contacts
    across the slab faces) must equal those of the same world stepped on one GPU with nans_step."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    exe = os.path.join(ROOT, "nans_projekat_b200", "host", "nans_slab_host")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "nans_slab_host"])
    args = ["--side-x", "16", "--ny", "12", "--nz", "16", "--steps", "20"]
    multi = subprocess.run([exe, "--gpus", "2"] + args, capture_output=True, text=True, timeout=600)
    assert multi.returncode == 0, multi.stderr[-2000:]
    single = subprocess.run([exe, "--single", "2"] + args, capture_output=True, text=True, timeout=600)
    assert single.returncode == 0, single.stderr[-2000:]
    h = lambda out: {l.split()[1]: l.split()[3] for l in out.splitlines() if l.startswith("slab ")}
    hm, hs = h(multi.stdout), h(single.stdout)
    assert len(hm) == 2 and hm == hs, f"multi {multi.stdout} single {single.stdout}"
    ghosts = [int(l.split()[-1]) for l in multi.stdout.splitlines() if l.startswith("slab ")]
    assert max(ghosts) > 0, "no ghosts were exchanged"
