"""GPU, needs >= 2 devices (skipped otherwise): the real one-process-per-GPU exchange -- CUDA IPC mailboxes over NVLink
("p2p") and NCCL all_gather + send/recv ("nccl") -- several steps, merged lists compared with the oracle by rank 0.
Run by hand with:  gpurun --gpus 2 -- python -m pytest tests/test_slab_multigpu.py -m gpu -q"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_two_process_slab_exchange(oracle, transport):
    ng = _ngpu()
    if ng < 2:
        pytest.skip("needs 2 GPUs")
    world = 4 if ng >= 4 else 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(root, "tests", "slab_p2p_worker.py"), transport, "30000", "4"]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(": OK") == 4, out.stdout[-3000:]


@pytest.mark.parametrize("kind", [2, 3])
def test_rb3d_sphere_slabs_one_process_per_gpu(oracle, kind):
    """rigidbody3d slab mode (all-sphere scenes, BASELINE configs[3]) with one process per GPU and CUDA IPC mailboxes."""
    ng = _ngpu()
    if ng < 2:
        pytest.skip("needs 2 GPUs")
    world = 4 if ng >= 4 else 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", "29534",
           os.path.join(root, "tests", "slab_rb3d_worker.py"), "40000", "4", str(kind)]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(": OK") == 4, out.stdout[-3000:]
