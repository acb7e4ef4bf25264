"""Multi-GPU parity through the library's own collective (needs >= 2 B200s on the box: ``gpurun --gpus 2``; skipped on a
single-GPU box, where ``test_comm_world_1`` still exercises the exports)."""
import ctypes
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_comm_exports_world_1():
    """zett_comm_init / zett_allgather_rows / zett_comm_info / zett_comm_destroy through ctypes; a single rank needs no NCCL
    id and gathers in place."""
    import torch
    from zett_b200 import parallel
    comm = parallel.NativeComm(0, 1, None)
    assert comm.info() == {"rank": 0, "world": 1, "nccl_version": 0}
    x = torch.arange(12, dtype=torch.float32, device="cuda").reshape(3, 4)
    comm.allgather_rows(x, 3)
    assert torch.equal(x.cpu(), torch.arange(12, dtype=torch.float32).reshape(3, 4))
    assert len(parallel.NativeComm.unique_id()) == 128
    comm.close()


def test_sharded_equals_single_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(HERE, "multi_gpu_check.py")], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
