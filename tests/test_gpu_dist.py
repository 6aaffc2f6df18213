"""Multi-GPU parity (needs >= 2 GPUs on the box, e.g. `gpurun --gpus 2`): the sharded table from
vapc_b200.dist.voxel_statistics equals the single-GPU table.  Launches tests/dist_worker.py under
torch.distributed.run (NCCL)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_matches_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "dist_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "DIST_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
