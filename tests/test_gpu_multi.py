"""Real multi-GPU parity (NCCL, one process per GPU): skipped on a box with fewer than two GPUs. The multi-process
host logic is additionally covered on CPU over gloo (tests/test_sharded_cpu.py) and by single-GPU emulations
(tests/test_gpu_parity.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_and_replicated_search_equal_the_oracle_on_real_gpus(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and f"MULTI_GPU_PARITY_OK {world}" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
