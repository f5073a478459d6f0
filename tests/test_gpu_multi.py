"""Multi-GPU data-parallel equivalence (needs >= 2 GPUs; skipped on single-GPU boxes): G ranks x N/G envs must produce
the same parameters, normalizer statistics and advantage statistics as one process with N envs.
The CPU-only gloo test of the same host logic is tests/test_cpu_host.py::test_data_parallel_host_logic_gloo_world2."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_data_parallel_equivalence_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DP_EQUIVALENCE_OK" in res.stdout
