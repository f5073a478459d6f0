"""Multi-GPU data-parallel equivalence (needs >= 2 GPUs; skipped on single-GPU boxes): G ranks x N/G envs must produce
the same parameters, normalizer statistics and advantage statistics as one process with N envs.
The CPU-only gloo test of the same host logic is tests/test_cpu_host.py::test_data_parallel_host_logic_gloo_world2."""
import os
import signal
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,comm", [(2, "peer"), (2, "nccl")])
def test_data_parallel_equivalence(world, comm):
    """comm = peer: the default NVLink peer-memory exchanges (csrc/comm.cu), eager and as one CUDA graph;
    comm = nccl: the torch.distributed fallback (SFB200_DP_COMM=nccl)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py")]
    # own process group + hard limit: a stalled collective must not outlive the test (nor its pytest-timeout)
    env = dict(os.environ, SFB200_DP_COMM=comm)
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, start_new_session=True,
                            env=env)
    try:
        out, err = proc.communicate(timeout=300)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        pytest.fail("data-parallel worker did not finish in 300 s\n" + out[-2000:] + err[-2000:])
    assert proc.returncode == 0, out[-3000:] + err[-3000:]
    assert "DP_EQUIVALENCE_OK" in out
    assert comm != "peer" or "DP_GRAPH_OK" in out
