"""Multi-GPU parity (needs >= 2 devices; skipped on a one-GPU box -- run with `gpurun --gpus 2 -- python -m pytest tests -m gpu`).
One evaluation spread over the GPUs must return what a single GPU returns, on every rank: the replicated-factor layout
(scripts/dist_check.py: objective, gradient, alpha, sharded prediction) and the partitioned-storage layout of BASELINE config 5
(scripts/part_check.py: objective, alpha, K alpha, gradient, predictive mean and variance, residual against the definition)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, nproc, *args, timeout=900):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", script)] + [str(a) for a in args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def _ngpu():
    import gp_ss_ak_b200 as G
    return G.device_count()


@pytest.mark.parametrize("script,token,sizes", [("dist_check.py", "DIST CHECK OK", (1100, 3000, 9000)), ("part_check.py", "PART CHECK OK", (1100, 3000))])
def test_two_gpu_evaluation_matches_single_gpu(script, token, sizes):
    """n = 9000 (n_pad > 8192) puts the replicated handle on the int8 tensor-core pipe, the default at the sizes multi-GPU runs are for;
    part_check.py also compares the partitioned handle's predictive variance (variance_partitioned) with the single-GPU one."""
    if _ngpu() < 2:
        pytest.skip("needs 2 CUDA devices")
    out = _torchrun(script, 2, *sizes)
    assert out.returncode == 0 and token in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
