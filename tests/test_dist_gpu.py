"""Multi-GPU parity over NCCL (needs >= 2 GPUs on the box: `gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("wire", ["fp32", "bf16"])
def test_two_rank_nccl_gradient_equals_mean_of_oracle_gradients(wire):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, CDAE_TEST_BF16_WIRE="1" if wire == "bf16" else "0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "dist_grad_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("DISTGRAD")]
    print("\n".join(lines))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert len(lines) == 2
