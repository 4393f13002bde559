"""GPU, >= 2 devices: the point-range-sharded solve with the per-iteration all-reduce (peer-memory
one-shot all-reduce fused into the iteration kernel, and ncclAllReduce) against the oracle.
Skipped on a single-GPU box; there the same property is covered by the shard-linearity tests in
test_gpu_parity.py and the gloo protocol test in test_distributed_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("comm", ["peer", "nccl"])
@pytest.mark.timeout(600)
def test_sharded_solve_matches_oracle(comm):
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tests", "multi_gpu_worker.py"), comm]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=500)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_OK" in out.stdout
