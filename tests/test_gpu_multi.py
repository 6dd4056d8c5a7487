"""Multi-GPU parity (needs >= 2 CUDA devices; skipped on a single-GPU box): the data-parallel training step through the
library's own NCCL communicator (se_comm_init / SE_OP_ALLREDUCE: bucketed all-reduces overlapped with the backward pass,
captured in the step graph) and through torch.distributed, each against the CPU oracle run per shard."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize('comm', ['native', 'torch'])
def test_two_gpu_data_parallel_step_matches_per_shard_oracle(comm):
    if _gpus() < 2:
        pytest.skip('needs two GPUs')
    env = dict(os.environ)
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'dp_check.py'), comm], capture_output=True, text=True,
                         timeout=600, env=env)
    assert out.returncode == 0 and 'DP_CHECK_OK ' + comm in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
