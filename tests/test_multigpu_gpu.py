"""NCCL paths of the sharded workloads on >= 2 GPUs (skipped on a single-GPU box): sharded brute force with the min/gather merge,
replica batches over the ranks."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_bf_and_replicas_over_nccl(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', str(29540 + world), os.path.join(ROOT, 'tests', 'multigpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert 'MULTIGPU_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
