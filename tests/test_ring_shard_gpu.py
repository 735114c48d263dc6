"""Ring-sharded dense SQA over 2 GPUs (NVLink P2P hand-off) must reproduce the single-ring chain exactly."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ring_shard_matches_oracle_chain():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(ROOT, 'tests', 'ring_shard_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert 'RING_SHARD_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
