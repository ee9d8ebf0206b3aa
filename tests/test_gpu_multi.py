"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): frame-sharded decoder == unsharded, bit for bit."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, *args, port=29731):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc),
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'multi', 'frame_shard_check.py')] + list(args)
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize('config,T', [('tiny', 8), ('r50_704x256', 8)])
def test_frame_sharded_decoder_is_bit_identical(config, T):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    world = 2
    r = _run(world, config, str(T))
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-1500:]
    for ex in ('p2p', 'nccl'):
        assert r.stdout.count('exchange=%s' % ex) == world, r.stdout[-3000:]
