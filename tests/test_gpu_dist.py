"""Multi-GPU parity inside the pytest suite: tests/dist_check.py under torchrun on all visible GPUs (skipped on a one-GPU box).
R ranks x (B / R samples) with row-sharded tables must reproduce the reference's single-process golden vectors -- y_pred, loss,
every gradient, three dense-Adam steps -- on both arithmetic paths."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('precision', ['tf32x3', 'fp32'])
def test_sharded_step_matches_reference(precision):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    env = dict(os.environ, SAT_PRECISION=precision)
    port = 29511 + (os.getpid() % 200)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                        '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'dist_check.py')],
                       env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout[-3000:] + r.stderr[-3000:])
    assert r.returncode == 0 and 'dist_check OK (world=%d)' % world in r.stdout, tail
