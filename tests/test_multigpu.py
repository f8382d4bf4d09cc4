"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches tests/multigpu_check.py under torchrun, one
rank per GPU, NCCL.  The CPU suite covers the same driver logic with gloo (tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('world', [2, 8])
def test_sharded_inversion_over_nccl(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(REPO, 'tests', 'multigpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    assert 'gathered maps == single-GPU maps: True' in out.stdout


def test_every_rank_falls_back_to_nccl_when_ipc_is_unavailable():
    """CLEDB_B200_NO_IPC=1 makes the IPC export of the flag arrays fail on every rank: the ranks agree on the NCCL transport
    (cledb_comm_transport reports it) and the same check passes through ncclAllGather + k_unshard alone."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    port = 29300 + os.getpid() % 300
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(REPO, 'tests', 'multigpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, CLEDB_B200_NO_IPC='1'))
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    assert 'default transport nccl' in out.stdout and 'transport=peer' not in out.stdout
    assert 'gathered maps == single-GPU maps: True' in out.stdout
