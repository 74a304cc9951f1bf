"""Real-NCCL multi-GPU run of the path (needs >= 2 GPUs; skipped otherwise)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize('transport', ['peer-memory', 'nccl', 'multicast'])
@pytest.mark.parametrize('world_size', [2, 4, 8])
def test_multi_gpu_path(world_size, transport):
    if _n_gpus() < world_size:
        pytest.skip('needs {} GPUs'.format(world_size))
    env = dict(os.environ)
    env['CHAINER_B200_P2P'] = '0' if transport == 'nccl' else '1'
    env['CHAINER_B200_MULTICAST'] = '1' if transport == 'multicast' else '0'
    env['CHAINER_B200_PEER_TIMEOUT_S'] = '60'      # a dead peer fails the test, not the box
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node', str(world_size), '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', '_dist_gpu_worker.py')]
    out = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         timeout=900, text=True)
    assert out.returncode == 0, out.stdout[-6000:]
    for r in range(world_size):
        assert 'GPU RANK %d OK' % r in out.stdout
    for line in out.stdout.splitlines():      # the worker's milestones, for the -rA log
        if line.startswith(('MNBN statistics exchange', 'MNBN MIXED-WIDTH', 'ONE-LAUNCH STEP OK',
                            'MULTICAST', 'GPU RANK')):
            print(line)
    if transport != 'nccl':
        assert 'ONE-LAUNCH STEP OK' in out.stdout
        assert 'MNBN MIXED-WIDTH WORKSPACE OK' in out.stdout
    if transport == 'multicast':
        if 'MULTICAST UNSUPPORTED' in out.stdout:
            pytest.skip([ln for ln in out.stdout.splitlines() if 'MULTICAST UNSUPPORTED' in ln][0])
        assert 'MULTICAST KERNEL OK' in out.stdout and 'MULTICAST PATH OK' in out.stdout
