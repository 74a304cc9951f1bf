"""world_size-2 (and 3) runs of the N>1 host path on CPU: gloo process group,
oracle-backed library double (tests/_dist_worker.py)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize('world_size', [2, 3])
def test_multi_rank_host_path(world_size):
    env = dict(os.environ)
    env['OMP_NUM_THREADS'] = '1'
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node', str(world_size), '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', '_dist_worker.py')]
    out = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         timeout=600, text=True)
    assert out.returncode == 0, out.stdout[-4000:]
    for r in range(world_size):
        assert 'RANK %d OK' % r in out.stdout
