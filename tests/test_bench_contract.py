"""The reference arm of bench.py (`--impl reference`: the oracle port of the naive
communicator step on host cores) keeps the JSON contract of the driver, at one rank
and at N ranks."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = {'impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
        'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline',
        'e2e'}


@pytest.mark.parametrize('n', [1, 2])
def test_reference_arm_line(n):
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', str(n),
           '--steps', '2', '--warmup', '1', '--workload', 'mnist_mlp']
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK')}
    out = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert KEYS <= set(line)
    assert line['impl'] == 'reference' and line['n_gpus'] == n and line['vs_baseline'] is None
    assert line['value'] > 0 and line['ms_per_step'] > 0 and line['higher_is_better'] is True
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] == n
    assert line['cpu_baseline']['value'] == line['value'] == line['e2e']['value']
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in line['config'] and 'model' not in line['config']


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--gpus', '2', '--steps', '1', '--warmup', '1'], cwd=ROOT, env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ''
