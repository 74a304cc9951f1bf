"""The reference arm of bench.py (`--impl reference`: the unmodified chainermn naive
communicator + chainer update_core_cpu from baseline/_ref when that install is present,
else the oracle port, on host cores) keeps the JSON contract of the driver, at one rank
and at N ranks, and prints the same `config` object as the B200 arm would."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = {'impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
        'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline',
        'e2e'}


@pytest.mark.parametrize('kind', ['auto', 'port'])
@pytest.mark.parametrize('n', [1, 2])
def test_reference_arm_line(n, kind):
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', str(n),
           '--steps', '2', '--warmup', '1', '--workload', 'mnist_mlp', '--reference-kind', kind]
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
    have_ref = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'chainermn'))
    want_kind = 'reference' if (kind == 'auto' and have_ref) else 'port'
    assert line['cpu_baseline']['kind'] == want_kind and line['cpu_baseline']['cores'] == n
    assert line['steps'] == 2 and line['warmup'] == 1          # exactly what was asked for
    # the same `config` the B200 arm prints for these flags
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    args = argparse.Namespace(workload='mnist_mlp', allreduce_dtype='float32', no_write_grad=False,
                              zero_embedding_rows=0.0)
    _, sizes = bench.workload_sizes('mnist_mlp')
    assert line['config'] == bench.make_config(args, 'adam', n, sizes)
    assert line['cpu_baseline']['value'] == line['value'] == line['e2e']['value']
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in line['config'] and 'model' not in line['config']


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--gpus', '2', '--steps', '1', '--warmup', '1'], cwd=ROOT, env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ''


def _run_guard(body, rank=0):
    code = ('import sys, time, json; sys.path.insert(0, %r); import bench\n'
            'line = {"metric": "m", "value": 1.0}\n'
            'g = bench._LegsGuard(line, %d, 0.5); g.start()\n' % (ROOT, rank)) + body
    return subprocess.run([sys.executable, '-c', code], cwd=ROOT, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=120)


def test_bench_line_survives_a_leg_that_never_returns():
    """A leg beside the headline that stops making progress (a peer lost inside an exchange)
    must cost its own numbers, not the line: at the deadline rank 0 prints the line as it
    stands with `legs_error` and exits 0; the stacks go to stderr."""
    out = _run_guard('time.sleep(60)\nprint("not reached")\n')
    assert out.returncode == 0
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1 and 'not reached' not in out.stdout
    line = json.loads(lines[0])
    assert line['value'] == 1.0 and 'did not finish within 0.5 s' in line['legs_error']
    assert 'time.sleep' in out.stderr or 'File' in out.stderr       # the stack dump
    # other ranks leave quietly with the same exit code
    out = _run_guard('time.sleep(60)\n', rank=1)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_bench_line_is_printed_once_when_the_legs_finish():
    out = _run_guard('line["config3"] = {"x": 1}\ng.finish()\ng.finish()\ntime.sleep(1.0)\n')
    assert out.returncode == 0
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['config3'] == {'x': 1} and 'legs_error' not in line
