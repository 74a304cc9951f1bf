"""Pins the NumPy oracle against vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py) and against the reference's own known-answer
tests.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import gradpath as og
from tests.helpers import assert_bits_equal

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SHAPES = [(2, 3), (), (1, 0, 2), (257,), (130,), (8, 3, 3, 3)]
NAMES = ['/p%02d' % i for i in range(len(SHAPES))]


def _npz(name):
    return np.load(os.path.join(GOLD, name))


def test_layouts_match_workload_generators():
    """workloads.py regenerates the reference's sorted(namedparams()) layouts."""
    from chainer_b200 import workloads
    with open(os.path.join(GOLD, 'layouts.json')) as f:
        ref = json.load(f)
    for key, gen in workloads.WORKLOADS.items():
        mine = gen()
        theirs = [(n, tuple(s)) for n, s, _ in ref[key]]
        assert mine == theirs, key
        assert all(d == 'float32' for _, _, d in ref[key])
    assert workloads.n_elements(workloads.resnet50()) == 25557096
    assert workloads.n_elements(workloads.seq2seq()) == 173300800
    assert workloads.n_elements(workloads.mnist_mlp()) == 1796010


@pytest.mark.parametrize('dtype', ['float32', 'float16', 'float64'])
def test_momentum_sgd_matches_reference_bitwise(dtype):
    z = _npz('momentum_sgd.npz')
    for n in NAMES:
        p = z['%s|init%s' % (dtype, n)].copy()
        v = np.zeros_like(p)
        for step in range(3):
            g = z['%s|grad%d%s' % (dtype, step, n)]
            og.momentum_sgd_update(p, g, v, 0.01, 0.9)
            assert_bits_equal(p, z['%s|param%d%s' % (dtype, step, n)], 'param')
            assert_bits_equal(v, z['%s|state_v%d%s' % (dtype, step, n)], 'v')


ADAM_VARIANTS = {
    'adam': dict(),
    'adamw': dict(eta=0.5, weight_decay_rate=0.1),
    'amsgrad': dict(amsgrad=True),
    'adabound': dict(adabound=True),
    'amsbound': dict(amsgrad=True, adabound=True),
}


@pytest.mark.parametrize('variant', sorted(ADAM_VARIANTS))
@pytest.mark.parametrize('dtype', ['float32', 'float16', 'float64'])
def test_adam_cpu_restatement_matches_reference_bitwise(variant, dtype):
    z = _npz('adam.npz')
    kw = ADAM_VARIANTS[variant]
    for n in NAMES:
        key = '%s|%s|' % (variant, dtype)
        p = z[key + 'init' + n].copy()
        m, v, vh = np.zeros_like(p), np.zeros_like(p), np.zeros_like(p)
        for step in range(3):
            g = z[key + 'grad%d%s' % (step, n)]
            og.adam_update_cpu(p, g, m, v, step + 1, vhat=vh, **kw)
            assert_bits_equal(p, z[key + 'param%d%s' % (step, n)], 'param')
            assert_bits_equal(m, z[key + 'state_m%d%s' % (step, n)], 'm')
            assert_bits_equal(v, z[key + 'state_v%d%s' % (step, n)], 'v')
            if kw.get('amsgrad'):
                assert_bits_equal(vh, z[key + 'state_vhat%d%s' % (step, n)], 'vhat')


@pytest.mark.parametrize('variant', sorted(ADAM_VARIANTS))
@pytest.mark.parametrize('dtype,rtol', [('float32', 1e-6), ('float64', 1e-14), ('float16', 2.5e-3)])
def test_adam_gpu_restatement_within_tolerance_of_reference(variant, dtype, rtol):
    """The GPU-kernel operation order (what libgradpath implements bit-exactly)
    agrees with the reference CPU implementation within north_star's tolerance
    (float16: one half ulp is 2**-10 ~ 1e-3 relative, so three steps are allowed
    2.5 ulp)."""
    z = _npz('adam.npz')
    kw = ADAM_VARIANTS[variant]
    for n in NAMES:
        key = '%s|%s|' % (variant, dtype)
        p = z[key + 'init' + n].copy()
        m, v, vh = np.zeros_like(p), np.zeros_like(p), np.zeros_like(p)
        # float16: the reference's CPU path takes sqrt of the float16-ROUNDED v
        # (adam.py:213) while its GPU kernel uses the unrounded float v_
        # (adam.py:320-326); they diverge where (1-beta2) g^2 is subnormal in
        # float16.  Compare only elements whose v stays a normal half number.
        ok = np.ones(p.shape, dtype=bool)
        for step in range(3):
            g = z[key + 'grad%d%s' % (step, n)]
            if dtype == 'float16':
                ok &= np.abs(g.astype(np.float64)) >= 0.3
            og.adam_update_gpu(p, g, m, v, step + 1, vhat=vh, **kw)
            ref = z[key + 'param%d%s' % (step, n)]
            np.testing.assert_allclose(p.astype(np.float64)[ok], ref.astype(np.float64)[ok],
                                       rtol=rtol, atol=rtol * 1e-2)
            np.testing.assert_allclose(m.astype(np.float64),
                                       z[key + 'state_m%d%s' % (step, n)].astype(np.float64),
                                       rtol=rtol, atol=1e-9)


def test_adamw_known_answer():
    """tests/chainer_tests/optimizers_tests/test_optimizers.py:278-310 (TestAdamW)."""
    z = _npz('adam.npz')
    np.testing.assert_allclose(z['kat|adamw'], [0.9495], atol=1e-7, rtol=1e-7)
    for fn in (og.adam_update_cpu, og.adam_update_gpu):
        x = np.ones(1, dtype=np.float32)
        m, v = np.zeros_like(x), np.zeros_like(x)
        fn(x, np.ones_like(x), m, v, 1, eta=0.5, weight_decay_rate=0.1)
        np.testing.assert_allclose(x, [0.9495], atol=1e-7, rtol=1e-7)


def test_amsgrad_known_answer():
    """tests/chainer_tests/optimizers_tests/test_optimizers.py:328-366 (TestAMSGrad):
    Adam(alpha=0.01, beta2=0.7, amsgrad=True), x = 0, two steps."""
    for fn in (og.adam_update_cpu, og.adam_update_gpu):
        x = np.zeros(4, dtype=np.float32)
        m, v, vh = (np.zeros(4, np.float32) for _ in range(3))
        fn(x, np.array([1, -1, 10, -10], np.float32), m, v, 1, alpha=0.01, beta2=0.7,
           amsgrad=True, vhat=vh)
        np.testing.assert_allclose(v, [0.3, 0.3, 30, 30], atol=1e-7, rtol=1e-7)
        np.testing.assert_allclose(x, [-0.01, 0.01, -0.01, 0.01], atol=1e-7, rtol=1e-7)
        fn(x, np.array([-10, -10, -1, -1], np.float32), m, v, 2, alpha=0.01, beta2=0.7,
           amsgrad=True, vhat=vh)
        np.testing.assert_allclose(v, [30.21, 30.21, 21.3, 21.3], atol=1e-7, rtol=1e-7)
        np.testing.assert_allclose(vh, [30.21, 30.21, 30, 30], atol=1e-7, rtol=1e-7)
        np.testing.assert_allclose(x, [-0.00377703, 0.01745388, -0.01548985, 0.01686232],
                                   atol=1e-7, rtol=1e-7)


PNAMES = ['/lazy/b', '/p00', '/p01', '/p02', '/p03', '/p04']


@pytest.mark.parametrize('size', [2, 3])
@pytest.mark.parametrize('dtype', ['float32', 'float16', 'float64'])
def test_mean_grad_matches_reference_naive_communicator(size, dtype):
    """The pure_nccl restatement (pack, sum in the buffer dtype, x*(1.0/size),
    unpack) agrees with the reference NaiveCommunicator run on `size` ranks."""
    z = _npz('naive_mean_grad.npz')
    names = sorted(k.split('|')[-1] for k in z.files if k.startswith('%d|%s|in|0|' % (size, dtype)))
    assert names == PNAMES
    rank_grads = [[z['%d|%s|in|%d|%s' % (size, dtype, r, n)] for n in names]
                  for r in range(size)]
    buf_dtype = np.float32 if dtype == 'float16' else np.dtype(dtype)
    got = og.multi_node_mean_grad(rank_grads, buf_dtype)
    for r in range(size):
        for n, g in zip(names, got):
            ref = z['%d|%s|out|%d|%s' % (size, dtype, r, n)]
            assert g.dtype == ref.dtype and g.shape == ref.shape
            if dtype == 'float16' and n != '/lazy/b':
                # the naive path rounds the fp32 sum to fp16 BEFORE scaling
                # (mpi_communicator_base.py:770-775): 1 ulp of fp16
                np.testing.assert_allclose(g.astype(np.float64), ref.astype(np.float64),
                                           rtol=1e-3, atol=1e-7)
            elif size == 2:
                assert_bits_equal(g, ref, n)       # power-of-two size: exact
            else:
                rtol = 1e-6 if g.dtype == np.float32 else 1e-15
                np.testing.assert_allclose(g.astype(np.float64), ref.astype(np.float64),
                                           rtol=rtol, atol=1e-12)


def test_reference_analytic_mean_grad_vectors():
    """tests/chainermn_tests/communicator_tests/test_communicator.py:252-269:
    grads filled with rank, rank+1, rank+2 -> every element (size-1)/2 + k."""
    for size in (1, 2, 3, 4, 8):
        for buf_dtype in (np.float32, np.float16, np.float64, og.BF16):
            rank_grads = []
            for rank in range(size):
                rank_grads.append([np.full((3, 2), rank, np.float32),
                                   np.full((3,), rank, np.float32),
                                   np.full((4, 3), rank + 1, np.float32),
                                   np.full((4,), rank + 1, np.float32),
                                   np.full((5,), rank + 2, np.float32)])
            got = og.multi_node_mean_grad(rank_grads, buf_dtype)
            base = (size - 1) / 2.0
            tol = dict(rtol=1e-4, atol=1e-5) if buf_dtype in (np.float32, np.float64) \
                else dict(rtol=4e-3, atol=1e-5)
            for g, k in zip(got, [0, 0, 1, 1, 2]):
                np.testing.assert_allclose(g, base + k, **tol)


def test_mnbn_statistics_match_reference():
    """Oracle BN statistics -> mean/var -> y, gx, ggamma, gbeta equal the
    reference MultiNodeBatchNormalization (2 ranks, _MpiImpl) and the
    single-process BatchNormalization on the whole batch."""
    z = _npz('mnbn.npz')
    x_all, gy_all, gamma, beta = z['x'], z['gy'], z['gamma'], z['beta']
    size, nb = 2, 4
    eps = 2e-5
    xs = [x_all[r * nb:(r + 1) * nb] for r in range(size)]
    gys = [gy_all[r * nb:(r + 1) * nb] for r in range(size)]
    stats = [og.bn_fwd_stats(x, np.float32) for x in xs]
    mean, var = og.bn_mean_var_from_stats(stats, np.float32)
    inv_std = 1.0 / np.sqrt(var + np.float32(eps))
    e = (1, -1, 1, 1)
    bstats = [og.bn_bwd_stats(gy, og.x_hat(x, mean, inv_std), np.float32)
              for gy, x in zip(gys, xs)]
    s = og.scale_buffer(og.allreduce_sum(bstats, np.float32), np.float32, 1.0 / size)
    C = gamma.size
    gbeta, ggamma = s[:C], s[C:]
    for r in range(size):
        xh = og.x_hat(xs[r], mean, inv_std)
        y = gamma.reshape(e) * xh + beta.reshape(e)
        np.testing.assert_allclose(y, z['mn|%d|y' % r], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ggamma, z['mn|%d|ggamma' % r], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(gbeta, z['mn|%d|gbeta' % r], rtol=1e-5, atol=1e-6)
        inv_m = np.float32(1.0 / (xs[r].size // C))
        gx = (gamma * inv_std).reshape(e) * (
            gys[r] - (xh * ggamma.reshape(e) + gbeta.reshape(e)) * inv_m)
        np.testing.assert_allclose(gx, z['mn|%d|gx' % r], rtol=1e-4, atol=1e-6)
    # equivalence with single-process BN on the global batch: forward output
    # identical; MNBN's gamma/beta gradients are MEANS over ranks, the single
    # worker's are sums over the whole batch
    y_all = np.concatenate([z['mn|%d|y' % r] for r in range(size)])
    np.testing.assert_allclose(y_all, z['single|y'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ggamma * size, z['single|ggamma'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gbeta * size, z['single|gbeta'], rtol=1e-4, atol=1e-5)


def test_bf16_rounding_matches_bit_tricks():
    """bf16 has no reference; the oracle's RNE is checked against exact cases."""
    x = np.array([1.0, 1.00390625, 1.01171875, -3.0e38, 0.0, 2.0 ** -133 * 1.5, 65504.0],
                 dtype=np.float64)
    got = og.bf16_round(x)
    want = np.array([1.0, 1.0, 1.015625, -2.9973142e38, 0.0, 2.0 ** -132, 65536.0], np.float32)
    np.testing.assert_array_equal(got[[0, 1, 2, 4, 5, 6]], want[[0, 1, 2, 4, 5, 6]])
    assert got.dtype == np.float32
    assert (og.bf16_round(got) == got).all()          # idempotent
    f = np.float32(1.00390625)                        # exact tie -> even (down)
    assert og.bf16_round(np.array([f]))[0] == np.float32(1.0)
    f = np.float32(1.01171875)                        # exact tie -> even (up)
    assert og.bf16_round(np.array([f]))[0] == np.float32(1.015625)


# ------------------------------------------------- optimizer hooks + loss scaling --
HOOK_CASES = {
    # variant: (optimizer, [(hook, arg), ...] in registration order, loss_scale)
    'wd': ('sgd', [('wd', 0.05)], None),
    'clip': ('sgd', [('clip', 0.05)], None),
    'clip_noop': ('sgd', [('clip', 1e3)], None),
    'clip_wd': ('sgd', [('clip', 0.05), ('wd', 0.05)], None),
    'wd_clip': ('sgd', [('wd', 0.05), ('clip', 0.05)], None),
    'ls128': ('sgd', [], 128.0),
    'ls100_wd': ('sgd', [('wd', 0.05)], 100.0),
    'ls128_clip_wd': ('sgd', [('clip', 5.0), ('wd', 0.05)], 128.0),
    'adam_clip_wd': ('adam', [('clip', 0.05), ('wd', 0.05)], None),
    'adam_ls128_wd': ('adam', [('wd', 0.05)], 128.0),
}


def hooks_golden():
    return _npz('hooks.npz')


def replay_hooks(z, variant, dtype, check):
    """Replays one hooks.npz scenario with the oracle; `check(step, name, what, got, want)`."""
    opt_name, hooks, ls = HOOK_CASES[variant]
    pre = '%s|%s|' % (variant, dtype)
    names = sorted(k[len(pre) + 4:] for k in z.files if k.startswith(pre + 'init'))
    params = [z[pre + 'init' + n].copy() for n in names]
    st = [dict(m=np.zeros_like(p), v=np.zeros_like(p)) for p in params]
    for step in range(3):
        grads = [z[pre + 'grad%d%s' % (step, n)].copy() for n in names]
        # GradientMethod.update: optimizer-level hooks first (optimizer.py:881-883) ...
        for kind, arg in hooks:
            if kind == 'wd':
                for p, g in zip(params, grads):
                    og.weight_decay_hook(p, g, arg, ls)
            else:
                og.gradient_clipping_hook(grads, arg)
        # ... then every rule: loss-scale division, update_core (optimizer.py:286-295)
        for n, p, g, s in zip(names, params, grads, st):
            if ls is not None:
                og.loss_scale_divide(g, ls)
            if opt_name == 'sgd':
                og.momentum_sgd_update(p, g, s['v'])
            else:
                og.adam_update_cpu(p, g, s['m'], s['v'], step + 1)
            check(step, n, 'grad', g, z[pre + 'gradafter%d%s' % (step, n)])
            check(step, n, 'param', p, z[pre + 'param%d%s' % (step, n)])


@pytest.mark.parametrize('dtype', ['float32', 'float16'])
@pytest.mark.parametrize('variant', sorted(HOOK_CASES))
def test_hooks_and_loss_scale_match_reference(variant, dtype):
    if HOOK_CASES[variant][0] == 'adam' and dtype == 'float16':
        pytest.skip('no reference vector: float16 CPU Adam underflows (make_golden.py)')
    """The oracle's restatement of WeightDecay / GradientClipping / loss scaling in
    the order of GradientMethod.update is bit-identical to the unmodified reference."""
    def check(step, name, what, got, want):
        assert_bits_equal(got, want, (variant, dtype, step, name, what))
    replay_hooks(hooks_golden(), variant, dtype, check)


# ------------------------------------------ SGD, CorrectedMomentumSGD, NesterovAG --
FAMILY = {
    'sgd': (lambda p, g, s: og.sgd_update(p, g, 0.05)),
    'corrected': (lambda p, g, s: og.corrected_momentum_sgd_update(p, g, s['v'], 0.05, 0.8)),
    'nesterov': (lambda p, g, s: og.nesterov_ag_update(p, g, s['v'], 0.05, 0.8)),
}


def replay_family(z, rule, dtype, hooks, check):
    pre = '%s%s|%s|' % (rule, '_clip_wd' if hooks else '', dtype)
    names = sorted(k[len(pre) + 4:] for k in z.files if k.startswith(pre + 'init'))
    params = [z[pre + 'init' + n].copy() for n in names]
    st = [dict(v=np.zeros_like(p)) for p in params]
    for step in range(3):
        grads = [z[pre + 'grad%d%s' % (step, n)].copy() for n in names]
        if hooks:
            og.gradient_clipping_hook(grads, 0.05)
            for p, g in zip(params, grads):
                og.weight_decay_hook(p, g, 0.05)
        for n, p, g, s in zip(names, params, grads, st):
            FAMILY[rule](p, g, s)
            check(step, n, 'param', p, z[pre + 'param%d%s' % (step, n)])
            if rule != 'sgd':
                check(step, n, 'v', s['v'], z[pre + 'state_v%d%s' % (step, n)])


@pytest.mark.parametrize('dtype,hooks', [('float32', False), ('float16', False), ('float64', False),
                                         ('float32', True)])
@pytest.mark.parametrize('rule', sorted(FAMILY))
def test_sgd_family_matches_reference(rule, dtype, hooks):
    def check(step, name, what, got, want):
        assert_bits_equal(got, want, (rule, dtype, step, name, what))
    replay_family(_npz('sgd_family.npz'), rule, dtype, hooks, check)


# ------------------------------------------------------ fp32 master weights --
FP32_CASES = {'sgd': ('sgd', None, None), 'sgd_wd_ls128': ('sgd', 0.05, 128.0), 'adam': ('adam', None, None)}


@pytest.mark.parametrize('case', sorted(FP32_CASES))
def test_fp32_update_matches_reference(case):
    """use_fp32_update (chainer/optimizer.py:262-305) restated with the oracle's pieces:
    float32 master created once, gradient up-cast, hooks on the float16 gradient first,
    loss-scale division on the float32 gradient, update in float32, cast back."""
    z = _npz('fp32_update.npz')
    opt_name, wd, ls = FP32_CASES[case]
    pre = case + '|'
    names = sorted(k[len(pre) + 4:] for k in z.files if k.startswith(pre + 'init'))
    params = [z[pre + 'init' + n].copy() for n in names]
    masters = [p.astype(np.float32) for p in params]
    st = [dict(m=np.zeros_like(q), v=np.zeros_like(q)) for q in masters]
    for step in range(3):
        for i, n in enumerate(names):
            g = z[pre + 'grad%d%s' % (step, n)].copy()
            if wd is not None:
                og.weight_decay_hook(params[i], g, wd, ls)          # float16, optimizer-level
            g32 = g.astype(np.float32)
            if ls is not None:
                og.loss_scale_divide(g32, ls)
            if opt_name == 'sgd':
                og.momentum_sgd_update(masters[i], g32, st[i]['v'])
            else:
                og.adam_update_cpu(masters[i], g32, st[i]['m'], st[i]['v'], step + 1)
            params[i] = masters[i].astype(np.float16)
            assert_bits_equal(masters[i], z[pre + 'master%d%s' % (step, n)], (case, step, n, 'master'))
            assert_bits_equal(params[i], z[pre + 'param%d%s' % (step, n)], (case, step, n))
