"""The drop-in, proven against the UNMODIFIED reference classes.

chainer / chainermn v7.8.1 are imported as they are (from `baseline/_ref`, or from
/root/reference in the build container) through `baseline/ref_shims.py`;
`chainer_b200.integration.install` applies INTEGRATION.md to them at run time; and the
reference's own objects -- `chainer.Link`, `chainer.Parameter`, `chainer.optimizers.
MomentumSGD / Adam`, `chainermn.create_multi_node_optimizer` -- then drive this package's
`pure_nccl` communicator and fused kernels through the C-ABI (the oracle-backed double of
the library on CPU; arrays are NumPy here, CuPy on a GPU box).  The checker is the
reference itself: the same step with its `naive` communicator and `update_core_cpu`.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.helpers import assert_bits_equal  # noqa: E402


def _import_reference():
    from baseline import ref_shims
    if not ref_shims.available():
        if os.path.isdir('/root/reference/chainermn'):
            ref_shims.REF_DIR = '/root/reference'
        else:
            pytest.skip('the unmodified reference is not installed (baseline/_ref)')
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return ref_shims.import_reference()


@pytest.fixture
def ref(monkeypatch):
    from chainer_b200.communicators import _control_plane
    from tests import fake_lib
    chainer, chainermn = _import_reference()
    fake, prev = fake_lib.install()
    _control_plane.reset_world()
    from chainer_b200 import integration
    integration.install(chainer, chainermn)
    yield chainer, chainermn, fake
    fake_lib.uninstall(prev)


SHAPES = [(3, 4), (7,), (64, 9), (5, 1, 2), (1000,)]


def _link(chainer, seed=3, dtype=np.float32):
    rng = np.random.default_rng(seed)
    link = chainer.Link()
    with link.init_scope():
        for i, s in enumerate(SHAPES):
            setattr(link, 'p%02d' % i, chainer.Parameter((rng.standard_normal(s) * 0.1).astype(dtype)))
    return link


def _set_grads(link, step, scale=1e-2):
    rng = np.random.default_rng(100 + step)
    for _, p in sorted(link.namedparams()):
        p.grad = (rng.standard_normal(p.shape) * scale).astype(p.dtype)


def test_reference_wrapper_drives_our_communicator(ref):
    """Unpatched protocol first: the reference's own `_MultiNodeOptimizer` calls
    `bcast_data` / `multi_node_mean_grad` of our communicator on chainer objects, then its own
    optimizer; equals the all-reference run bit for bit."""
    chainer, chainermn, fake = ref
    import chainermn.optimizers as ref_opt_mod
    comm = chainermn.create_communicator('pure_nccl')
    from chainer_b200.communicators.pure_nccl_communicator import PureNcclCommunicator
    assert type(comm) is PureNcclCommunicator
    assert isinstance(comm, chainermn.CommunicatorBase)
    a, b = _link(chainer), _link(chainer)
    oa = chainermn.create_multi_node_optimizer(chainer.optimizers.MomentumSGD(lr=0.1),
                                               chainermn.create_communicator('naive'))
    ob = ref_opt_mod._MultiNodeOptimizer.__new__(ref_opt_mod._MultiNodeOptimizer)
    ref_opt_mod._MultiNodeOptimizer.__init__(ob, chainer.optimizers.MomentumSGD(lr=0.1), comm, True)
    oa.setup(a)
    ob.setup(b)
    for step in range(4):
        _set_grads(a, step)
        _set_grads(b, step)
        oa.update()
        fake.calls[:] = []
        # the reference's original two-call sequence (what `update` does without section 4)
        if ob.is_changed(b):
            comm.bcast_data(b)
        else:
            comm.multi_node_mean_grad(b, ob.zero_fill)
            ob.actual_optimizer.update(None)
        names = [c[0] for c in fake.calls]
        assert 'gp_pack' in names and 'gp_unpack_scale' in names
        for (n, p), (_, q) in zip(sorted(a.namedparams()), sorted(b.namedparams())):
            assert_bits_equal(q.data, p.data, (step, n))
            assert_bits_equal(q.grad, p.grad, (step, n, 'grad'))


@pytest.mark.parametrize('opt_name', ['momentum_sgd', 'adam'])
@pytest.mark.parametrize('hook', [None, 'wd'])
def test_fused_step_with_reference_objects(ref, opt_name, hook):
    """`chainermn.create_multi_node_optimizer(chainer.optimizers.X(), create_communicator(
    'pure_nccl')).update()` with INTEGRATION.md applied: ONE fused library launch per step
    (gp_step_* / gp_unpack_*_hooked) on the reference's Link / Parameter / UpdateRule
    objects; parameters, states, step counters and param.grad equal the all-reference run."""
    chainer, chainermn, fake = ref
    mk = (lambda: chainer.optimizers.MomentumSGD(lr=0.1, momentum=0.9)) \
        if opt_name == 'momentum_sgd' else (lambda: chainer.optimizers.Adam(alpha=0.01))
    a, b = _link(chainer), _link(chainer)
    oa = chainermn.create_multi_node_optimizer(mk(), chainermn.create_communicator('naive'))
    ob = chainermn.create_multi_node_optimizer(mk(), chainermn.create_communicator('pure_nccl'))
    oa.setup(a)
    ob.setup(b)
    if hook == 'wd':
        oa.add_hook(chainer.optimizer_hooks.WeightDecay(0.05))
        ob.add_hook(chainer.optimizer_hooks.WeightDecay(0.05))
    exact = opt_name == 'momentum_sgd'       # Adam: the GPU kernel's formula vs update_core_cpu
    for step in range(4):
        _set_grads(a, step)
        _set_grads(b, step)
        oa.update()
        fake.calls[:] = []
        ob.update()
        names = [c[0] for c in fake.calls]
        if step == 0:
            assert 'gp_nccl_bcast' in names or 'gp_pack' in names        # bcast_data only
        else:
            want = {('momentum_sgd', None): 'gp_step_momentum_sgd', ('adam', None): 'gp_step_adam',
                    ('momentum_sgd', 'wd'): 'gp_unpack_momentum_sgd_hooked',
                    ('adam', 'wd'): 'gp_unpack_adam_hooked'}[(opt_name, hook)]
            assert want in names, names
            assert 'gp_unpack_scale' not in names            # not the two-call sequence
        assert oa.t == ob.t
        for (n, p), (_, q) in zip(sorted(a.namedparams()), sorted(b.namedparams())):
            assert p.update_rule.t == q.update_rule.t
            if exact:
                assert_bits_equal(q.data, p.data, (step, n))
                assert_bits_equal(q.grad, p.grad, (step, n, 'grad'))
                if step > 0:
                    assert_bits_equal(q.update_rule.state['v'], p.update_rule.state['v'], (step, n, 'v'))
            else:
                np.testing.assert_allclose(q.data, p.data, rtol=1e-6, atol=1e-7)


def test_reference_mnbn_link_uses_our_statistics(ref):
    """`chainermn.links.MultiNodeBatchNormalization(size, comm)` of the reference with our
    communicator: its impl selector picks `_NcclImpl`, which INTEGRATION.md section 5 replaces
    by the statistics kernels; forward / backward equal the reference's own
    `chainer.links.BatchNormalization` on the same (single-worker) batch."""
    chainer, chainermn, fake = ref
    comm = chainermn.create_communicator('pure_nccl')
    rng = np.random.default_rng(5)
    x = rng.standard_normal((6, 5, 4, 3)).astype(np.float32)
    gy = rng.standard_normal((6, 5, 4, 3)).astype(np.float32) * 0.1
    mn = chainermn.links.MultiNodeBatchNormalization(5, comm)
    bn = chainer.links.BatchNormalization(5)
    outs = []
    for link in (mn, bn):
        v = chainer.Variable(x.copy())
        fake.calls[:] = []
        with chainer.using_config('train', True):
            y = link(v)
        y.grad = gy.copy()
        link.cleargrads()
        y.backward()
        outs.append((y.data, v.grad, link.gamma.grad, link.beta.grad, link.avg_mean, link.avg_var,
                     [c[0] for c in fake.calls]))
    assert 'gp_bn_fwd_stats' in outs[0][6] and 'gp_bn_bwd_stats' in outs[0][6]
    assert outs[1][6] == []
    for got, want in zip(outs[0][:5], outs[1][:5]):
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)
    # the multi-node link starts avg_var at zeros, the stock link at ones
    # (chainermn/links/batch_normalization.py:57-60): decay * 1 apart after one step
    np.testing.assert_allclose(outs[0][5], outs[1][5] - 0.9, rtol=2e-5, atol=2e-6)
