"""GPU tests through the public Python API (torch CUDA tensors and DeviceArray
as device buffers): the reference's communicator / optimizer / MNBN tests with
results compared against the oracle."""
import numpy as np
import pytest

import chainer_b200
from chainer_b200 import config
from oracle import gradpath as og
from tests.helpers import assert_bits_equal

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _model(plist, rng, dtype=np.float32, scale=0.05):
    from chainer_b200.core.link import link_from_named_arrays
    host = [np.asarray(rng.standard_normal(s) * scale).astype(dtype).reshape(s) for _, s in plist]
    return link_from_named_arrays([(n, _t(a)) for (n, _), a in zip(plist, host)]), host


PLIST = [('/a/W', (3, 2)), ('/a/b', (3,)), ('/b/W', (4, 3)), ('/b/b', (4,)), ('/c/b', (5,)),
         ('/d/W', (257, 129)), ('/d/b', (1,)), ('/e/W', (64, 3, 7, 7)), ('/f/W', (0,))]


@pytest.fixture
def comm():
    c = chainer_b200.create_communicator('pure_nccl')
    yield c
    c.finalize()
    config.set_debug(False)
    config.set_dtype(None)


def _set_grads(model, host_grads):
    for (_, p), g in zip(sorted(model.namedparams()), host_grads):
        p.grad = _t(g)


@pytest.mark.parametrize('allreduce_dtype', [None, np.float16, 'bfloat16', np.float64])
def test_mean_grad_single_gpu(comm, allreduce_dtype):
    import torch
    comm.set_config('allreduce_grad_dtype', allreduce_dtype)
    rng = np.random.default_rng(0)
    model, _ = _model(sorted(PLIST), rng)
    grads = [np.asarray(rng.standard_normal(s)).astype(np.float32).reshape(s) for _, s in sorted(PLIST)]
    _set_grads(model, grads)
    comm.multi_node_mean_grad(model)
    torch.cuda.synchronize()
    bd = og.BF16 if allreduce_dtype == 'bfloat16' else (allreduce_dtype or np.float32)
    want = og.multi_node_mean_grad([grads], bd)
    for (_, p), w in zip(sorted(model.namedparams()), want):
        assert_bits_equal(p.grad.cpu().numpy(), w, 'mean grad')


@pytest.mark.parametrize('opt_name,kw', [
    ('momentum_sgd', dict(lr=0.01, momentum=0.9)), ('adam', dict()),
    ('adam', dict(eta=0.5, weight_decay_rate=0.1)), ('adam', dict(amsgrad=True)),
    ('adam', dict(adabound=True)), ('adam', dict(amsgrad=True, adabound=True))])
@pytest.mark.parametrize('pdtype', [np.float32, np.float16, np.float64])
@pytest.mark.parametrize('write_grad', [True, False])
def test_multi_node_optimizer_fused_bit_exact(comm, opt_name, kw, pdtype, write_grad):
    import torch
    comm.write_grad = write_grad
    rng = np.random.default_rng(1)
    gscale = 0.5 if pdtype == np.float16 else 1e-2
    model, host_p = _model(sorted(PLIST), rng, pdtype)
    actual = chainer_b200.MomentumSGD(**kw) if opt_name == 'momentum_sgd' else chainer_b200.Adam(**kw)
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(model)
    opt.update()
    assert actual.t == 0
    # The first update() is bcast_data, whose transfer dtype is chainer.get_dtype()
    # = float32 (pure_nccl_communicator.py:85-99): float64 parameters come back
    # rounded to float32, exactly as in the reference.
    host_p = [a.astype(np.float32).astype(pdtype) for a in host_p]
    st = [dict(m=np.zeros_like(a), v=np.zeros_like(a), vhat=np.zeros_like(a)) for a in host_p]
    for step in range(1, 4):
        grads = [np.asarray(rng.standard_normal(a.shape) * gscale).astype(pdtype).reshape(a.shape)
                 for a in host_p]
        _set_grads(model, grads)
        held = [p.grad for _, p in sorted(model.namedparams())]
        opt.update()
        torch.cuda.synchronize()
        assert actual.t == step
        mean = og.multi_node_mean_grad([grads], np.float32 if pdtype != np.float64 else np.float32)
        for (name, p), q, g, s, h, g_in in zip(sorted(model.namedparams()), host_p, mean, st, held,
                                               grads):
            assert p.update_rule.t == step
            if opt_name == 'momentum_sgd':
                og.momentum_sgd_update(q, g, s['v'], kw['lr'], kw['momentum'])
            else:
                og.adam_update_gpu(q, g, s['m'], s['v'], step, vhat=s['vhat'], **kw)
            assert_bits_equal(p.data.cpu().numpy(), q, name)
            assert_bits_equal(h.cpu().numpy(), g if write_grad else g_in, name + ' grad')


def test_fused_equals_unfused_on_gpu(comm):
    import torch
    res = []
    for fused in (True, False):
        rng = np.random.default_rng(5)
        model, _ = _model(sorted(PLIST), rng)
        actual = chainer_b200.Adam()
        opt = chainer_b200.create_multi_node_optimizer(actual, comm)
        opt.setup(model)
        if not fused:
            actual.add_hook(lambda o: None, name='noop')
        opt.update()
        for step in range(3):
            grads = [np.asarray(rng.standard_normal(s) * 1e-2).astype(np.float32).reshape(s)
                     for _, s in sorted(PLIST)]
            _set_grads(model, grads)
            opt.update()
        torch.cuda.synchronize()
        res.append([p.data.cpu().numpy() for _, p in sorted(model.namedparams())])
    for a, b in zip(*res):
        assert_bits_equal(a, b, 'fused vs unfused')


def test_device_array_buffers(comm):
    """Any device buffer works: the package's own DeviceArray instead of torch."""
    from chainer_b200.core.link import link_from_named_arrays
    from chainer_b200.device import DeviceArray
    rng = np.random.default_rng(2)
    plist = sorted(PLIST)[:6]
    host = [np.asarray(rng.standard_normal(s)).astype(np.float32).reshape(s) for _, s in plist]
    model = link_from_named_arrays([(n, DeviceArray.from_numpy(a)) for (n, _), a in zip(plist, host)])
    grads = [np.asarray(rng.standard_normal(s)).astype(np.float32).reshape(s) for _, s in plist]
    for (_, p), g in zip(sorted(model.namedparams()), grads):
        p.grad = DeviceArray.from_numpy(g)
    opt = chainer_b200.create_multi_node_optimizer(chainer_b200.MomentumSGD(lr=0.1), comm)
    opt.setup(model)
    opt.update()
    opt.update()
    for (_, p), q, g in zip(sorted(model.namedparams()), host, grads):
        v = np.zeros_like(q)
        og.momentum_sgd_update(q, g, v, 0.1, 0.9)
        assert_bits_equal(p.data.get(), q, 'DeviceArray param')


def test_debug_mode_divergence_on_gpu(comm):
    rng = np.random.default_rng(3)
    model, _ = _model(sorted(PLIST), rng)
    grads = [np.asarray(rng.standard_normal(s)).astype(np.float32).reshape(s) for _, s in sorted(PLIST)]
    grads[3][0] = np.nan
    _set_grads(model, grads)
    config.set_debug(True)
    with pytest.raises(ValueError, match='.* diverged .*'):
        comm.multi_node_mean_grad(model)


BN_SHAPES = [(32, 64, 112, 112), (32, 256, 56, 56), (32, 512, 28, 28), (32, 1024, 14, 14),
             (32, 2048, 7, 7), (8, 3, 5, 7), (5, 17, 1, 1), (2, 4, 3)]


@pytest.mark.parametrize('shape', BN_SHAPES)
@pytest.mark.parametrize('xdtype', [np.float32, np.float16])
def test_bn_statistics_match_oracle(comm, shape, xdtype):
    """BASELINE config 3 layer shapes: [mean | E[x^2]] -> mean, var; and
    [sum gy | sum gy*x_hat].  fp32 accumulation order differs from NumPy's
    pairwise sum: 1e-6 relative on mean-of-squares / sums of O(1) magnitude, with
    an absolute floor for the near-zero means (SURVEY.md hard parts)."""
    import torch
    from chainer_b200.functions.batch_normalization import _NcclImpl
    rng = np.random.default_rng(11)
    x = rng.standard_normal(shape).astype(xdtype)
    gy = (rng.standard_normal(shape) * 1e-3).astype(xdtype)
    C = shape[1]
    gamma = np.ones(C, np.float32)
    impl = _NcclImpl(comm)
    mean, var = impl.get_mean_and_var(None, _t(gamma), _t(x))
    x64 = x.astype(np.float64)
    axis = (0,) + tuple(range(2, x.ndim))
    m_ref = x64.mean(axis=axis)
    sq_ref = np.square(x64).mean(axis=axis)
    np.testing.assert_allclose(mean.cpu().numpy(), m_ref, rtol=1e-6, atol=2e-7)
    np.testing.assert_allclose(var.cpu().numpy(), sq_ref - m_ref ** 2, rtol=3e-6, atol=1e-6)
    # oracle (float32 NumPy, the reference's own arithmetic) within its own error
    o = og.bn_fwd_stats(x, np.float32)
    np.testing.assert_allclose(mean.cpu().numpy(), o[:C], rtol=1e-5, atol=1e-6)

    inv_std = 1.0 / np.sqrt((sq_ref - m_ref ** 2) + 2e-5)
    xh = og.x_hat(x.astype(np.float32), m_ref.astype(np.float32), inv_std.astype(np.float32))
    gbeta, ggamma = impl.get_ggamma_and_gbeta(None, _t(gamma), _t(gy), _t(xh))
    gb_ref = gy.astype(np.float64).sum(axis=axis)
    gg_ref = (gy.astype(np.float64) * xh.astype(np.float64)).sum(axis=axis)
    scale = np.abs(gy.astype(np.float64)).sum(axis=axis).max()
    np.testing.assert_allclose(gbeta.cpu().numpy(), gb_ref, rtol=1e-5, atol=1e-6 * scale)
    np.testing.assert_allclose(ggamma.cpu().numpy(), gg_ref, rtol=1e-5, atol=1e-6 * scale)
    # on-the-fly x_hat variant agrees with the materialised one
    gb2, gg2 = impl.get_ggamma_and_gbeta_from_x(None, _t(gamma), _t(gy), _t(x.astype(xdtype)),
                                                _t(m_ref.astype(np.float32)),
                                                _t(inv_std.astype(np.float32)))
    np.testing.assert_allclose(gg2.cpu().numpy(), gg_ref, rtol=1e-4, atol=2e-6 * scale)
    np.testing.assert_allclose(gb2.cpu().numpy(), gb_ref, rtol=1e-5, atol=1e-6 * scale)


def test_bn_statistics_layers_of_different_width_share_one_workspace(comm):
    """A model's BN layers all use the communicator's ONE statistics workspace.  Walk the
    distinct ResNet-50 layer shapes (C = 64 ... 2048, split counts S = 8 ... 1) twice, forward
    and backward, after the workspace has reached its final size: every call must still be
    right.  (With a C-dependent workspace layout the split partials of the narrow layers landed
    on the tickets of the wider ones and those channels were never written.)"""
    import torch
    from chainer_b200 import workloads
    from chainer_b200.functions.batch_normalization import _NcclImpl
    impl = _NcclImpl(comm)
    shapes = []
    for _, s in workloads.resnet50_bn_layers(32):
        if s not in shapes:
            shapes.append(s)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(5)
    data = {}
    for s in shapes:
        x = torch.randn(*s, device='cuda', generator=gen) + 0.25
        gy = torch.randn(*s, device='cuda', generator=gen) * 1e-3
        x64 = x.double()
        m = x64.mean(dim=(0, 2, 3))
        v = (x64 * x64).mean(dim=(0, 2, 3)) - m * m
        inv = torch.rsqrt(v + 2e-5)
        gb = gy.double().sum(dim=(0, 2, 3))
        gg = (gy.double() * (x64 - m[None, :, None, None]) * inv[None, :, None, None]).sum(dim=(0, 2, 3))
        data[s] = (x, gy, torch.ones(s[1], device='cuda'), m, v, inv, gb, gg,
                   gy.double().abs().sum(dim=(0, 2, 3)).max().item())
        del x64
    # the workspace takes its final size here (widest layer last) ...
    impl.get_mean_and_var(None, data[shapes[-1]][2], data[shapes[-1]][0])
    # ... and is then shared by every width, narrow and wide in turn
    for order in (shapes, shapes[::-1], shapes):
        for s in order:
            x, gy, gamma, m, v, inv, gb, gg, scale = data[s]
            mean, var = impl.get_mean_and_var(None, gamma, x)
            torch.testing.assert_close(mean.double(), m, rtol=1e-5, atol=2e-6, msg=str(s))
            torch.testing.assert_close(var.double(), v, rtol=1e-5, atol=2e-6, msg=str(s))
            gbeta, ggamma = impl.get_ggamma_and_gbeta_from_x(None, gamma, gy, x, m.float(),
                                                             inv.float())
            torch.testing.assert_close(gbeta.double(), gb, rtol=1e-4, atol=2e-6 * scale, msg=str(s))
            torch.testing.assert_close(ggamma.double(), gg, rtol=1e-4, atol=4e-6 * scale, msg=str(s))


def test_bn_statistics_deterministic(comm):
    import torch
    from chainer_b200.functions.batch_normalization import _NcclImpl
    x = torch.randn(32, 64, 56, 56, device='cuda')
    gamma = torch.ones(64, device='cuda')
    impl = _NcclImpl(comm)
    a = [t.clone() for t in impl.get_mean_and_var(None, gamma, x)]
    for _ in range(5):
        b = impl.get_mean_and_var(None, gamma, x)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_mnbn_link_matches_golden_and_torch_batchnorm(comm):
    """tests/chainermn_tests/links_tests/test_batch_normalization.py:54-186 at one
    rank: MNBN == plain BN on the same batch (golden single-process reference
    values from the unmodified reference)."""
    import os
    import torch
    from chainer_b200.links import MultiNodeBatchNormalization
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'mnbn.npz'))
    x = torch.from_numpy(z['x']).cuda().requires_grad_(True)
    C = z['gamma'].size
    bn = MultiNodeBatchNormalization(C, comm)
    bn.gamma.data.copy_(torch.from_numpy(z['gamma']))
    bn.beta.data.copy_(torch.from_numpy(z['beta']))
    bn.gamma.data.requires_grad_(True)
    bn.beta.data.requires_grad_(True)
    y = bn(x)
    y.backward(torch.from_numpy(z['gy']).cuda())
    np.testing.assert_allclose(y.detach().cpu().numpy(), z['single|y'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), z['single|gx'], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(bn.gamma.data.grad.cpu().numpy(), z['single|ggamma'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(bn.beta.data.grad.cpu().numpy(), z['single|gbeta'], rtol=1e-4, atol=1e-5)
    assert float(bn.avg_var.abs().sum()) > 0 and float(bn.avg_mean.abs().sum()) > 0


def test_full_size_resnet50_through_api(comm):
    """BASELINE configs[1] at full size through the public API; checked against
    torch elementwise arithmetic (size-independent property: the result of the
    fused path equals param - lr * mean_grad after one step from v = 0)."""
    import torch
    from chainer_b200 import workloads
    from chainer_b200.core.link import link_from_named_arrays
    plist = workloads.resnet50()
    torch.manual_seed(0)
    data = [torch.randn(s, device='cuda') * 0.05 for _, s in plist]
    keep = [d.clone() for d in data]
    model = link_from_named_arrays([(n, d) for (n, _), d in zip(plist, data)])
    opt = chainer_b200.create_multi_node_optimizer(chainer_b200.MomentumSGD(lr=0.01), comm)
    opt.setup(model)
    opt.update()
    grads = [torch.randn(s, device='cuda') * 1e-2 for _, s in plist]
    for (_, p), g in zip(sorted(model.namedparams()), grads):
        p.grad = g.clone()
    opt.update()
    torch.cuda.synchronize()
    lr = torch.tensor(0.01, device='cuda')
    for (name, p), d0, g in zip(sorted(model.namedparams()), keep, grads):
        assert torch.equal(p.update_rule.state['v'], -(lr * g)), name
        assert torch.equal(p.data, d0 + (-(lr * g))), name
        assert torch.equal(p.grad, g), name


def test_double_buffering_optimizer_on_gpu(comm):
    """tests/chainermn_tests/optimizer_tests/test_double_buffering_optimizer.py:43-90
    on the device: the mean lands in communicated_target one call late, computed
    on the non-blocking side stream."""
    import torch
    rng = np.random.default_rng(4)
    model, host = _model(sorted(PLIST)[:6], rng)
    actual = chainer_b200.MomentumSGD(lr=0.1)
    opt = chainer_b200.create_multi_node_optimizer(actual, comm, double_buffering=True)
    opt.setup(model)
    grads = [np.asarray(rng.standard_normal(a.shape)).astype(np.float32).reshape(a.shape) for a in host]
    _set_grads(model, grads)
    opt.update()                                        # bcast + deep copy
    assert actual.t == 0 and opt.communicated_target is not None
    _set_grads(model, grads)
    opt.update()                                        # swap + async mean; no update yet
    assert actual.t == 0
    opt.wait()
    for (_, p), g in zip(sorted(opt.communicated_target.namedparams()), grads):
        assert_bits_equal(p.grad.cpu().numpy(), g, 'communicated_target grad')
    _set_grads(model, grads)
    opt.update()
    opt.wait()
    assert actual.t == 1


@pytest.mark.parametrize('opt_name', ['momentum_sgd', 'adam'])
def test_standalone_multi_tensor_update_on_gpu(opt_name):
    """optimizer.update() without a communicator (buffer == NULL kernels), mixed
    float16 / float32 parameters, full-size-ish tensors: bit-exact vs the oracle."""
    import torch
    from chainer_b200.core.link import link_from_named_arrays
    rng = np.random.default_rng(21)
    spec = [('/a/W', (300, 200), np.float32), ('/a/b', (7,), np.float32),
            ('/b/W', (128, 64), np.float16), ('/b/b', (64,), np.float16),
            ('/c/W', (100000,), np.float32)]
    host = [np.asarray(rng.standard_normal(s) * 0.05).astype(dt) for _, s, dt in spec]
    model = link_from_named_arrays([(n, _t(a)) for (n, _, _), a in zip(spec, host)])
    opt = chainer_b200.MomentumSGD(lr=0.01) if opt_name == 'momentum_sgd' else chainer_b200.Adam()
    opt.setup(model)
    st = [dict(m=np.zeros_like(a), v=np.zeros_like(a)) for a in host]
    for step in range(1, 4):
        grads = [np.asarray(rng.standard_normal(a.shape) * (0.5 if a.dtype == np.float16 else 1e-2))
                 .astype(a.dtype) for a in host]
        for (_, p), g in zip(sorted(model.namedparams()), grads):
            p.grad = _t(g)
        opt.update()
        torch.cuda.synchronize()
        for (name, p), q, g, s in zip(sorted(model.namedparams()), host, grads, st):
            if opt_name == 'momentum_sgd':
                og.momentum_sgd_update(q, g, s['v'], 0.01, 0.9)
            else:
                og.adam_update_gpu(q, g, s['m'], s['v'], step)
            assert_bits_equal(p.data.cpu().numpy(), q, name)
            assert p.update_rule.t == step


def test_allreduce_persistent_on_gpu(comm):
    """AllreducePersistent over a model with MultiNodeBatchNormalization links: one
    packed allreduce per buffer dtype; with one worker the mean is the identity,
    bit-for-bit, and the integer persistent N is left alone."""
    import torch
    from chainer_b200 import _lib
    from chainer_b200.core import link as L
    from chainer_b200.extensions import AllreducePersistent
    from chainer_b200.links.batch_normalization import MultiNodeBatchNormalization

    class Net(L.Chain):
        def __init__(self):
            super(Net, self).__init__()
            with self.init_scope():
                self.bn1 = MultiNodeBatchNormalization(64, comm)
                self.bn2 = MultiNodeBatchNormalization(1000, comm)
                self.bn3 = MultiNodeBatchNormalization(7, comm, dtype=np.float16)

    net = Net()
    gen = torch.Generator(device='cuda').manual_seed(5)
    want = {}
    for name, bn in (('bn1', net.bn1), ('bn2', net.bn2), ('bn3', net.bn3)):
        for attr in ('avg_mean', 'avg_var'):
            t = getattr(bn, attr)
            t.copy_(torch.randn(t.shape, device='cuda', generator=gen).to(t.dtype))
            want[name, attr] = t.clone()
        bn.N = 11
    assert sorted(n for n, _ in net.namedlinks()) == ['/', '/bn1', '/bn2', '/bn3']
    lib = _lib.get()
    before = lib.launches
    AllreducePersistent(net, comm)()
    torch.cuda.synchronize()
    assert lib.launches - before <= 4            # pack, scale, unpack (+ nothing per array)
    for (name, attr), t in want.items():
        assert torch.equal(getattr(getattr(net, name), attr), t), (name, attr)
    assert net.bn1.N == 11
