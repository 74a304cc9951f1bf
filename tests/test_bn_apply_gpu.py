"""GPU parity tests of the elementwise halves of batch normalisation
(csrc/gp_bn_apply.cu: gp_bn_fwd_apply, gp_bn_bwd_apply) through the C-ABI against the NumPy
oracle -- the reference's `bn_fwd`, `update_mean_var` and `bn_bwd` kernels restated
operation by operation -- and against a float64 evaluation.  Tolerance (written here): the
north star's 1e-6 relative in float32 (1e-3 in float16), with an absolute floor of the same
size times the magnitude of the terms that cancel.  Shapes: the ResNet-50 planes (112x112
... 7x7, BASELINE configs[2]) plus ragged / tiny ones."""
import numpy as np
import pytest

from tests.helpers import to_dev, to_host

pytestmark = pytest.mark.gpu

SHAPES = [(4, 3, 5, 7), (8, 64, 14, 14), (2, 5, 1, 1), (32, 2048, 7, 7), (3, 16, 8, 8),
          (2, 64, 112, 112), (1, 1, 1, 3), (5, 7, 2, 2)]


def _tol(dtype):
    return {'float32': 2e-6, 'float16': 2e-3, 'float64': 1e-12}[dtype]


def _floor(dtype):
    """Absolute floor: the spacing of the output type's subnormals (float16: 6e-8)."""
    return {'float32': 1e-30, 'float16': 1e-7, 'float64': 1e-300}[dtype]


@pytest.mark.parametrize('dtype', ['float32', 'float16', 'float64'])
@pytest.mark.parametrize('shape', SHAPES, ids=[str(s) for s in SHAPES])
@pytest.mark.parametrize('with_running', [False, True])
def test_bn_fwd_apply(dtype, shape, with_running):
    import torch
    from chainer_b200.functions import batch_normalization as bnf
    from oracle import gradpath as og
    rng = np.random.default_rng(3)
    sdt = np.float64 if dtype == 'float64' else np.float32
    C = shape[1]
    x = (rng.standard_normal(shape) * 2 + 0.5).astype(dtype)
    mean = rng.standard_normal(C).astype(sdt) * 0.3
    var = (rng.random(C) + 0.1).astype(sdt)
    gamma = (rng.standard_normal(C) + 1).astype(sdt)
    beta = rng.standard_normal(C).astype(sdt)
    rm = rng.standard_normal(C).astype(sdt)
    rv = (rng.random(C) + 0.5).astype(sdt)
    eps, decay, adjust = 2e-5, 0.9, 1.25
    d_rm, d_rv = to_dev(rm), to_dev(rv)
    y, inv_std = bnf.fwd_apply(to_dev(x), to_dev(mean), to_dev(var), to_dev(gamma), to_dev(beta),
                               eps, d_rm if with_running else None,
                               d_rv if with_running else None, decay, adjust)
    torch.cuda.synchronize()
    want = og.bn_fwd_apply(x, mean, var, gamma, beta, eps)
    sh = (1, -1) + (1,) * (x.ndim - 2)
    x64 = x.astype(np.float64)
    istd64 = 1.0 / np.sqrt(var.astype(np.float64) + eps)
    y64 = gamma.reshape(sh) * (x64 - mean.reshape(sh)) * istd64.reshape(sh) + beta.reshape(sh)
    mag = np.abs(gamma.reshape(sh) * (x64 - mean.reshape(sh)) * istd64.reshape(sh)) + np.abs(beta.reshape(sh))
    t = _tol(dtype)
    got = to_host(y).astype(np.float64)
    assert np.all(np.abs(got - y64) <= 4 * t * mag + _floor(dtype)), float(np.abs(got - y64).max())
    assert np.all(np.abs(got - want.astype(np.float64)) <= 2 * t * mag + _floor(dtype))
    np.testing.assert_allclose(to_host(inv_std), istd64, rtol=2e-7 if sdt is np.float32 else 1e-15)
    if with_running:
        og.bn_running_update(rm, rv, mean, var, decay, adjust)
        np.testing.assert_allclose(to_host(d_rm), rm, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(to_host(d_rv), rv, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('dtype', ['float32', 'float16', 'float64'])
@pytest.mark.parametrize('shape', SHAPES, ids=[str(s) for s in SHAPES])
def test_bn_bwd_apply(dtype, shape):
    import torch
    from chainer_b200.functions import batch_normalization as bnf
    from oracle import gradpath as og
    rng = np.random.default_rng(4)
    sdt = np.float64 if dtype == 'float64' else np.float32
    C = shape[1]
    x = (rng.standard_normal(shape) * 2 + 0.5).astype(dtype)
    gy = (rng.standard_normal(shape) * 1e-2).astype(dtype)
    mean = rng.standard_normal(C).astype(sdt) * 0.3
    inv_std = (1.0 / np.sqrt(rng.random(C) + 0.1)).astype(sdt)
    gamma = (rng.standard_normal(C) + 1).astype(sdt)
    ggamma = rng.standard_normal(C).astype(sdt) * 1e-2
    gbeta = rng.standard_normal(C).astype(sdt) * 1e-2
    gx = bnf.bwd_apply(to_dev(gy), to_dev(x), to_dev(mean), to_dev(inv_std), to_dev(gamma),
                       to_dev(ggamma), to_dev(gbeta))
    torch.cuda.synchronize()
    m = x.size // C
    inv_m = sdt(1.0 / m)
    want = og.bn_bwd_apply(gy, x, mean, inv_std, gamma, ggamma, gbeta, inv_m)
    sh = (1, -1) + (1,) * (x.ndim - 2)
    x64, g64 = x.astype(np.float64), gy.astype(np.float64)
    xh = (x64 - mean.reshape(sh)) * inv_std.reshape(sh)
    t64 = (xh * ggamma.reshape(sh) + gbeta.reshape(sh)) * float(inv_m)
    gx64 = (gamma * inv_std).reshape(sh) * (g64 - t64)
    mag = np.abs((gamma * inv_std).reshape(sh)) * (np.abs(g64) + np.abs(t64))
    t = _tol(dtype)
    got = to_host(gx).astype(np.float64)
    assert np.all(np.abs(got - gx64) <= 6 * t * mag + _floor(dtype)), float(np.abs(got - gx64).max())
    assert np.all(np.abs(got - want.astype(np.float64)) <= 2 * t * mag + _floor(dtype))


@pytest.mark.parametrize('shape', [(8, 32, 14, 14), (4, 6, 5, 5)])
def test_batch_normalization_link_matches_torch(shape):
    """The single-worker link (statistics + apply kernels, 2 launches each way) against
    torch.nn.functional.batch_norm: y, gx, ggamma, gbeta and the running statistics."""
    import torch
    from chainer_b200.links import BatchNormalization
    torch.manual_seed(0)
    C = shape[1]
    x = torch.randn(*shape, device='cuda', requires_grad=True)
    gy = torch.randn(*shape, device='cuda')
    bn = BatchNormalization(C)
    bn.gamma.data.copy_(torch.rand(C) + 0.5)
    bn.beta.data.copy_(torch.randn(C))
    bn.gamma.data.requires_grad_(True)
    bn.beta.data.requires_grad_(True)
    y = bn(x)
    y.backward(gy)
    x2 = x.detach().clone().requires_grad_(True)
    g2 = bn.gamma.data.detach().clone().requires_grad_(True)
    b2 = bn.beta.data.detach().clone().requires_grad_(True)
    rm, rv = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    y2 = torch.nn.functional.batch_norm(x2, rm, rv, g2, b2, training=True, momentum=0.1, eps=2e-5)
    y2.backward(gy)
    torch.testing.assert_close(y, y2, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(x.grad, x2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(bn.gamma.data.grad, g2.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(bn.beta.data.grad, b2.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(bn.avg_mean, rm, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn.avg_var, rv, rtol=1e-5, atol=1e-6)


def test_bn_link_replays_from_a_cuda_graph():
    """Forward + backward of the BN link (4 library launches) captured once into a CUDA graph
    on a side stream and replayed: same y / gx / running statistics as eager calls.  The
    kernels launch on torch's current stream and keep no host-side per-call state."""
    import torch
    from chainer_b200.links import BatchNormalization
    torch.manual_seed(1)
    shape = (8, 32, 14, 14)
    bn_e, bn_g = BatchNormalization(32), BatchNormalization(32)
    for bn in (bn_e, bn_g):
        bn.gamma.data.copy_(torch.linspace(0.5, 1.5, 32))
        bn.beta.data.copy_(torch.linspace(-1, 1, 32))
    xs = [torch.randn(*shape, device='cuda') for _ in range(3)]
    gys = [torch.randn(*shape, device='cuda') for _ in range(3)]
    # eager reference
    want = []
    for x, gy in zip(xs, gys):
        xv = x.clone().requires_grad_(True)
        y = bn_e(xv)
        y.backward(gy)
        want.append((y.detach().clone(), xv.grad.clone()))
    # graph: static input / output buffers
    sx = torch.zeros(*shape, device='cuda', requires_grad=True)
    sgy = torch.zeros(*shape, device='cuda')
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        warm = BatchNormalization(32)        # allocates the statistics workspace outside the capture
        warm(sx).backward(sgy)
        sx.grad = None
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        sy = bn_g(sx)
        sy.backward(sgy)
    for (x, gy), (wy, wgx) in zip(zip(xs, gys), want):
        sx.data.copy_(x)
        sgy.copy_(gy)
        g.replay()
        torch.cuda.synchronize()
        torch.testing.assert_close(sy, wy, rtol=0, atol=0)
        torch.testing.assert_close(sx.grad, wgx, rtol=0, atol=0)
    # (one extra step went into bn_g's running statistics during the capture-free warm-up of
    # `warm`, none into bn_g: capture does not execute) -> 3 replays == 3 eager steps
    torch.testing.assert_close(bn_g.avg_mean, bn_e.avg_mean, rtol=0, atol=0)
    torch.testing.assert_close(bn_g.avg_var, bn_e.avg_var, rtol=0, atol=0)
