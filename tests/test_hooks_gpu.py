"""GPU parity tests of the optimizer-hook kernels (csrc/gp_hooks.cu, gp_sgd_hooks.cu,
gp_adam_hooks.cu) against the NumPy oracle, whose hook / loss-scale restatement is
pinned bit-for-bit to the unmodified reference by tests/test_oracle_golden.py.

Elementwise work (decay, rate, loss-scale division inside the fused updates and
as stand-alone kernels) is BIT-EXACT.  The squared norm is a double-precision
deterministic reduction: compared with NumPy's float64 sum to 1e-12 relative, and
the rate derived from it is checked exactly from the returned sum.
"""
import ctypes
import struct

import numpy as np
import pytest

from tests.helpers import P, assert_bits_equal, to_dev, to_host
from tests.test_kernels_gpu import ALIGNED, RAGGED, _odt, _torch_buf
from tests.test_oracle_golden import HOOK_CASES

pytestmark = pytest.mark.gpu

HOOKSETS = {
    #          rate (device float) | decay | loss_scale
    'wd': (None, 0.05, None),
    'ls': (None, None, 100.0),
    'wd_ls128': (None, 0.05 * 128.0, 128.0),
    'rate': (0.37, None, None),
    'all': (0.61, 0.05 * 100.0, 100.0),
}


@pytest.fixture(params=[(256, 0, 0), (256, 4, 1), (128, 2, 0)], ids=['default', 't256u4p', 't128u2'])
def walker_tuning(request):
    from chainer_b200 import _lib
    lib = _lib.get()
    threads, unroll, persistent = request.param
    lib.gp_set_tuning(b'threads', threads)
    lib.gp_set_tuning(b'unroll', unroll)
    lib.gp_set_tuning(b'persistent', persistent)
    yield request.param
    lib.gp_set_tuning(b'threads', 256)
    lib.gp_set_tuning(b'unroll', 0)
    lib.gp_set_tuning(b'persistent', 0)


class _Hooks(object):
    """A gp_hooks_t with its device-side rate."""

    def __init__(self, rate, decay, loss_scale):
        import torch
        from chainer_b200 import _lib
        self.rate_dev = None
        self.struct = _lib.GpHooks()
        if rate is not None:
            self.rate_dev = torch.tensor([0.0, rate, 0.0], dtype=torch.float32, device='cuda')
            self.struct.clip_rate = self.rate_dev.data_ptr() + 4
        self.struct.weight_decay = decay if decay is not None else 0.0
        self.struct.loss_scale = loss_scale if loss_scale is not None else 0.0
        self.addr = ctypes.addressof(self.struct)
        self.rate, self.decay, self.loss_scale = rate, decay, loss_scale

    def oracle(self, g, p):
        """The oracle's hook sequence on a fresh gradient array (dtype of p)."""
        from oracle import gradpath as og
        g = np.array(g)
        if self.rate is not None:
            g *= g.dtype.type(np.float32(self.rate))
        if self.decay is not None:
            og.weight_decay_hook(p, g, self.decay)
        if self.loss_scale is not None:
            og.loss_scale_divide(g, self.loss_scale)
        return g


@pytest.mark.parametrize('buf_dtype', ['float32', 'float16', 'bfloat16'])
@pytest.mark.parametrize('sizes,pdtype', [(ALIGNED, 'float32'), (RAGGED, 'float32'),
                                          (RAGGED, 'float16'), (RAGGED, 'float64')])
@pytest.mark.parametrize('hookset', sorted(HOOKSETS))
@pytest.mark.parametrize('n_ranks', [1, 3])
def test_hooked_momentum_sgd_bit_exact(buf_dtype, sizes, pdtype, hookset, n_ranks, walker_tuning):
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from chainer_b200.communicators import _memory_utility as mu
    from oracle import gradpath as og
    lib = _lib.get()
    rng = np.random.default_rng(123)
    pdt = np.dtype(pdtype)
    hp = [(rng.standard_normal(n) * 0.05).astype(pdt) for n in sizes]
    hv = [np.zeros_like(p) for p in hp]
    d_p = [to_dev(a) for a in hp]
    d_v = [to_dev(a) for a in hv]
    n = sum(sizes)
    hooks = _Hooks(*HOOKSETS[hookset])
    gscale = 1e-2 * (hooks.loss_scale or 1.0)
    for step in range(2):
        summed = og.cast(rng.standard_normal(n) * gscale * n_ranks, _odt(buf_dtype))
        buf = _torch_buf(n, buf_dtype)
        buf[:n] = to_dev(summed).to(buf.dtype)
        params = [P(data=d_p[i], grad=torch.full_like(d_p[i], 3.0)) for i in range(len(sizes))]
        pd = mu.ParamsData(params, 'grad', False,
                           extra_ptrs=[(d_p[i], [d_v[i]]) for i in range(len(sizes))])
        lib.gp_unpack_momentum_sgd_hooked(buf.data_ptr(), dev.dtype_id(_odt(buf_dtype)), pd.d_csum,
                                          pd.d_segs, pd.n_params, 0, n, 1.0 / n_ranks, 0.01, 0.9,
                                          1, pd.layout_hint(_odt(buf_dtype)), hooks.addr, 0)
        torch.cuda.synchronize()
        g = og.mean_grad_value(summed, _odt(buf_dtype), n_ranks, pdt)
        cs = og.size_csum(hp)
        for i in range(len(sizes)):
            gi = hooks.oracle(g[cs[i]:cs[i + 1]], hp[i])
            og.momentum_sgd_update(hp[i], gi, hv[i], 0.01, 0.9)
            assert_bits_equal(to_host(params[i].grad), gi, 'grad step %d' % step)
            assert_bits_equal(to_host(d_p[i]), hp[i], 'param step %d' % step)
            assert_bits_equal(to_host(d_v[i]), hv[i], 'v step %d' % step)


@pytest.mark.parametrize('buf_dtype', ['float32', 'float16'])
@pytest.mark.parametrize('sizes,pdtype', [(ALIGNED, 'float32'), (RAGGED, 'float32'),
                                          (RAGGED, 'float16'), (RAGGED, 'float64')])
@pytest.mark.parametrize('amsgrad', [False, True])
@pytest.mark.parametrize('hookset', ['wd', 'all'])
def test_hooked_adam_bit_exact(buf_dtype, sizes, pdtype, amsgrad, hookset, walker_tuning):
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from chainer_b200.communicators import _memory_utility as mu
    from oracle import gradpath as og
    lib = _lib.get()
    rng = np.random.default_rng(321)
    pdt = np.dtype(pdtype)
    # float16: gradients of O(1) so that v stays a normal half number
    base = 0.5 if pdtype == 'float16' else 1e-2
    hooks = _Hooks(*HOOKSETS[hookset])
    gscale = base * (hooks.loss_scale or 1.0)
    if pdtype == 'float16':
        gscale = min(gscale, 40.0)          # |g| * loss_scale must stay below 65504
    hp = [(rng.standard_normal(n) * 0.05).astype(pdt) for n in sizes]
    st = [dict(m=np.zeros_like(p), v=np.zeros_like(p), vhat=np.zeros_like(p)) for p in hp]
    d_p = [to_dev(a) for a in hp]
    d_s = [[to_dev(s['m']), to_dev(s['v']), to_dev(s['vhat'])] for s in st]
    n = sum(sizes)
    n_ranks = 2
    alpha, b1, b2, eps = 0.001, 0.9, 0.999, 1e-8
    for t in range(1, 3):
        summed = og.cast(rng.standard_normal(n) * gscale * n_ranks, _odt(buf_dtype))
        buf = _torch_buf(n, buf_dtype)
        buf[:n] = to_dev(summed).to(buf.dtype)
        params = [P(data=d_p[i], grad=torch.full_like(d_p[i], 3.0)) for i in range(len(sizes))]
        extra = [(d_p[i], d_s[i][:3 if amsgrad else 2]) for i in range(len(sizes))]
        pd = mu.ParamsData(params, 'grad', False, extra_ptrs=extra)
        alpha_t = og.adam_alpha_t(alpha, b1, b2, t)
        lib.gp_unpack_adam_hooked(buf.data_ptr(), dev.dtype_id(_odt(buf_dtype)), pd.d_csum,
                                  pd.d_segs, pd.n_params, 0, n, 1.0 / n_ranks, alpha_t, 1 - b1,
                                  1 - b2, eps, 1.0, 0.0, 0.0, 0.0, 1 if amsgrad else 0, 1,
                                  pd.layout_hint(_odt(buf_dtype)), hooks.addr, 0)
        torch.cuda.synchronize()
        g = og.mean_grad_value(summed, _odt(buf_dtype), n_ranks, pdt)
        cs = og.size_csum(hp)
        for i in range(len(sizes)):
            gi = hooks.oracle(g[cs[i]:cs[i + 1]], hp[i])
            og.adam_update_gpu(hp[i], gi, st[i]['m'], st[i]['v'], t, alpha=alpha, beta1=b1,
                               beta2=b2, eps=eps, amsgrad=amsgrad,
                               vhat=st[i]['vhat'] if amsgrad else None)
            assert_bits_equal(to_host(params[i].grad), gi, 'grad t %d' % t)
            assert_bits_equal(to_host(d_p[i]), hp[i], 'param t %d' % t)
            assert_bits_equal(to_host(d_s[i][0]), st[i]['m'], 'm t %d' % t)
            assert_bits_equal(to_host(d_s[i][1]), st[i]['v'], 'v t %d' % t)
            if amsgrad:
                assert_bits_equal(to_host(d_s[i][2]), st[i]['vhat'], 'vhat t %d' % t)


def _read_out(lib, out_ptr):
    host = (ctypes.c_char * 16)()
    lib.gp_memcpy_async(ctypes.addressof(host), out_ptr, 16, 1, 0)
    lib.gp_stream_synchronize(0)
    return struct.unpack('dff', bytes(host))      # sqsum, rate, norm


def _expected_rate(sqsum, threshold):
    norm = np.float32(np.sqrt(np.float64(sqsum)))
    with np.errstate(divide='ignore'):
        rate = np.float32(threshold) / norm
    return min(rate, np.float32(1.0)), norm


@pytest.mark.parametrize('dtype', ['float32', 'float16', 'bfloat16', 'float64'])
@pytest.mark.parametrize('n', [1, 7, 4096, 100003, 3000000])
@pytest.mark.parametrize('scale', [1.0, 0.125, 1.0 / 3.0])
def test_sqnorm_and_rate(dtype, n, scale):
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from chainer_b200.optimizer_hooks import _NormScratch
    from oracle import gradpath as og
    lib = _lib.get()
    rng = np.random.default_rng(5 + n)
    host = og.cast(rng.standard_normal(n) * 0.3, _odt(dtype))
    buf = _torch_buf(n, dtype)
    buf[:n] = to_dev(host).to(buf.dtype)
    sc = _NormScratch()
    did = dev.dtype_id(_odt(dtype))
    mean = np.asarray(og.scale_buffer(np.array(host), _odt(dtype), scale), dtype=np.float64)
    want = float(np.sum(mean * mean))
    for threshold in (0.05, 1e6):
        lib.gp_sqnorm(buf.data_ptr(), did, n, scale, 0, threshold, sc.ws, sc.out, 0)
        sqsum, rate, norm = _read_out(lib, sc.out)
        assert abs(sqsum - want) <= 1e-12 * want
        erate, enorm = _expected_rate(sqsum, threshold)
        assert np.float32(norm) == enorm and np.float32(rate) == erate
        assert rate <= 1.0
        # deterministic: same bits on every run (and so on every rank)
        lib.gp_sqnorm(buf.data_ptr(), did, n, scale, 0, threshold, sc.ws, sc.out, 0)
        assert _read_out(lib, sc.out)[0] == sqsum
    # accumulate: per-array use by the unfused hook
    lib.gp_sqnorm(buf.data_ptr(), did, n, scale, 1, 0.05, sc.ws, sc.out, 0)
    sq2, rate2, _ = _read_out(lib, sc.out)
    assert sq2 == sqsum + sqsum
    assert np.float32(rate2) == _expected_rate(sq2, 0.05)[0]
    torch.cuda.synchronize()


def test_sqnorm_of_zeros_gives_rate_one():
    from chainer_b200 import _lib
    from chainer_b200.optimizer_hooks import _NormScratch
    lib = _lib.get()
    buf = _torch_buf(1000, 'float32')
    sc = _NormScratch()
    lib.gp_sqnorm(buf.data_ptr(), 7, 1000, 1.0, 0, 0.5, sc.ws, sc.out, 0)
    sqsum, rate, norm = _read_out(lib, sc.out)
    assert sqsum == 0.0 and norm == 0.0 and rate == 1.0


@pytest.mark.parametrize('dtype', ['float32', 'float16', 'float64'])
@pytest.mark.parametrize('n', [1, 5, 4096, 70001])
def test_flat_hook_kernels_bit_exact(dtype, n):
    """gp_weight_decay, gp_scale_by_device, gp_divide: the unfused hook kernels."""
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from oracle import gradpath as og
    lib = _lib.get()
    dt = np.dtype(dtype)
    rng = np.random.default_rng(n)
    g = (rng.standard_normal(n) * 0.7).astype(dt)
    p = (rng.standard_normal(n) * 0.05).astype(dt)
    d_g, d_p = to_dev(g), to_dev(p)
    did = dev.dtype_id(dt)
    lib.gp_weight_decay(d_g.data_ptr(), d_p.data_ptr(), did, n, 0.05, 0)
    og.weight_decay_hook(p, g, 0.05)
    assert_bits_equal(to_host(d_g), g, 'weight decay')
    rate = torch.tensor([0.4321], dtype=torch.float32, device='cuda')
    lib.gp_scale_by_device(d_g.data_ptr(), did, n, rate.data_ptr(), 0)
    g *= dt.type(np.float32(0.4321))
    assert_bits_equal(to_host(d_g), g, 'scale by device factor')
    lib.gp_divide(d_g.data_ptr(), did, n, 100.0, 0)
    og.loss_scale_divide(g, 100.0)
    assert_bits_equal(to_host(d_g), g, 'divide')


@pytest.mark.parametrize('dtype', ['float32', 'float16'])
@pytest.mark.parametrize('variant', sorted(HOOK_CASES))
def test_hooks_through_the_public_api(variant, dtype):
    """hooks.npz (the unmodified reference with WeightDecay / GradientClipping / loss
    scaling) through create_multi_node_optimizer on the GPU."""
    from tests.hooks_scenario import run_hooks_scenario
    opt_name, spec, ls = HOOK_CASES[variant]
    if opt_name == 'adam' and dtype == 'float16':
        pytest.skip('no reference vector: float16 CPU Adam underflows (make_golden.py)')
    fusable = [k for k, _ in spec] != ['wd', 'clip']

    def after_step(comm):
        assert (comm._fused_plan is not None) == fusable

    run_hooks_scenario(variant, dtype, lambda a: to_dev(np.array(a)), to_host, None, after_step)


def test_clipping_hook_standalone_and_norm_readback():
    """optimizer.update() without a communicator: the unfused hook kernels, and the
    norm the hook saw."""
    import chainer_b200
    from chainer_b200 import optimizer_hooks as H
    from chainer_b200.core import link as L
    from oracle import gradpath as og
    rng = np.random.default_rng(3)
    shapes = [(33,), (4, 5), (1000,)]
    hp = [(rng.standard_normal(s) * 0.05).astype(np.float32) for s in shapes]
    hg = [(rng.standard_normal(s) * 0.1).astype(np.float32) for s in shapes]
    model = L.link_from_named_arrays([('/p%d' % i, to_dev(a)) for i, a in enumerate(hp)])
    opt = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9)
    opt.setup(model)
    clip = H.GradientClipping(0.5)
    opt.add_hook(clip)
    opt.add_hook(H.WeightDecay(0.01))
    for (_, p), g in zip(sorted(model.namedparams()), hg):
        p.grad = to_dev(g)
    opt.update()
    want_norm = float(np.sqrt(sum(float(np.sum(g.astype(np.float64) ** 2)) for g in hg)))
    assert abs(clip.last_norm() - want_norm) <= 1e-6 * want_norm
    og.gradient_clipping_hook(hg, 0.5)
    hv = [np.zeros_like(p) for p in hp]
    for (_, p), q, g, v in zip(sorted(model.namedparams()), hp, hg, hv):
        og.weight_decay_hook(q, g, 0.01)
        og.momentum_sgd_update(q, g, v)
        np.testing.assert_allclose(to_host(p.data), q, rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(to_host(p.grad), g, rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize('multi_node', [True, False])
@pytest.mark.parametrize('dtype', ['float32', 'float16'])
def test_dynamic_loss_scaling_gpu(dtype, multi_node):
    from tests.hooks_scenario import run_dynamic_loss_scale
    run_dynamic_loss_scale(dtype, multi_node, lambda a: to_dev(np.array(a)), to_host)


@pytest.mark.parametrize('multi_node', [True, False])
@pytest.mark.parametrize('case', ['sgd', 'sgd_wd_ls128', 'adam'])
def test_fp32_master_weights_gpu(case, multi_node):
    from tests.hooks_scenario import run_fp32_update
    run_fp32_update(case, multi_node, lambda a: to_dev(np.array(a)), to_host)


@pytest.mark.parametrize('case', ['sgd_wd', 'adam'])
@pytest.mark.parametrize('multi_node', [False, True])
def test_fp32_master_with_dynamic_loss_scaling_gpu(case, multi_node):
    """float16 parameters + float32 masters + dynamic loss scaling against the reference's
    vectors (fp32_dynamic.npz): the fused master kernels with the device-side skip word."""
    from tests.hooks_scenario import run_fp32_dynamic
    run_fp32_dynamic(case, multi_node, lambda a: to_dev(np.array(a)), to_host)
