"""GPU parity tests of the libgradpath kernels against the NumPy oracle.

Every call goes through the C-ABI (ctypes); integer/byte-level results (the
pack layout, casts, the fused updates which avoid FMA contraction) are compared
BIT-EXACTLY with the oracle.
"""
import ctypes

import numpy as np
import pytest

from tests.helpers import P, assert_bits_equal, to_dev, to_host

pytestmark = pytest.mark.gpu

BUF_DTYPES = ['float32', 'float16', 'bfloat16', 'float64']
RAGGED = [7, 1, 0, 1000, 4096, 12345, 64, 3, 513, 2048, 70001]
ALIGNED = [64, 64, 9408, 64, 2048, 1000, 256, 16384, 36864, 4, 131072]
ALIGNED8 = [64, 8, 9408, 64, 2048, 1000, 256, 16384, 36864, 8, 131072, 0, 24, 70000]


def _torch_buf(n, dtype):
    import torch
    td = {'float32': torch.float32, 'float16': torch.float16, 'bfloat16': torch.bfloat16,
          'float64': torch.float64}[dtype]
    return torch.zeros(max(n, 1), dtype=td, device='cuda')


def _odt(dtype):
    from oracle import gradpath as og
    return og.BF16 if dtype == 'bfloat16' else np.dtype(dtype)


def _make_arrays(sizes, dtypes, rng, scale=1.0):
    out = []
    for i, n in enumerate(sizes):
        dt = np.dtype(dtypes[i % len(dtypes)])
        out.append((rng.standard_normal(n) * scale).astype(dt))
    return out


class _Buf(object):
    def __init__(self, t):
        self.t = t

    def ptr(self):
        return self.t.data_ptr()


@pytest.fixture(params=[(256, 0, 0, 0, 4096, 4), (256, 4, 1, 1, 2048, 4), (128, 2, 1, 0, 2048, 4),
                        (512, 2, 0, 1, 512, 2), (256, 4, 0, 1, 4096, 3)],
                ids=['default', 't256u4p-bulk', 't128u2p-nobulk', 't512u2np-bulk512x2',
                     't256u4np-bulk4096x3'])
def tuning(request):
    """(threads, unroll, persistent) of the register-path walker and
    (enable, tile, stages) of the TMA-staged kernels."""
    from chainer_b200 import _lib
    lib = _lib.get()
    threads, unroll, persistent, bulk, tile, stages = request.param
    lib.gp_set_tuning(b'threads', threads)
    lib.gp_set_tuning(b'unroll', unroll)
    lib.gp_set_tuning(b'persistent', persistent)
    lib.gp_set_tuning(b'bulk', bulk)
    lib.gp_set_tuning(b'bulk_tile', tile)
    lib.gp_set_tuning(b'bulk_stages', stages)
    yield request.param
    lib.gp_set_tuning(b'threads', 256)
    lib.gp_set_tuning(b'unroll', 0)
    lib.gp_set_tuning(b'persistent', 0)
    lib.gp_set_tuning(b'bulk', 0)
    lib.gp_set_tuning(b'bulk_tile', 4096)
    lib.gp_set_tuning(b'bulk_stages', 4)


@pytest.mark.parametrize('buf_dtype', BUF_DTYPES)
@pytest.mark.parametrize('sizes,dtypes', [
    (RAGGED, ['float32']), (ALIGNED, ['float32']), (ALIGNED, ['float16']),
    (RAGGED, ['float32', 'float16', 'float64']), (ALIGNED, ['float16', 'float32']),
    (ALIGNED, ['float64']),
], ids=['ragged32', 'aligned32', 'aligned16', 'raggedmix', 'alignedmix', 'aligned64'])
@pytest.mark.parametrize('scale', [1.0, 0.125, 1.0 / 3.0])
def test_pack_unpack_bit_exact(buf_dtype, sizes, dtypes, scale, tuning):
    import torch
    from chainer_b200.communicators import _memory_utility as mu
    from oracle import gradpath as og
    rng = np.random.default_rng(1234)
    host = _make_arrays(sizes, dtypes, rng)
    params = [P(data=to_dev(a), grad=to_dev(a)) for a in host]
    n = sum(sizes)
    buf = _torch_buf(n, buf_dtype)
    pd = mu.ParamsData(params, 'grad', False)
    assert pd.n_elems == n
    mu._batched_pack_params(pd, _Buf(buf), _odt(buf_dtype), scale=scale)
    torch.cuda.synchronize()
    want = og.pack(host, _odt(buf_dtype), scale)
    assert_bits_equal(to_host(buf)[:n], want, 'pack')

    # unpack with descale back into fresh arrays (poisoned first)
    outs = [P(data=p.data, grad=torch.full_like(p.grad, 777.0)) for p in params]
    pd2 = mu.ParamsData(outs, 'grad', False)
    mu._batched_unpack_params(pd2, _Buf(buf), _odt(buf_dtype), scale=scale)
    torch.cuda.synchronize()
    scaled = og.scale_buffer(want, _odt(buf_dtype), scale)
    want_arrays = og.unpack(scaled, host, _odt(buf_dtype))
    for o, w in zip(outs, want_arrays):
        assert_bits_equal(to_host(o.grad), w, 'unpack')


def test_pack_subranges_cover_everything(tuning):
    """Bucketed launches over [begin, end) ranges produce the same buffer."""
    import torch
    from chainer_b200.communicators import _memory_utility as mu
    from oracle import gradpath as og
    rng = np.random.default_rng(5)
    host = _make_arrays(ALIGNED + RAGGED, ['float32'], rng)
    params = [P(data=to_dev(a), grad=to_dev(a)) for a in host]
    n = sum(a.size for a in host)
    buf = _torch_buf(n, 'float32')
    buf.fill_(-5.0)
    pd = mu.ParamsData(params, 'grad', False)
    cuts = [0, 4096, 4100, 100000, (n // 2) // 4 * 4, n]
    for b, e in zip(cuts[:-1], cuts[1:]):
        mu._batched_pack_params(pd, _Buf(buf), np.float32, elem_begin=b, elem_end=e)
    torch.cuda.synchronize()
    assert_bits_equal(to_host(buf)[:n], og.pack(host, np.float32), 'bucketed pack')


def _states(host_params, names):
    return [{k: np.zeros_like(p) for k in names} for p in host_params]


@pytest.mark.parametrize('buf_dtype', ['float32', 'float16', 'bfloat16'])
@pytest.mark.parametrize('sizes,pdtype', [(ALIGNED, 'float32'), (RAGGED, 'float32'),
                                          (ALIGNED, 'float16'), (RAGGED, 'float64'),
                                          (RAGGED, 'float16'), (ALIGNED8, 'float32')])
@pytest.mark.parametrize('n_ranks', [1, 8, 3])
@pytest.mark.parametrize('write_grad', [0, 1])
def test_fused_momentum_sgd_bit_exact(buf_dtype, sizes, pdtype, n_ranks, write_grad, tuning):
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from chainer_b200.communicators import _memory_utility as mu
    from oracle import gradpath as og
    lib = _lib.get()
    rng = np.random.default_rng(99)
    pdt = np.dtype(pdtype)
    hp = [(rng.standard_normal(n) * 0.05).astype(pdt) for n in sizes]
    hv = [np.zeros_like(p) for p in hp]
    d_p = [to_dev(a) for a in hp]
    d_v = [to_dev(a) for a in hv]
    n = sum(sizes)
    lr, mom = 0.01, 0.9
    for step in range(3):
        # the "allreduced" buffer: a sum over n_ranks of gradients, in buffer dtype
        summed = og.cast(rng.standard_normal(n) * 1e-2 * n_ranks, _odt(buf_dtype))
        buf = _torch_buf(n, buf_dtype)
        buf[:n] = to_dev(summed).to(buf.dtype)
        params = [P(data=d_p[i], grad=torch.full_like(d_p[i], 3.0)) for i in range(len(sizes))]
        pd = mu.ParamsData(params, 'grad', False,
                           extra_ptrs=[(d_p[i], [d_v[i]]) for i in range(len(sizes))])
        lib.gp_unpack_momentum_sgd(buf.data_ptr(), dev.dtype_id(_odt(buf_dtype)), pd.d_csum,
                                   pd.d_segs, pd.n_params, 0, n, 1.0 / n_ranks, lr, mom,
                                   write_grad, pd.layout_hint(_odt(buf_dtype)), 0)
        torch.cuda.synchronize()
        g = og.mean_grad_value(summed, _odt(buf_dtype), n_ranks, pdt)
        cs = og.size_csum(hp)
        for i in range(len(sizes)):
            gi = g[cs[i]:cs[i + 1]]
            og.momentum_sgd_update(hp[i], gi, hv[i], lr, mom)
            assert_bits_equal(to_host(d_p[i]), hp[i], 'param step %d' % step)
            assert_bits_equal(to_host(d_v[i]), hv[i], 'v step %d' % step)
            if write_grad:
                assert_bits_equal(to_host(params[i].grad), gi, 'grad')
            else:
                assert (to_host(params[i].grad) == 3.0).all()


ADAM_VARIANTS = {
    'adam': dict(),
    'adamw': dict(eta=0.5, weight_decay_rate=0.1),
    'amsgrad': dict(amsgrad=True),
    'adabound': dict(adabound=True),
    'amsbound': dict(amsgrad=True, adabound=True),
}


@pytest.mark.parametrize('buf_dtype', ['float32', 'float16'])
@pytest.mark.parametrize('sizes,pdtype', [(ALIGNED, 'float32'), (RAGGED, 'float32'),
                                          (ALIGNED, 'float16'), (RAGGED, 'float64'),
                                          (ALIGNED8, 'float32')])
@pytest.mark.parametrize('variant', sorted(ADAM_VARIANTS))
def test_fused_adam_bit_exact(buf_dtype, sizes, pdtype, variant, tuning):
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from chainer_b200.communicators import _memory_utility as mu
    from oracle import gradpath as og
    lib = _lib.get()
    kw = dict(alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-8, eta=1.0, weight_decay_rate=0.0,
              amsgrad=False, adabound=False, final_lr=0.1, gamma=1e-3)
    kw.update(ADAM_VARIANTS[variant])
    rng = np.random.default_rng(7)
    pdt = np.dtype(pdtype)
    n_ranks = 4
    hp = [(rng.standard_normal(n) * 0.05).astype(pdt) for n in sizes]
    hm = [np.zeros_like(p) for p in hp]
    hv = [np.zeros_like(p) for p in hp]
    hh = [np.zeros_like(p) for p in hp]
    d_p, d_m, d_v, d_h = ([to_dev(a) for a in x] for x in (hp, hm, hv, hh))
    n = sum(sizes)
    flags = (1 if kw['amsgrad'] else 0) | (2 if kw['adabound'] else 0)
    for t in range(1, 4):
        summed = og.cast(rng.standard_normal(n) * 1e-2 * n_ranks, _odt(buf_dtype))
        buf = _torch_buf(n, buf_dtype)
        buf[:n] = to_dev(summed).to(buf.dtype)
        params = [P(data=d_p[i], grad=torch.zeros_like(d_p[i])) for i in range(len(sizes))]
        states = [[d_m[i], d_v[i]] + ([d_h[i]] if kw['amsgrad'] else [])
                  for i in range(len(sizes))]
        pd = mu.ParamsData(params, 'grad', False,
                           extra_ptrs=[(d_p[i], states[i]) for i in range(len(sizes))])
        alpha_t = og.adam_alpha_t(kw['alpha'], kw['beta1'], kw['beta2'], t)
        lower, upper = (og.adam_bounds(kw['final_lr'], kw['alpha'], kw['alpha'], kw['gamma'], t)
                        if kw['adabound'] else (0.0, 0.0))
        lib.gp_unpack_adam(buf.data_ptr(), dev.dtype_id(_odt(buf_dtype)), pd.d_csum, pd.d_segs,
                           pd.n_params, 0, n, 1.0 / n_ranks, alpha_t, 1 - kw['beta1'],
                           1 - kw['beta2'], kw['eps'], kw['eta'], kw['weight_decay_rate'],
                           lower, upper, flags, 1, pd.layout_hint(_odt(buf_dtype)), 0)
        torch.cuda.synchronize()
        g = og.mean_grad_value(summed, _odt(buf_dtype), n_ranks, pdt)
        cs = og.size_csum(hp)
        for i in range(len(sizes)):
            gi = g[cs[i]:cs[i + 1]]
            og.adam_update_gpu(hp[i], gi, hm[i], hv[i], t, vhat=hh[i], **kw)
            assert_bits_equal(to_host(d_p[i]), hp[i], 'param t=%d' % t)
            assert_bits_equal(to_host(d_m[i]), hm[i], 'm t=%d' % t)
            assert_bits_equal(to_host(d_v[i]), hv[i], 'v t=%d' % t)
            if kw['amsgrad']:
                assert_bits_equal(to_host(d_h[i]), hh[i], 'vhat t=%d' % t)
            assert_bits_equal(to_host(params[i].grad), gi, 'grad')


@pytest.mark.parametrize('dtype', BUF_DTYPES)
@pytest.mark.parametrize('n', [1, 3, 4, 1023, 1 << 20])
@pytest.mark.parametrize('scale', [0.5, 1.0 / 3.0, 1.0 / 6.0])
def test_scale_bit_exact(dtype, n, scale):
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from oracle import gradpath as og
    rng = np.random.default_rng(3)
    h = og.cast(rng.standard_normal(n) * 10, _odt(dtype))
    buf = _torch_buf(n, dtype)
    buf[:n] = to_dev(h).to(buf.dtype)
    _lib.get().gp_scale(buf.data_ptr(), dev.dtype_id(_odt(dtype)), n, scale, 0)
    torch.cuda.synchronize()
    assert_bits_equal(to_host(buf)[:n], og.scale_buffer(h, _odt(dtype), scale), 'scale')


@pytest.mark.parametrize('dtype', BUF_DTYPES)
@pytest.mark.parametrize('bad', [None, np.nan, np.inf, -np.inf])
def test_check_finite(dtype, bad):
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    n = 100003
    buf = _torch_buf(n, dtype)
    buf.normal_()
    if bad is not None:
        buf[n - 2] = bad
    flag = torch.zeros(1, dtype=torch.int32, device='cuda')
    _lib.get().gp_check_finite(buf.data_ptr(), dev.dtype_id(_odt(dtype)), n, flag.data_ptr(), 0)
    assert int(flag.item()) == (0 if bad is None else 1)


def test_resnet50_full_size_roundtrip_and_update():
    """BASELINE config 2 at full size: pack -> unpack(1/8) round trip equals a
    plain scaled copy (size-independent property), and the fused update equals
    the unfused unpack + per-tensor torch update."""
    import torch
    from chainer_b200 import _lib, workloads
    from chainer_b200.communicators import _memory_utility as mu
    lib = _lib.get()
    plist = workloads.resnet50()
    torch.manual_seed(7)
    grads = [torch.randn(int(np.prod(s)), device='cuda') * 1e-2 for _, s in plist]
    data = [torch.randn(int(np.prod(s)), device='cuda') * 0.05 for _, s in plist]
    vs = [torch.zeros_like(d) for d in data]
    n = sum(g.numel() for g in grads)
    assert n == 25557096
    params = [P(data=d, grad=g) for d, g in zip(data, grads)]
    buf = torch.empty(n, device='cuda')
    pd = mu.ParamsData(params, 'grad', False, extra_ptrs=[(d, [v]) for d, v in zip(data, vs)])
    mu._batched_pack_params(pd, _Buf(buf), np.float32)
    flat = torch.cat(grads)
    assert torch.equal(buf, flat)                     # layout: bit-exact
    want_p = [d - 0.01 * (g * 0.125) for d, g in zip(data, grads)]
    assert pd.layout_hint(np.float32) == 7          # ResNet-50 takes the TMA-staged kernel
    lib.gp_unpack_momentum_sgd(buf.data_ptr(), 7, pd.d_csum, pd.d_segs, pd.n_params, 0, n,
                               0.125, 0.01, 0.9, 1, 7, 0)
    torch.cuda.synchronize()
    for g, g0 in zip(grads, torch.split(flat, [x.numel() for x in grads])):
        assert torch.equal(g, g0 * 0.125)             # written-back mean gradient
    for d, w, v, g in zip(data, want_p, vs, grads):
        assert torch.equal(v, -(torch.tensor(0.01, device='cuda') * g))
        torch.testing.assert_close(d, w, rtol=1e-6, atol=1e-7)
