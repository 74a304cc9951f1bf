"""GPU parity tests of the one-launch step kernels (csrc/gp_step.cu) at world size 1:
pack -> fused update in ONE kernel (fused at the register level: the packed buffer is
written once and never read back), through the C-ABI, BIT-EXACT against the NumPy
oracle (pack layout and casts, the 1/N rounding sequence, MomentumSGD / Adam
arithmetic) and therefore against the separate launches.  The N-rank kernels are
covered by tests/test_multi_gpu.py (tests/_dist_gpu_worker.py).
"""
import numpy as np
import pytest

from tests.helpers import P, assert_bits_equal, to_dev, to_host

pytestmark = pytest.mark.gpu

RAGGED = [7, 1, 0, 1000, 4096, 12345, 64, 3, 513, 2048, 70001]
ALIGNED = [64, 64, 9408, 64, 2048, 1000, 256, 16384, 36864, 4, 131072]
BIG = [1 << 20, 12, 300000, 64, 64, 2359296 // 4, 5]          # many tiles, many CTAs


def _odt(dtype):
    from oracle import gradpath as og
    return og.BF16 if dtype == 'bfloat16' else np.dtype(dtype)


def _torch_dt(dtype):
    import torch
    return {'float32': torch.float32, 'float16': torch.float16, 'bfloat16': torch.bfloat16}[dtype]


@pytest.fixture(params=[(256, 0, 0), (256, 4, 0), (128, 2, 1), (512, 4, 1)],
                ids=['default', 't256u4', 't128u2-persistent', 't512u4-persistent'])
def step_tuning(request):
    """The one-rank step is a walker launch: (threads, unroll, persistent) of gp_set_tuning."""
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


@pytest.mark.parametrize('buf_dtype', ['float32', 'float16', 'bfloat16'])
@pytest.mark.parametrize('sizes', [ALIGNED, RAGGED, BIG], ids=['aligned', 'ragged', 'big'])
@pytest.mark.parametrize('scale_ranks', [1, 2])
@pytest.mark.parametrize('write_grad', [0, 1])
def test_step1_momentum_sgd_bit_exact(buf_dtype, sizes, scale_ranks, write_grad, step_tuning):
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from chainer_b200.communicators import _memory_utility as mu
    from oracle import gradpath as og
    lib = _lib.get()
    bdt = _odt(buf_dtype)
    rng = np.random.default_rng(11)
    hp = [(rng.standard_normal(n) * 0.05).astype(np.float32) for n in sizes]
    hv = [np.zeros_like(p) for p in hp]
    d_p = [to_dev(a) for a in hp]
    d_v = [to_dev(a) for a in hv]
    n = sum(sizes)
    lr, mom = 0.01, 0.9
    assert lib.gp_step_supported(1, dev.dtype_id(bdt), 7, 1.0 / scale_ranks, 0) == 1
    buf = torch.zeros(max(n, 4) + 4, dtype=_torch_dt(buf_dtype), device='cuda')
    for step in range(3):
        hg = [(rng.standard_normal(k) * 1e-2 * scale_ranks).astype(np.float32) for k in sizes]
        d_g = [to_dev(g) for g in hg]
        params = [P(data=d_p[i], grad=d_g[i]) for i in range(len(sizes))]
        pd = mu.ParamsData(params, 'grad', False,
                           extra_ptrs=[(d_p[i], [d_v[i]]) for i in range(len(sizes))])
        assert pd.all_float32
        lib.gp_step_momentum_sgd(None, None, buf.data_ptr(), dev.dtype_id(bdt), pd.d_csum,
                                 pd.d_segs, pd.n_params, n, 1.0 / scale_ranks, lr, mom,
                                 write_grad, 7, 0)
        torch.cuda.synchronize()
        packed = og.pack(hg, bdt)
        assert_bits_equal(to_host(buf)[:n], packed, 'packed buffer')
        g = og.mean_grad_value(packed, bdt, scale_ranks, np.float32)
        cs = og.size_csum(hp)
        for i in range(len(sizes)):
            gi = g[cs[i]:cs[i + 1]]
            og.momentum_sgd_update(hp[i], gi, hv[i], lr, mom)
            assert_bits_equal(to_host(d_p[i]), hp[i], 'param step %d' % step)
            assert_bits_equal(to_host(d_v[i]), hv[i], 'v step %d' % step)
            if write_grad:
                assert_bits_equal(to_host(d_g[i]), gi, 'grad')
            else:
                assert_bits_equal(to_host(d_g[i]), hg[i], 'grad untouched')


@pytest.mark.parametrize('buf_dtype', ['float32', 'float16', 'bfloat16'])
@pytest.mark.parametrize('sizes', [ALIGNED, RAGGED, BIG], ids=['aligned', 'ragged', 'big'])
@pytest.mark.parametrize('variant', ['adam', 'adamw', 'adabound'])
def test_step1_adam_bit_exact(buf_dtype, sizes, variant, step_tuning):
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from chainer_b200.communicators import _memory_utility as mu
    from oracle import gradpath as og
    lib = _lib.get()
    kw = dict(alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-8, eta=1.0, weight_decay_rate=0.0,
              amsgrad=False, adabound=False, final_lr=0.1, gamma=1e-3)
    kw.update({'adam': {}, 'adamw': dict(eta=0.5, weight_decay_rate=0.1),
               'adabound': dict(adabound=True)}[variant])
    bdt = _odt(buf_dtype)
    rng = np.random.default_rng(5)
    hp = [(rng.standard_normal(n) * 0.05).astype(np.float32) for n in sizes]
    hm = [np.zeros_like(p) for p in hp]
    hv = [np.zeros_like(p) for p in hp]
    d_p, d_m, d_v = ([to_dev(a) for a in x] for x in (hp, hm, hv))
    n = sum(sizes)
    flags = 2 if kw['adabound'] else 0
    buf = torch.zeros(max(n, 4) + 4, dtype=_torch_dt(buf_dtype), device='cuda')
    for t in range(1, 4):
        hg = [(rng.standard_normal(k) * 1e-2).astype(np.float32) for k in sizes]
        d_g = [to_dev(g) for g in hg]
        params = [P(data=d_p[i], grad=d_g[i]) for i in range(len(sizes))]
        pd = mu.ParamsData(params, 'grad', False,
                           extra_ptrs=[(d_p[i], [d_m[i], d_v[i]]) for i in range(len(sizes))])
        alpha_t = og.adam_alpha_t(kw['alpha'], kw['beta1'], kw['beta2'], t)
        lower, upper = (og.adam_bounds(kw['final_lr'], kw['alpha'], kw['alpha'], kw['gamma'], t)
                        if kw['adabound'] else (0.0, 0.0))
        lib.gp_step_adam(None, None, buf.data_ptr(), dev.dtype_id(bdt), pd.d_csum, pd.d_segs,
                         pd.n_params, n, 1.0, alpha_t, 1 - kw['beta1'], 1 - kw['beta2'], kw['eps'],
                         kw['eta'], kw['weight_decay_rate'], lower, upper, flags, 1, 7, 0)
        torch.cuda.synchronize()
        packed = og.pack(hg, bdt)
        assert_bits_equal(to_host(buf)[:n], packed, 'packed buffer')
        g = og.mean_grad_value(packed, bdt, 1, np.float32)
        cs = og.size_csum(hp)
        for i in range(len(sizes)):
            gi = g[cs[i]:cs[i + 1]]
            og.adam_update_gpu(hp[i], gi, hm[i], hv[i], t, vhat=None, **kw)
            assert_bits_equal(to_host(d_p[i]), hp[i], 'param t=%d' % t)
            assert_bits_equal(to_host(d_m[i]), hm[i], 'm t=%d' % t)
            assert_bits_equal(to_host(d_v[i]), hv[i], 'v t=%d' % t)
            assert_bits_equal(to_host(d_g[i]), gi, 'grad')


def test_step_not_covered_is_refused():
    """float64 buffers, non-float32 arrays, general (non power-of-two) scales and AMSGrad
    stay on the separate launches: gp_step_supported says so and the entry points refuse."""
    from chainer_b200 import _lib
    lib = _lib.get()
    assert lib.gp_step_supported(1, 8, 7, 1.0, 0) == 0          # float64 buffer
    assert lib.gp_step_supported(1, 7, 0, 1.0, 0) == 0          # mixed / non-float32 arrays
    assert lib.gp_step_supported(1, 7, 7, 1.0 / 3.0, 0) == 0    # general scale
    assert lib.gp_step_supported(3, 7, 7, 1.0 / 3.0, 0) == 0    # 3 ranks
    assert lib.gp_step_supported(1, 7, 7, 1.0, 1) == 0          # AMSGrad
    assert lib.gp_step_supported(8, 6, 7, 0.125, 2) == 1
    with pytest.raises(_lib.GradpathError):
        lib.gp_step_momentum_sgd(None, None, 256, 8, 256, 256, 1, 16, 1.0, 0.01, 0.9, 1, 7, 0)


def test_step1_through_public_api_equals_separate_launches():
    """create_multi_node_optimizer(...).update() with and without the one-launch step gives
    identical bits (ResNet-50 size histogram, MomentumSGD and Adam, float16 buffer too)."""
    import torch
    import chainer_b200
    from chainer_b200 import workloads
    from chainer_b200.core.link import link_from_named_arrays
    plist = workloads.scaled_histogram(700000)
    for opt_name in ('momentum_sgd', 'adam'):
        for adt in (None, np.float16):
            results = []
            for use_step in (True, False):
                comm = chainer_b200.create_communicator('pure_nccl', allreduce_grad_dtype=adt)
                comm.use_step = use_step
                rng = np.random.default_rng(3)
                model = link_from_named_arrays(
                    [(nm, to_dev((rng.standard_normal(s) * 0.05).astype(np.float32)))
                     for nm, s in plist])
                actual = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9) \
                    if opt_name == 'momentum_sgd' else chainer_b200.Adam()
                opt = chainer_b200.create_multi_node_optimizer(actual, comm)
                opt.setup(model)
                opt.update()
                from chainer_b200 import _lib
                before = _lib.get().launches
                for step in range(3):
                    for _, p in sorted(model.namedparams()):
                        p.grad = to_dev((rng.standard_normal(tuple(p.data.shape)) * 1e-2)
                                        .astype(np.float32))
                    opt.update()
                torch.cuda.synchronize()
                per_step = (_lib.get().launches - before) // 3
                assert per_step == (1 if use_step else 2), (use_step, per_step)
                results.append([to_host(p.data) for _, p in sorted(model.namedparams())] +
                               [to_host(p.grad) for _, p in sorted(model.namedparams())])
                comm.finalize()
            for a, b in zip(*results):
                assert_bits_equal(a, b, '%s %s' % (opt_name, adt))
