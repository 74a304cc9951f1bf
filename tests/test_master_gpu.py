"""GPU parity tests of the float32-master-weight kernels (csrc/gp_master.cu) through the
C-ABI: BIT-EXACT against the NumPy oracle's pieces chained as the reference chains them
(chainer/optimizer.py:262-305: mean gradient in float16 -> optimizer hooks on the float16
arrays -> float32 copy / loss scale -> update_core on the float32 master -> float16
parameter), for ragged / aligned layouts, three buffer dtypes, with and without hooks, and
with the dynamic-loss-scaling skip word set."""
import ctypes

import numpy as np
import pytest

from tests.helpers import P, assert_bits_equal, to_dev, to_host

pytestmark = pytest.mark.gpu

RAGGED = [7, 1, 0, 1000, 4096, 12345, 64, 3, 513, 2048, 70001]
ALIGNED = [64, 64, 9408, 64, 2048, 1000, 256, 16384, 36864, 8, 131072]


def _odt(dtype):
    from oracle import gradpath as og
    return og.BF16 if dtype == 'bfloat16' else np.dtype(dtype)


@pytest.mark.parametrize('buf_dtype', ['float16', 'float32', 'bfloat16'])
@pytest.mark.parametrize('sizes', [ALIGNED, RAGGED], ids=['aligned', 'ragged'])
@pytest.mark.parametrize('rule', ['sgd', 'adam'])
@pytest.mark.parametrize('hooks', ['none', 'wd+ls', 'skip'])
@pytest.mark.parametrize('n_ranks', [1, 8, 3])
def test_master_kernels_bit_exact(buf_dtype, sizes, rule, hooks, n_ranks):
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from chainer_b200.communicators import _memory_utility as mu
    from oracle import gradpath as og
    lib = _lib.get()
    bdt = _odt(buf_dtype)
    f16, f32 = np.float16, np.float32
    rng = np.random.default_rng(21)
    hm = [(rng.standard_normal(n) * 0.05).astype(f32) for n in sizes]        # masters
    hp = [m.astype(f16) for m in hm]
    hs1 = [np.zeros_like(m) for m in hm]
    hs2 = [np.zeros_like(m) for m in hm]
    d_m, d_p, d_s1, d_s2 = ([to_dev(a) for a in x] for x in (hm, hp, hs1, hs2))
    n = sum(sizes)
    ls = 128.0 if hooks == 'wd+ls' else None
    wd = 0.05 if hooks == 'wd+ls' else None
    hk = None
    if hooks == 'wd+ls':
        hk = _lib.GpHooks()
        hk.clip_rate = None
        hk.weight_decay = wd * ls
        hk.loss_scale = ls
    skip = torch.tensor([1 if hooks == 'skip' else 0], dtype=torch.int32, device='cuda')
    td = {'float16': torch.float16, 'float32': torch.float32, 'bfloat16': torch.bfloat16}[buf_dtype]
    for t in range(1, 4):
        gscale = 1e-2 * n_ranks * (ls or 1.0) if rule == 'sgd' else 0.5 * n_ranks
        summed = og.cast(rng.standard_normal(n) * gscale, bdt)
        buf = torch.zeros(max(n, 4), dtype=td, device='cuda')
        buf[:n] = to_dev(summed).to(td)
        d_g = [torch.full((k,), 3.0, dtype=torch.float16, device='cuda') for k in sizes]
        params = [P(data=d_p[i], grad=d_g[i]) for i in range(len(sizes))]
        states = [[d_s1[i]] if rule == 'sgd' else [d_s1[i], d_s2[i]] for i in range(len(sizes))]
        pd = mu.ParamsData(params, 'grad', False,
                           extra_ptrs=[(d_m[i], states[i], d_p[i]) for i in range(len(sizes))])
        hk_addr = ctypes.addressof(hk) if hk is not None else None
        if rule == 'sgd':
            lib.gp_unpack_momentum_sgd_master(buf.data_ptr(), dev.dtype_id(bdt), pd.d_csum, pd.d_segs,
                                              pd.n_params, 0, n, 1.0 / n_ranks, 0.01, 0.9, 1, hk_addr,
                                              skip.data_ptr(), 0)
        else:
            alpha_t = og.adam_alpha_t(0.001, 0.9, 0.999, t)
            lib.gp_unpack_adam_master(buf.data_ptr(), dev.dtype_id(bdt), pd.d_csum, pd.d_segs,
                                      pd.n_params, 0, n, 1.0 / n_ranks, alpha_t, 1 - 0.9, 1 - 0.999,
                                      1e-8, 1.0, 0.0, 0.0, 0.0, 0, 1, hk_addr, skip.data_ptr(), 0)
        torch.cuda.synchronize()
        g16_all = og.mean_grad_value(summed, bdt, n_ranks, f16)
        cs = og.size_csum(hm)
        for i in range(len(sizes)):
            g16 = np.array(g16_all[cs[i]:cs[i + 1]])
            if hooks != 'skip':
                if wd is not None:
                    og.weight_decay_hook(hm[i].astype(f16), g16, wd, ls)
                g32 = g16.astype(f32)
                if ls is not None:
                    og.loss_scale_divide(g32, ls)
                if rule == 'sgd':
                    og.momentum_sgd_update(hm[i], g32, hs1[i], 0.01, 0.9)
                else:
                    og.adam_update_gpu(hm[i], g32, hs1[i], hs2[i], t)
                hp[i] = hm[i].astype(f16)
            assert_bits_equal(to_host(d_m[i]), hm[i], 'master t=%d' % t)
            assert_bits_equal(to_host(d_p[i]), hp[i], 'param16 t=%d' % t)
            assert_bits_equal(to_host(d_s1[i]), hs1[i], 'state1 t=%d' % t)
            if rule == 'adam':
                assert_bits_equal(to_host(d_s2[i]), hs2[i], 'state2 t=%d' % t)
            assert_bits_equal(to_host(d_g[i]), g16, 'grad16 t=%d' % t)
