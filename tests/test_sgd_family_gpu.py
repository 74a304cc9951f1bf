"""GPU parity tests of gp_unpack_sgd_family (SGD, CorrectedMomentumSGD, NesterovAG;
csrc/gp_sgd_family.cu): BIT-EXACT against the oracle, which is pinned bit-for-bit
to the unmodified reference's update_core_cpu (tests/golden/sgd_family.npz), and the
same vectors through the public API."""
import numpy as np
import pytest

from tests.helpers import P, assert_bits_equal, to_dev, to_host
from tests.test_hooks_gpu import HOOKSETS, _Hooks, walker_tuning  # noqa: F401
from tests.test_kernels_gpu import ALIGNED, RAGGED, _odt, _torch_buf

pytestmark = pytest.mark.gpu

RULES = {'sgd': 0, 'corrected': 1, 'nesterov': 2}


def _oracle_update(rule, p, g, v, lr, mom):
    from oracle import gradpath as og
    if rule == 'sgd':
        og.sgd_update(p, g, lr)
    elif rule == 'corrected':
        og.corrected_momentum_sgd_update(p, g, v, lr, mom)
    else:
        og.nesterov_ag_update(p, g, v, lr, mom)


@pytest.mark.parametrize('buf_dtype', ['float32', 'float16', 'bfloat16'])
@pytest.mark.parametrize('sizes,pdtype', [(ALIGNED, 'float32'), (RAGGED, 'float32'),
                                          (RAGGED, 'float16'), (RAGGED, 'float64')])
@pytest.mark.parametrize('hookset', [None, 'all'])
@pytest.mark.parametrize('write_grad', [0, 1])
@pytest.mark.parametrize('rule', sorted(RULES))
def test_sgd_family_bit_exact(rule, buf_dtype, sizes, pdtype, hookset, write_grad, walker_tuning):  # noqa: F811
    import torch
    from chainer_b200 import _lib
    from chainer_b200 import device as dev
    from chainer_b200.communicators import _memory_utility as mu
    from oracle import gradpath as og
    lib = _lib.get()
    rng = np.random.default_rng(77)
    pdt = np.dtype(pdtype)
    hp = [(rng.standard_normal(n) * 0.05).astype(pdt) for n in sizes]
    hv = [np.zeros_like(p) for p in hp]
    d_p = [to_dev(a) for a in hp]
    d_v = [to_dev(a) for a in hv]
    n = sum(sizes)
    hooks = _Hooks(*HOOKSETS[hookset]) if hookset else None
    gscale = 1e-2 * ((hooks.loss_scale or 1.0) if hooks else 1.0)
    n_ranks, lr, mom = 3, 0.05, 0.8
    for step in range(3):
        summed = og.cast(rng.standard_normal(n) * gscale * n_ranks, _odt(buf_dtype))
        buf = _torch_buf(n, buf_dtype)
        buf[:n] = to_dev(summed).to(buf.dtype)
        params = [P(data=d_p[i], grad=torch.full_like(d_p[i], 3.0)) for i in range(len(sizes))]
        extra = [(d_p[i], [d_v[i]] if rule != 'sgd' else []) for i in range(len(sizes))]
        pd = mu.ParamsData(params, 'grad', False, extra_ptrs=extra)
        lib.gp_unpack_sgd_family(buf.data_ptr(), dev.dtype_id(_odt(buf_dtype)), pd.d_csum,
                                 pd.d_segs, pd.n_params, 0, n, 1.0 / n_ranks, RULES[rule], lr, mom,
                                 write_grad, pd.layout_hint(_odt(buf_dtype)),
                                 hooks.addr if hooks else None, 0)
        torch.cuda.synchronize()
        g = og.mean_grad_value(summed, _odt(buf_dtype), n_ranks, pdt)
        cs = og.size_csum(hp)
        for i in range(len(sizes)):
            gi = np.array(g[cs[i]:cs[i + 1]])
            if hooks:
                gi = hooks.oracle(gi, hp[i])
            _oracle_update(rule, hp[i], gi, hv[i], lr, mom)
            assert_bits_equal(to_host(d_p[i]), hp[i], 'param step %d' % step)
            if rule != 'sgd':
                assert_bits_equal(to_host(d_v[i]), hv[i], 'v step %d' % step)
            if write_grad:
                assert_bits_equal(to_host(params[i].grad), gi, 'grad')
            else:
                assert (to_host(params[i].grad) == 3.0).all()


@pytest.mark.parametrize('multi_node', [True, False])
@pytest.mark.parametrize('dtype,hooks', [('float32', False), ('float16', False), ('float64', False),
                                         ('float32', True)])
@pytest.mark.parametrize('rule', sorted(RULES))
def test_sgd_family_through_the_public_api(rule, dtype, hooks, multi_node):
    """sgd_family.npz (the unmodified reference) on the GPU: behind
    create_multi_node_optimizer (fused) and as a stand-alone optimizer.update()."""
    from tests.hooks_scenario import run_family_scenario
    if multi_node and dtype == 'float64':
        pytest.skip('bcast_data and the allreduce buffer are float32 (chainer.get_dtype())')

    def after_step(comm):
        if comm is not None:
            assert comm._fused_plan is not None

    run_family_scenario(rule, dtype, hooks, multi_node, lambda a: to_dev(np.array(a)), to_host,
                        after_step)
