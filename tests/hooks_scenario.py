"""The hooks.npz scenarios (WeightDecay / GradientClipping / static loss scaling of
the unmodified reference) replayed through ``create_multi_node_optimizer``; shared
by the CPU host-logic test (oracle-backed library double) and the GPU API test."""
import numpy as np

import chainer_b200
from chainer_b200.core import link as L
from tests.helpers import assert_bits_equal
from tests.test_oracle_golden import HOOK_CASES, _npz, hooks_golden  # noqa: F401


def hook_objects(spec):
    from chainer_b200 import optimizer_hooks as H
    return [H.WeightDecay(arg) if kind == 'wd' else H.GradientClipping(arg) for kind, arg in spec]


def run_hooks_scenario(variant, dtype, to_arr, to_np, before_step=None, after_step=None):
    z = hooks_golden()
    opt_name, spec, ls = HOOK_CASES[variant]
    pre = '%s|%s|' % (variant, dtype)
    names = sorted(k[len(pre) + 4:] for k in z.files if k.startswith(pre + 'init'))
    model = L.link_from_named_arrays([(n, to_arr(z[pre + 'init' + n])) for n in names])
    comm = chainer_b200.create_communicator('pure_nccl')
    actual = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9) if opt_name == 'sgd' \
        else chainer_b200.Adam()
    opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    opt.setup(model)
    for h in hook_objects(spec):
        opt.add_hook(h)
    if ls is not None:
        actual.loss_scaling(scale=ls)
    opt.update()                                       # first call: broadcast only
    params = dict(sorted(model.namedparams()))
    clip = any(k == 'clip' for k, _ in spec)
    # exact: the arithmetic is the oracle's (pinned bit-for-bit to the reference by
    # test_oracle_golden).  The norm of the clipping hook is accumulated in double
    # here and in float32 dot products there; Adam runs the GPU formula where the
    # vectors come from the reference's CPU formula.
    exact = opt_name == 'sgd' and not clip
    for step in range(3):
        for n in names:
            params[n].grad = to_arr(z[pre + 'grad%d%s' % (step, n)])
            params[n]._loss_scale = ls
        if before_step is not None:
            before_step()
        opt.update()
        if after_step is not None:
            after_step(comm)
        for n in names:
            want_p, want_g = z[pre + 'param%d%s' % (step, n)], z[pre + 'gradafter%d%s' % (step, n)]
            got_p, got_g = to_np(params[n].data), to_np(params[n].grad)
            if exact:
                assert_bits_equal(got_p.reshape(want_p.shape), want_p, (variant, step, n))
                assert_bits_equal(got_g.reshape(want_g.shape), want_g, (variant, step, n, 'grad'))
            else:
                tol = 2e-3 if dtype == 'float16' else 2e-6
                np.testing.assert_allclose(np.asarray(got_g, dtype=np.float64).reshape(want_g.shape),
                                           np.asarray(want_g, dtype=np.float64), rtol=tol,
                                           atol=tol * 1e-2, err_msg=str((variant, step, n, 'grad')))
                ptol = tol if opt_name == 'sgd' else max(tol, 2e-5)
                np.testing.assert_allclose(np.asarray(got_p, dtype=np.float64).reshape(want_p.shape),
                                           np.asarray(want_p, dtype=np.float64), rtol=ptol,
                                           atol=ptol * 1e-2, err_msg=str((variant, step, n)))
    assert actual.t == 3
    comm.finalize()


def _family_optimizer(rule):
    return {'sgd': lambda: chainer_b200.SGD(lr=0.05),
            'corrected': lambda: chainer_b200.CorrectedMomentumSGD(lr=0.05, momentum=0.8),
            'nesterov': lambda: chainer_b200.NesterovAG(lr=0.05, momentum=0.8)}[rule]()


def run_family_scenario(rule, dtype, hooks, multi_node, to_arr, to_np, after_step=None):
    """sgd_family.npz (the unmodified reference) replayed through the product."""
    z = _npz('sgd_family.npz')
    pre = '%s%s|%s|' % (rule, '_clip_wd' if hooks else '', dtype)
    names = sorted(k[len(pre) + 4:] for k in z.files if k.startswith(pre + 'init'))
    model = L.link_from_named_arrays([(n, to_arr(z[pre + 'init' + n])) for n in names])
    actual = _family_optimizer(rule)
    comm = None
    if multi_node:
        comm = chainer_b200.create_communicator('pure_nccl')
        opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    else:
        opt = actual
    opt.setup(model)
    if hooks:
        for h in hook_objects([('clip', 0.05), ('wd', 0.05)]):
            opt.add_hook(h)
    if multi_node:
        opt.update()                                   # first call: broadcast only
    params = dict(sorted(model.namedparams()))
    for step in range(3):
        for n in names:
            params[n].grad = to_arr(z[pre + 'grad%d%s' % (step, n)])
        opt.update()
        if after_step is not None:
            after_step(comm)
        for n in names:
            want = z[pre + 'param%d%s' % (step, n)]
            got = to_np(params[n].data).reshape(want.shape)
            if hooks:       # the norm is accumulated in double here, float32 dots there
                np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-8)
            else:
                assert_bits_equal(got, want, (rule, dtype, step, n))
            if rule != 'sgd':
                want_v = z[pre + 'state_v%d%s' % (step, n)]
                got_v = to_np(params[n].update_rule.state['v']).reshape(want_v.shape)
                if hooks:
                    np.testing.assert_allclose(got_v, want_v, rtol=2e-6, atol=2e-8)
                else:
                    assert_bits_equal(got_v, want_v, (rule, dtype, step, n, 'v'))
    assert actual.t == 3
    if comm is not None:
        comm.finalize()


def run_dynamic_loss_scale(dtype, multi_node, to_arr, to_np):
    """dynamic_loss_scale.npz: MomentumSGD with dynamic loss scaling; step 2 carries a
    non-finite gradient (update skipped, scale halved, then grown by 2**(1/interval))."""
    import warnings
    z = _npz('dynamic_loss_scale.npz')
    pre = dtype + '|'
    dt = np.dtype(dtype)
    names = sorted(k[len(pre) + 4:] for k in z.files if k.startswith(pre + 'init'))
    model = L.link_from_named_arrays([(n, to_arr(z[pre + 'init' + n])) for n in names])
    actual = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9)
    comm = None
    if multi_node:
        comm = chainer_b200.create_communicator('pure_nccl')
        opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    else:
        opt = actual
    opt.setup(model)
    actual.loss_scaling(interval=2)
    if multi_node:
        opt.update()
    params = dict(sorted(model.namedparams()))
    scales = []
    for step in range(6):
        ls = actual._loss_scale
        for n in names:
            g = z[pre + 'grad%d%s' % (step, n)]
            params[n].grad = to_arr(np.asarray(g * dt.type(ls)).astype(dt).reshape(g.shape))
            params[n]._loss_scale = ls
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter('always')
            opt.update()
        assert any('Non finite number found in param.grad of /p03' in str(x.message)
                   for x in w) == (step == 2)
        scales.append(actual._loss_scale)
        for n in names:
            want = z[pre + 'param%d%s' % (step, n)]
            assert_bits_equal(to_np(params[n].data).reshape(want.shape), want, (dtype, step, n))
    np.testing.assert_array_equal(np.asarray(scales, dtype=np.float64), z[pre + 'scales'])
    ts = [actual.t] + [p.update_rule.t for _, p in sorted(model.namedparams())]
    np.testing.assert_array_equal(np.asarray(ts), z[pre + 't'])
    if comm is not None:
        comm.finalize()


def run_fp32_update(case, multi_node, to_arr, to_np):
    """fp32_update.npz: float16 parameters updated through float32 master weights."""
    from tests.test_oracle_golden import FP32_CASES
    z = _npz('fp32_update.npz')
    opt_name, wd, ls = FP32_CASES[case]
    pre = case + '|'
    names = sorted(k[len(pre) + 4:] for k in z.files if k.startswith(pre + 'init'))
    model = L.link_from_named_arrays([(n, to_arr(z[pre + 'init' + n])) for n in names])
    actual = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9) if opt_name == 'sgd' \
        else chainer_b200.Adam()
    actual.use_fp32_update()
    comm = None
    if multi_node:
        comm = chainer_b200.create_communicator('pure_nccl', allreduce_grad_dtype=np.float16)
        opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    else:
        opt = actual
    opt.setup(model)
    if wd is not None:
        from chainer_b200 import optimizer_hooks as H
        opt.add_hook(H.WeightDecay(wd))
    if ls is not None:
        actual.loss_scaling(scale=ls)
    if multi_node:
        opt.update()
    params = dict(sorted(model.namedparams()))
    exact = opt_name == 'sgd'           # Adam: GPU formula here, CPU formula in the vectors
    for step in range(3):
        for n in names:
            params[n].grad = to_arr(z[pre + 'grad%d%s' % (step, n)])
            params[n]._loss_scale = ls
        opt.update()
        for n in names:
            rule = params[n].update_rule
            want_p, want_m = z[pre + 'param%d%s' % (step, n)], z[pre + 'master%d%s' % (step, n)]
            got_p = to_np(params[n].data).reshape(want_p.shape)
            got_m = to_np(rule._fp32_param.data).reshape(want_m.shape)
            assert got_p.dtype == np.float16 and got_m.dtype == np.float32
            assert all(to_np(s).dtype == np.float32 for s in rule.state.values())
            if exact:
                assert_bits_equal(got_m, want_m, (case, step, n, 'master'))
                assert_bits_equal(got_p, want_p, (case, step, n))
            else:
                np.testing.assert_allclose(got_m, want_m, rtol=2e-5, atol=1e-7)
                np.testing.assert_allclose(got_p.astype(np.float32), want_p.astype(np.float32),
                                           rtol=2e-3, atol=1e-6)
    assert actual.t == 3
    if comm is not None:
        comm.finalize()


def run_fp32_dynamic(case, multi_node, to_arr, to_np, fake=None):
    """fp32_dynamic.npz (generated from the unmodified reference): float16 parameters,
    float32 master weights and DYNAMIC loss scaling together; step 2 carries a non-finite
    gradient -> update skipped, scale halved, step counters of the rules not advanced.
    Behind the multi-node optimizer this is the fused master path (gp_unpack_*_master with
    the device-side skip word)."""
    import warnings
    z = _npz('fp32_dynamic.npz')
    pre = case + '|'
    dt = np.dtype(np.float16)
    names = sorted(k[len(pre) + 4:] for k in z.files if k.startswith(pre + 'init'))
    model = L.link_from_named_arrays([(n, to_arr(z[pre + 'init' + n])) for n in names])
    actual = chainer_b200.MomentumSGD(lr=0.01, momentum=0.9) if case == 'sgd_wd' \
        else chainer_b200.Adam()
    actual.use_fp32_update()
    comm = None
    if multi_node:
        comm = chainer_b200.create_communicator('pure_nccl', allreduce_grad_dtype=np.float16)
        opt = chainer_b200.create_multi_node_optimizer(actual, comm)
    else:
        opt = actual
    opt.setup(model)
    if case == 'sgd_wd':
        from chainer_b200 import optimizer_hooks as H
        opt.add_hook(H.WeightDecay(0.05))
    actual.loss_scaling(interval=2)
    if multi_node:
        opt.update()
    params = dict(sorted(model.namedparams()))
    exact = case == 'sgd_wd'            # Adam: GPU formula here, CPU formula in the vectors
    scales = []
    for step in range(6):
        ls = actual._loss_scale
        for n in names:
            g = z[pre + 'grad%d%s' % (step, n)]
            params[n].grad = to_arr(np.asarray(g * dt.type(ls)).astype(dt).reshape(g.shape))
            params[n]._loss_scale = ls
        if fake is not None:
            fake.calls[:] = []
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter('always')
            opt.update()
        assert any('Non finite number found in param.grad of /p03' in str(x.message)
                   for x in w) == (step == 2)
        if fake is not None and multi_node:
            called = [c[0] for c in fake.calls]
            kernel = 'gp_unpack_momentum_sgd_master' if case == 'sgd_wd' else 'gp_unpack_adam_master'
            assert kernel in called and 'gp_check_finite' in called, called
            assert 'gp_divide' not in called and 'gp_weight_decay' not in called
        scales.append(actual._loss_scale)
        for n in names:
            want_p, want_m = z[pre + 'param%d%s' % (step, n)], z[pre + 'master%d%s' % (step, n)]
            got_p = to_np(params[n].data).reshape(want_p.shape)
            got_m = to_np(params[n].update_rule._fp32_param.data).reshape(want_m.shape)
            if exact:
                assert_bits_equal(got_m, want_m, (case, step, n, 'master'))
                assert_bits_equal(got_p, want_p, (case, step, n))
            else:
                np.testing.assert_allclose(got_m, want_m, rtol=2e-5, atol=1e-7)
                np.testing.assert_allclose(got_p.astype(np.float32), want_p.astype(np.float32),
                                           rtol=2e-3, atol=1e-6)
    np.testing.assert_array_equal(np.asarray(scales, dtype=np.float64), z[pre + 'scales'])
    ts = [actual.t] + [p.update_rule.t for _, p in sorted(model.namedparams())]
    np.testing.assert_array_equal(np.asarray(ts), z[pre + 't'])
    if comm is not None:
        comm.finalize()
