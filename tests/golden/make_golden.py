"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (chainer/chainer v7.8.1, imported from /root/reference through
`_ref_shims`) on seeded inputs.

    python tests/golden/make_golden.py

Outputs (committed):
  layouts.json          sorted(model.namedparams()) of the example models
  momentum_sgd.npz      chainer.optimizers.MomentumSGD, 3 steps, f16/f32/f64
  adam.npz              chainer.optimizers.Adam (+AdamW, AMSGrad, AdaBound, AMSBound)
  naive_mean_grad.npz   chainermn NaiveCommunicator.multi_node_mean_grad, 2 and 3 ranks
  mnbn.npz              MultiNodeBatchNormalization (_MpiImpl) forward/backward, 2 ranks
  sgd_family.npz        SGD, CorrectedMomentumSGD, NesterovAG (+ hooks), f16/f32/f64
  fp32_update.npz       float16 parameters with fp32 master weights (use_fp32_update)
  dynamic_loss_scale.npz  MomentumSGD with dynamic loss scaling and a non-finite gradient
  hooks.npz             MomentumSGD / Adam with WeightDecay, GradientClipping hooks and static
                        loss scaling, f16/f32
The reference tree is not available on the GPU box, hence fixtures.
"""
import importlib.util
import json
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_shims  # noqa: E402

chainer, chainermn = _ref_shims.import_reference()
import chainer.functions as F  # noqa: E402
import chainer.links as L  # noqa: E402
from chainer import optimizers  # noqa: E402

EX = os.path.join(_ref_shims.REFERENCE_ROOT, 'examples', 'chainermn')


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _layout(model):
    return [[name, list(p.shape), str(p.dtype)] for name, p in sorted(model.namedparams())]


def make_layouts():
    out = {}
    resnet = _load(os.path.join(EX, 'imagenet', 'models', 'resnet50.py'), 'ref_resnet50')
    out['resnet50'] = _layout(resnet.ResNet50())
    mnist = _load(os.path.join(EX, 'mnist', 'train_mnist.py'), 'ref_train_mnist')
    mlp = mnist.MLP(1000, 10)
    mlp(np.zeros((1, 784), dtype=np.float32))          # resolve the lazy input sizes
    out['mnist_mlp'] = _layout(mlp)
    # seq2seq.py imports nltk (absent) and `europal`: stub / path them for the import only
    nltk = types.ModuleType('nltk')
    tr = types.ModuleType('nltk.translate')
    tr.bleu_score = types.ModuleType('nltk.translate.bleu_score')
    nltk.translate = tr
    sys.modules.update({'nltk': nltk, 'nltk.translate': tr,
                        'nltk.translate.bleu_score': tr.bleu_score,
                        'progressbar': types.ModuleType('progressbar')})
    sys.path.insert(0, os.path.join(EX, 'seq2seq'))
    s2s = _load(os.path.join(EX, 'seq2seq', 'seq2seq.py'), 'ref_seq2seq')
    model = s2s.Seq2seq(3, 40000, 40000, 1024)
    out['seq2seq'] = _layout(model)
    with open(os.path.join(HERE, 'layouts.json'), 'w') as f:
        json.dump(out, f, indent=0)
    for k, v in out.items():
        print('layout', k, len(v), sum(int(np.prod(s)) for _, s, _ in v))


class _Net(chainer.Chain):
    """Parameters of the given shapes, registered as a, b, c, ..."""

    def __init__(self, shapes, dtype, rng):
        super(_Net, self).__init__()
        with self.init_scope():
            for i, shape in enumerate(shapes):
                init = np.asarray(rng.standard_normal(shape) * 0.05).astype(dtype).reshape(shape)
                setattr(self, 'p%02d' % i, chainer.Parameter(init))


SHAPES = [(2, 3), (), (1, 0, 2), (257,), (130,), (8, 3, 3, 3)]


def _run_optimizer(make_opt, dtype, n_steps, seed, grad_scale=1e-2, hooks=(), loss_scale=None):
    rng = np.random.default_rng(seed)
    net = _Net(SHAPES, dtype, rng)
    opt = make_opt()
    opt.setup(net)
    for h in hooks:
        opt.add_hook(h)
    rec = {}
    names = [n for n, _ in sorted(net.namedparams())]
    for n, p in sorted(net.namedparams()):
        rec['init' + n] = p.data.copy()
    for step in range(n_steps):
        for n, p in sorted(net.namedparams()):
            g = np.asarray(rng.standard_normal(p.shape) * grad_scale).astype(dtype).reshape(p.shape)
            if loss_scale is not None:
                # what backward(loss_scale=...) leaves behind (variable.py: _loss_scale)
                g = np.asarray(g * dtype.type(loss_scale)).astype(dtype).reshape(p.shape)
                p._loss_scale = loss_scale
            p.grad = g
            rec['grad%d%s' % (step, n)] = g.copy()
        opt.update()
        for n, p in sorted(net.namedparams()):
            rec['param%d%s' % (step, n)] = p.data.copy()
            if hooks or loss_scale is not None:
                rec['gradafter%d%s' % (step, n)] = p.grad.copy()
            for k, s in p.update_rule.state.items():
                rec['state_%s%d%s' % (k, step, n)] = np.array(s, copy=True)
    return names, rec


def make_momentum_sgd():
    out = {}
    for dtype in ('float32', 'float16', 'float64'):
        names, rec = _run_optimizer(lambda: optimizers.MomentumSGD(lr=0.01, momentum=0.9),
                                    np.dtype(dtype), 3, 11)
        for k, v in rec.items():
            out['%s|%s' % (dtype, k)] = v
    np.savez_compressed(os.path.join(HERE, 'momentum_sgd.npz'), **out)
    print('momentum_sgd', len(out))


HOOK_VARIANTS = {
    # name: (optimizer, hook list factory, loss_scale)
    'wd': ('sgd', lambda H: [H.WeightDecay(0.05)], None),
    'clip': ('sgd', lambda H: [H.GradientClipping(0.05)], None),
    'clip_noop': ('sgd', lambda H: [H.GradientClipping(1e3)], None),
    'clip_wd': ('sgd', lambda H: [H.GradientClipping(0.05), H.WeightDecay(0.05)], None),
    'wd_clip': ('sgd', lambda H: [H.WeightDecay(0.05), H.GradientClipping(0.05)], None),
    'ls128': ('sgd', lambda H: [], 128.0),
    'ls100_wd': ('sgd', lambda H: [H.WeightDecay(0.05)], 100.0),
    'ls128_clip_wd': ('sgd', lambda H: [H.GradientClipping(5.0), H.WeightDecay(0.05)], 128.0),
    'adam_clip_wd': ('adam', lambda H: [H.GradientClipping(0.05), H.WeightDecay(0.05)], None),
    'adam_ls128_wd': ('adam', lambda H: [H.WeightDecay(0.05)], 128.0),
}


def make_hooks():
    """MomentumSGD / Adam with optimizer hooks (chainer/optimizer_hooks/weight_decay.py,
    gradient_clipping.py) and static loss scaling (chainer/optimizer.py:286-291), CPU path."""
    from chainer import optimizer_hooks as H
    out = {}
    for variant, (opt_name, mk, ls) in HOOK_VARIANTS.items():
        for dtype in ('float32', 'float16'):
            if opt_name == 'adam' and dtype == 'float16':
                # the reference's float16 CPU Adam underflows v for clipped gradients and
                # diverges from its own GPU formula (see make_adam): not a usable vector
                continue
            if opt_name == 'sgd':
                make_opt = lambda: optimizers.MomentumSGD(lr=0.01, momentum=0.9)   # noqa: E731
                gs = 1e-2
            else:
                make_opt = lambda: optimizers.Adam()                               # noqa: E731
                gs = 0.5 if dtype == 'float16' else 1e-2
            names, rec = _run_optimizer(make_opt, np.dtype(dtype), 3, 17, grad_scale=gs,
                                        hooks=mk(H), loss_scale=ls)
            for k, v in rec.items():
                out['%s|%s|%s' % (variant, dtype, k)] = v
    np.savez_compressed(os.path.join(HERE, 'hooks.npz'), **out)
    print('hooks', len(out))


def make_sgd_family():
    """chainer.optimizers.SGD / CorrectedMomentumSGD / NesterovAG, 3 steps, f16/f32/f64, and
    each once more with [GradientClipping, WeightDecay] hooks (float32)."""
    from chainer import optimizer_hooks as H
    out = {}
    makers = {
        'sgd': lambda: optimizers.SGD(lr=0.05),
        'corrected': lambda: optimizers.CorrectedMomentumSGD(lr=0.05, momentum=0.8),
        'nesterov': lambda: optimizers.NesterovAG(lr=0.05, momentum=0.8),
    }
    for rule, mk in makers.items():
        for dtype in ('float32', 'float16', 'float64'):
            names, rec = _run_optimizer(mk, np.dtype(dtype), 3, 19)
            for k, v in rec.items():
                out['%s|%s|%s' % (rule, dtype, k)] = v
        names, rec = _run_optimizer(mk, np.dtype('float32'), 3, 23,
                                    hooks=[H.GradientClipping(0.05), H.WeightDecay(0.05)])
        for k, v in rec.items():
            out['%s_clip_wd|float32|%s' % (rule, k)] = v
    np.savez_compressed(os.path.join(HERE, 'sgd_family.npz'), **out)
    print('sgd_family', len(out))


def make_fp32_update():
    """float16 parameters with fp32 master weights (use_fp32_update,
    chainer/optimizer.py:262-305): MomentumSGD, MomentumSGD + WeightDecay + loss scale 128,
    Adam.  Records the float16 parameters, the float32 masters and states."""
    from chainer import optimizer_hooks as H
    out = {}
    cases = {
        'sgd': (lambda: optimizers.MomentumSGD(lr=0.01, momentum=0.9), [], None, 1e-2),
        'sgd_wd_ls128': (lambda: optimizers.MomentumSGD(lr=0.01, momentum=0.9),
                         [H.WeightDecay(0.05)], 128.0, 1e-2),
        'adam': (lambda: optimizers.Adam(), [], None, 0.5),
    }
    dt = np.dtype('float16')
    for case, (mk, hooks, ls, gs) in cases.items():
        rng = np.random.default_rng(31)
        net = _Net(SHAPES, dt, rng)
        opt = mk()
        opt.use_fp32_update()
        opt.setup(net)
        for h in hooks:
            opt.add_hook(h)
        for n, p in sorted(net.namedparams()):
            out['%s|init%s' % (case, n)] = p.data.copy()
        for step in range(3):
            for n, p in sorted(net.namedparams()):
                g = np.asarray(rng.standard_normal(p.shape) * gs).astype(dt).reshape(p.shape)
                if ls is not None:
                    g = np.asarray(g * dt.type(ls)).astype(dt).reshape(p.shape)
                    p._loss_scale = ls
                p.grad = g
                out['%s|grad%d%s' % (case, step, n)] = g.copy()
            opt.update()
            for n, p in sorted(net.namedparams()):
                assert p.data.dtype == np.float16
                out['%s|param%d%s' % (case, step, n)] = p.data.copy()
                out['%s|master%d%s' % (case, step, n)] = p.update_rule._fp32_param.data.copy()
                for k, st in p.update_rule.state.items():
                    assert st.dtype == np.float32
                    out['%s|state_%s%d%s' % (case, k, step, n)] = np.array(st, copy=True)
    np.savez_compressed(os.path.join(HERE, 'fp32_update.npz'), **out)
    print('fp32_update', len(out))


def make_dynamic_loss_scale():
    """MomentumSGD with dynamic loss scaling (chainer/optimizer.py:736-791, 881-894): 6
    steps, a non-finite gradient in step 2; records the loss scale and the parameters."""
    out = {}
    for dtype in ('float32', 'float16'):
        dt = np.dtype(dtype)
        rng = np.random.default_rng(29)
        net = _Net(SHAPES, dt, rng)
        opt = optimizers.MomentumSGD(lr=0.01, momentum=0.9)
        opt.setup(net)
        opt.loss_scaling(interval=2)
        for n, p in sorted(net.namedparams()):
            out['%s|init%s' % (dtype, n)] = p.data.copy()
        scales = []
        for step in range(6):
            ls = opt._loss_scale
            for n, p in sorted(net.namedparams()):
                g = np.asarray(rng.standard_normal(p.shape) * 1e-2).astype(dt).reshape(p.shape)
                if step == 2 and n == '/p03':
                    g[5] = np.inf
                out['%s|grad%d%s' % (dtype, step, n)] = g.copy()        # unscaled
                p.grad = np.asarray(g * dt.type(ls)).astype(dt).reshape(p.shape)
                p._loss_scale = ls
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                opt.update()
            scales.append(opt._loss_scale)
            for n, p in sorted(net.namedparams()):
                out['%s|param%d%s' % (dtype, step, n)] = p.data.copy()
        out['%s|scales' % dtype] = np.asarray(scales, dtype=np.float64)
        out['%s|t' % dtype] = np.asarray([opt.t] + [p.update_rule.t for _, p in sorted(net.namedparams())])
    np.savez_compressed(os.path.join(HERE, 'dynamic_loss_scale.npz'), **out)
    print('dynamic_loss_scale', len(out), out['float32|scales'], out['float32|t'])


def make_fp32_dynamic():
    """float16 parameters, float32 master weights AND dynamic loss scaling together (the
    float16 training recipe; chainer/optimizer.py:262-305, 736-791, 857-894): MomentumSGD
    with WeightDecay, and Adam; 6 steps, a non-finite gradient in step 2.  Records the scaled
    gradients fed in, the float16 parameters, the float32 masters, the loss scales and t."""
    from chainer import optimizer_hooks as H
    out = {}
    cases = {
        'sgd_wd': (lambda: optimizers.MomentumSGD(lr=0.01, momentum=0.9), [H.WeightDecay(0.05)], 1e-2),
        'adam': (lambda: optimizers.Adam(), [], 0.5),
    }
    dt = np.dtype('float16')
    for case, (mk, hooks, gs) in cases.items():
        rng = np.random.default_rng(37)
        net = _Net(SHAPES, dt, rng)
        opt = mk()
        opt.use_fp32_update()
        opt.setup(net)
        for h in hooks:
            opt.add_hook(h)
        opt.loss_scaling(interval=2)
        for n, p in sorted(net.namedparams()):
            out['%s|init%s' % (case, n)] = p.data.copy()
        scales = []
        for step in range(6):
            ls = opt._loss_scale
            for n, p in sorted(net.namedparams()):
                g = np.asarray(rng.standard_normal(p.shape) * gs).astype(dt).reshape(p.shape)
                if step == 2 and n == '/p03':
                    g[5] = np.inf
                out['%s|grad%d%s' % (case, step, n)] = g.copy()        # unscaled
                p.grad = np.asarray(g * dt.type(ls)).astype(dt).reshape(p.shape)
                p._loss_scale = ls
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                opt.update()
            scales.append(opt._loss_scale)
            for n, p in sorted(net.namedparams()):
                out['%s|param%d%s' % (case, step, n)] = p.data.copy()
                out['%s|master%d%s' % (case, step, n)] = p.update_rule._fp32_param.data.copy()
        out['%s|scales' % case] = np.asarray(scales, dtype=np.float64)
        out['%s|t' % case] = np.asarray([opt.t] + [p.update_rule.t for _, p in sorted(net.namedparams())])
    np.savez_compressed(os.path.join(HERE, 'fp32_dynamic.npz'), **out)
    print('fp32_dynamic', len(out), out['sgd_wd|scales'], out['sgd_wd|t'])


ADAM_VARIANTS = {
    'adam': dict(),
    'adamw': dict(eta=0.5, weight_decay_rate=0.1),
    'amsgrad': dict(amsgrad=True),
    'adabound': dict(adabound=True),
    'amsbound': dict(amsgrad=True, adabound=True),
}


def make_adam():
    out = {}
    for variant, kw in ADAM_VARIANTS.items():
        for dtype in ('float32', 'float16', 'float64'):
            # float16: gradients of O(1) so that v = (1-beta2) g^2 stays a normal half
            # number (with 1e-2 it underflows in the reference's float16 state and
            # its CPU and GPU code paths legitimately diverge)
            names, rec = _run_optimizer(lambda: optimizers.Adam(**kw), np.dtype(dtype), 3, 13,
                                        grad_scale=0.5 if dtype == 'float16' else 1e-2)
            for k, v in rec.items():
                out['%s|%s|%s' % (variant, dtype, k)] = v
    # the reference's own known-answer test: tests/chainer_tests/optimizers_tests/
    # test_optimizers.py:278-310 (TestAdamW) x=1, g=1, eta=.5, wd=.1 -> 0.9495
    link = chainer.Link(x=(1,))
    link.x.data.fill(1)
    link.x.grad = np.ones_like(link.x.data)
    opt = optimizers.Adam(eta=0.5, weight_decay_rate=0.1)
    opt.setup(link)
    opt.update()
    out['kat|adamw'] = link.x.data.copy()
    np.savez_compressed(os.path.join(HERE, 'adam.npz'), **out)
    print('adam', len(out), 'AdamW KAT', out['kat|adamw'])


def make_naive_mean_grad():
    """NaiveCommunicator.multi_node_mean_grad (naive_communicator.py:10-17 ->
    mpi_communicator_base.py:735-778) on `size` ranks, rank-dependent grads,
    including a None grad with zero_fill and an uninitialised parameter."""
    from chainermn.communicators.naive_communicator import NaiveCommunicator
    out = {}
    for size in (2, 3):
        for dtype in ('float32', 'float16', 'float64'):
            def fn(mpi_comm, rank, dtype=dtype):
                comm = NaiveCommunicator(mpi_comm)
                rng = np.random.default_rng(7)
                net = _Net([(2, 3), (5,), (3, 4), (0,), (33,)], np.dtype(dtype), rng)
                with net.init_scope():
                    net.lazy = L.Linear(None, 5)              # W uninitialised: skipped
                grng = np.random.default_rng(1000 + rank)
                grads = {}
                for n, p in sorted(net.namedparams()):
                    if p.data is None:
                        continue
                    g = np.asarray(grng.standard_normal(p.shape) * 1e-2).astype(p.dtype).reshape(p.shape)
                    if n == '/p01' and rank % 2 == 1:
                        p.grad = None                          # zero_fill case
                        g = np.zeros(p.shape, dtype=p.dtype)
                    else:
                        p.grad = g.copy()
                    grads[n] = g
                comm.multi_node_mean_grad(net, zero_fill=True)
                res = {n: p.grad.copy() for n, p in sorted(net.namedparams())
                       if p.data is not None}
                return grads, res
            results = _ref_shims.run_ranks(size, fn)
            for r, (grads, res) in enumerate(results):
                for n, g in grads.items():
                    out['%d|%s|in|%d|%s' % (size, dtype, r, n)] = g
                for n, g in res.items():
                    out['%d|%s|out|%d|%s' % (size, dtype, r, n)] = g
    np.savez_compressed(os.path.join(HERE, 'naive_mean_grad.npz'), **out)
    print('naive_mean_grad', len(out))


def make_mnbn():
    """MultiNodeBatchNormalization with the MPI backend (`_MpiImpl`,
    chainermn/functions/batch_normalization.py:7-32) on 2 ranks: forward output,
    running statistics, and gradients; plus the single-process BatchNormalization
    on the concatenated batch (the equivalence the reference tests,
    tests/chainermn_tests/links_tests/test_batch_normalization.py:54-186)."""
    from chainermn.communicators.naive_communicator import NaiveCommunicator
    from chainermn.links import MultiNodeBatchNormalization
    out = {}
    size, nb, C, H, W = 2, 4, 6, 5, 3
    rng = np.random.default_rng(71)
    x_all = rng.standard_normal((size * nb, C, H, W)).astype(np.float32)
    gy_all = (rng.standard_normal((size * nb, C, H, W)) * 1e-1).astype(np.float32)
    gamma0 = rng.uniform(0.5, 1.5, C).astype(np.float32)
    beta0 = rng.uniform(-0.5, 0.5, C).astype(np.float32)
    out['x'], out['gy'], out['gamma'], out['beta'] = x_all, gy_all, gamma0, beta0

    def fn(mpi_comm, rank):
        comm = NaiveCommunicator(mpi_comm)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            bn = MultiNodeBatchNormalization(C, comm, communication_backend='mpi')
        bn.gamma.data[...] = gamma0
        bn.beta.data[...] = beta0
        x = chainer.Variable(x_all[rank * nb:(rank + 1) * nb].copy())
        with chainer.using_config('train', True):
            y = bn(x)
        y.grad = gy_all[rank * nb:(rank + 1) * nb].copy()
        bn.cleargrads()
        y.backward()
        return dict(y=y.data.copy(), gx=x.grad.copy(), ggamma=bn.gamma.grad.copy(),
                    gbeta=bn.beta.grad.copy(), avg_mean=bn.avg_mean.copy(),
                    avg_var=bn.avg_var.copy())
    for r, res in enumerate(_ref_shims.run_ranks(size, fn)):
        for k, v in res.items():
            out['mn|%d|%s' % (r, k)] = v
    # single process, whole batch, stock BatchNormalization
    bn = L.BatchNormalization(C)
    bn.gamma.data[...] = gamma0
    bn.beta.data[...] = beta0
    x = chainer.Variable(x_all.copy())
    with chainer.using_config('train', True):
        y = bn(x)
    y.grad = gy_all.copy()
    bn.cleargrads()
    y.backward()
    out['single|y'] = y.data.copy()
    out['single|gx'] = x.grad.copy()
    out['single|ggamma'] = bn.gamma.grad.copy()
    out['single|gbeta'] = bn.beta.grad.copy()
    np.savez_compressed(os.path.join(HERE, 'mnbn.npz'), **out)
    print('mnbn', len(out))


if __name__ == '__main__':
    make_layouts()
    make_momentum_sgd()
    make_adam()
    make_naive_mean_grad()
    make_mnbn()
    make_hooks()
    make_sgd_family()
    make_dynamic_loss_scale()
    make_fp32_update()
    make_fp32_dynamic()
