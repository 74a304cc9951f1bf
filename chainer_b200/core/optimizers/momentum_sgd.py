"""MomentumSGD: mirror of ``chainer/optimizers/momentum_sgd.py``."""
from chainer_b200 import _lib
from chainer_b200 import device as _dev
from chainer_b200.core import optimizer
from chainer_b200.core.optimizers import _single

_default_hyperparam = optimizer.Hyperparameter()
_default_hyperparam.lr = 0.01
_default_hyperparam.momentum = 0.9


class MomentumSGDRule(optimizer.UpdateRule):
    """``momentum_sgd.py:25-88``: ``v = momentum * v - lr * grad; param += v``."""

    is_elementwise = True
    fused_kind = 'momentum_sgd'     # handled by gp_unpack_momentum_sgd
    state_names = ('v',)

    def __init__(self, parent_hyperparam=None, lr=None, momentum=None):
        super(MomentumSGDRule, self).__init__(parent_hyperparam or _default_hyperparam)
        if lr is not None:
            self.hyperparam.lr = lr
        if momentum is not None:
            self.hyperparam.momentum = momentum

    def init_state(self, param):
        self.state['v'] = _dev.zeros_like(param.data)

    def fused_key(self):
        """Launch-group key: parameters with equal keys share one fused launch."""
        hp = self.hyperparam
        return ('momentum_sgd', float(hp.lr), float(hp.momentum))

    def fused_signature(self):
        return self.fused_key() + (self.t,)

    def update_core_gpu(self, param):
        grad = param.grad
        if grad is None:
            return
        hp = self.hyperparam
        pd = _single.single_param_table(param, [self.state['v']])
        _lib.get().gp_unpack_momentum_sgd(
            _dev.device_ptr(grad), _dev.dtype_id(_dev.array_dtype(grad)), pd.d_csum, pd.d_segs,
            1, 0, pd.n_elems, 1.0, float(hp.lr), float(hp.momentum), 0, 0, 0)


class MomentumSGD(optimizer.GradientMethod):
    """``momentum_sgd.py:91-116``."""

    def __init__(self, lr=_default_hyperparam.lr, momentum=_default_hyperparam.momentum):
        super(MomentumSGD, self).__init__()
        self.hyperparam.lr = lr
        self.hyperparam.momentum = momentum

    lr = optimizer.HyperparameterProxy('lr')
    momentum = optimizer.HyperparameterProxy('momentum')

    def create_update_rule(self):
        return MomentumSGDRule(self.hyperparam)
