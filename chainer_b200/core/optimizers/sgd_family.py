"""SGD, CorrectedMomentumSGD and NesterovAG: mirrors of ``chainer/optimizers/sgd.py``,
``corrected_momentum_sgd.py`` and ``nesterov_ag.py``.  One kernel family
(``csrc/gp_sgd_family.cu``) serves the three rules, fused behind the multi-node
optimizer like MomentumSGD and Adam, and multi-tensor when used stand-alone."""
from chainer_b200 import _lib
from chainer_b200 import device as _dev
from chainer_b200.core import optimizer
from chainer_b200.core.optimizers import _single

_sgd_default = optimizer.Hyperparameter()
_sgd_default.lr = 0.01

_momentum_default = optimizer.Hyperparameter()
_momentum_default.lr = 0.01
_momentum_default.momentum = 0.9


class _FamilyRule(optimizer.UpdateRule):
    is_elementwise = True
    fused_kind = 'sgd_family'       # handled by gp_unpack_sgd_family
    rule_id = None
    state_names = ('v',)

    def _momentum(self):
        return float(self.hyperparam.momentum)

    def init_state(self, param):
        for name in self.state_names:
            self.state[name] = _dev.zeros_like(param.data)

    def fused_key(self):
        return ('sgd_family', self.rule_id, float(self.hyperparam.lr), self._momentum())

    def fused_signature(self):
        return self.fused_key() + (self.t,)

    def update_core_gpu(self, param):
        grad = param.grad
        if grad is None:
            return
        pd = _single.single_param_table(param, [self.state[k] for k in self.state_names])
        _lib.get().gp_unpack_sgd_family(
            _dev.device_ptr(grad), _dev.dtype_id(_dev.array_dtype(grad)), pd.d_csum, pd.d_segs,
            1, 0, pd.n_elems, 1.0, self.rule_id, float(self.hyperparam.lr), self._momentum(),
            0, 0, None, 0)


class SGDRule(_FamilyRule):
    """``sgd.py:22-63``: ``param -= lr * grad``."""
    rule_id = _lib.RULE_SGD
    state_names = ()

    def __init__(self, parent_hyperparam=None, lr=None):
        super(SGDRule, self).__init__(parent_hyperparam or _sgd_default)
        if lr is not None:
            self.hyperparam.lr = lr

    def _momentum(self):
        return 0.0


class SGD(optimizer.GradientMethod):
    """``sgd.py:66-83``."""

    def __init__(self, lr=_sgd_default.lr):
        super(SGD, self).__init__()
        self.hyperparam.lr = lr

    lr = optimizer.HyperparameterProxy('lr')

    def create_update_rule(self):
        return SGDRule(self.hyperparam)


class CorrectedMomentumSGDRule(_FamilyRule):
    """``corrected_momentum_sgd.py:22-89``: ``v = momentum * v - grad; param += lr * v``."""
    rule_id = _lib.RULE_CORRECTED_MOMENTUM

    def __init__(self, parent_hyperparam=None, lr=None, momentum=None):
        super(CorrectedMomentumSGDRule, self).__init__(parent_hyperparam or _momentum_default)
        if lr is not None:
            self.hyperparam.lr = lr
        if momentum is not None:
            self.hyperparam.momentum = momentum


class CorrectedMomentumSGD(optimizer.GradientMethod):
    """``corrected_momentum_sgd.py:92-150``."""

    def __init__(self, lr=_momentum_default.lr, momentum=_momentum_default.momentum):
        super(CorrectedMomentumSGD, self).__init__()
        self.hyperparam.lr = lr
        self.hyperparam.momentum = momentum

    lr = optimizer.HyperparameterProxy('lr')
    momentum = optimizer.HyperparameterProxy('momentum')

    def create_update_rule(self):
        return CorrectedMomentumSGDRule(self.hyperparam)


class NesterovAGRule(_FamilyRule):
    """``nesterov_ag.py:22-84`` (the arithmetic of ``update_core_cpu``)."""
    rule_id = _lib.RULE_NESTEROV_AG

    def __init__(self, parent_hyperparam=None, lr=None, momentum=None):
        super(NesterovAGRule, self).__init__(parent_hyperparam or _momentum_default)
        if lr is not None:
            self.hyperparam.lr = lr
        if momentum is not None:
            self.hyperparam.momentum = momentum


class NesterovAG(optimizer.GradientMethod):
    """``nesterov_ag.py:87-112``."""

    def __init__(self, lr=_momentum_default.lr, momentum=_momentum_default.momentum):
        super(NesterovAG, self).__init__()
        self.hyperparam.lr = lr
        self.hyperparam.momentum = momentum

    lr = optimizer.HyperparameterProxy('lr')
    momentum = optimizer.HyperparameterProxy('momentum')

    def create_update_rule(self):
        return NesterovAGRule(self.hyperparam)
