"""Adam / AdamW / AMSGrad / AdaBound: mirror of ``chainer/optimizers/adam.py``."""
import math
import warnings

import numpy as np

from chainer_b200 import _lib
from chainer_b200 import device as _dev
from chainer_b200.core import optimizer
from chainer_b200.core.optimizers import _single

_default_hyperparam = optimizer.Hyperparameter()
_default_hyperparam.alpha = 0.001
_default_hyperparam.beta1 = 0.9
_default_hyperparam.beta2 = 0.999
_default_hyperparam.eps = 1e-8
_default_hyperparam.eta = 1.0
_default_hyperparam.weight_decay_rate = 0
_default_hyperparam.amsgrad = False
_default_hyperparam.adabound = False
_default_hyperparam.final_lr = 0.1
_default_hyperparam.gamma = 1e-3


def _learning_rate(hp, t):
    """``adam.py:47-54``."""
    if t == 0:
        raise RuntimeError(
            'Can\'t determine the learning rate of Adam optimizer '
            'because the update steps have not been started.')
    fix1 = 1. - math.pow(hp.beta1, t)
    fix2 = 1. - math.pow(hp.beta2, t)
    return hp.alpha * math.sqrt(fix2) / fix1


def _get_intermediate_dtype(dtype):
    """``adam.py:57-63``."""
    if dtype == np.float16:
        return np.float32
    return dtype


class AdamRule(optimizer.UpdateRule):
    """``adam.py:77-358``."""

    is_elementwise = True
    fused_kind = 'adam'             # handled by gp_unpack_adam

    def __init__(self, parent_hyperparam=None, alpha=None, beta1=None, beta2=None, eps=None,
                 eta=None, weight_decay_rate=None, amsgrad=None, adabound=None, final_lr=None,
                 gamma=None):
        super(AdamRule, self).__init__(parent_hyperparam or _default_hyperparam)
        for name, value in (('alpha', alpha), ('beta1', beta1), ('beta2', beta2), ('eps', eps),
                            ('eta', eta), ('weight_decay_rate', weight_decay_rate),
                            ('amsgrad', amsgrad), ('adabound', adabound),
                            ('final_lr', final_lr), ('gamma', gamma)):
            if value is not None:
                setattr(self.hyperparam, name, value)
        if self.hyperparam.adabound:
            self.initial_alpha = self.hyperparam.alpha

    @property
    def state_names(self):
        return ('m', 'v', 'vhat') if self.hyperparam.amsgrad else ('m', 'v')

    def init_state(self, param):
        self.state['m'] = _dev.zeros_like(param.data)
        self.state['v'] = _dev.zeros_like(param.data)
        if self.hyperparam.amsgrad:
            self.state['vhat'] = _dev.zeros_like(param.data)

    def _check_eps(self, interm_dtype):
        """``adam.py:176-187``: eps must not underflow in the intermediate dtype."""
        hp = self.hyperparam
        eps = interm_dtype(hp.eps)
        if hp.eps != 0 and eps == 0:
            raise ValueError(
                'eps of Adam optimizer is too small for {} ({})'.format(
                    np.dtype(interm_dtype).name, hp.eps))

    @property
    def alpha_t(self):
        return _learning_rate(self.hyperparam, self.t)

    @property
    def lr(self):
        warnings.warn(
            'AdamRule.lr has been renamed to AdamRule.alpha_t. '
            'Use of AdamRule.lr is deprecated in Chainer v6.',
            DeprecationWarning)
        return self.alpha_t

    @property
    def bounds(self):
        """``adam.py:346-358``."""
        if self.t == 0:
            raise RuntimeError(
                'Can\'t determine the bounds of AdaBound optimizer '
                'because the update steps have not been started.')
        hp = self.hyperparam
        final_lr = hp.final_lr * hp.alpha / self.initial_alpha
        lower = final_lr * (1.0 - 1.0 / (hp.gamma * self.t + 1))
        upper = final_lr * (1.0 + 1.0 / (hp.gamma * self.t))
        return lower, upper

    def kernel_args(self):
        """(alpha_t, 1-beta1, 1-beta2, eps, eta, wd, lower, upper, flags) for
        gp_unpack_adam, evaluated with this rule's own ``t``."""
        hp = self.hyperparam
        lower, upper = self.bounds if hp.adabound else (0.0, 0.0)
        flags = (_lib.GP_ADAM_AMSGRAD if hp.amsgrad else 0) | \
            (_lib.GP_ADAM_ADABOUND if hp.adabound else 0)
        return (float(self.alpha_t), float(1 - hp.beta1), float(1 - hp.beta2), float(hp.eps),
                float(hp.eta), float(hp.weight_decay_rate), float(lower), float(upper), flags)

    def fused_key(self):
        return ('adam',) + self.kernel_args()

    def fused_signature(self):
        """Grouping key that is defined before the first step (t == 0)."""
        d = self.hyperparam.get_dict()
        return ('adam', self.t, getattr(self, 'initial_alpha', None)) + \
            tuple(sorted((k, float(v)) for k, v in d.items()))

    def update_core_gpu(self, param):
        grad = param.grad
        if grad is None:
            return
        dtype = _get_intermediate_dtype(param.dtype.type)
        self._check_eps(dtype)
        states = [self.state['m'], self.state['v']]
        if self.hyperparam.amsgrad:
            states.append(self.state['vhat'])
        pd = _single.single_param_table(param, states)
        a = self.kernel_args()
        _lib.get().gp_unpack_adam(
            _dev.device_ptr(grad), _dev.dtype_id(_dev.array_dtype(grad)), pd.d_csum, pd.d_segs,
            1, 0, pd.n_elems, 1.0, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], 0, 0, 0)


class Adam(optimizer.GradientMethod):
    """``adam.py:361-446``."""

    def __init__(self, alpha=_default_hyperparam.alpha, beta1=_default_hyperparam.beta1,
                 beta2=_default_hyperparam.beta2, eps=_default_hyperparam.eps,
                 eta=_default_hyperparam.eta,
                 weight_decay_rate=_default_hyperparam.weight_decay_rate,
                 amsgrad=_default_hyperparam.amsgrad, adabound=_default_hyperparam.adabound,
                 final_lr=_default_hyperparam.final_lr, gamma=_default_hyperparam.gamma):
        super(Adam, self).__init__()
        self.hyperparam.alpha = alpha
        self.hyperparam.beta1 = beta1
        self.hyperparam.beta2 = beta2
        self.hyperparam.eps = eps
        self.hyperparam.eta = eta
        self.hyperparam.weight_decay_rate = weight_decay_rate
        self.hyperparam.amsgrad = amsgrad
        self.hyperparam.adabound = adabound
        self.hyperparam.final_lr = final_lr
        self.hyperparam.gamma = gamma

    alpha = optimizer.HyperparameterProxy('alpha')
    beta1 = optimizer.HyperparameterProxy('beta1')
    beta2 = optimizer.HyperparameterProxy('beta2')
    eps = optimizer.HyperparameterProxy('eps')
    eta = optimizer.HyperparameterProxy('eta')
    weight_decay_rate = optimizer.HyperparameterProxy('weight_decay_rate')
    amsgrad = optimizer.HyperparameterProxy('amsgrad')
    adabound = optimizer.HyperparameterProxy('adabound')
    final_lr = optimizer.HyperparameterProxy('final_lr')
    gamma = optimizer.HyperparameterProxy('gamma')

    def create_update_rule(self):
        return AdamRule(self.hyperparam)

    @property
    def alpha_t(self):
        return _learning_rate(self.hyperparam, self.t)

    @property
    def lr(self):
        warnings.warn(
            'Adam.lr has been renamed to AdamRule.alpha_t. '
            'Use of Adam.lr is deprecated in Chainer v6.',
            DeprecationWarning)
        return self.alpha_t
