"""Adam / AdamW / AMSGrad / AdaBound: mirror of ``chainer/optimizers/adam.py``."""
import math
import warnings

import numpy as np

from chainer_b200 import _lib
from chainer_b200 import device as _dev
from chainer_b200.core import optimizer
from chainer_b200.core.optimizers import _single

# name -> default of every hyperparameter of the Adam family (values: adam.py:34-44)
_DEFAULTS = (('alpha', 0.001), ('beta1', 0.9), ('beta2', 0.999), ('eps', 1e-8), ('eta', 1.0),
             ('weight_decay_rate', 0), ('amsgrad', False), ('adabound', False),
             ('final_lr', 0.1), ('gamma', 1e-3))
_NAMES = tuple(name for name, _ in _DEFAULTS)
_default_hyperparam = optimizer.Hyperparameter()
for _name, _value in _DEFAULTS:
    setattr(_default_hyperparam, _name, _value)


def _learning_rate(hp, t):
    """Bias-corrected step size ``alpha * sqrt(1 - beta2^t) / (1 - beta1^t)``, evaluated in
    double with ``math.pow`` / ``math.sqrt`` like the reference (``adam.py:47-54``) so that
    the scalar handed to the kernel has the same bits."""
    if t == 0:
        raise RuntimeError(
            'Can\'t determine the learning rate of Adam optimizer '
            'because the update steps have not been started.')
    return hp.alpha * math.sqrt(1. - math.pow(hp.beta2, t)) / (1. - math.pow(hp.beta1, t))


def _renamed_lr(obj, owner):
    """The deprecated ``lr`` alias of ``alpha_t`` (``adam.py:334-344, 437-446``)."""
    warnings.warn('{0}.lr has been renamed to AdamRule.alpha_t. '
                  'Use of {0}.lr is deprecated in Chainer v6.'.format(owner), DeprecationWarning)
    return obj.alpha_t


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
        eps = self.hyperparam.eps
        if eps != 0 and interm_dtype(eps) == 0:
            raise ValueError('eps of Adam optimizer is too small for {} ({})'.format(
                np.dtype(interm_dtype).name, eps))

    alpha_t = property(lambda self: _learning_rate(self.hyperparam, self.t))
    lr = property(lambda self: _renamed_lr(self, 'AdamRule'))

    @property
    def bounds(self):
        """AdaBound's (lower, upper) clip of the per-element step at this rule's ``t``
        (``adam.py:346-358``): both converge to ``final_lr`` (rescaled by how far alpha has
        moved from its initial value) as ``gamma * t`` grows."""
        t = self.t
        if t == 0:
            raise RuntimeError('Can\'t determine the bounds of AdaBound optimizer '
                               'because the update steps have not been started.')
        hp = self.hyperparam
        target = hp.final_lr * hp.alpha / self.initial_alpha
        gt = hp.gamma * t
        return target * (1.0 - 1.0 / (gt + 1)), target * (1.0 + 1.0 / gt)

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
    """``chainer.optimizers.Adam`` (``adam.py:361-446``): keyword arguments and attribute
    proxies for every entry of ``_DEFAULTS``."""

    def __init__(self, *args, **kwargs):
        if len(args) > len(_NAMES):
            raise TypeError('Adam() takes at most {} arguments'.format(len(_NAMES)))
        for name, value in zip(_NAMES, args):        # positional, in the reference's order
            if name in kwargs:
                raise TypeError('Adam() got multiple values for argument {!r}'.format(name))
            kwargs[name] = value
        unknown = set(kwargs) - set(_NAMES)
        if unknown:
            raise TypeError('Adam() got unexpected keyword arguments {}'.format(sorted(unknown)))
        super(Adam, self).__init__()
        for name, default in _DEFAULTS:
            setattr(self.hyperparam, name, kwargs.get(name, default))

    def create_update_rule(self):
        return AdamRule(self.hyperparam)

    alpha_t = property(lambda self: _learning_rate(self.hyperparam, self.t))
    lr = property(lambda self: _renamed_lr(self, 'Adam'))


for _name in _NAMES:
    setattr(Adam, _name, optimizer.HyperparameterProxy(_name))
