"""``Hyperparameter`` / ``UpdateRule`` / ``Optimizer`` / ``GradientMethod``.

Mirror of the parts of ``chainer/optimizer.py`` that the gradient path drives:

* ``Hyperparameter`` with parent chain              -- ``optimizer.py:92-145``
* ``UpdateRule.update / __update / _init_states / update_core``
                                                    -- ``optimizer.py:236-336, 473-484``
* ``Optimizer.setup / add_hook / call_hooks / loss scaling bookkeeping``
                                                    -- ``optimizer.py:560-791``
* ``GradientMethod.update / reallocate_cleared_grads`` -- ``optimizer.py:834-894``

``update_core`` always dispatches to ``update_core_gpu``: this package has no
CPU path (the reference's ``update_core_cpu`` lives in ``oracle/`` as the
checker only).
"""
import collections
import copy
import math
import warnings

import numpy as np

from chainer_b200 import device as _dev


# Bumped whenever something that decides HOW parameters are updated changes:
# per-rule hyperparameter overrides, enabled flags, hooks, fp32-update,
# loss scaling, state (de)serialisation.  Values of the optimizer-level
# hyperparameters (lr, alpha, ...) are NOT covered: they are re-read every step.
_rules_version = [0]


def rules_version():
    return _rules_version[0]


class _Tracked(object):
    """Attribute whose assignment bumps the rules version."""

    def __init__(self, name, default):
        self._name = '_tracked_' + name
        self._default = default

    def __get__(self, obj, type=None):
        if obj is None:
            return self
        return obj.__dict__.get(self._name, self._default)

    def __set__(self, obj, value):
        _rules_version[0] += 1
        obj.__dict__[self._name] = value


class Hyperparameter(object):
    """Set of hyperparameter entries with a parent to fall back to
    (``optimizer.py:92-145``)."""

    def __init__(self, parent=None):
        self._parent = parent

    def __setattr__(self, name, value):
        # an override on a rule-level (child) hyperparameter changes the launch
        # grouping of the fused update
        if self.__dict__.get('_parent') is not None or name == '_parent':
            _rules_version[0] += 1
        object.__setattr__(self, name, value)

    def __getattr__(self, name):
        if '_parent' not in self.__dict__:
            raise AttributeError('_parent is not set up yet')
        parent = self.__dict__['_parent']
        if parent is None:
            raise AttributeError(name)
        return getattr(parent, name)

    def __repr__(self):
        d = self.get_dict()
        keys = sorted(d.keys())
        values_repr = ', '.join('%s=%s' % (k, d[k]) for k in keys)
        return 'Hyperparameter(%s)' % values_repr

    @property
    def parent(self):
        return self._parent

    def get_dict(self):
        d = {} if self._parent is None else self._parent.get_dict()
        for k, v in self.__dict__.items():
            if k != '_parent':
                d[k] = v
        return d


class HyperparameterProxy(object):
    """``optimizer.HyperparameterProxy``: alias of ``self.hyperparam.<name>``."""

    def __init__(self, attr_name):
        self._attr_name = attr_name
        self.__doc__ = 'Alias to ``self.hyperparam.{}``'.format(attr_name)

    def __get__(self, obj, type=None):
        if obj is None:
            return self
        return getattr(obj.hyperparam, self._attr_name)

    def __set__(self, obj, value):
        setattr(obj.hyperparam, self._attr_name, value)


class _Hookable(object):
    def __init__(self):
        self._pre = collections.OrderedDict()
        self._post = collections.OrderedDict()

    def add_hook(self, hook, name=None, timing='auto'):
        if not callable(hook):
            raise TypeError('hook function must be callable')
        if timing not in ('pre', 'post', 'auto'):
            raise ValueError("timing must be one of ('pre', 'post', 'auto')")
        if timing == 'auto':
            timing = getattr(hook, 'timing', 'pre')
        if name is None:
            name = getattr(hook, 'name', getattr(hook, '__name__', None))
            if name is None:
                raise ValueError('the name of the hook function is not specified')
        if name in self._pre or name in self._post:
            raise KeyError('hook "{}" already exists'.format(name))
        (self._pre if timing == 'pre' else self._post)[name] = hook
        _rules_version[0] += 1

    def remove_hook(self, name):
        if name in self._pre:
            del self._pre[name]
        elif name in self._post:
            del self._post[name]
        else:
            raise KeyError('hook "{}" does not exist'.format(name))
        _rules_version[0] += 1

    def has_hooks(self):
        return bool(self._pre) or bool(self._post)

    def call_hooks(self, timing, args):
        for hook in list((self._pre if timing == 'pre' else self._post).values()):
            hook(*args)


class UpdateRule(object):
    """Base class of all update rules (``optimizer.py:148-530``)."""

    is_elementwise = False
    _b200_versioned = True
    enabled = _Tracked('enabled', True)
    _use_fp32_update = _Tracked('use_fp32_update', False)
    hyperparam = _Tracked('hyperparam', None)

    def __init__(self, parent_hyperparam=None):
        self._state = None
        self.enabled = True
        self.hyperparam = Hyperparameter(parent_hyperparam)
        self.t = 0
        self._use_fp32_update = False
        self._fp32_param = None
        self._hookable = _Hookable()

    @property
    def state(self):
        return self._state

    def add_hook(self, hook, name=None, timing='auto'):
        self._hookable.add_hook(hook, name, timing)

    def remove_hook(self, name):
        self._hookable.remove_hook(name)

    def update(self, param):
        """``UpdateRule.update`` (``optimizer.py:236-250``)."""
        if not self.enabled:
            return
        self.t += 1
        self.__update(param)

    def __update(self, param):
        # ``optimizer.py:252-305`` without the ChainerX branches
        is_initialized = param.data is not None
        loss_scale = getattr(param, '_loss_scale', None)
        fp32_converted = False
        param_ = param
        if self._use_fp32_update and is_initialized and param.dtype == np.float16:
            # fp32 master weights (``optimizer.py:262-282``): the update runs on a
            # float32 copy of the parameter with the gradient up-cast to float32
            from chainer_b200.core.optimizers import _single
            fp32_param = self.fp32_param_for(param)
            if param.grad is not None:
                g32 = _single.new_like(param.grad, np.float32)
                _single.cast_copy(g32, param.grad)
                fp32_param.grad = g32
            else:
                fp32_param.grad = None
            param_ = fp32_param
            fp32_converted = True
        if is_initialized:
            self._init_states(param_)
            if loss_scale is not None and param_.grad is not None:
                from chainer_b200 import _lib
                g = param_.grad
                _lib.get().gp_divide(_dev.device_ptr(g), _dev.dtype_id(_dev.array_dtype(g)),
                                     _dev.array_size(g), float(loss_scale), 0)
        self._hookable.call_hooks('pre', (self, param_))
        self.update_core(param_)
        self._hookable.call_hooks('post', (self, param_))
        if fp32_converted:
            # ``optimizer.py:297-305``: back to the parameter's dtype (written in place:
            # same values as the reference's ``param.array = fp32.astype(float16)``)
            from chainer_b200.core.optimizers import _single
            _single.cast_copy(param.data, param_.data)
            param_.grad = None

    def fp32_param_for(self, param):
        """The float32 master copy of a float16 parameter, created on first use
        (``optimizer.py:262-273``)."""
        if self._fp32_param is None:
            from chainer_b200.core import link as _link
            from chainer_b200.core.optimizers import _single
            master = _single.new_like(param.data, np.float32)
            _single.cast_copy(master, param.data)
            self._fp32_param = _link.Parameter(master, name=param.name)
        return self._fp32_param

    def update_core(self, param):
        self.update_core_gpu(param)

    def update_core_gpu(self, param):
        raise NotImplementedError

    def init_state(self, param):
        pass

    def _init_states(self, param):
        """``optimizer.py:473-484``: lazy ``init_state`` on first use."""
        if self._state is None:
            self._state = {}
            self.init_state(param)

    def use_fp32_update(self, flag=True):
        self._use_fp32_update = flag

    def serialize(self, serializer):
        """``t`` and the state arrays by name (``optimizer.py:433-471``).

        Loading into a FRESH rule (no state yet -- the normal resume flow: new optimizer,
        ``setup()``, then load) restores the state arrays too: the keys come from
        ``state_names`` (or from ``init_state`` on a one-element dummy, as the reference
        does), each is read from the snapshot, and a snapshot that lacks them leaves the
        state ``None`` so that ``_init_states`` creates it on first use -- silently for a
        disabled rule, with the ``KeyError`` for an enabled one."""
        self.t = serializer('t', self.t)
        if self._state is not None:
            for key in self._state:
                self._state[key] = serializer(key, self._state[key])
        elif _is_deserializer(serializer):
            loaded = {}
            for key in self._state_keys():
                try:
                    value = serializer(key, None)
                except KeyError:
                    if self.enabled:
                        raise
                    value = None
                if value is None:
                    loaded = None
                    break
                loaded[key] = value
            self._state = loaded
        _rules_version[0] += 1      # state arrays may have been replaced

    def _state_keys(self):
        names = getattr(self, 'state_names', None)
        if names is not None:
            return tuple(names)
        probe = copy.copy(self)
        probe._state = {}
        from chainer_b200.core import link as _link
        probe.init_state(_link.Parameter(np.empty(1, dtype=np.float32)))
        return tuple(probe._state)


def _is_deserializer(serializer):
    """Serializers are the caller's objects (``chainer.serializers`` in the reference):
    recognised by ``chainer.Deserializer`` when chainer is imported, else by an
    ``is_deserializer`` attribute or a class name ending in ``Deserializer``."""
    import sys
    ch = sys.modules.get('chainer')
    base = getattr(getattr(ch, 'serializer', None), 'Deserializer', None) if ch else None
    if base is not None and isinstance(serializer, base):
        return True
    if getattr(serializer, 'is_deserializer', False):
        return True
    return type(serializer).__name__.endswith('Deserializer')


class Optimizer(object):
    """Base class of all numerical optimizers (``optimizer.py:533-791``)."""

    _b200_versioned = True
    target = _Tracked('target', None)
    t = 0
    epoch = 0
    _loss_scale = _Tracked('loss_scale', None)
    _loss_scale_max = 65504
    _loss_scaling_is_dynamic = _Tracked('loss_scaling_is_dynamic', False)
    use_auto_new_epoch = False

    def __init__(self):
        self._hookable = _Hookable()

    def setup(self, link):
        if not hasattr(link, 'namedparams'):
            raise TypeError('optimization target must be a link')
        self.target = link
        self.t = 0
        self.epoch = 0
        self._hookable = _Hookable()
        return self

    def update(self, lossfun=None, *args, **kwds):
        raise NotImplementedError

    def new_epoch(self, auto=False):
        self.epoch += 1

    def add_hook(self, hook, name=None, timing='auto'):
        if self.target is None:
            raise RuntimeError('call `setup` method before `add_hook` method')
        self._hookable.add_hook(hook, name, timing)

    def remove_hook(self, name):
        self._hookable.remove_hook(name)

    def call_hooks(self, timing='pre'):
        for hook in list((self._hookable._pre if timing == 'pre'
                          else self._hookable._post).values()):
            self.call_hook(hook)

    def call_hook(self, hook):
        # ``optimizer.py:706-711``
        if getattr(hook, 'call_for_each_param', False):
            for param in self.target.params():
                hook(param.update_rule, param)
        else:
            hook(self)

    def loss_scaling(self, interval=1000, scale=None):
        """``optimizer.py:736-761``."""
        if scale is None:
            self._loss_scaling_is_dynamic = True
            if interval < 1:
                raise ValueError('interval must be greater than or equal to 1.'
                                 ' Actual: {}'.format(interval))
            self._loss_scale = 1.0
            self._loss_scaling_multiplier = math.pow(2.0, 1.0 / interval)
            self._loss_scaling_isnan_ever = False
        else:
            if scale <= 0:
                raise ValueError('loss_scale must be a positive number. '
                                 'Actual: {}'.format(scale))
            self._loss_scale = scale

    def set_loss_scale(self, loss_scale):
        self.loss_scaling(scale=loss_scale)

    def check_nan_in_grads(self):
        """``optimizer.py:763-776``: with dynamic loss scaling, look for non-finite
        gradients.  One finiteness kernel per gradient writes its own flag word; ONE
        small read-back tells which parameters overflowed (the reference
        synchronises once per parameter)."""
        self._loss_scaling_isnan = False
        if not self._loss_scaling_is_dynamic:
            return
        from chainer_b200 import _lib
        lib = _lib.get()
        named = [(name, p) for name, p in self.target.namedparams() if p.grad is not None]
        if not named:
            return
        flags = _dev.DeviceArray.zeros((len(named),), np.int32)
        for i, (_, p) in enumerate(named):
            g = p.grad
            lib.gp_check_finite(_dev.device_ptr(g), _dev.dtype_id(_dev.array_dtype(g)),
                                _dev.array_size(g), flags.data.ptr + 4 * i, 0)
        host = flags.get()
        for (name, _), bad in zip(named, host):
            if bad:
                self._loss_scaling_isnan = True
                self._loss_scaling_isnan_ever = True
                warnings.warn(
                    'Non finite number found in param.grad of {}'
                    ' (iteration: {}, loss_scale: {})'
                    .format(name, self.t, self._loss_scale))

    def is_safe_to_update(self):
        return not getattr(self, '_loss_scaling_isnan', False)

    def update_loss_scale(self):
        """``optimizer.py:781-791``."""
        if not self._loss_scaling_is_dynamic:
            return
        if self._loss_scaling_isnan:
            multiplier = 0.5
        elif self._loss_scaling_isnan_ever:
            multiplier = self._loss_scaling_multiplier
        else:
            multiplier = 2.0
        self._loss_scale = max(1, min(self._loss_scale_max,
                                      self._loss_scale * multiplier))

    def serialize(self, serializer):
        self.t = serializer('t', self.t)
        self.epoch = serializer('epoch', self.epoch)
        for name, param in self.target.namedparams():
            rule = getattr(param, 'update_rule', None)
            if rule is not None:
                rule.serialize(serializer[name])


class GradientMethod(Optimizer):
    """``optimizer.py:794-894``."""

    def __init__(self):
        super(GradientMethod, self).__init__()
        self.hyperparam = Hyperparameter()
        self._use_fp32_update = False

    def setup(self, link):
        super(GradientMethod, self).setup(link)
        for param in link.params():
            param.update_rule = self.create_update_rule()
            if self._use_fp32_update:
                param.update_rule.use_fp32_update()
        return self

    def reallocate_cleared_grads(self):
        """``optimizer.py:834-851``: zeros for every gradient that is None."""
        for name, param in self.target.namedparams(False):
            if param.grad is None:
                param.grad = _dev.zeros_like(param.data)

    def call_hook(self, hook):
        super(GradientMethod, self).call_hook(hook)
        self.reallocate_cleared_grads()

    def update(self, lossfun=None, *args, **kwds):
        """``optimizer.py:857-894``."""
        if lossfun is not None:
            use_cleargrads = getattr(self, '_use_cleargrads', True)
            loss = lossfun(*args, **kwds)
            if use_cleargrads:
                self.target.cleargrads()
            else:
                self.target.zerograds()
            loss.backward(loss_scale=self._loss_scale)
            del loss

        self.reallocate_cleared_grads()
        self.check_nan_in_grads()
        self.call_hooks('pre')

        self.t += 1
        if self.is_safe_to_update():
            if not self._multi_tensor_update():
                for param in self.target.params():
                    param.update()

        self.reallocate_cleared_grads()
        self.call_hooks('post')
        self.update_loss_scale()

    def _multi_tensor_update(self):
        """All parameter updates of this step as ONE launch per (dtype, hyperparameter)
        group instead of one kernel per parameter, when every rule is a stock
        MomentumSGD / Adam rule without hooks, fp32-update or loss scaling.
        Returns False (nothing done) otherwise."""
        from chainer_b200.core import _multi_tensor
        return _multi_tensor.update(self)

    def use_cleargrads(self, use=True):
        warnings.warn('GradientMethod.use_cleargrads is deprecated.', DeprecationWarning)
        self._use_cleargrads = use

    def use_fp32_update(self, flag=True):
        self._use_fp32_update = flag
        link = getattr(self, 'target', None)
        if link is not None:
            for param in link.params():
                param.update_rule.use_fp32_update()

    def create_update_rule(self):
        raise NotImplementedError
