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
        return 'Hyperparameter(%s)' % ', '.join(
            '%s=%s' % item for item in sorted(self.get_dict().items()))

    parent = property(lambda self: self._parent)

    def get_dict(self):
        """Every entry visible from here: ancestors first, nearer ones override."""
        chain = []
        node = self
        while node is not None:
            chain.append(node)
            node = node._parent
        merged = {}
        for node in reversed(chain):
            merged.update((k, v) for k, v in node.__dict__.items() if k != '_parent')
        return merged


class HyperparameterProxy(property):
    """``optimizer.HyperparameterProxy``: ``obj.<name>`` reads and writes
    ``obj.hyperparam.<name>`` (a ``property`` whose accessors close over the name)."""

    def __init__(self, attr_name):
        super(HyperparameterProxy, self).__init__(
            lambda obj: getattr(obj.hyperparam, attr_name),
            lambda obj, value: setattr(obj.hyperparam, attr_name, value),
            doc='Alias to ``self.hyperparam.{}``'.format(attr_name))
        self._attr_name = attr_name


class _Hookable(object):
    def __init__(self):
        self._pre = collections.OrderedDict()
        self._post = collections.OrderedDict()

    def _registry(self, timing):
        return self._pre if timing == 'pre' else self._post

    def _holder_of(self, name):
        for reg in (self._pre, self._post):
            if name in reg:
                return reg
        return None

    def add_hook(self, hook, name=None, timing='auto'):
        """Same contract as ``Optimizer.add_hook`` / ``UpdateRule.add_hook``
        (``optimizer.py:188-224, 693-731``): callable, a known timing ('auto' takes the
        hook's own ``timing``, 'pre' if it has none), a name that is not taken."""
        if not callable(hook):
            raise TypeError('hook function must be callable')
        if timing == 'auto':
            timing = getattr(hook, 'timing', 'pre')
        elif timing not in ('pre', 'post'):
            raise ValueError("timing must be one of ('pre', 'post', 'auto')")
        if name is None:
            for attr in ('name', '__name__'):
                name = getattr(hook, attr, None)
                if name is not None:
                    break
            else:
                raise ValueError('the name of the hook function is not specified')
        if self._holder_of(name) is not None:
            raise KeyError('hook "{}" already exists'.format(name))
        self._registry(timing)[name] = hook
        _rules_version[0] += 1

    def remove_hook(self, name):
        reg = self._holder_of(name)
        if reg is None:
            raise KeyError('hook "{}" does not exist'.format(name))
        del reg[name]
        _rules_version[0] += 1

    def has_hooks(self):
        return bool(self._pre or self._post)

    def call_hooks(self, timing, args):
        # a snapshot: hooks may add or remove hooks
        for hook in tuple(self._registry(timing).values()):
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
        # ``optimizer.py:252-305`` without the ChainerX branches, in three stages:
        # choose the parameter the rule works on, run hooks + rule on it, write back
        work = self._working_param(param)
        if param.data is not None:
            self._init_states(work)
            self._undo_loss_scale(work, getattr(param, '_loss_scale', None))
        hooks = self._hookable
        hooks.call_hooks('pre', (self, work))
        self.update_core(work)
        hooks.call_hooks('post', (self, work))
        if work is not param:
            # ``optimizer.py:297-305``: back to the parameter's dtype (written in place:
            # same values as the reference's ``param.array = fp32.astype(float16)``)
            from chainer_b200.core.optimizers import _single
            _single.cast_copy(param.data, work.data)
            work.grad = None

    def _working_param(self, param):
        """`param` itself, or -- fp32 update of an initialised float16 parameter
        (``optimizer.py:262-282``) -- its float32 master copy carrying the gradient
        up-cast to float32."""
        if not self._use_fp32_update or param.data is None or param.dtype != np.float16:
            return param
        from chainer_b200.core.optimizers import _single
        master = self.fp32_param_for(param)
        grad = param.grad
        if grad is not None:
            wide = _single.new_like(grad, np.float32)
            _single.cast_copy(wide, grad)
            grad = wide
        master.grad = grad
        return master

    @staticmethod
    def _undo_loss_scale(work, loss_scale):
        """``grad /= loss_scale`` in place (``optimizer.py:286-291``)."""
        g = work.grad
        if loss_scale is None or g is None:
            return
        from chainer_b200 import _lib
        _lib.get().gp_divide(_dev.device_ptr(g), _dev.dtype_id(_dev.array_dtype(g)),
                             _dev.array_size(g), float(loss_scale), 0)

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
        """``optimizer.py:736-761``: a fixed `scale`, or (none given) dynamic scaling that
        starts at 1 and, once an overflow has been seen, doubles over `interval` steps."""
        if scale is not None:
            if not scale > 0:
                raise ValueError('loss_scale must be a positive number. '
                                 'Actual: {}'.format(scale))
            self._loss_scale = scale
            return
        if interval < 1:
            raise ValueError('interval must be greater than or equal to 1.'
                             ' Actual: {}'.format(interval))
        self._loss_scaling_is_dynamic = True
        self._loss_scaling_isnan_ever = False
        self._loss_scaling_multiplier = math.pow(2.0, 1.0 / interval)
        self._loss_scale = 1.0

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
        """``optimizer.py:781-791``: halve after an overflow; otherwise grow -- by 2 per
        step until the first overflow ever, by 2^(1/interval) after it; clamp to
        [1, 65504]."""
        if not self._loss_scaling_is_dynamic:
            return
        if self._loss_scaling_isnan:
            factor = 0.5
        else:
            factor = self._loss_scaling_multiplier if self._loss_scaling_isnan_ever else 2.0
        grown = self._loss_scale * factor
        self._loss_scale = max(1, grown if grown < self._loss_scale_max else self._loss_scale_max)

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
        fp32 = self._use_fp32_update
        for param in link.params():
            rule = param.update_rule = self.create_update_rule()
            if fp32:
                rule.use_fp32_update()
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
        """``optimizer.py:857-894``: (forward, clear, backward) when a loss function is
        given; then overflow check, pre hooks, the parameter updates unless a gradient
        overflowed, post hooks, and the loss-scale schedule."""
        if lossfun is not None:
            self._backward(lossfun, args, kwds)
        self.reallocate_cleared_grads()
        self.check_nan_in_grads()
        self.call_hooks('pre')
        self.t += 1
        if self.is_safe_to_update() and not self._multi_tensor_update():
            for param in self.target.params():
                param.update()
        self.reallocate_cleared_grads()
        self.call_hooks('post')
        self.update_loss_scale()

    def _backward(self, lossfun, args, kwds):
        target = self.target
        loss = lossfun(*args, **kwds)
        clear = target.cleargrads if getattr(self, '_use_cleargrads', True) else target.zerograds
        clear()
        loss.backward(loss_scale=self._loss_scale)

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
        """Switch on for this optimizer and for the rules it has already created
        (``optimizer.py:817-827``: the rules are switched ON whatever `flag` says)."""
        self._use_fp32_update = flag
        target = getattr(self, 'target', None)
        for param in (target.params() if target is not None else ()):
            param.update_rule.use_fp32_update()

    def create_update_rule(self):
        raise NotImplementedError
