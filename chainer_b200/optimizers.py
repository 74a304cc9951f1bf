"""Multi-node optimizer wrappers behind ``create_multi_node_optimizer``.

Behavioural mirror of ``chainermn/optimizers.py:5-182`` (same public names,
same protocol, same attribute forwarding), organised around one proxy base:

* ``update(lossfun=None, ...)`` runs forward / backward when a loss function is
  given, then EITHER broadcasts the parameters (first call, or the parameter set
  changed -- and performs NO update on that call, ``:29-30``) OR averages the
  gradients over the workers and updates (``:31-33``).
* The plain wrapper first offers the step to the communicator's fused pipeline
  (``PureNcclCommunicator.multi_node_mean_grad_and_update``: pack -> allreduce ->
  fused unpack + update); when that declines (custom hooks or rules, dynamic loss
  scaling ...) the two reference steps run one after the other.
* The double-buffering wrapper (``:59-146``) averages the gradients of step k on a
  side stream while step k + 1 computes, and applies them one call late.

Every attribute that is not the wrapper's own is read from / written to the
wrapped optimizer (``:52-56``), so ``opt.lr = ...``, ``opt.add_hook(...)``,
``opt.t`` behave as on the optimizer itself.
"""
import copy

from chainer_b200 import device as _dev


def _layout_signature(link):
    """What `is_changed` compares: the sorted parameter names and whether each
    parameter is initialised (``chainermn/optimizers.py:35-46``)."""
    return [(name, param.data is not None) for name, param in sorted(link.namedparams())]


class _OptimizerProxy(object):
    """Holds the wrapper's own fields in ``__dict__`` and forwards every other
    attribute access to ``actual_optimizer``."""

    def __init__(self, actual_optimizer, communicator, zero_fill, **own):
        fields = dict(actual_optimizer=actual_optimizer, communicator=communicator,
                      zero_fill=zero_fill)
        fields.update(own)
        for key, value in fields.items():
            self._own(key, value)

    def _own(self, name, value):
        object.__setattr__(self, name, value)

    def __getattr__(self, name):
        # only reached for names that are not the wrapper's own
        if name == 'actual_optimizer':          # half-built instance (copy / pickle)
            raise AttributeError(name)
        return getattr(self.actual_optimizer, name)

    def __setattr__(self, name, value):
        setattr(self.actual_optimizer, name, value)

    def setup(self, link):
        self.actual_optimizer.setup(link)
        return self

    def _compute_gradients(self, lossfun, args, kwds):
        """The forward / backward half of ``Optimizer.update(lossfun, ...)``."""
        if lossfun is None:
            return
        target = self.target
        loss = lossfun(*args, **kwds)
        if getattr(self, '_use_cleargrads', True):
            target.cleargrads()
        else:
            target.zerograds()
        loss.backward(loss_scale=self.actual_optimizer._loss_scale)


class _MultiNodeOptimizer(_OptimizerProxy):

    def __init__(self, actual_optimizer, communicator, zero_fill):
        super(_MultiNodeOptimizer, self).__init__(actual_optimizer, communicator, zero_fill,
                                                  target_params=[], _stamp=None)

    def update(self, lossfun=None, *args, **kwds):
        self._compute_gradients(lossfun, args, kwds)
        target = self.target
        comm = self.communicator
        if self.is_changed(target):
            comm.bcast_data(target)
            return
        if not args and not kwds:
            fused = getattr(comm, 'multi_node_mean_grad_and_update', None)
            if fused is not None and fused(target, self.actual_optimizer, self.zero_fill):
                return
        comm.multi_node_mean_grad(target, self.zero_fill)
        self.actual_optimizer.update(None, *args, **kwds)

    def is_changed(self, target):
        """True when the set of (name, initialised?) pairs differs from the one seen
        at the previous call; remembers the new one."""
        if getattr(target, '_b200_versioned', False):
            # Link classes of this package count structural changes: an unchanged
            # counter means an unchanged layout, without walking the model
            from chainer_b200.core import link as _link
            stamp = (id(target), _link.structure_version())
            if stamp == self._stamp:
                return False
            self._own('_stamp', stamp)
        seen = self.target_params
        now = _layout_signature(target)
        self._own('target_params', now)
        return now != seen


class _DoubleBufferingOptimizer(_OptimizerProxy):
    """One-step-stale overlap: the gradients of the model are swapped into a deep
    copy (``communicated_target``), averaged there on a non-blocking side stream
    while the next forward / backward runs, and applied at the next call."""

    def __init__(self, actual_optimizer, communicator, zero_fill):
        super(_DoubleBufferingOptimizer, self).__init__(
            actual_optimizer, communicator, zero_fill,
            needs_update=False, communicated_target=None, target_params_list=[[], []],
            allreduce_grad_stream=_dev.Stream(non_blocking=True))

    def update(self, lossfun=None, *args, **kwds):
        self._compute_gradients(lossfun, args, kwds)
        target = self.target
        mine, shadow = self.target_params_list
        changed = self.is_changed(target, mine)
        self.wait()
        if changed:
            self.communicator.bcast_data(target)
            twin = copy.deepcopy(target)
            self._own('communicated_target', twin)
            self._own('target_params_list', [sorted(target.namedparams()),
                                             sorted(twin.namedparams())])
            self._own('needs_update', False)
            return
        self.swap_grad(mine, shadow)
        self.multi_node_mean_grad_async()
        if self.needs_update:
            # the gradients averaged during the PREVIOUS call are in the model now
            self.actual_optimizer.update(None, *args, **kwds)
        else:
            self._own('needs_update', True)

    def multi_node_mean_grad_async(self):
        self.communicator._multi_node_mean_grad_async(
            self.communicated_target, self.zero_fill, self.allreduce_grad_stream)

    def is_changed(self, target, previous_params):
        before = [(name, p.data is not None) for name, p in previous_params]
        return _layout_signature(target) != before

    def swap_grad(self, target1_params, target2_params):
        for (_, a), (_, b) in zip(target1_params, target2_params):
            a.grad, b.grad = b.grad, a.grad

    def wait(self):
        self.allreduce_grad_stream.synchronize()
        _dev.Stream.null.synchronize()


def create_multi_node_optimizer(actual_optimizer, communicator,
                                double_buffering=False, zero_fill=True):
    """``chainermn.create_multi_node_optimizer`` (``chainermn/optimizers.py:149-182``):
    wraps ``actual_optimizer`` so that ``update()`` averages gradients over the
    workers of ``communicator`` first.  ``double_buffering`` needs the ``pure_nccl``
    communicator; ``zero_fill`` treats missing gradients as zeros in the mean."""
    if not double_buffering:
        return _MultiNodeOptimizer(actual_optimizer, communicator, zero_fill)
    from chainer_b200.communicators.pure_nccl_communicator import PureNcclCommunicator
    if not isinstance(communicator, PureNcclCommunicator):
        raise ValueError('This communicator does not support double buffering.')
    return _DoubleBufferingOptimizer(actual_optimizer, communicator, zero_fill)
