"""Multi-node optimizer wrappers: mirror of ``chainermn/optimizers.py:5-182``.

``_MultiNodeOptimizer.update`` keeps the reference protocol -- forward /
backward, then ``bcast_data`` on the first call or after the parameter set
changed (and NO update on that call), else mean-gradient + update -- but asks
the communicator for the fused pipeline first
(``PureNcclCommunicator.multi_node_mean_grad_and_update``); when that declines
(hooks, loss scaling, custom rules ...) the two reference steps run unfused.
"""
import copy

from chainer_b200 import device as _dev


class _MultiNodeOptimizer(object):

    def __init__(self, actual_optimizer, communicator, zero_fill):
        super(_MultiNodeOptimizer, self).__setattr__(
            'communicator', communicator)
        super(_MultiNodeOptimizer, self).__setattr__(
            'actual_optimizer', actual_optimizer)
        super(_MultiNodeOptimizer, self).__setattr__(
            'target_params', [])
        super(_MultiNodeOptimizer, self).__setattr__(
            'zero_fill', zero_fill)

    def update(self, lossfun=None, *args, **kwds):
        target = self.target
        if lossfun is not None:
            use_cleargrads = getattr(self, '_use_cleargrads', True)
            loss = lossfun(*args, **kwds)
            if use_cleargrads:
                target.cleargrads()
            else:
                target.zerograds()
            loss.backward(loss_scale=self.actual_optimizer._loss_scale)
            del loss

        if self.is_changed(target):
            self.communicator.bcast_data(target)
        else:
            fused = getattr(self.communicator, 'multi_node_mean_grad_and_update', None)
            if fused is not None and not args and not kwds and \
                    fused(target, self.actual_optimizer, self.zero_fill):
                return
            self.communicator.multi_node_mean_grad(target, self.zero_fill)
            self.actual_optimizer.update(None, *args, **kwds)

    def is_changed(self, target):
        # Fast path: Link classes that count structural changes
        # (chainer_b200.core.link) need not be re-walked every step.
        if getattr(target, '_b200_versioned', False):
            from chainer_b200.core import link as _link
            stamp = (id(target), _link.structure_version())
            if self.__dict__.get('_stamp') == stamp:
                return False
            super(_MultiNodeOptimizer, self).__setattr__('_stamp', stamp)
        previous_params = self.target_params
        super(_MultiNodeOptimizer, self).__setattr__(
            'target_params', [(name, param.data is not None)
                              for name, param in sorted(target.namedparams())])
        if len(previous_params) != len(self.target_params):
            return True

        for param1, param2 in zip(self.target_params, previous_params):
            if (param1[0] != param2[0]) or param1[1] != param2[1]:
                return True
        return False

    def setup(self, link):
        self.actual_optimizer.setup(link)
        return self

    def __getattr__(self, attr_name):
        return getattr(self.actual_optimizer, attr_name)

    def __setattr__(self, attr_name, value):
        setattr(self.actual_optimizer, attr_name, value)


class _DoubleBufferingOptimizer(object):
    """1-step-stale overlap (``chainermn/optimizers.py:59-146``): gradients are
    swapped into a deep copy of the model (``communicated_target``), averaged on
    a non-blocking side stream while the next forward/backward runs, and applied
    one call late."""

    def __init__(self, actual_optimizer, communicator, zero_fill):
        super(_DoubleBufferingOptimizer, self).__setattr__(
            'communicator', communicator)
        super(_DoubleBufferingOptimizer, self).__setattr__(
            'actual_optimizer', actual_optimizer)
        super(_DoubleBufferingOptimizer, self).__setattr__(
            'needs_update', False)
        super(_DoubleBufferingOptimizer, self).__setattr__(
            'communicated_target', None)
        super(_DoubleBufferingOptimizer, self).__setattr__(
            'target_params_list', [[], []])
        super(_DoubleBufferingOptimizer, self).__setattr__(
            'allreduce_grad_stream', _dev.Stream(non_blocking=True))
        super(_DoubleBufferingOptimizer, self).__setattr__(
            'zero_fill', zero_fill)

    def update(self, lossfun=None, *args, **kwds):
        target = self.target
        if lossfun is not None:
            use_cleargrads = getattr(self, '_use_cleargrads', True)
            loss = lossfun(*args, **kwds)
            if use_cleargrads:
                target.cleargrads()
            else:
                target.zerograds()
            loss.backward(loss_scale=self.actual_optimizer._loss_scale)
            del loss

        if self.is_changed(target, self.target_params_list[0]):
            self.wait()
            self.communicator.bcast_data(target)
            super(_DoubleBufferingOptimizer, self).__setattr__(
                'communicated_target', copy.deepcopy(target))
            super(_DoubleBufferingOptimizer, self).__setattr__(
                'target_params_list', [
                    list(sorted(self.target.namedparams())),
                    list(sorted(self.communicated_target.namedparams()))])
            super(_DoubleBufferingOptimizer, self).__setattr__(
                'needs_update', False)
        else:
            self.wait()
            self.swap_grad(self.target_params_list[0],
                           self.target_params_list[1])
            self.multi_node_mean_grad_async()
            if self.needs_update:
                self.actual_optimizer.update(None, *args, **kwds)
            else:
                super(_DoubleBufferingOptimizer, self).__setattr__(
                    'needs_update', True)

    def multi_node_mean_grad_async(self):
        self.communicator._multi_node_mean_grad_async(
            self.communicated_target, self.zero_fill,
            self.allreduce_grad_stream)

    def is_changed(self, target, previous_params):
        target_params = list(sorted(target.namedparams()))
        if len(previous_params) != len(target_params):
            return True

        for param1, param2 in zip(target_params, previous_params):
            name1, var1 = param1
            name2, var2 = param2
            if (name1 != name2) or (var1.data is None) != (var2.data is None):
                return True
        return False

    def swap_grad(self, target1_params, target2_params):
        for param1, param2 in zip(target1_params, target2_params):
            _, var1 = param1
            _, var2 = param2
            var1.grad, var2.grad = var2.grad, var1.grad

    def wait(self):
        self.allreduce_grad_stream.synchronize()
        _dev.Stream.null.synchronize()

    def setup(self, link):
        self.actual_optimizer.setup(link)
        return self

    def __getattr__(self, attr_name):
        return getattr(self.actual_optimizer, attr_name)

    def __setattr__(self, attr_name, value):
        setattr(self.actual_optimizer, attr_name, value)


def create_multi_node_optimizer(actual_optimizer, communicator,
                                double_buffering=False, zero_fill=True):
    """Create a multi node optimizer from a Chainer optimizer
    (``chainermn/optimizers.py:149-182``; same arguments)."""
    if double_buffering:
        from chainer_b200.communicators.pure_nccl_communicator \
            import PureNcclCommunicator
        if not isinstance(communicator, PureNcclCommunicator):
            raise ValueError(
                'This communicator does not support double buffering.')
        return _DoubleBufferingOptimizer(actual_optimizer, communicator,
                                         zero_fill)
    return _MultiNodeOptimizer(actual_optimizer, communicator,
                               zero_fill)
