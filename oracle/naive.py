"""CPU baseline: NumPy port of the reference's `naive` communicator step.

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py).  This is what the
reference executes on CPUs for the same path, restated without Chainer so that
it can run on the GPU box (where /root/reference does not exist):

  NaiveCommunicator.multi_node_mean_grad   chainermn/communicators/naive_communicator.py:10-17
    -> MpiCommunicatorBase._multi_node_mean  mpi_communicator_base.py:735-778
       per parameter: in-place Allreduce(SUM) (float16 up-cast to float32), then
       `recvbuf *= 1.0 / size`
  GradientMethod.update                      chainer/optimizer.py:857-894
    -> per parameter MomentumSGDRule.update_core_cpu  momentum_sgd.py:61-73
       or AdamRule.update_core_cpu                      adam.py:189-222

It is bit-exact with the unmodified reference for float32/float64
(tests/test_oracle_golden.py pins the update rules; the mean is pinned through
tests/golden/naive_mean_grad.npz).  `allreduce` is a callable
(array -> None, in place) standing in for MPI_Allreduce: identity for one rank,
a gloo all_reduce for several.
"""
import numpy as np

from oracle import gradpath as og


def multi_node_mean_grad(grads, size, allreduce=None):
    """In place on the list of gradient arrays of THIS rank."""
    for g in grads:
        is_float16 = g.dtype == np.float16
        work = g.astype(np.float32) if is_float16 else g
        if allreduce is not None:
            allreduce(work)
        if is_float16:
            g[...] = work.astype(np.float16)
        g *= 1.0 / size


class MomentumSGD(object):
    def __init__(self, params, lr=0.01, momentum=0.9):
        self.lr, self.momentum = lr, momentum
        self.v = [np.zeros_like(p) for p in params]
        self.t = 0

    def update(self, params, grads):
        self.t += 1
        for p, g, v in zip(params, grads, self.v):
            og.momentum_sgd_update(p, g, v, self.lr, self.momentum)


class Adam(object):
    def __init__(self, params, **hyper):
        self.hyper = hyper
        self.m = [np.zeros_like(p) for p in params]
        self.v = [np.zeros_like(p) for p in params]
        self.vhat = [np.zeros_like(p) for p in params] if hyper.get('amsgrad') else None
        self.t = 0

    def update(self, params, grads):
        self.t += 1
        for i, (p, g) in enumerate(zip(params, grads)):
            og.adam_update_cpu(p, g, self.m[i], self.v[i], self.t,
                               vhat=None if self.vhat is None else self.vhat[i], **self.hyper)


def step(params, grads, optimizer, size=1, allreduce=None):
    """One reference training-step tail: mean of gradients, then update."""
    multi_node_mean_grad(grads, size, allreduce)
    optimizer.update(params, grads)
