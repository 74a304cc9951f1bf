"""MultiNodeBatchNormalization link: mirror of
``chainermn/links/batch_normalization.py:15-147``.

Constructor arguments, parameters (``gamma``, ``beta``), persistents
(``avg_mean``, ``avg_var`` -- initialised to ZEROS like the reference, ``:57-60``
--, ``N``), ``decay`` / ``eps`` / ``finetune`` semantics and backend validation
follow the reference.  The statistics (+ their exchange over the workers) and the
elementwise halves -- ``y = gamma * (x - mean) * inv_std + beta`` with ``inv_std`` and
the running-statistics update, and ``gx`` with ``x_hat`` formed on the fly
(``chainer/functions/normalization/batch_normalization.py:40-77, 105-133``) -- are
this package's kernels (``functions/batch_normalization``): 2 launches forward and 2
backward per layer; ``torch.autograd`` only carries the graph, since torch is the array
library of this image.
"""
import numpy as np

from chainer_b200 import config
from chainer_b200 import device as _dev
from chainer_b200.core import link
from chainer_b200.functions import batch_normalization as mnbn_functions


def _torch():
    import torch
    return torch


class _MNBNFunction(object):
    """Builds the torch.autograd.Function lazily (torch import on first use)."""

    _fn = None

    @classmethod
    def get(cls):
        if cls._fn is not None:
            return cls._fn
        torch = _torch()

        class MNBN(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, gamma, beta, impl, eps, running_mean, running_var, decay):
                axis = (0,) + tuple(range(2, x.dim()))
                x = x.contiguous()
                # launch on torch's current stream (capturable in a CUDA graph)
                st = torch.cuda.current_stream().cuda_stream if x.is_cuda else 0
                impl.stream = st
                mean, var = impl.get_mean_and_var(axis, gamma, x)
                # y, inv_std and the running statistics in ONE launch (gp_bn_fwd_apply); m is
                # the LOCAL element count per channel (chainer/functions/normalization/
                # batch_normalization.py:50-52)
                m = x.numel() // gamma.numel()
                adjust = m / max(m - 1., 1.)
                y, inv_std = mnbn_functions.fwd_apply(x, mean, var, gamma, beta, eps,
                                                      running_mean, running_var, decay, adjust, st)
                ctx.impl = impl
                ctx.save_for_backward(x, gamma, mean, inv_std)
                ctx.mark_non_differentiable(mean, var)
                return y, mean, var

            @staticmethod
            def backward(ctx, gy, _gm, _gv):
                x, gamma, mean, inv_std = ctx.saved_tensors
                axis = (0,) + tuple(range(2, x.dim()))
                gy = gy.contiguous()
                st = torch.cuda.current_stream().cuda_stream if x.is_cuda else 0
                ctx.impl.stream = st
                gbeta, ggamma = ctx.impl.get_ggamma_and_gbeta_from_x(axis, gamma, gy, x, mean,
                                                                     inv_std)
                # gx in ONE launch, x_hat formed on the fly (gp_bn_bwd_apply)
                gx = mnbn_functions.bwd_apply(gy, x, mean, inv_std, gamma, ggamma, gbeta, st)
                return gx, ggamma, gbeta, None, None, None, None, None

        cls._fn = MNBN
        return MNBN


class MultiNodeBatchNormalization(link.Link):

    def __init__(self, size, comm, decay=0.9, eps=2e-5, dtype=None,
                 use_gamma=True, use_beta=True,
                 initial_gamma=None, initial_beta=None,
                 communication_backend='auto', device='cuda'):
        super(MultiNodeBatchNormalization, self).__init__()
        backend = mnbn_functions.get_communication_backend(comm, communication_backend)
        self._setup(size, comm, backend, decay, eps, dtype, use_gamma, use_beta,
                    initial_gamma, initial_beta, device)

    def _setup(self, size, comm, backend, decay, eps, dtype, use_gamma, use_beta,
               initial_gamma, initial_beta, device):
        torch = _torch()
        if isinstance(size, (tuple, list)):
            size, = size
        self._highprec_dtype = config.get_dtype(dtype, map_mixed16=np.float32)
        tdt = {np.dtype(np.float16): torch.float16, np.dtype(np.float32): torch.float32,
               np.dtype(np.float64): torch.float64}[np.dtype(self._highprec_dtype)]
        self.comm = comm
        # persistents, as in the reference (links/batch_normalization.py:57-61)
        self.add_persistent('avg_mean', torch.zeros(size, dtype=tdt, device=device))
        self.add_persistent('avg_var', torch.zeros(size, dtype=tdt, device=device))
        self.add_persistent('N', 0)
        self.decay = decay
        self.eps = eps
        self._device = device
        self._tdt = tdt
        self._communication_backend = backend
        with self.init_scope():
            if use_gamma:
                g = torch.full((size,), 1.0 if initial_gamma is None else float(initial_gamma),
                               dtype=tdt, device=device)
                self.gamma = link.Parameter(g)
            if use_beta:
                b = torch.full((size,), 0.0 if initial_beta is None else float(initial_beta),
                               dtype=tdt, device=device)
                self.beta = link.Parameter(b)

    def __deepcopy__(self, memo):
        # a copy of the link shares the communicator (handles of live collectives
        # cannot be duplicated); parameters and persistents are copied
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for key, value in self.__dict__.items():
            new.__dict__[key] = value if key == 'comm' else copy.deepcopy(value, memo)
        return new

    def _impl(self):
        return mnbn_functions.MultiNodeBNImplSelector(
            self.comm, self._communication_backend)(None, None)

    def __call__(self, x, finetune=False, train=True):
        torch = _torch()
        size = self.avg_mean.shape[0]
        gamma = self.gamma.data if hasattr(self, 'gamma') else \
            torch.ones(size, dtype=self._tdt, device=self._device)
        beta = self.beta.data if hasattr(self, 'beta') else \
            torch.zeros(size, dtype=self._tdt, device=self._device)
        if train:
            if finetune:
                self.N += 1
                decay = 1. - 1. / self.N
            else:
                decay = self.decay
            y, mean, var = _MNBNFunction.get().apply(x, gamma, beta, self._impl(), self.eps,
                                                     self.avg_mean, self.avg_var, decay)
            return y
        # fixed statistics (evaluation): the same elementwise kernel, one launch
        with torch.no_grad():
            st = torch.cuda.current_stream().cuda_stream if x.is_cuda else 0
            y, _ = mnbn_functions.fwd_apply(x.contiguous(), self.avg_mean, self.avg_var, gamma,
                                            beta, self.eps, stream=st)
        return y

    def start_finetuning(self):
        self.N = 0


class _SingleWorker(object):
    """The communicator of a link that normalises over the local batch only (one
    shared instance: it only carries the statistics kernels' scratch)."""
    size = 1
    rank = 0

    def __deepcopy__(self, memo):
        return self

    def __copy__(self):
        return self


_LOCAL = _SingleWorker()


class BatchNormalization(MultiNodeBatchNormalization):
    """Single-worker batch normalisation on the same statistics kernels: the
    stand-in of ``chainer.links.BatchNormalization`` that ``create_mnbn_model``
    replaces (``chainermn/links/create_mnbn_model.py:27-41``).  Same parameters and
    persistents as the multi-node link; statistics over the local batch."""

    def __init__(self, size, decay=0.9, eps=2e-5, dtype=None, use_gamma=True, use_beta=True,
                 initial_gamma=None, initial_beta=None, device='cuda'):
        link.Link.__init__(self)
        self._setup(size, _LOCAL, 'nccl', decay, eps, dtype, use_gamma, use_beta,
                    initial_gamma, initial_beta, device)
