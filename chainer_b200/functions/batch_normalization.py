"""MultiNodeBatchNormalization statistics: mirror of
``chainermn/functions/batch_normalization.py:35-127``.

``_NcclImpl`` keeps the two hook signatures that
``chainer.functions.normalization.batch_normalization.GeneralBatchNormalizationImpl``
calls (``get_mean_and_var(axis, gamma, x, xp, interm_dtype)``,
``get_ggamma_and_gbeta(axis, gamma, gy, x_hat, xp)``), so it can be handed to
the reference's ``BatchNormalization(impl_selector=...)`` unchanged; everything
else of BN (y, gx, running statistics) stays on the caller's path.

Per call the reference launches 2 reductions + 1 elementwise (materialising
``square(x)`` / ``gy * x_hat``), 2 fresh device allocations, an out-of-place
allreduce and ``div_by_size``.  Here: ONE statistics kernel that reads the
activation once and writes ``[mean | sqmean]`` (or ``[sum gy | sum gy*x_hat]``)
straight into the message buffer, an in-place allreduce of those 2C values,
and one tiny kernel that applies 1/size and forms ``var = sqmean - mean**2``.
"""
import numpy as np

from chainer_b200 import _lib
from chainer_b200 import device as _dev
from chainer_b200 import nccl
from chainer_b200.communicators import _communication_utility


def _new_like(like, n, dtype):
    """Uninitialised 1-d array of `n` elements in the module of `like`."""
    if _dev.is_torch(like):
        import torch
        td = {np.dtype(np.float16): torch.float16, np.dtype(np.float32): torch.float32,
              np.dtype(np.float64): torch.float64}[np.dtype(dtype)]
        return torch.empty(n, dtype=td, device=like.device)
    if isinstance(like, np.ndarray):
        return np.empty(n, dtype=dtype)
    return _dev.DeviceArray.empty((n,), dtype)


def _halves(buf, C):
    if isinstance(buf, _dev.DeviceArray):
        return buf.view1d(0, C), buf.view1d(C, C)
    return buf[:C], buf[C:]


def _check_layout(axis, x, C):
    shape = _dev.array_shape(x)
    want = (0,) + tuple(range(2, len(shape)))
    if axis is not None and tuple(axis) != want:
        raise NotImplementedError(
            'MultiNodeBatchNormalization statistics are implemented for the default '
            'channel axis 1 (aggregation axes {}), got {}'.format(want, tuple(axis)))
    if len(shape) < 2 or shape[1] != C:
        raise ValueError('x.shape[1] must equal gamma.size')
    if _dev.is_torch(x) and not x.is_contiguous():
        raise ValueError('x must be C-contiguous')
    hw = 1
    for s in shape[2:]:
        hw *= s
    return shape[0], hw


class _Workspace(object):
    """Per-communicator scratch of the statistics kernels, shared by every BN layer of the
    model whatever its width (zero-initialised once; the kernels leave its header zeroed,
    see ``gp_bn_workspace_bytes`` in include/gradpath.h).  Sized for 4096 channels up front
    (about 4 MB) so that its address does not change under a captured CUDA graph; a wider
    layer still grows it (outside a capture)."""

    MIN_CHANNELS = 4096

    def __init__(self):
        self.mem = None
        self.C = 0

    def get(self, C):
        if self.mem is None or C > self.C:
            lib = _lib.get()
            C = max(int(C), self.MIN_CHANNELS)
            nbytes = lib.gp_bn_workspace_bytes(C)
            self.mem = _dev._Allocation(nbytes)
            lib.gp_memset_async(self.mem.ptr, 0, nbytes, 0)
            lib.gp_stream_synchronize(0)      # zeroed before any stream uses it
            self.C = C
        return self.mem.ptr


def _workspace(comm, C):
    ws = getattr(comm, '_bn_workspace', None)
    if ws is None:
        ws = comm._bn_workspace = _Workspace()
    return ws.get(C)


def _allreduce_in_place(comm, buf, n_elems, dtype, stream_ptr=0):
    comm._init_comms()
    if comm.size > 1:
        type_id = _communication_utility._get_nccl_type_id(dtype)
        ptr = _dev.device_ptr(buf)
        comm.nccl_comm.allReduce(ptr, ptr, n_elems, type_id, nccl.NCCL_SUM, stream_ptr)


def _small_p2p(comm, gdt, n_elems):
    """The peer-memory one-shot allreduce, if this message qualifies."""
    comm._init_comms()
    p2p = getattr(comm, '_p2p', None)
    if p2p is None or isinstance(gdt, str) or np.dtype(gdt) != np.float32:
        return None
    return p2p if n_elems <= p2p.small_cap else None


class _NcclImpl(object):
    """``chainermn/functions/batch_normalization.py:35-93``."""

    def __init__(self, comm, stream=0):
        self.comm = comm
        #: raw cudaStream_t every launch of this object goes to (0: the legacy default
        #: stream, what ``chainer.cuda.Stream.null`` is; the torch-based link passes
        #: torch's current stream, which also makes the calls capturable in a CUDA graph)
        self.stream = stream

    def get_mean_and_var(self, axis, gamma, x, xp=None, interm_dtype=None):
        lib = _lib.get()
        st = getattr(self, 'stream', 0)
        C = _dev.array_size(gamma)
        N, HW = _check_layout(axis, x, C)
        gdt = _dev.array_dtype(gamma)
        buf = _new_like(gamma, 2 * C, gdt)
        if self.comm.size == 1:
            # one rank: mean and var straight from the statistics kernel (one launch)
            lib.gp_bn_fwd_mean_var(_dev.device_ptr(x), _dev.array_dtype_id(x), N, C,
                                   HW, _dev.device_ptr(buf), _dev.dtype_id(gdt),
                                   _workspace(self.comm, C), st)
            return _halves(buf, C)
        p2p = _small_p2p(self.comm, gdt, 2 * C)
        if p2p is not None:
            # statistics + allReduce + div_by_size + var in ONE kernel: the CTA that finishes
            # the last channel exchanges the 2C values over NVLink peer memory
            with _lib.nvtx_range('mnbn.fwd_stats+allreduce'):
                lib.gp_bn_fwd_stats_allreduce(p2p.handle, _dev.device_ptr(x),
                                              _dev.array_dtype_id(x), N, C, HW,
                                              _dev.device_ptr(buf), _workspace(self.comm, C), st)
            return _halves(buf, C)
        lib.gp_bn_fwd_stats(_dev.device_ptr(x), _dev.array_dtype_id(x), N, C, HW,
                            _dev.device_ptr(buf), _dev.dtype_id(gdt), _workspace(self.comm, C), st)
        _allreduce_in_place(self.comm, buf, 2 * C, gdt, st)
        mean, var = _halves(buf, C)
        # buf *= 1/size; var = sqmean - mean**2 (written over sqmean)
        lib.gp_bn_finish_mean_var(_dev.device_ptr(buf), _dev.dtype_id(gdt), C,
                                  1.0 / self.comm.size, _dev.device_ptr(var), st)
        return mean, var

    def get_ggamma_and_gbeta(self, axis, gamma, gy, x_hat, xp=None):
        lib = _lib.get()
        st = getattr(self, 'stream', 0)
        C = _dev.array_size(gamma)
        N, HW = _check_layout(axis, gy, C)
        gdt = _dev.array_dtype(gamma)
        buf = _new_like(gamma, 2 * C, gdt)
        lib.gp_bn_bwd_stats(_dev.device_ptr(gy), _dev.array_dtype_id(gy),
                            _dev.device_ptr(x_hat), _dev.array_dtype_id(x_hat),
                            None, None, _lib.GP_F32, N, C, HW, _dev.device_ptr(buf),
                            _dev.dtype_id(gdt), _workspace(self.comm, C), st)
        buf = self._mean_over_ranks(buf, 2 * C, gdt, gamma)
        gbeta, ggamma = _halves(buf, C)
        return gbeta, ggamma

    def _mean_over_ranks(self, buf, n, gdt, like):
        if self.comm.size == 1:
            return buf                     # the mean over one rank
        st = getattr(self, 'stream', 0)
        p2p = _small_p2p(self.comm, gdt, n)
        if p2p is not None:
            out = _new_like(like, n, gdt)
            p2p.allreduce_small(_dev.device_ptr(buf), _dev.device_ptr(out), n, 0,
                                1.0 / self.comm.size, st)
            return out
        _allreduce_in_place(self.comm, buf, n, gdt, st)
        _lib.get().gp_scale(_dev.device_ptr(buf), _dev.dtype_id(gdt), n, 1.0 / self.comm.size, st)
        return buf

    def get_ggamma_and_gbeta_from_x(self, axis, gamma, gy, x, mean, inv_std):
        """Same statistics with ``x_hat = (x - mean) * inv_std`` formed on the
        fly (``_x_hat``), so the caller need not materialise x_hat."""
        lib = _lib.get()
        st = getattr(self, 'stream', 0)
        C = _dev.array_size(gamma)
        N, HW = _check_layout(axis, gy, C)
        gdt = _dev.array_dtype(gamma)
        buf = _new_like(gamma, 2 * C, gdt)
        p2p = _small_p2p(self.comm, gdt, 2 * C) if self.comm.size > 1 else None
        if p2p is not None:
            # statistics + allReduce + div_by_size in ONE kernel (see get_mean_and_var)
            lib.gp_bn_bwd_stats_allreduce(p2p.handle, _dev.device_ptr(gy),
                                          _dev.array_dtype_id(gy), _dev.device_ptr(x),
                                          _dev.array_dtype_id(x), _dev.device_ptr(mean),
                                          _dev.device_ptr(inv_std),
                                          _dev.array_dtype_id(mean), N, C, HW,
                                          _dev.device_ptr(buf), _workspace(self.comm, C), st)
            gbeta, ggamma = _halves(buf, C)
            return gbeta, ggamma
        lib.gp_bn_bwd_stats(_dev.device_ptr(gy), _dev.array_dtype_id(gy),
                            _dev.device_ptr(x), _dev.array_dtype_id(x),
                            _dev.device_ptr(mean), _dev.device_ptr(inv_std),
                            _dev.array_dtype_id(mean), N, C, HW,
                            _dev.device_ptr(buf), _dev.dtype_id(gdt),
                            _workspace(self.comm, C), st)
        buf = self._mean_over_ranks(buf, 2 * C, gdt, gamma)
        gbeta, ggamma = _halves(buf, C)
        return gbeta, ggamma


class _MpiImpl(object):
    """``chainermn/functions/batch_normalization.py:7-32``: the same statistics
    through ``comm._multi_node_mean`` (host-staged control plane).  Kept for API
    completeness; ``nccl`` is what ``auto`` selects for PureNcclCommunicator."""

    def __init__(self, comm):
        self.comm = comm
        self._local = _NcclImpl(_SizeOne(comm))

    def get_mean_and_var(self, axis, gamma, x, xp=None, interm_dtype=None):
        lib = _lib.get()
        C = _dev.array_size(gamma)
        N, HW = _check_layout(axis, x, C)
        gdt = _dev.array_dtype(gamma)
        tmp = _new_like(gamma, 2 * C, gdt)
        lib.gp_bn_fwd_stats(_dev.device_ptr(x), _dev.array_dtype_id(x), N, C, HW,
                            _dev.device_ptr(tmp), _dev.dtype_id(gdt), _workspace(self.comm, C), 0)
        _dev.Stream.null.synchronize()
        self.comm._multi_node_mean(None, tmp)
        mean, var = _halves(tmp, C)
        lib.gp_bn_finish_mean_var(_dev.device_ptr(tmp), _dev.dtype_id(gdt), C, 1.0,
                                  _dev.device_ptr(var), 0)
        return mean, var

    def get_ggamma_and_gbeta(self, axis, gamma, gy, x_hat, xp=None):
        gbeta, ggamma = self._local.get_ggamma_and_gbeta(axis, gamma, gy, x_hat)
        C = _dev.array_size(gamma)
        # gbeta and ggamma are the two halves of one buffer: average it whole
        _dev.Stream.null.synchronize()
        for part in (gbeta, ggamma):
            self.comm._multi_node_mean(None, part)
        return gbeta, ggamma


class _SizeOne(object):
    """View of a communicator as a world of one (local statistics only)."""

    size = 1

    def __init__(self, comm):
        self._comm = comm

    def _init_comms(self):
        pass

    def __getattr__(self, name):
        return getattr(self._comm, name)

    def __setattr__(self, name, value):
        if name == '_comm':
            object.__setattr__(self, name, value)
        else:
            setattr(self._comm, name, value)


def get_communication_backend(comm, communication_backend='auto'):
    """``chainermn/functions/batch_normalization.py:96-114``."""
    if communication_backend not in ['mpi', 'nccl', 'auto']:
        raise ValueError('MultiNodeBatchNormalization does not support '
                         '{}.'.format(communication_backend))
    from chainer_b200.communicators.pure_nccl_communicator \
        import PureNcclCommunicator
    if communication_backend != 'auto':
        if 'nccl' == communication_backend:
            if not isinstance(comm, PureNcclCommunicator):
                raise ValueError('{} is not supported in '
                                 'MultiNodeBatchNormalization when using '
                                 '{}.'.format(communication_backend,
                                              type(comm)))
        selected_communication_backend = communication_backend
    else:
        if isinstance(comm, PureNcclCommunicator):
            selected_communication_backend = 'nccl'
        else:
            selected_communication_backend = 'mpi'
    return selected_communication_backend


class MultiNodeBNImplSelector:
    """``chainermn/functions/batch_normalization.py:117-127``."""

    def __init__(self, comm, communication_backend_name):
        self.comm = comm
        self.communication_backend_name = communication_backend_name

    def __call__(self, batch_norm_func, inputs):
        if self.communication_backend_name == 'nccl':
            return _NcclImpl(self.comm)
        else:
            return _MpiImpl(self.comm)


def fwd_apply(x, mean, var, gamma, beta, eps, running_mean=None, running_var=None, decay=0.9,
              adjust=1.0, stream=0):
    """Everything of the BN forward that follows the statistics, ONE launch
    (``chainer/functions/normalization/batch_normalization.py:40-77``): ``inv_std =
    rsqrt(var + eps)``, ``y = gamma * (x - mean) * inv_std + beta`` and, when given, the
    in-place update of the running statistics.  Returns ``(y, inv_std)``."""
    lib = _lib.get()
    C = _dev.array_size(gamma)
    N, HW = _check_layout(None, x, C)
    sdt = _dev.array_dtype(gamma)
    y = _dev.empty_like(x)
    inv_std = _new_like(gamma, C, sdt)
    rdt = _dev.array_dtype(running_mean) if running_mean is not None else sdt
    lib.gp_bn_fwd_apply(_dev.device_ptr(x), _dev.array_dtype_id(x), N, C, HW,
                        _dev.device_ptr(mean), _dev.device_ptr(var), _dev.device_ptr(gamma),
                        _dev.device_ptr(beta), _dev.dtype_id(sdt), float(eps), _dev.device_ptr(y),
                        _dev.device_ptr(inv_std),
                        _dev.device_ptr(running_mean) if running_mean is not None else None,
                        _dev.device_ptr(running_var) if running_var is not None else None,
                        _dev.dtype_id(rdt), float(decay), float(adjust), stream)
    return y, inv_std


def bwd_apply(gy, x, mean, inv_std, gamma, ggamma, gbeta, stream=0):
    """``gx`` of the BN backward, ONE launch with ``x_hat`` formed on the fly
    (``chainer/functions/normalization/batch_normalization.py:105-133``)."""
    lib = _lib.get()
    C = _dev.array_size(gamma)
    N, HW = _check_layout(None, x, C)
    sdt = _dev.array_dtype(gamma)
    gx = _dev.empty_like(x)
    inv_m = float(np.dtype(sdt).type(1.0 / (N * HW))) if not isinstance(sdt, str) else 1.0 / (N * HW)
    lib.gp_bn_bwd_apply(_dev.device_ptr(gy), _dev.array_dtype_id(gy),
                        _dev.device_ptr(x), _dev.array_dtype_id(x), N, C, HW,
                        _dev.device_ptr(mean), _dev.device_ptr(inv_std), _dev.device_ptr(gamma),
                        _dev.device_ptr(ggamma), _dev.device_ptr(gbeta), _dev.dtype_id(sdt),
                        inv_m, _dev.device_ptr(gx), stream)
    return gx


def mean_and_var(comm, x, gamma):
    """Convenience: whole-batch (all ranks) per-channel mean and biased variance."""
    return _NcclImpl(comm).get_mean_and_var(None, gamma, x)
