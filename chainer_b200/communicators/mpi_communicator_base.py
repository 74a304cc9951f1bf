"""MpiCommunicatorBase: mirror of
``chainermn/communicators/mpi_communicator_base.py:99-815``.

The control-plane methods (``*_obj``, ``split``, rank properties, the
``batched_copy`` config) behave as in the reference on any mpi4py-like
``mpi_comm`` (see ``_control_plane``).  The ndarray collectives
(``send/recv/bcast/gather/allgather/allreduce/scatter/alltoall``) serve the
reference's model-parallel functions, which are outside this path; they are
provided host-staged so that the class is complete, not as a fast path.
"""
import collections

import numpy as np

from chainer_b200 import config
from chainer_b200 import device as _dev
from chainer_b200.communicators import _communication_utility
from chainer_b200.communicators import _memory_utility
from chainer_b200.communicators import communicator_base


def _to_host(x):
    return np.ascontiguousarray(_dev.to_numpy(x))


def _like(x, host):
    """`host` (numpy) as an array of the module of `x`."""
    if isinstance(x, np.ndarray):
        return host
    if _dev.is_torch(x):
        import torch
        return torch.from_numpy(host).to(x.device)
    if isinstance(x, _dev.DeviceArray):
        return _dev.DeviceArray.from_numpy(host)
    return host


_Topology = collections.namedtuple(
    '_Topology', ['global_rank', 'intra_rank', 'intra_size', 'inter_rank', 'inter_size'])


def _from_topology(field):
    return property(lambda self: getattr(self._topology, field),
                    doc='``%s`` of this process (``init_ranks``)' % field)


class MpiCommunicatorBase(communicator_base.CommunicatorBase):

    #: configuration keys this class owns (name -> default); others go to the base class
    _OWN_CONFIG = {'batched_copy': False}

    def __init__(self, mpi_comm):
        self.mpi_comm = mpi_comm
        self._init_ranks()
        with self.config_scope():
            for key, default in self._OWN_CONFIG.items():
                setattr(self, key, default)

    rank = property(lambda self: self.mpi_comm.rank)
    size = property(lambda self: self.mpi_comm.size)
    intra_rank = _from_topology('intra_rank')
    intra_size = _from_topology('intra_size')
    inter_rank = _from_topology('inter_rank')
    inter_size = _from_topology('inter_size')

    def set_config(self, name, value=True, **kwargs):
        if name not in self._OWN_CONFIG:
            return super(MpiCommunicatorBase, self).set_config(name, **kwargs)
        with self.config_scope():
            setattr(self, name, value)

    def get_config(self, name=None):
        if name in self._OWN_CONFIG:
            return getattr(self, name)
        return super(MpiCommunicatorBase, self).get_config(name)

    def split(self, color, key):
        return type(self)(mpi_comm=self.mpi_comm.Split(color, key))

    # -- ndarray collectives (host staged; outside the gradient path) ---------
    def alltoall(self, xs):
        if len(xs) != self.size:
            raise ValueError('The length of data must be same as communicator size.')
        hosts = [_to_host(x) for x in xs]
        all_parts = self.mpi_comm.allgather(hosts)
        return tuple(_like(xs[0], all_parts[src][self.rank]) for src in range(self.size))

    def send(self, data, dest, tag):
        self.mpi_comm.send(_to_host(data), dest=dest, tag=tag)

    def recv(self, source, tag):
        return self.mpi_comm.recv(source=source, tag=tag)

    def bcast(self, x, root=0):
        host = self.mpi_comm.bcast(_to_host(x) if self.rank == root else None, root)
        return _like(x, host) if x is not None else host

    def gather(self, x, root=0):
        parts = self.mpi_comm.gather(_to_host(x), root)
        if self.rank == root:
            return tuple(_like(x, p) for p in parts)
        return None

    def allgather(self, x):
        parts = self.mpi_comm.allgather(_to_host(x))
        return tuple(_like(x, p) for p in parts)

    def allreduce(self, x):
        parts = self.mpi_comm.allgather(_to_host(x))
        acc = parts[0].copy()
        for p in parts[1:]:
            acc += p
        return _like(x, acc)

    def scatter(self, xs, root=0):
        if self.rank == root:
            hosts = [_to_host(x) for x in xs]
        else:
            hosts = None
        return self.mpi_comm.scatter(hosts, root)

    # -- objects ---------------------------------------------------------------
    def send_obj(self, obj, dest, tag=0):
        self.mpi_comm.send(obj, dest=dest, tag=tag)

    def recv_obj(self, source, status=None, tag=0):
        return self.mpi_comm.recv(source=source, tag=tag)

    def bcast_obj(self, obj, max_buf_len=256 * 1024 * 1024, root=0):
        return self.mpi_comm.bcast(obj, root)

    def gather_obj(self, obj, root=0):
        return self.mpi_comm.gather(obj, root=root)

    def allreduce_obj(self, obj):
        return self.mpi_comm.allreduce(obj)

    def bcast_data(self, model):
        """``mpi_communicator_base.py:694-707`` (host-staged here)."""
        for _, param in sorted(model.namedparams()):
            if param.data is not None:
                host = self.mpi_comm.bcast(_to_host(param.data) if self.rank == 0 else None, 0)
                if self.rank != 0:
                    _copy_into(param.data, host)

    # -- private ---------------------------------------------------------------
    def _init_ranks(self):
        topo = _Topology(*_communication_utility.init_ranks(self.mpi_comm)[:5])
        assert topo.global_rank == self.mpi_comm.rank
        self._topology = topo
        # the reference's attribute names, for code that reads them directly
        self._intra_rank, self._intra_size = topo.intra_rank, topo.intra_size
        self._inter_rank, self._inter_size = topo.inter_rank, topo.inter_size

    def _check_ready_to_allreduce(self, array_a, array_b):
        """Debug mode (``mpi_communicator_base.py:709-728``): rank 0 compares every
        rank's (shape, dtype) pair of the two buffers with its own."""
        def describe(a):
            return (tuple(a.shape), str(_dev.array_dtype(a)))
        mine = ((None, None) if array_a is None else describe(array_a),) + describe(array_b)
        gathered = self.gather_obj((self.rank, mine))
        if self.rank != 0:
            return
        for rank, theirs in gathered:
            if theirs != mine:
                raise ValueError('Shape does not match: {} at rank 0 while {} at rank {}'
                                 .format(mine, theirs, rank))

    def _ensure_all_finite(self, array):
        if not np.isfinite(_to_host(array)).all():
            raise ValueError('Parameters diverged after allreduce.')

    def _multi_node_mean(self, sendbuf, recvbuf):
        """``mpi_communicator_base.py:735-778``: mean over ranks through the
        control plane (host staged).  Used by ``AllreducePersistent`` and the
        ``mpi`` MNBN backend in the reference -- epoch-rate callers, not the
        per-step path."""
        if config.is_debug():
            self._check_ready_to_allreduce(sendbuf, recvbuf)
        src = recvbuf if sendbuf is None else sendbuf
        host = _to_host(src)
        is_float16 = host.dtype == np.float16
        work = host.astype(np.float32) if is_float16 else host.copy()
        from chainer_b200.communicators import _control_plane
        self.mpi_comm.Allreduce(_control_plane.IN_PLACE, work)
        if is_float16:
            work = work.astype(np.float16)
        work *= 1.0 / self.mpi_comm.size
        _copy_into(recvbuf, work)
        if config.is_debug():
            self._ensure_all_finite(recvbuf)

    def _pack_params_to_buffer(self, params, attr_name, buffer, allreduce_grad_dtype,
                               zero_fill, stream=None):
        """``mpi_communicator_base.py:780-796``.  Both settings of
        ``batched_copy`` run the batched kernel here (the per-parameter memcpy
        path of the reference is only slower, never different)."""
        params_data = _memory_utility.ParamsData(params, attr_name, zero_fill, stream=stream)
        _memory_utility._batched_pack_params(params_data, buffer, allreduce_grad_dtype,
                                             stream=stream)
        self.params_data = params_data

    def _unpack_params_from_buffer(self, params, attr_name, buffer, allreduce_grad_dtype,
                                   zero_fill, stream=None):
        """``mpi_communicator_base.py:798-815``."""
        if getattr(self, 'params_data', None) is not None:
            params_data = self.params_data
            self.params_data = None
        else:
            params_data = _memory_utility.ParamsData(params, attr_name, zero_fill,
                                                     stream=stream)
        _memory_utility._batched_unpack_params(params_data, buffer, allreduce_grad_dtype,
                                               stream=stream)


def _copy_into(dst, host):
    """dst[...] = host for any supported array module."""
    if isinstance(dst, np.ndarray):
        dst[...] = host.reshape(dst.shape)
    elif _dev.is_torch(dst):
        import torch
        dst.copy_(torch.from_numpy(np.ascontiguousarray(host)).reshape(dst.shape))
    elif isinstance(dst, _dev.DeviceArray):
        dst.set(host)
    else:
        raise ValueError('{} is from an unsupported array module'.format(type(dst)))
