"""``AllreducePersistent``: average the float persistents of a model (the running
mean / variance of batch normalisation) over the workers -- mirror of
``chainermn/extensions/allreduce_persistent.py:15-47``.

The reference calls ``comm._multi_node_mean(None, array)`` once per persistent: a
device-to-host copy, an MPI_Allreduce and a host-to-device copy EACH (106 round
trips for ResNet-50's 53 BN layers).  With a ``PureNcclCommunicator`` the
persistents are gathered by the batched pack kernel into one buffer, reduced by
ONE NCCL allreduce, scaled by 1/size and scattered back by the batched unpack
kernel: 4 launches in total, nothing leaves the device.  Other communicators run
the reference sequence.
"""
import numpy as np

from chainer_b200 import device as _dev
from chainer_b200.communicators import _memory_utility

PRIORITY_WRITER = 300          # chainer/training/extension.py


def _namedpersistents(model):
    for lname, link in model.namedlinks():
        for pname in link._persistent:
            yield lname + '/' + pname, link.__dict__[pname]


class _Holder(object):
    """What ParamsData needs of a parameter, for a bare array."""

    def __init__(self, array):
        self.data = array
        self.grad = None


class AllreducePersistent(object):
    """Trainer extension (``trigger = 1, 'epoch'``) -- or plain callable -- that
    replaces every float persistent of ``model`` by its mean over the workers.
    Integer persistents (``N``) are ignored, as in the reference."""

    trigger = 1, 'epoch'
    # called earlier than evaluators (allreduce_persistent.py:36-37)
    priority = PRIORITY_WRITER + 1

    def __init__(self, model, comm):
        self.model = model
        self.comm = comm
        self._buffers = {}

    def _float_persistents(self):
        out = []
        for _, value in sorted(_namedpersistents(self.model), key=lambda kv: kv[0]):
            if not hasattr(value, 'dtype'):
                continue                     # python ints and the like
            dt = _dev.array_dtype(value)
            if isinstance(dt, str) or np.dtype(dt).kind != 'f':
                continue
            out.append(value)
        return out

    def __call__(self, trainer=None):
        arrays = self._float_persistents()
        if not arrays:
            return
        comm = self.comm
        from chainer_b200.communicators.pure_nccl_communicator import PureNcclCommunicator
        if not isinstance(comm, PureNcclCommunicator):
            for a in arrays:
                comm._multi_node_mean(None, a)
            return
        # float16 / float32 persistents travel in one float32 buffer (the reference
        # up-casts float16 for the sum, mpi_communicator_base.py:757-763); float64 in
        # a buffer of its own
        groups = {}
        for a in arrays:
            wide = np.dtype(np.float64) if _dev.array_dtype(a) == np.float64 else np.dtype(np.float32)
            groups.setdefault(wide, []).append(a)
        stream = _dev.Stream.null
        for dt, group in groups.items():
            holders = [_Holder(a) for a in group]
            pd = _memory_utility.ParamsData(holders, 'data', False, stream=stream)
            if pd.n_elems == 0:
                continue
            buf = self._buffers.get(dt)
            if buf is None:
                buf = self._buffers[dt] = _memory_utility.DeviceMemory()
            buf.assign(pd.n_elems * dt.itemsize)
            _memory_utility._batched_pack_params(pd, buf, dt, stream)
            comm._multi_node_mean_nccl(buf, buf, pd.n_elems, dt, stream)
            _memory_utility._batched_unpack_params(pd, buf, dt, stream)
