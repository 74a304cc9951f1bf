"""Make the UNMODIFIED reference (chainer/chainer v7.8.1 under /root/reference)
importable on Python 3.12 / NumPy 2 without copying or editing it, and provide
an in-process stand-in for ``mpi4py`` so that ``chainermn``'s CPU communicators
run as N threads of this process.

Used only by ``make_golden.py`` in the build container (the reference tree does
not exist on the GPU box; the generated vectors are committed instead).
"""
import sys
import threading
import types

import numpy as np

REFERENCE_ROOT = '/root/reference'


def install_numpy_shims():
    if not hasattr(np, 'sctypes'):   # removed in NumPy 2; chainer/functions/array/as_strided.py:10
        np.sctypes = {
            'int': [np.int8, np.int16, np.int32, np.int64],
            'uint': [np.uint8, np.uint16, np.uint32, np.uint64],
            'float': [np.float16, np.float32, np.float64],
            'complex': [np.complex64, np.complex128],
            'others': [bool, object, bytes, str, np.void],
        }
    for name, typ in dict(bool=bool, int=int, float=float, complex=complex, object=object,
                          str=str).items():
        if name not in np.__dict__:
            setattr(np, name, typ)
    # chainer/_environment_check.py:6 imports numpy.distutils.system_info
    if 'numpy.distutils' not in sys.modules:
        nd = types.ModuleType('numpy.distutils')
        si = types.ModuleType('numpy.distutils.system_info')
        si.get_info = lambda *a, **k: {}
        nd.system_info = si
        sys.modules['numpy.distutils'] = nd
        sys.modules['numpy.distutils.system_info'] = si
    # chainerx/__init__.py:4-19 needs a build-info module (build_chainerx = False)
    if 'chainerx._build_info' not in sys.modules:
        bi = types.ModuleType('chainerx._build_info')
        bi.build_chainerx = False
        sys.modules['chainerx._build_info'] = bi


# ------------------------------------------------------------ fake mpi4py --
class _World(object):
    """Shared state of `size` ranks living in threads of this process."""

    def __init__(self, size):
        self.size = size
        self.barrier = threading.Barrier(size)
        self.slots = [None] * size
        self.lock = threading.Lock()


class _InPlace(object):
    pass


IN_PLACE = _InPlace()


def _as_array(buf):
    if isinstance(buf, (tuple, list)):
        buf = buf[0]
    return buf


class FakeComm(object):
    """The subset of mpi4py's Intracomm that MpiCommunicatorBase touches."""

    def __init__(self, world, rank):
        self.world = world
        self.rank = rank
        self.size = world.size

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def _exchange(self, obj):
        w = self.world
        w.slots[self.rank] = obj
        w.barrier.wait()
        out = list(w.slots)
        w.barrier.wait()
        return out

    def barrier(self):
        self.world.barrier.wait()

    Barrier = barrier

    def bcast(self, obj, root=0):
        return self._exchange(obj)[root]

    def gather(self, obj, root=0):
        out = self._exchange(obj)
        return out if self.rank == root else None

    def allgather(self, obj):
        return self._exchange(obj)

    def scatter(self, objs, root=0):
        return self._exchange(objs)[root][self.rank]

    def allreduce(self, obj, op=None):
        return sum(self._exchange(obj))

    def Allreduce(self, sendbuf, recvbuf, op=None):
        recv = _as_array(recvbuf)
        send = recv if sendbuf is IN_PLACE else _as_array(sendbuf)
        parts = self._exchange(np.array(send, copy=True))
        # rank-ordered summation in the buffer's dtype (MPI does not pin the order)
        acc = parts[0].copy()
        for p in parts[1:]:
            acc += p
        recv[...] = acc.reshape(recv.shape)

    def Bcast(self, buf, root=0):
        arr = _as_array(buf)
        parts = self._exchange(np.array(arr, copy=True))
        arr[...] = parts[root]

    def Split(self, color, key):
        infos = self._exchange((color, key, self.rank))
        mine = sorted([(k, r) for c, k, r in infos if c == color])
        # worlds must be shared objects: build them deterministically on rank order
        key_ = ('split', color, tuple(r for _, r in mine))
        with self.world.lock:
            reg = self.world.__dict__.setdefault('_splits', {})
            if key_ not in reg:
                reg[key_] = _World(len(mine))
        return FakeComm(reg[key_], [r for _, r in mine].index(self.rank))


def install_fake_mpi4py():
    if 'mpi4py' in sys.modules and getattr(sys.modules['mpi4py'], '_is_fake', False):
        return sys.modules['mpi4py']
    mpi4py = types.ModuleType('mpi4py')
    mpi4py._is_fake = True
    MPI = types.ModuleType('mpi4py.MPI')
    MPI.IN_PLACE = IN_PLACE
    MPI.ANY_TAG = -1
    MPI.FLOAT = 'FLOAT'
    MPI.DOUBLE = 'DOUBLE'
    MPI.INT = 'INT'
    MPI.LONG = 'LONG'
    MPI._typedict = {'i': 'INT', 'l': 'LONG', 'f': 'FLOAT', 'd': 'DOUBLE'}
    MPI.Get_processor_name = lambda: 'localhost'
    MPI.COMM_WORLD = FakeComm(_World(1), 0)

    class Status(object):
        pass
    MPI.Status = Status
    mpi4py.MPI = MPI
    sys.modules['mpi4py'] = mpi4py
    sys.modules['mpi4py.MPI'] = MPI
    return mpi4py


def import_reference():
    """Import chainer + chainermn from /root/reference; returns (chainer, chainermn)."""
    install_numpy_shims()
    install_fake_mpi4py()
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import chainer
        import chainermn
    return chainer, chainermn


def run_ranks(size, fn):
    """Run fn(comm_world_like, rank) on `size` threads; returns the list of results."""
    world = _World(size)
    out = [None] * size
    err = []

    def target(r):
        try:
            out[r] = fn(FakeComm(world, r), r)
        except BaseException as e:  # noqa
            err.append(e)
            try:
                world.barrier.abort()
            except Exception:
                pass
    ts = [threading.Thread(target=target, args=(r,)) for r in range(size)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if err:
        raise err[0]
    return out
