"""Import shims for the reference arm of bench.py (`--impl reference`).

`baseline/_ref/` holds the UNMODIFIED reference (chainer/chainer v7.8.1: `chainer`,
`chainermn`, `chainerx` stubs) installed with

    python -m pip install --no-index --no-build-isolation --no-deps \
        --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>

(git-ignored; it travels to the GPU box with the snapshot).  Chainer 7 predates
Python 3.12 / NumPy 2 and ChainerMN needs mpi4py, which is not installed: this module
supplies, WITHOUT editing the installed files, (1) the NumPy names Chainer still
imports, (2) the `chainerx._build_info` module a source install lacks, and (3) an
`mpi4py` module whose COMM_WORLD runs the collectives the `naive` communicator uses
(`gather`/`scatter`/`bcast` of objects, in-place `Allreduce`, `Bcast`, `Split`) on a
torch.distributed gloo group -- host memory only; transport is sockets, not MPI.
None of the B200 package is imported here.
"""
import os
import socket
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def available():
    return os.path.isdir(os.path.join(REF_DIR, 'chainer')) and \
        os.path.isdir(os.path.join(REF_DIR, 'chainermn'))


def install_numpy_shims():
    if not hasattr(np, 'sctypes'):   # chainer/functions/array/as_strided.py:10
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
    if 'numpy.distutils' not in sys.modules:   # chainer/_environment_check.py:6
        nd = types.ModuleType('numpy.distutils')
        si = types.ModuleType('numpy.distutils.system_info')
        si.get_info = lambda *a, **k: {}
        nd.system_info = si
        sys.modules['numpy.distutils'] = nd
        sys.modules['numpy.distutils.system_info'] = si
    if 'chainerx._build_info' not in sys.modules:   # chainerx/__init__.py:4-19
        bi = types.ModuleType('chainerx._build_info')
        bi.build_chainerx = False
        sys.modules['chainerx._build_info'] = bi


class _InPlace(object):
    pass


IN_PLACE = _InPlace()


def _arr(buf):
    return buf[0] if isinstance(buf, (tuple, list)) else buf


class SelfComm(object):
    """COMM_WORLD of a one-process job."""
    rank, size = 0, 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def barrier(self):
        pass

    Barrier = barrier

    def bcast(self, obj, root=0):
        return obj

    def gather(self, obj, root=0):
        return [obj]

    def allgather(self, obj):
        return [obj]

    def scatter(self, objs, root=0):
        return objs[0]

    def allreduce(self, obj, op=None):
        return obj

    def Allreduce(self, sendbuf, recvbuf, op=None):
        if sendbuf is not IN_PLACE:
            np.copyto(_arr(recvbuf), _arr(sendbuf))

    def Bcast(self, buf, root=0):
        pass

    def Split(self, color=0, key=0):
        return SelfComm()


class GlooComm(object):
    """The mpi4py Intracomm calls of MpiCommunicatorBase on a gloo group."""

    def __init__(self, group=None, ranks=None):
        import torch.distributed as dist
        self._d = dist
        self._g = group
        self._ranks = ranks
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def _glob(self, r):
        return r if self._ranks is None else self._ranks[r]

    def barrier(self):
        self._d.barrier(group=self._g)

    Barrier = barrier

    def bcast(self, obj, root=0):
        box = [obj if self.rank == root else None]
        self._d.broadcast_object_list(box, src=self._glob(root), group=self._g)
        return box[0]

    def allgather(self, obj):
        out = [None] * self.size
        self._d.all_gather_object(out, obj, group=self._g)
        return out

    def gather(self, obj, root=0):
        out = self.allgather(obj)
        return out if self.rank == root else None

    def scatter(self, objs, root=0):
        return self.bcast(objs, root)[self.rank]

    def allreduce(self, obj, op=None):
        parts = self.allgather(obj)
        acc = parts[0]
        for p in parts[1:]:
            acc = acc + p
        return acc

    def Allreduce(self, sendbuf, recvbuf, op=None):
        import torch
        recv = _arr(recvbuf)
        if sendbuf is not IN_PLACE:
            np.copyto(recv, _arr(sendbuf))
        if recv.size:
            self._d.all_reduce(torch.from_numpy(recv), group=self._g)

    def Bcast(self, buf, root=0):
        import torch
        a = _arr(buf)
        if a.size:
            self._d.broadcast(torch.from_numpy(a), src=self._glob(root), group=self._g)

    def Split(self, color=0, key=0):
        infos = self.allgather((color, key, self.rank))
        mine = None
        for c in sorted(set(i[0] for i in infos)):
            members = [self._glob(r) for _, r in sorted((k, r) for cc, k, r in infos if cc == c)]
            g = self._d.new_group(ranks=members, backend='gloo')
            if c == color:
                mine = GlooComm(g, members)
        return mine


def install_mpi4py(world):
    mpi4py = types.ModuleType('mpi4py')
    MPI = types.ModuleType('mpi4py.MPI')
    MPI.IN_PLACE = IN_PLACE
    MPI.ANY_TAG = -1
    MPI.FLOAT, MPI.DOUBLE, MPI.INT, MPI.LONG = 'FLOAT', 'DOUBLE', 'INT', 'LONG'
    MPI._typedict = {'i': 'INT', 'l': 'LONG', 'f': 'FLOAT', 'd': 'DOUBLE'}
    MPI.Get_processor_name = socket.gethostname
    MPI.COMM_WORLD = world
    MPI.Status = type('Status', (object,), {})
    mpi4py.MPI = MPI
    sys.modules['mpi4py'] = mpi4py
    sys.modules['mpi4py.MPI'] = MPI


def import_reference(world=None):
    """(chainer, chainermn) of the unmodified reference under baseline/_ref."""
    if not available():
        raise ImportError('baseline/_ref does not hold the reference (see the module docstring)')
    install_numpy_shims()
    install_mpi4py(world if world is not None else SelfComm())
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import chainer
        import chainermn
    return chainer, chainermn
