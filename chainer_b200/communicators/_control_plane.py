"""Control plane: the ``mpi_comm`` of this implementation.

The reference uses mpi4py for rank discovery, the NCCL-id broadcast and the
``*_obj`` collectives (``chainermn/communicators/mpi_communicator_base.py``,
``_communication_utility.py:9-76``) and is launched with ``mpiexec``.  There is
no MPI on a B200 box; processes are launched one per GPU (``torchrun`` /
``python -m torch.distributed.run``) and rendezvous through the environment
(``RANK``, ``WORLD_SIZE``, ``LOCAL_RANK``, ``MASTER_ADDR``, ``MASTER_PORT``).
:class:`TorchDistComm` offers the slice of the mpi4py ``Intracomm`` surface the
communicators use (``rank``, ``size``, ``bcast``, ``gather``, ``allgather``,
``scatter``, ``allreduce``, ``barrier``, ``send``/``recv``, ``Split``,
``Allreduce``/``Bcast`` on host buffers) on a gloo process group -- control
messages only; gradients never travel through it.

A real ``mpi4py`` communicator can be passed to ``create_communicator`` instead
and is used unchanged.
"""
import os
import socket

import numpy as np


class SingleProcessComm(object):
    """World of one process (no launcher): every collective is the identity."""

    rank = 0
    size = 1

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

    def send(self, obj, dest, tag=0):
        raise RuntimeError('send in a single-process world')

    def recv(self, source=0, tag=0, status=None):
        raise RuntimeError('recv in a single-process world')

    def Allreduce(self, sendbuf, recvbuf, op=None):
        if sendbuf is not IN_PLACE:
            np.copyto(_host(recvbuf), _host(sendbuf))

    def Bcast(self, buf, root=0):
        pass

    def Split(self, color=0, key=0):
        return SingleProcessComm()


class _InPlace(object):
    def __repr__(self):
        return 'IN_PLACE'


IN_PLACE = _InPlace()


def _host(buf):
    if isinstance(buf, (tuple, list)):
        buf = buf[0]
    return buf


class TorchDistComm(object):
    """mpi4py-like communicator over a ``torch.distributed`` (gloo) group."""

    def __init__(self, group=None, ranks=None):
        import torch.distributed as dist
        self._dist = dist
        self._group = group
        self._ranks = ranks                      # global ranks of this group, or None = world
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def _global(self, r):
        return r if self._ranks is None else self._ranks[r]

    def barrier(self):
        self._dist.barrier(group=self._group)

    Barrier = barrier

    def bcast(self, obj, root=0):
        box = [obj if self.rank == root else None]
        self._dist.broadcast_object_list(box, src=self._global(root), group=self._group)
        return box[0]

    def allgather(self, obj):
        out = [None] * self.size
        self._dist.all_gather_object(out, obj, group=self._group)
        return out

    def gather(self, obj, root=0):
        out = self.allgather(obj)
        return out if self.rank == root else None

    def scatter(self, objs, root=0):
        objs = self.bcast(objs, root)
        return objs[self.rank]

    def allreduce(self, obj, op=None):
        parts = self.allgather(obj)
        acc = parts[0]
        for p in parts[1:]:
            acc = acc + p
        return acc

    def send(self, obj, dest, tag=0):
        import pickle
        import torch
        data = pickle.dumps(obj)
        n = torch.tensor([len(data)], dtype=torch.int64)
        self._dist.send(n, self._global(dest), group=self._group, tag=tag)
        self._dist.send(torch.frombuffer(bytearray(data), dtype=torch.uint8),
                        self._global(dest), group=self._group, tag=tag)

    ssend = send

    def recv(self, source=0, tag=0, status=None):
        import pickle
        import torch
        n = torch.zeros(1, dtype=torch.int64)
        self._dist.recv(n, self._global(source), group=self._group, tag=tag)
        buf = torch.zeros(int(n.item()), dtype=torch.uint8)
        self._dist.recv(buf, self._global(source), group=self._group, tag=tag)
        return pickle.loads(buf.numpy().tobytes())

    def Allreduce(self, sendbuf, recvbuf, op=None):
        import torch
        recv = _host(recvbuf)
        if sendbuf is not IN_PLACE:
            np.copyto(recv, _host(sendbuf))
        t = torch.from_numpy(recv)
        self._dist.all_reduce(t, group=self._group)

    def Bcast(self, buf, root=0):
        import torch
        t = torch.from_numpy(_host(buf))
        self._dist.broadcast(t, src=self._global(root), group=self._group)

    def Split(self, color=0, key=0):
        infos = self.allgather((color, key, self.rank))
        colors = sorted(set(c for c, _, _ in infos))
        mine = None
        for c in colors:   # every rank must create every group, in the same order
            members = [r for _, r in sorted((k, r) for cc, k, r in infos if cc == c)]
            global_members = [self._global(r) for r in members]
            g = self._dist.new_group(ranks=global_members, backend='gloo')
            if c == color:
                mine = TorchDistComm(g, global_members)
        return mine


def get_processor_name():
    """``mpi4py.MPI.Get_processor_name()`` stand-in used by ``init_ranks``."""
    return os.environ.get('CHAINER_B200_HOSTNAME', socket.gethostname())


_world = None


def get_world():
    """The default ``mpi_comm``: COMM_WORLD of this launch.

    * inside a ``torchrun`` launch (``WORLD_SIZE`` > 1) a gloo process group is
      initialised (if the caller has not done so) and wrapped;
    * otherwise a single-process world.
    """
    global _world
    if _world is not None:
        return _world
    world_size = int(os.environ.get('WORLD_SIZE', '1'))
    try:
        import torch.distributed as dist
        have_dist = dist.is_available()
    except Exception:  # pragma: no cover
        have_dist = False
    if have_dist and dist.is_initialized():
        # the control plane needs host tensors: a gloo group beside a nccl default group
        if dist.get_backend() == 'gloo':
            _world = TorchDistComm(None)
        else:
            g = dist.new_group(backend='gloo')
            _world = TorchDistComm(g)
    elif world_size > 1 and have_dist:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        dist.init_process_group(backend='gloo', rank=int(os.environ['RANK']),
                                world_size=world_size)
        _world = TorchDistComm(None)
    else:
        _world = SingleProcessComm()
    return _world


def reset_world():
    global _world
    _world = None
