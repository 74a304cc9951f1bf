"""Rank discovery and NCCL bootstrap over the control plane.

Same results as ``chainermn/communicators/_communication_utility.py:9-76, 177-186``
(``init_ranks``, ``init_nccl_comm``, ``_get_nccl_type_id``); written for this package's
control plane: every rank learns all host names with ONE ``allgather`` and derives its own
five numbers locally (the reference gathers to rank 0, computes everybody's tuple there
and scatters).
"""
import numpy as np

from chainer_b200 import nccl
from chainer_b200.communicators import _control_plane

_NCCL_IDS = {
    np.dtype(np.float16): nccl.NCCL_FLOAT16,
    np.dtype(np.float32): nccl.NCCL_FLOAT32,
    np.dtype(np.float64): nccl.NCCL_FLOAT64,
}


def init_ranks(mpi_comm):
    """``(rank, intra_rank, intra_size, inter_rank, inter_size)`` of this process.

    A *node* is a distinct processor (host) name; nodes are numbered by their lowest
    global rank, the processes of a node by ascending global rank -- the numbering the
    reference produces."""
    me = mpi_comm.rank
    hosts = mpi_comm.allgather(_control_plane.get_processor_name())
    first_rank_of = {}
    for r, host in enumerate(hosts):
        first_rank_of.setdefault(host, r)
    node_order = sorted(first_rank_of, key=first_rank_of.get)
    mates = [r for r, host in enumerate(hosts) if host == hosts[me]]
    return (me, mates.index(me), len(mates), node_order.index(hosts[me]), len(node_order))


def init_nccl_comm(mpi_comm):
    """NCCL communicator over all ranks of ``mpi_comm``: rank 0 draws the unique id, the
    control plane carries it, ``ncclCommInitRank`` binds to the CURRENT CUDA device."""
    uid = mpi_comm.bcast(nccl.get_unique_id() if mpi_comm.rank == 0 else None)
    return nccl.NcclCommunicator(mpi_comm.size, uid, mpi_comm.rank)


def _get_nccl_type_id(dtype):
    """NCCL datatype of a float dtype (``'bfloat16'`` is this package's extension);
    ``ValueError`` for anything else, as the reference."""
    if isinstance(dtype, str) and dtype == 'bfloat16':
        return nccl.NCCL_BFLOAT16
    try:
        return _NCCL_IDS[np.dtype(dtype)]
    except (KeyError, TypeError):
        raise ValueError('dtype must be float16, float32, or float64.')
