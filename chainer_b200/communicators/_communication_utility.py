"""Rank discovery and NCCL bootstrap: mirror of
``chainermn/communicators/_communication_utility.py:9-76, 177-186``."""
import collections

import numpy as np

from chainer_b200 import nccl
from chainer_b200.communicators import _control_plane


def init_ranks(mpi_comm):
    """Returns (rank, intra_rank, intra_size, inter_rank, inter_size), derived
    from the processor (host) names exactly as the reference does
    (``_communication_utility.py:9-58``)."""
    global_names = mpi_comm.gather(_control_plane.get_processor_name())

    if mpi_comm.rank == 0:
        name_to_global_ranks = collections.defaultdict(list)
        for global_rank, name in enumerate(global_names):
            name_to_global_ranks[name].append(global_rank)

        for global_ranks in name_to_global_ranks.values():
            global_ranks.sort()

        inter_names = sorted(
            set(global_names), key=lambda name: name_to_global_ranks[name])
        name_to_inter_rank = {
            name: inter_rank
            for inter_rank, name in enumerate(inter_names)
        }
        inter_size = len(inter_names)

        all_ranks = []
        for global_rank, name in enumerate(global_names):
            ranks = name_to_global_ranks[name]
            intra_rank = ranks.index(global_rank)
            intra_size = len(ranks)
            inter_rank = name_to_inter_rank[name]
            all_ranks.append((
                global_rank, intra_rank, intra_size,
                inter_rank, inter_size))
        my_ranks = mpi_comm.scatter(all_ranks)
    else:
        my_ranks = mpi_comm.scatter(None)

    assert my_ranks[0] == mpi_comm.rank
    return my_ranks


def init_nccl_comm(mpi_comm):
    """``_communication_utility.py:69-76``: unique id from rank 0, broadcast over
    the control plane, then ncclCommInitRank on the CURRENT CUDA device."""
    if mpi_comm.rank == 0:
        nccl_comm_id = nccl.get_unique_id()
    else:
        nccl_comm_id = None
    nccl_comm_id = mpi_comm.bcast(nccl_comm_id)
    return nccl.NcclCommunicator(mpi_comm.size, nccl_comm_id, mpi_comm.rank)


def _get_nccl_type_id(dtype):
    """``_communication_utility.py:177-186`` (+ bfloat16)."""
    if isinstance(dtype, str) and dtype == 'bfloat16':
        return nccl.NCCL_BFLOAT16
    dtype = np.dtype(dtype)
    if dtype == np.float16:
        return nccl.NCCL_FLOAT16
    elif dtype == np.float32:
        return nccl.NCCL_FLOAT32
    elif dtype == np.float64:
        return nccl.NCCL_FLOAT64
    else:
        raise ValueError(
            'dtype must be float16, float32, or float64.')
