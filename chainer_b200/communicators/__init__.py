"""``chainermn.communicators`` mirror: ``create_communicator``
(``chainermn/communicators/__init__.py:8-132``)."""
import numpy as np

from chainer_b200.communicators.communicator_base import CommunicatorBase  # NOQA


def create_communicator(
        communicator_name='pure_nccl', mpi_comm=None,
        allreduce_grad_dtype=None, batched_copy=True):
    """Create a communicator.

    Same signature and argument meaning as the reference.  Only ``pure_nccl``
    is implemented on this path -- the other names of the reference (``naive``,
    ``flat``, ``non_cuda_aware``, ``dummy``, the deprecated hierarchical ones)
    are legacy transports outside the scope of this package and raise
    ``ValueError`` like any unrecognised name.

    ``mpi_comm``: an mpi4py communicator if MPI is in use; by default the
    torchrun/env based control plane (``_control_plane.get_world()``).

    ``allreduce_grad_dtype``: ``numpy.float16``, ``numpy.float32``,
    ``numpy.float64``, ``'bfloat16'`` (extension) or ``None`` (then
    ``chainer.get_dtype()`` decides, as in the reference's table,
    ``__init__.py:43-53``).
    """
    if mpi_comm is None:
        from chainer_b200.communicators import _control_plane
        mpi_comm = _control_plane.get_world()

    if communicator_name != 'pure_nccl' and allreduce_grad_dtype is not None:
        raise ValueError(
            'allreduce_grad_dtype is only available'
            'at \'pure_nccl\' communicator.')

    if communicator_name == 'pure_nccl':
        from chainer_b200.communicators.pure_nccl_communicator import PureNcclCommunicator
        comm = PureNcclCommunicator(mpi_comm=mpi_comm)
        comm.set_config('allreduce_grad_dtype', allreduce_grad_dtype)
    else:
        raise ValueError(
            'Unrecognized communicator: "{}"'.format(communicator_name))

    comm.set_config('batched_copy', batched_copy)
    return comm
