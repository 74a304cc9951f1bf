"""``create_mnbn_model``: a copy of a model in which every single-worker
``BatchNormalization`` link is replaced by ``MultiNodeBatchNormalization`` with
the same parameters and running statistics -- mirror of
``chainermn/links/create_mnbn_model.py:7-66`` for the link containers of this
package (``Chain``, ``ChainList``)."""
import copy

from chainer_b200.core import link as _link
from chainer_b200.links.batch_normalization import BatchNormalization
from chainer_b200.links.batch_normalization import MultiNodeBatchNormalization


def _to_multi_node(bn, comm, communication_backend):
    mnbn = MultiNodeBatchNormalization(
        size=bn.avg_mean.shape, comm=comm, decay=bn.decay, eps=bn.eps,
        dtype=bn._highprec_dtype, use_gamma=hasattr(bn, 'gamma'), use_beta=hasattr(bn, 'beta'),
        communication_backend=communication_backend, device=bn._device)
    mnbn.copyparams(bn)           # gamma, beta and the persistents avg_mean, avg_var, N
    mnbn.name = bn.name
    return mnbn


def create_mnbn_model(link, comm, communication_backend='auto'):
    """Returns a copy of ``link`` where BatchNormalization is replaced by
    MultiNodeBatchNormalization (``communication_backend``: ``mpi``, ``nccl`` or
    ``auto``, as for the link itself).  The original model is left untouched."""
    if isinstance(link, BatchNormalization):
        return _to_multi_node(link, comm, communication_backend)
    if isinstance(link, _link.Chain):
        children = {name: create_mnbn_model(link.__dict__[name], comm, communication_backend)
                    for name in link._children}
        clone = copy.deepcopy(link)
        for name, child in children.items():
            clone.__dict__[name] = child
        return clone
    if isinstance(link, _link.ChainList):
        children = [create_mnbn_model(child, comm, communication_backend) for child in link]
        clone = copy.deepcopy(link)
        for i, child in enumerate(children):
            clone._children[i] = child
        return clone
    return copy.deepcopy(link)
