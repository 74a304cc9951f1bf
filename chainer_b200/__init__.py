"""chainer_b200 -- B200-native (sm_100a) implementation of ChainerMN's
data-parallel gradient path: ``PureNcclCommunicator.multi_node_mean_grad``
(pack, cast, allreduce, 1/N scale, unpack), the following MomentumSGD / Adam
update, and MultiNodeBatchNormalization's statistics allreduce, behind the
reference's API names.  See DESIGN.md and INTEGRATION.md.

Host code is Python; every kernel is hand-written CUDA in ``libgradpath.so``
reached through ctypes (``chainer_b200._lib``).  There is no CPU fallback.
"""
__version__ = '0.1.0'

from chainer_b200 import config  # NOQA
from chainer_b200 import optimizer_hooks  # NOQA
from chainer_b200 import extensions  # NOQA
from chainer_b200.communicators import CommunicatorBase  # NOQA
from chainer_b200.communicators import create_communicator  # NOQA
from chainer_b200.optimizers import create_multi_node_optimizer  # NOQA
from chainer_b200.core import Chain, ChainList, Link, Parameter  # NOQA
from chainer_b200.core.optimizers import Adam, MomentumSGD  # NOQA
from chainer_b200.core.optimizers import SGD, CorrectedMomentumSGD, NesterovAG  # NOQA
from chainer_b200.config import get_dtype, is_debug, set_debug  # NOQA
