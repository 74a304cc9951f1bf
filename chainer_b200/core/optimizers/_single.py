"""One-parameter launches of the fused kernels: what ``update_core_gpu`` of a
single ``UpdateRule`` does (the reference launches one ElementwiseKernel per
parameter: ``momentum_sgd.py:75-88``, ``adam.py:224-332``).  The gradient array
itself plays the role of the packed buffer (offset 0, scale 1)."""
import numpy as np

from chainer_b200 import _lib
from chainer_b200 import device as _dev
from chainer_b200.communicators import _memory_utility as _mu

_table = [None]


def single_param_table(param, states):
    """ParamsData of one parameter with `states` attached (ptr[2..])."""
    if _table[0] is None or _table[0]._lib is not _lib.get():
        _table[0] = _mu.DeviceTable()
    return _mu.ParamsData([param], 'grad', False, extra_ptrs=[(param.data, list(states))],
                          table=_table[0])
