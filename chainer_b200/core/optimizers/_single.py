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


class _Src(object):
    def __init__(self, array):
        self.data = array
        self.grad = None


def cast_copy(dst, src):
    """``dst[...] = src.astype(dst.dtype)`` on the device: a one-segment gather + cast
    (gp_pack with `dst` as the "buffer").  Used by the fp32 master-weight path
    (``chainer/optimizer.py:262-305``)."""
    n = _dev.array_size(src)
    if n != _dev.array_size(dst):
        raise ValueError('cast_copy: size mismatch')
    if n == 0:
        return
    if _table[0] is None or _table[0]._lib is not _lib.get():
        _table[0] = _mu.DeviceTable()
    pd = _mu.ParamsData([_Src(src)], 'data', False, table=_table[0])
    ddt = _dev.array_dtype(dst)
    _lib.get().gp_pack(_dev.device_ptr(dst), _dev.dtype_id(ddt), pd.d_csum, pd.d_segs, 1, 0, n,
                       1.0, pd.layout_hint(ddt), 0)


def new_like(array, dtype):
    """Uninitialised array of `array`'s shape and module with another dtype."""
    if _dev.is_torch(array):
        import torch
        td = {np.dtype(np.float16): torch.float16, np.dtype(np.float32): torch.float32,
              np.dtype(np.float64): torch.float64}[np.dtype(dtype)]
        return torch.empty(array.shape, dtype=td, device=array.device)
    if isinstance(array, np.ndarray):
        return np.empty(array.shape, dtype=dtype)
    return _dev.DeviceArray.empty(tuple(array.shape), dtype)
