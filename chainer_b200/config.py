"""The two global switches of Chainer that the gradient path reads.

* ``get_dtype()``  -- ``chainer.get_dtype`` (``chainer/__init__.py:293-314``):
  default dtype from ``CHAINER_DTYPE`` (``float32`` by default; ``mixed16``
  maps to float16), used as the default allreduce dtype
  (``pure_nccl_communicator.py:111-114``) and the ``bcast_data`` transfer dtype.
* ``is_debug()`` / ``set_debug()`` -- ``chainer.is_debug``: enables the
  shape-agreement and finiteness checks around the allreduce
  (``pure_nccl_communicator.py:170-174, 191-193``).

When the real ``chainer`` package is importable its functions are used, so that
``chainer.using_config('dtype', ...)`` and ``chainer.set_debug`` keep working.
"""
import os
import sys

import numpy as np

mixed16 = 'mixed16'
_debug = [os.environ.get('CHAINER_DEBUG', '0') not in ('0', '')]
_dtype = [None]
# read once at import, like chainer.global_config.dtype (chainer/__init__.py:217-218)
_env_dtype = os.environ.get('CHAINER_DTYPE', 'float32')
_modules = sys.modules


def _real_chainer():
    return _modules.get('chainer')


def get_dtype(dtype=None, map_mixed16=None):
    ch = _modules.get('chainer')
    if ch is not None and hasattr(ch, 'get_dtype'):
        return ch.get_dtype(dtype, map_mixed16)
    if dtype is None:
        dtype = _dtype[0] if _dtype[0] is not None else _env_dtype
    if isinstance(dtype, str) and dtype == mixed16:
        dtype = np.float16 if map_mixed16 is None else map_mixed16
    return np.dtype(dtype)


def set_dtype(dtype):
    """Set the global default dtype (None restores CHAINER_DTYPE); with the real ``chainer``
    imported this is ``chainer.global_config.dtype``."""
    _dtype[0] = dtype
    ch = _modules.get('chainer')
    if ch is not None and hasattr(ch, 'global_config'):
        value = _env_dtype if dtype is None else dtype
        if isinstance(value, str) and value == mixed16:
            ch.global_config.dtype = getattr(ch, 'mixed16', np.dtype(np.float16))
        else:
            ch.global_config.dtype = np.dtype(value)


def is_debug():
    ch = _real_chainer()
    if ch is not None and hasattr(ch, 'is_debug'):
        return ch.is_debug()
    return _debug[0]


def set_debug(flag):
    ch = _real_chainer()
    if ch is not None and hasattr(ch, 'set_debug'):
        ch.set_debug(flag)
    _debug[0] = bool(flag)
