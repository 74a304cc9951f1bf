"""Device-buffer plumbing: array protocol helpers, streams, events and a tiny
CuPy-like device array.

The reference passes ``cupy.ndarray`` objects around and takes their
``.data.ptr`` (``chainermn/communicators/_memory_utility.py:44-57``).  CuPy is
only a container of device memory on this path, so any object that exposes a
device pointer is accepted here:

* ``cupy.ndarray`` / :class:`DeviceArray` -- ``a.data.ptr``
* ``torch.Tensor`` on a CUDA device        -- ``a.data_ptr()``
* anything with ``__cuda_array_interface__``

Host (NumPy) arrays are rejected, exactly as the reference rejects arrays of an
unsupported module (``_memory_utility.py:50-52``) -- except when the test-suite
has swapped in its host-memory double of the C library (``_lib.set_backend_for_testing``).
"""
import ctypes
import operator
import sys as _sys

import numpy as np

from chainer_b200 import _lib

BF16 = 'bfloat16'

_NP_TO_ID = {
    np.dtype(np.float16): _lib.GP_F16,
    np.dtype(np.float32): _lib.GP_F32,
    np.dtype(np.float64): _lib.GP_F64,
}


def _torch():
    try:
        import torch
        return torch
    except Exception:  # pragma: no cover
        return None


_TT = [None, None, None]      # torch.Tensor, {torch dtype: numpy dtype | 'bfloat16'}, {torch dtype: id}


def _torch_tables():
    torch = _torch()
    if torch is None:
        return None
    _TT[0] = torch.Tensor
    _TT[1] = {torch.float32: np.dtype(np.float32), torch.float16: np.dtype(np.float16),
              torch.float64: np.dtype(np.float64), torch.bfloat16: BF16,
              torch.int32: np.dtype(np.int32), torch.int64: np.dtype(np.int64)}
    _TT[2] = {torch.float32: _lib.GP_F32, torch.float16: _lib.GP_F16,
              torch.float64: _lib.GP_F64, torch.bfloat16: _lib.GP_BF16}
    return _TT[0]


def is_torch(a):
    t = _TT[0]
    if t is None:
        if 'torch' not in _sys.modules:        # nobody can hold a tensor yet
            return False
        t = _torch_tables()
    return t is not None and isinstance(a, t)


def array_dtype(a):
    """numpy dtype of an array of any supported module; the string 'bfloat16'
    for torch.bfloat16 (NumPy has no such dtype)."""
    if is_torch(a):
        try:
            return _TT[1][a.dtype]
        except KeyError:
            raise ValueError('unsupported torch dtype {}'.format(a.dtype))
    return np.dtype(a.dtype)


def array_dtype_id(a):
    """``dtype_id(array_dtype(a))`` in one step (the hot callers' form)."""
    if is_torch(a):
        try:
            return _TT[2][a.dtype]
        except KeyError:
            raise ValueError('dtype must be float16, float32, or float64.')
    return dtype_id(np.dtype(a.dtype))


def array_size(a):
    if is_torch(a):
        return a.numel()
    return int(a.size)


def array_shape(a):
    return tuple(a.shape)


def dtype_id(dtype):
    """NCCL dtype id (== libgradpath dtype id); ValueError for non-float dtypes
    like `_get_nccl_type_id` (``_communication_utility.py:177-186``)."""
    if isinstance(dtype, str) and dtype == BF16:
        return _lib.GP_BF16
    try:
        return _NP_TO_ID[np.dtype(dtype)]
    except (KeyError, TypeError):
        raise ValueError('dtype must be float16, float32, or float64.')


def dtype_itemsize(dtype):
    if isinstance(dtype, str) and dtype == BF16:
        return 2
    return np.dtype(dtype).itemsize


_DATA_PTR = operator.methodcaller('data_ptr')


def ptr_key(arrays):
    """Tuple of the raw addresses of `arrays` -- the identity of a gradient set as the
    kernels see it.  One C call per torch tensor (no validation: callers validate a set
    the first time they see it); the generic route otherwise."""
    try:
        return tuple(map(_DATA_PTR, arrays))
    except AttributeError:
        return tuple(device_ptr(a) for a in arrays)


def device_ptr(a):
    """Raw device address of `a` (int)."""
    if is_torch(a):                                      # (first: `tensor.data` is not free)
        if not a.is_cuda and not _lib.get().accepts_host_pointers:
            raise ValueError('{} is not on a CUDA device'.format(type(a)))
        if not a.is_contiguous():
            raise ValueError('non-contiguous array')
        return int(a.data_ptr())
    data = getattr(a, 'data', None)
    if data is not None and hasattr(data, 'ptr'):        # cupy / DeviceArray
        return int(data.ptr)
    cai = getattr(a, '__cuda_array_interface__', None)
    if cai is not None:
        return int(cai['data'][0])
    if isinstance(a, np.ndarray) and _lib.get().accepts_host_pointers:
        if not a.flags.c_contiguous:
            raise ValueError('non-contiguous array')
        return int(a.ctypes.data)
    raise ValueError('{} is from an unsupported array module'.format(type(a)))


def zeros_like(a):
    """`xp.zeros_like(a)` for whichever module `a` comes from."""
    if isinstance(a, DeviceArray):
        return DeviceArray.zeros(a.shape, a.dtype)
    if is_torch(a):
        return _torch().zeros_like(a)
    if isinstance(a, np.ndarray):
        return np.zeros_like(a)
    mod = type(a).__module__.split('.')[0]
    if mod == 'cupy':  # pragma: no cover - CuPy is absent from this image
        import cupy
        return cupy.zeros_like(a)
    raise ValueError('{} is from an unsupported array module'.format(type(a)))


def empty_like(a):
    if isinstance(a, DeviceArray):
        return DeviceArray.empty(a.shape, a.dtype)
    if is_torch(a):
        return _torch().empty_like(a)
    if isinstance(a, np.ndarray):
        return np.empty_like(a)
    return zeros_like(a)


def to_numpy(a):
    """Host copy (synchronises)."""
    if isinstance(a, np.ndarray):
        return a
    if isinstance(a, DeviceArray):
        return a.get()
    if is_torch(a):
        torch = _torch()
        if a.dtype == torch.bfloat16:
            return a.detach().to(torch.float32).cpu().numpy()
        return a.detach().cpu().numpy()
    if hasattr(a, 'get'):
        return a.get()
    raise ValueError('{} is from an unsupported array module'.format(type(a)))


# ------------------------------------------------------------------ streams --
class Stream(object):
    """CUDA stream handle (mirror of ``chainer.cuda.Stream``: ``.ptr``,
    ``.synchronize()``).  ``Stream.null`` is the legacy default stream."""

    null = None

    def __init__(self, null=False, non_blocking=False, ptr=None):
        self._owned = False
        if null:
            self.ptr = 0
        elif ptr is not None:
            self.ptr = int(ptr)
        else:
            p = ctypes.c_void_p()
            self._lib = _lib.get()
            self._lib.gp_stream_create(ctypes.byref(p), 1 if non_blocking else 0)
            self.ptr = p.value or 0
            self._owned = True

    def synchronize(self):
        _lib.get().gp_stream_synchronize(self.ptr)

    def wait_event(self, event):
        _lib.get().gp_stream_wait_event(self.ptr, event.ptr)

    def record(self, event=None):
        if event is None:
            event = Event()
        event.record(self)
        return event

    def __eq__(self, other):
        return isinstance(other, Stream) and self.ptr == other.ptr

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.ptr)

    def __del__(self):
        if getattr(self, '_owned', False) and self.ptr:
            try:
                self._lib.gp_stream_destroy(self.ptr)
            except Exception:
                pass


Stream.null = Stream(null=True)


def stream_ptr(stream):
    """cudaStream_t (int) of a Stream / torch stream / None (= null stream)."""
    if stream is None:
        return 0
    if isinstance(stream, int):
        return stream
    p = getattr(stream, 'ptr', None)
    if p is not None:
        return int(p)
    p = getattr(stream, 'cuda_stream', None)  # torch.cuda.Stream
    if p is not None:
        return int(p)
    raise TypeError('not a stream: {!r}'.format(stream))


class Event(object):
    def __init__(self, timing=False):
        p = ctypes.c_void_p()
        self._lib = _lib.get()
        self._lib.gp_event_create(ctypes.byref(p), 1 if timing else 0)
        self.ptr = p.value

    def record(self, stream=None):
        _lib.get().gp_event_record(self.ptr, stream_ptr(stream))

    def synchronize(self):
        _lib.get().gp_event_synchronize(self.ptr)

    def elapsed_ms(self, end):
        ms = ctypes.c_float()
        _lib.get().gp_event_elapsed_ms(ctypes.byref(ms), self.ptr, end.ptr)
        return ms.value

    def __del__(self):
        if getattr(self, 'ptr', None):
            try:
                self._lib.gp_event_destroy(self.ptr)
            except Exception:
                pass


# -------------------------------------------------------------- DeviceArray --
class _MemPtr(object):
    __slots__ = ('ptr', '_owner')

    def __init__(self, ptr, owner=None):
        self.ptr = ptr
        self._owner = owner

    # a raw address cannot be duplicated by copying the Python object: DeviceArray copies
    # the MEMORY (below); copying the handle alone would alias it
    def __copy__(self):
        raise TypeError('device memory handles are not copyable; copy the DeviceArray')

    def __deepcopy__(self, memo):
        raise TypeError('device memory handles are not copyable; copy the DeviceArray')


class _Allocation(object):
    def __init__(self, nbytes):
        p = ctypes.c_void_p()
        self._lib = _lib.get()     # free with the library that allocated
        self._lib.gp_malloc(ctypes.byref(p), max(int(nbytes), 1))
        self.ptr = p.value
        self.nbytes = nbytes

    def __del__(self):
        if getattr(self, 'ptr', None):
            try:
                self._lib.gp_free(self.ptr)
            except Exception:
                pass
            self.ptr = None

    # one owner per cudaMalloc: a copied owner would free the same pointer twice
    def __copy__(self):
        raise TypeError('device allocations are not copyable')

    def __deepcopy__(self, memo):
        raise TypeError('device allocations are not copyable')


class DeviceArray(object):
    """A minimal C-contiguous device array with the slice of the
    ``cupy.ndarray`` surface this path touches: ``.data.ptr``, ``.shape``,
    ``.size``, ``.dtype``, ``.nbytes``, ``.get()``, ``.set()``, ``fill``.

    It exists so that the path can run where CuPy is not installed; where it
    is, pass ``cupy.ndarray`` objects instead -- both are only containers of
    device memory here."""

    def __init__(self, shape, dtype, memptr=None):
        if isinstance(shape, int):
            shape = (shape,)
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        if memptr is None:
            alloc = _Allocation(self.nbytes)
            memptr = _MemPtr(alloc.ptr, alloc)
        self.data = memptr

    @property
    def ndim(self):
        return len(self.shape)

    @classmethod
    def empty(cls, shape, dtype=np.float32):
        return cls(shape, dtype)

    @classmethod
    def zeros(cls, shape, dtype=np.float32):
        a = cls(shape, dtype)
        _lib.get().gp_memset_async(a.data.ptr, 0, a.nbytes, 0)
        return a

    @classmethod
    def from_numpy(cls, arr, stream=None):
        arr = np.ascontiguousarray(arr)
        a = cls(arr.shape, arr.dtype)
        a.set(arr, stream)
        return a

    def set(self, arr, stream=None):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        assert arr.size == self.size
        lib = _lib.get()
        lib.gp_memcpy_async(self.data.ptr, arr.ctypes.data, self.nbytes, 0, stream_ptr(stream))
        lib.gp_stream_synchronize(stream_ptr(stream))  # pageable source

    def get(self, stream=None):
        out = np.empty(self.shape, dtype=self.dtype)
        lib = _lib.get()
        lib.gp_memcpy_async(out.ctypes.data, self.data.ptr, self.nbytes, 1, stream_ptr(stream))
        lib.gp_stream_synchronize(stream_ptr(stream))
        return out

    def fill(self, value):
        self.set(np.full(self.shape, value, dtype=self.dtype))

    def view1d(self, offset_elems, count):
        """A 1-d view of `count` elements starting at `offset_elems` (shares memory)."""
        return DeviceArray((count,), self.dtype,
                           _MemPtr(self.data.ptr + offset_elems * self.dtype.itemsize, self.data))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        out = DeviceArray(shape, self.dtype, self.data)
        assert out.size == self.size
        return out

    def copy(self):
        out = DeviceArray(self.shape, self.dtype)
        _lib.get().gp_memcpy_async(out.data.ptr, self.data.ptr, self.nbytes, 2, 0)
        return out

    # copy.copy / copy.deepcopy (Link.copyparams, the double-buffering optimizer's
    # deepcopy(target), optimizer state): a NEW allocation with the same contents
    def __copy__(self):
        return self.copy()

    def __deepcopy__(self, memo):
        out = self.copy()
        memo[id(self)] = out
        return out

    def __repr__(self):
        return 'DeviceArray(shape={}, dtype={})'.format(self.shape, self.dtype)
