"""Packed-buffer utilities: the host side of the pack / unpack kernels.

Mirror of ``chainermn/communicators/_memory_utility.py`` (same names, argument
meaning and error behaviour):

* :class:`ParamsData`   -- ``_memory_utility.py:31-62``
* :class:`DeviceMemory` / :class:`HostPinnedMemory` -- ``:65-151``
* ``extract_params_set_data`` / ``extract_params_set_grad`` /
  ``count_grad_elements`` -- ``:154-172``
* ``pack_params`` / ``unpack_params`` -- ``:175-216`` (the reference does one
  ``astype`` + D2D memcpy per parameter there; here they run the same single
  batched kernel)
* ``_batched_pack_params`` / ``_batched_unpack_params`` -- ``:253-286``

Differences that are not observable through the reference API:
the device tables are one blob ``[csum int64[n+1] | gp_seg_t[n]]`` uploaded with
ONE asynchronous copy from a pinned staging ring (the reference does three
synchronous ``cupy.asarray`` uploads per step), and cumulative sizes are int64
(int32 in the reference, which overflows past 2**31 elements).
"""
import ctypes

import numpy as np

from chainer_b200 import _lib
from chainer_b200 import device as _dev

_ALIGN_ELEMS = 4  # the kernels move 4 elements per lane


def _align_up(n, a):
    return (n + a - 1) // a * a


class ParamsData(object):
    """Device tables describing a list of parameters for the batched kernels.

    ``ParamsData(params, attr_name, zero_fill)`` has the reference signature.
    Extra keyword arguments are used by the fused optimizer path to attach
    ``param.data`` and optimizer-state pointers to each segment.
    """

    def __init__(self, params, attr_name, zero_fill, extra_ptrs=None, stream=None,
                 table=None, buf_offsets=None):
        n_params = len(params)
        segs = np.zeros(n_params, dtype=_lib.SEG_DTYPE)
        csum = np.zeros(n_params + 1, dtype=np.int64)
        arrays, ptrs, ids, sizes = [], [], [], []
        for param in params:
            v = getattr(param, attr_name)
            if attr_name == 'grad' and v is None and zero_fill:
                v = _dev.zeros_like(param.data)
                setattr(param, attr_name, v)
            ptrs.append(_dev.device_ptr(v))  # ValueError: unsupported array module
            dtype = _dev.array_dtype(v)
            if isinstance(dtype, str) or dtype not in (np.float16, np.float32, np.float64):
                raise ValueError('dtype must be float16, float32 or float64.')
            ids.append(_dev.dtype_id(dtype))
            sizes.append(_dev.array_size(v))
            arrays.append(v)
        if n_params:
            np.cumsum(sizes, out=csum[1:])
            segs['ptr'][:, 0] = np.asarray(ptrs, dtype=np.uint64)
            segs['dtype0'] = ids
            segs['dtype1'] = ids
        if buf_offsets is None:
            segs['buf_off'] = csum[:-1]
        else:
            segs['buf_off'] = np.asarray(buf_offsets, dtype=np.int64)
        self._hint4 = self._hint8 = False
        self.all_float32 = False
        self.n_params = n_params
        self.n_elems = int(csum[n_params])
        self.attr_name = attr_name
        self.arrays = arrays          # keeps the arrays alive while kernels may run
        self.host_csum = csum
        self.host_segs = segs
        if extra_ptrs is not None:
            self.attach(extra_ptrs)
        self._finish_flags()
        self._table = table
        self.d_csum = None
        self.d_segs = None
        if n_params > 0:
            self.upload(stream)

    # -- fused-update support -------------------------------------------------
    def attach(self, extra_ptrs):
        """extra_ptrs: list (one per param) of (data_array, [state arrays...])."""
        segs = self.host_segs
        for i, item in enumerate(extra_ptrs):
            data, states = item[0], item[1]
            segs['ptr'][i, 1] = _dev.device_ptr(data)
            segs['dtype1'][i] = _dev.dtype_id(_dev.array_dtype(data))
            for k, s in enumerate(states):
                segs['ptr'][i, 2 + k] = _dev.device_ptr(s)
            if len(item) > 2:
                # float32 master weights: ptr[4] is the float16 parameter array behind the
                # master in ptr[1] (csrc/gp_master.cu); aligned to 4 of ITS elements
                segs['ptr'][i, 4] = _dev.device_ptr(item[2])
                self._aux4_align = 8
            self.arrays.append(item)

    def _finish_flags(self):
        segs, csum = self.host_segs, self.host_csum
        n = self.n_params
        if n == 0:
            return
        ok = (csum[:-1] % _ALIGN_ELEMS == 0) & (segs['buf_off'] % _ALIGN_ELEMS == 0)
        # every array must be aligned to 4 elements of its own type (<= 16 B:
        # the widest single access of the kernels)
        isz0 = np.where(segs['dtype0'] == _lib.GP_F64, 8, np.where(segs['dtype0'] == _lib.GP_F32, 4, 2))
        isz1 = np.where(segs['dtype1'] == _lib.GP_F64, 8, np.where(segs['dtype1'] == _lib.GP_F32, 4, 2))
        a0 = np.minimum(isz0 * 4, 16).astype(np.uint64)
        a1 = np.minimum(isz1 * 4, 16).astype(np.uint64)
        ok &= (segs['ptr'][:, 0] % a0) == 0
        for k in range(1, 4):
            ok &= (segs['ptr'][:, k] % a1) == 0
        aux = getattr(self, '_aux4_align', None)
        ok &= (segs['ptr'][:, 4] % (a1 if aux is None else np.uint64(aux))) == 0
        segs['flags'] = np.where(ok, _lib.GP_SEG_VEC_OK, 0).astype(np.uint32)
        # layout promise for the TMA-staged kernels (include/gradpath.h "layout_hint"):
        # uniform float32 arrays, 16-byte aligned pointers, offsets multiple of 8
        # (of 4 when the packed buffer is a 4-byte type too)
        f32 = bool(np.all(segs['dtype0'] == _lib.GP_F32) and np.all(segs['dtype1'] == _lib.GP_F32))
        al16 = bool(np.all(segs['ptr'] % np.uint64(16) == 0))
        self.all_float32 = f32        # what the one-launch step needs (any alignment)
        self._hint4 = f32 and al16 and bool(np.all(csum % 4 == 0) and np.all(segs['buf_off'] % 4 == 0))
        self._hint8 = self._hint4 and bool(np.all(csum % 8 == 0) and np.all(segs['buf_off'] % 8 == 0))

    def upload(self, stream=None):
        lib = _lib.get()
        if self._table is None:
            self._table = DeviceTable()
        n = self.n_params
        csum_bytes = _align_up((n + 1) * 8, 64)
        blob = np.zeros(csum_bytes + n * _lib.SEG_DTYPE.itemsize, dtype=np.uint8)
        blob[:(n + 1) * 8] = self.host_csum.view(np.uint8)
        blob[csum_bytes:] = self.host_segs.view(np.uint8)
        base = self._table.upload(blob, stream)
        self.d_csum = base
        self.d_segs = base + csum_bytes
        self._blob = blob
        return lib

    def clone(self):
        """A copy with its own host tables and device table (same parameters)."""
        new = ParamsData.__new__(ParamsData)
        new.__dict__.update(self.__dict__)
        new.host_segs = self.host_segs.copy()
        new.arrays = list(self.arrays)
        new._table = DeviceTable()
        new.d_csum = new.d_segs = None
        return new

    def set_ptr0(self, ptrs, stream=None):
        """Replace the ptr[0] column (the gradient arrays) and re-upload."""
        self.host_segs['ptr'][:, 0] = ptrs
        self._finish_flags()
        if self.n_params > 0:
            self.upload(stream)

    def layout_hint(self, buf_dtype):
        """layout_hint argument of the fused update kernels for this table."""
        if self.n_params == 0:
            return 0
        four_byte = not isinstance(buf_dtype, str) and np.dtype(buf_dtype) == np.float32
        ok = self._hint4 if four_byte else self._hint8
        return _lib.GP_F32 if ok else 0

    # reference attribute names (device arrays there; device addresses here)
    @property
    def size_csum(self):
        return self.d_csum

    @property
    def dptr(self):
        return self.d_segs


class DeviceTable(object):
    """Owner of a ``gp_table`` handle (pinned staging ring + device copies)."""

    def __init__(self):
        h = ctypes.c_void_p()
        self._lib = _lib.get()
        self._lib.gp_table_create(ctypes.byref(h))
        self.handle = h.value

    def upload(self, blob, stream=None):
        out = ctypes.c_void_p()
        self._lib.gp_table_upload(self.handle, blob.ctypes.data, blob.nbytes,
                                   _dev.stream_ptr(stream), ctypes.byref(out))
        return out.value

    def __del__(self):
        if getattr(self, 'handle', None):
            try:
                self._lib.gp_table_destroy(self.handle)
            except Exception:
                pass
            self.handle = None


class HostPinnedMemory(object):
    """``_memory_utility.py:65-91``."""

    def __init__(self):
        self.size = 0
        self.memory = None

    def assign(self, size):
        if size > self.size:
            self._free()
            p = ctypes.c_void_p()
            self._lib = _lib.get()
            self._lib.gp_malloc_host(ctypes.byref(p), size)
            self.memory = p.value
            self.size = size

    def ptr(self, offset=0):
        return ctypes.c_void_p(self.memory + offset)

    def buffer(self, size):
        return ctypes.cast(self.memory, ctypes.POINTER(ctypes.c_ubyte * size)).contents

    def array(self, count, offset=0, dtype=np.float32):
        if dtype is None:
            raise TypeError('dtype must be an instance of numpy.dtype class')
        dtype = np.dtype(dtype)
        buf = (ctypes.c_ubyte * (count * dtype.itemsize)).from_address(self.memory + offset)
        return np.frombuffer(buf, dtype=dtype, count=count)

    def _free(self):
        if self.memory:
            try:
                self._lib.gp_free_host(self.memory)
            except Exception:
                pass
            self.memory = None
            self.size = 0

    def __del__(self):
        self._free()


class DeviceMemory(object):
    """Grow-only raw device buffer (``_memory_utility.py:94-151``)."""

    def __init__(self):
        self.size = 0
        self.memory = None
        self._alloc = None
        # optional callable(nbytes) -> object with .ptr (or None: use cudaMalloc);
        # the communicator installs the multicast allocator here
        self.allocator = None

    def assign(self, size):
        if size > self.size:
            alloc = self.allocator(size) if self.allocator is not None else None
            # capacity in whole 16-byte vectors: the collective kernels round the tail up
            self._alloc = alloc if alloc is not None else _dev._Allocation((size + 15) // 16 * 16)
            self.memory = _dev._MemPtr(self._alloc.ptr, self._alloc)
            self.size = size

    def from_device(self, src, size, offset=0, stream=None):
        _lib.get().gp_memcpy_async(self.memory.ptr + offset, _dev.device_ptr(src), size, 2,
                                   _dev.stream_ptr(stream))

    def to_device(self, dst, size, offset=0, stream=None):
        _lib.get().gp_memcpy_async(_dev.device_ptr(dst), self.memory.ptr + offset, size, 2,
                                   _dev.stream_ptr(stream))

    def ptr(self):
        return self.memory.ptr

    def buffer(self, size):
        return ctypes.cast(self.memory.ptr, ctypes.POINTER(ctypes.c_ubyte * size)).contents

    def array(self, shape, offset=0, dtype=np.float32):
        if dtype is None:
            raise TypeError('dtype must be an instance of numpy.dtype class')
        return _dev.DeviceArray(shape, dtype, _dev._MemPtr(self.memory.ptr + offset, self._alloc))


def extract_params_set_data(model):
    return [param for _, param in sorted(model.namedparams())
            if param.data is not None]


def extract_params_set_grad(model, zero_fill):
    if zero_fill:
        return [param for _, param in sorted(model.namedparams())
                if param.data is not None]
    else:
        return [param for _, param in sorted(model.namedparams())
                if param.data is not None and param.grad is not None]


def count_grad_elements(params, zero_fill):
    if zero_fill:
        return sum(_dev.array_size(param.data) for param in params)
    else:
        return sum(_dev.array_size(param.grad) for param in params)


def _batched_pack_params(params_data, buffer, dtype, stream=None, scale=1.0,
                         elem_begin=0, elem_end=None):
    if params_data.n_params == 0:
        return
    if elem_end is None:
        elem_end = params_data.n_elems
    _lib.get().gp_pack(buffer.ptr(), _dev.dtype_id(dtype), params_data.d_csum,
                       params_data.d_segs, params_data.n_params, elem_begin, elem_end,
                       float(scale), params_data.layout_hint(dtype), _dev.stream_ptr(stream))


def _batched_unpack_params(params_data, buffer, dtype, stream=None, scale=1.0,
                           elem_begin=0, elem_end=None):
    if params_data.n_params == 0:
        return
    if elem_end is None:
        elem_end = params_data.n_elems
    _lib.get().gp_unpack_scale(buffer.ptr(), _dev.dtype_id(dtype), params_data.d_csum,
                               params_data.d_segs, params_data.n_params, elem_begin, elem_end,
                               float(scale), params_data.layout_hint(dtype),
                               _dev.stream_ptr(stream))


def pack_params(params, attr_name, buffer, transfer_dtype, zero_fill, stream=None):
    """Reference signature (``_memory_utility.py:175-193``); implemented with the
    batched kernel instead of one cast + memcpy per parameter."""
    if len(params) == 0:
        return
    pd = ParamsData(params, attr_name, zero_fill, stream=stream)
    _batched_pack_params(pd, buffer, transfer_dtype, stream)
    return pd


def unpack_params(params, attr_name, buffer, transfer_dtype, zero_fill, stream=None):
    """Reference signature (``_memory_utility.py:196-216``)."""
    if len(params) == 0:
        return
    for param in params:
        v = getattr(param, attr_name)
        if attr_name == 'grad' and v is None and zero_fill:
            setattr(param, attr_name, _dev.empty_like(param.data))
    pd = ParamsData(params, attr_name, False, stream=stream)
    _batched_unpack_params(pd, buffer, transfer_dtype, stream)
    return pd
