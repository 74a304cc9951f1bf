"""Peer-memory allreduce of the packed buffer (host side of ``csrc/gp_p2p.cu``).

Used by ``PureNcclCommunicator`` in place of ``nccl_comm.allReduce``
(``pure_nccl_communicator.py:180-182``) when every rank lives on the same
NVSwitch box (world size 2, 4 or 8): the packed buffers and a small flag block
of every rank are shared through CUDA IPC handles exchanged over the control
plane, and ONE kernel per rank does reduce-scatter (peer loads) + all-gather
(peer stores) + both cross-GPU barriers.
"""
import ctypes
import os

import numpy as np

from chainer_b200 import _lib
from chainer_b200 import device as _dev


def enabled_by_env():
    return os.environ.get('CHAINER_B200_P2P', '1') not in ('0', '', 'false', 'False')


class PeerAllreduce(object):

    def __init__(self, mpi_comm):
        lib = _lib.get()
        self.lib = lib
        self.mpi_comm = mpi_comm
        self.rank = mpi_comm.rank
        self.size = mpi_comm.size
        if self.size not in (2, 4, 8):
            raise ValueError('peer-memory allreduce supports 2, 4 or 8 ranks')
        nbytes = lib.gp_p2p_flag_bytes()
        self._flag_alloc = _dev._Allocation(nbytes)
        lib.gp_memset_async(self._flag_alloc.ptr, 0, nbytes, 0)
        lib.gp_stream_synchronize(0)
        self._flag_ptrs, self._flag_maps = self._exchange(self._flag_alloc.ptr)
        self._buf_alloc = None          # keeps the exported allocation alive
        self._buf_maps = []
        self._buf_ptrs = None
        # the comm handle exists from the start (packed buffers are attached later)
        arr_f = (ctypes.c_void_p * self.size)(*self._flag_ptrs)
        arr_b = (ctypes.c_void_p * self.size)(*([None] * self.size))
        h = ctypes.c_void_p()
        lib.gp_p2p_create(ctypes.byref(h), self.rank, self.size, arr_b, arr_f)
        self.handle = h.value
        # small-message area (MNBN statistics): capacity 2 * 4096 channels
        self.small_cap = 8192
        nb = lib.gp_p2p_small_bytes(self.size, self.small_cap)
        self._small_alloc = _dev._Allocation(nb)
        self._small_flag_alloc = _dev._Allocation(nbytes)
        lib.gp_memset_async(self._small_alloc.ptr, 0, nb, 0)
        lib.gp_memset_async(self._small_flag_alloc.ptr, 0, nbytes, 0)
        lib.gp_stream_synchronize(0)
        sp, self._small_maps = self._exchange(self._small_alloc.ptr)
        fp, self._small_flag_maps = self._exchange(self._small_flag_alloc.ptr)
        lib.gp_p2p_set_small(self.handle, (ctypes.c_void_p * self.size)(*sp),
                             (ctypes.c_void_p * self.size)(*fp), self.small_cap)

    # -- IPC plumbing -------------------------------------------------------------
    def _exchange(self, ptr):
        """All-gather an IPC handle of `ptr`; returns (pointers by rank, opened mappings)."""
        lib = self.lib
        h = ctypes.create_string_buffer(64)
        lib.gp_ipc_get_handle(ptr, h)
        handles = self.mpi_comm.allgather(bytes(h.raw))
        ptrs, maps = [], []
        for r, raw in enumerate(handles):
            if r == self.rank:
                ptrs.append(ptr)
            else:
                out = ctypes.c_void_p()
                lib.gp_ipc_open_handle(raw, ctypes.byref(out))
                ptrs.append(out.value)
                maps.append(out.value)
        return ptrs, maps

    def _close(self, maps):
        for p in maps:
            try:
                self.lib.gp_ipc_close_handle(p)
            except Exception:
                pass

    def ensure(self, device_memory, stream=None):
        """(Re)exchange the packed-buffer handles after a (re)allocation.  Collective:
        every rank resizes its buffer at the same step (same element count)."""
        alloc = device_memory._alloc
        if alloc is self._buf_alloc:
            return
        lib = self.lib
        # peers may still be reading the old buffer through their mappings
        lib.gp_device_synchronize()
        self._close(self._buf_maps)
        self._buf_maps = []
        self.mpi_comm.barrier()          # everybody unmapped: the old allocation may go
        self._buf_alloc = alloc
        self._buf_ptrs, self._buf_maps = self._exchange(alloc.ptr)
        arr_b = (ctypes.c_void_p * self.size)(*self._buf_ptrs)
        lib.gp_p2p_set_buffers(self.handle, arr_b)

    def allreduce(self, dtype, offset_elems, n_elems, stream):
        self.lib.gp_p2p_allreduce(self.handle, _dev.dtype_id(dtype), offset_elems, n_elems,
                                  _dev.stream_ptr(stream))

    def allreduce_small(self, in_ptr, out_ptr, n_elems, C, scale, stream):
        """out = scale * sum over ranks of in (float32, n_elems <= small_cap); with
        C > 0 additionally out[C:] -= out[:C]**2 (mean | var)."""
        self.lib.gp_p2p_allreduce_small(self.handle, in_ptr, out_ptr, n_elems, C, float(scale),
                                        _dev.stream_ptr(stream))

    def destroy(self):
        if self.handle is not None:
            self.lib.gp_device_synchronize()
            self.mpi_comm.barrier()
            self.lib.gp_p2p_destroy(self.handle)
            self.handle = None
        self._close(self._buf_maps)
        self._close(self._flag_maps)
        self._close(self._small_maps)
        self._close(self._small_flag_maps)
        self._buf_maps, self._flag_maps = [], []
        self._small_maps, self._small_flag_maps = [], []
        self.mpi_comm.barrier()
        self._buf_alloc = None
