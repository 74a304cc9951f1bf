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
import socket

import numpy as np

from chainer_b200 import _lib
from chainer_b200 import device as _dev


def enabled_by_env():
    return os.environ.get('CHAINER_B200_P2P', '1') not in ('0', '', 'false', 'False')


def multicast_enabled_by_env():
    """CHAINER_B200_MULTICAST: '0' never, '1' whenever supported, unset: 4 / 8 ranks."""
    return os.environ.get('CHAINER_B200_MULTICAST')


class _McAllocation(object):
    """The packed buffer as a multicast-bound allocation (``csrc/gp_mc.cu``):
    ``ptr`` is the unicast address, ``mc_ptr`` the NVSwitch multicast one.  Owned by
    ``PeerAllreduce`` (freeing is collective), not by the garbage collector."""

    def __init__(self, handle, ptr, mc_ptr, nbytes):
        self.handle = handle
        self.ptr = ptr
        self.mc_ptr = mc_ptr
        self.nbytes = nbytes


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
        # NVSwitch multicast: supported by every device of the job?
        sup = ctypes.c_int(0)
        try:
            lib.gp_mc_supported(ctypes.byref(sup))
        except Exception:
            sup = ctypes.c_int(0)
        self.multicast_supported = all(mpi_comm.allgather(bool(sup.value)))
        self.multicast_error = None
        self._mc = None                 # current _McAllocation
        self._mc_serial = 0
        # per-tile words of the one-launch step (csrc/gp_step.cu)
        self._step_alloc = None
        self._step_maps = []
        self._step_cap = 0
        self._step_tile = 0

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
        if isinstance(alloc, _McAllocation):
            raise RuntimeError('multicast-bound buffers are not shared through CUDA IPC')
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

    # -- multicast buffers --------------------------------------------------------
    def _pass_fd(self, fd):
        """Rank 0 hands the POSIX fd of the multicast object to every peer over an
        abstract-namespace Unix socket (SCM_RIGHTS); returns the local fd."""
        comm = self.mpi_comm
        self._mc_serial += 1
        if self.rank == 0:
            name = '\0chainer_b200_mc_%d_%d' % (os.getpid(), self._mc_serial)
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.settimeout(60)
            try:
                srv.bind(name)
                srv.listen(self.size)
                comm.bcast(name, root=0)
                for _ in range(self.size - 1):
                    conn, _addr = srv.accept()
                    conn.settimeout(60)
                    try:
                        socket.send_fds(conn, [b'm'], [fd])
                        conn.recv(1)                      # peer has the fd
                    finally:
                        conn.close()
            finally:
                srv.close()
            return fd
        name = comm.bcast(None, root=0)
        cl = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        cl.settimeout(60)
        try:
            cl.connect(name)
            _msg, fds, _flags, _addr = socket.recv_fds(cl, 1, 1)
            cl.send(b'k')
        finally:
            cl.close()
        if not fds:
            raise RuntimeError('multicast handle was not received')
        return fds[0]

    def mc_allocate(self, nbytes):
        """Collective: (re)allocate the packed buffer as multicast-bound memory and
        return it (``_McAllocation``).  Every rank calls with the same size.  On
        failure anywhere, every rank gets ``None`` (caller falls back to cudaMalloc +
        peer-memory kernel) and multicast is disabled for this communicator."""
        lib = self.lib
        comm = self.mpi_comm
        self.mc_release()
        h = ctypes.c_void_p()
        err = None
        stage_ok = True

        def stage(fn):
            # run fn on this rank, then agree on success
            nonlocal err, stage_ok
            if stage_ok:
                try:
                    fn()
                except Exception as e:      # noqa: BLE001 -- reported, then collective fallback
                    err = e
                    stage_ok = False
            stage_ok = all(comm.allgather(stage_ok))
            return stage_ok

        fd0 = ctypes.c_int(-1)

        def export():
            if self.rank == 0:
                lib.gp_mc_export_fd(h.value, ctypes.byref(fd0))

        def share():
            if self.rank == 0:
                try:
                    self._pass_fd(fd0.value)
                finally:
                    os.close(fd0.value)
            else:
                fd = self._pass_fd(None)
                try:
                    lib.gp_mc_import_fd(h.value, fd)
                finally:
                    os.close(fd)

        ok = stage(lambda: lib.gp_mc_create(ctypes.byref(h), self.rank, self.size, int(nbytes)))
        # the fd hand-off blocks on its peers: only enter it when every rank is ready
        ok = ok and stage(export)
        ok = ok and stage(share)
        ok = ok and stage(lambda: lib.gp_mc_add_device(h.value))
        ok = ok and stage(lambda: lib.gp_mc_bind(h.value))
        if not ok:
            if h.value:
                lib.gp_mc_destroy(h.value)
            self.multicast_supported = False
            self.multicast_error = str(err) if err is not None else 'failed on another rank'
            return None
        uc, mc, size = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t()
        lib.gp_mc_pointers(h.value, ctypes.byref(uc), ctypes.byref(mc), ctypes.byref(size))
        lib.gp_memset_async(uc.value, 0, size.value, 0)
        lib.gp_stream_synchronize(0)
        comm.barrier()
        self._mc = _McAllocation(h.value, uc.value, mc.value, size.value)
        return self._mc

    def mc_release(self):
        """Collective: drop the current multicast buffer."""
        if self._mc is None:
            return
        self.lib.gp_device_synchronize()
        self.mpi_comm.barrier()          # nobody is still reducing through it
        self.lib.gp_mc_destroy(self._mc.handle)
        self._mc.handle = None
        self._mc.ptr = None
        self._mc = None
        self.mpi_comm.barrier()

    def mc_allreduce(self, dtype, offset_elems, n_elems, stream):
        self.lib.gp_mc_allreduce(self.handle, self._mc.handle, _dev.dtype_id(dtype), offset_elems,
                                 n_elems, _dev.stream_ptr(stream))

    def allreduce(self, dtype, offset_elems, n_elems, stream):
        self.lib.gp_p2p_allreduce(self.handle, _dev.dtype_id(dtype), offset_elems, n_elems,
                                  _dev.stream_ptr(stream))

    # -- one-launch step -----------------------------------------------------------
    def step_prepare(self, n_elems):
        """Collective when it (re)allocates: per-tile counters / flags of the one-launch
        step for a packed buffer of `n_elems` elements.  Every rank takes the same
        decision (same element count, same tuning)."""
        lib = self.lib
        tile = lib.gp_step_tile_elems()
        need = (n_elems + tile - 1) // tile
        if self._step_alloc is not None and tile == self._step_tile and need <= self._step_cap:
            return
        lib.gp_device_synchronize()
        self._close(self._step_maps)
        self._step_maps = []
        self.mpi_comm.barrier()          # nobody still signals into the old words
        cap = max(need * 2, 1024)
        nbytes = lib.gp_step_words_bytes(cap)
        self._step_alloc = _dev._Allocation(nbytes)
        lib.gp_memset_async(self._step_alloc.ptr, 0, nbytes, 0)
        lib.gp_stream_synchronize(0)
        ptrs, self._step_maps = self._exchange(self._step_alloc.ptr)
        lib.gp_p2p_set_step_words(self.handle, (ctypes.c_void_p * self.size)(*ptrs), cap, tile)
        self._step_cap, self._step_tile = cap, tile
        self.mpi_comm.barrier()          # every rank's words are zeroed and mapped

    def allreduce_small(self, in_ptr, out_ptr, n_elems, C, scale, stream):
        """out = scale * sum over ranks of in (float32, n_elems <= small_cap); with
        C > 0 additionally out[C:] -= out[:C]**2 (mean | var)."""
        self.lib.gp_p2p_allreduce_small(self.handle, in_ptr, out_ptr, n_elems, C, float(scale),
                                        _dev.stream_ptr(stream))

    def destroy(self):
        self.mc_release()
        if self.handle is not None:
            self.lib.gp_device_synchronize()
            self.mpi_comm.barrier()
            self.lib.gp_p2p_destroy(self.handle)
            self.handle = None
        self._close(self._buf_maps)
        self._close(self._flag_maps)
        self._close(self._small_maps)
        self._close(self._small_flag_maps)
        self._close(self._step_maps)
        self._step_maps, self._step_alloc = [], None
        self._buf_maps, self._flag_maps = [], []
        self._small_maps, self._small_flag_maps = [], []
        self.mpi_comm.barrier()
        self._buf_alloc = None
