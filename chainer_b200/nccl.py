"""NCCL binding: mirror of ``chainermn/nccl.py:1-14`` (which re-exports
``cupy.cuda.nccl``).  The calls go to libnccl.so.2 through libgradpath's
``gp_nccl_*`` entry points (dlopen at first use)."""
import ctypes

from chainer_b200 import _lib

NCCL_FLOAT16 = 6
NCCL_FLOAT32 = 7
NCCL_FLOAT64 = 8
NCCL_BFLOAT16 = 9
NCCL_SUM = 0

_loaded = [False]
_available = True   # resolved lazily: see _ensure_loaded


def _ensure_loaded():
    if not _loaded[0]:
        lib = _lib.get()
        lib.gp_nccl_load(_lib.find_libnccl().encode())
        _loaded[0] = True


def get_version():
    _ensure_loaded()
    v = ctypes.c_int()
    _lib.get().gp_nccl_version(ctypes.byref(v))
    return v.value


def get_build_version():
    return get_version()


def get_unique_id():
    _ensure_loaded()
    buf = ctypes.create_string_buffer(_lib.GP_NCCL_UNIQUE_ID_BYTES)
    _lib.get().gp_nccl_get_unique_id(buf)
    return bytes(buf.raw)


class NcclCommunicator(object):
    """``cupy.cuda.nccl.NcclCommunicator`` surface used by the reference:
    ``allReduce(sendbuf, recvbuf, count, datatype, op, stream)``,
    ``bcast(buff, count, datatype, root, stream)``,
    ``reduce(...)``, ``destroy()`` -- all with raw integer pointers."""

    def __init__(self, ndev, commId, rank):
        _ensure_loaded()
        h = ctypes.c_void_p()
        self._lib = _lib.get()
        self._lib.gp_nccl_comm_init_rank(ctypes.byref(h), ndev, bytes(commId), rank)
        self.handle = h.value
        self.size = ndev
        self.rank = rank

    def allReduce(self, sendbuf, recvbuf, count, datatype, op, stream):
        _lib.get().gp_nccl_allreduce(self.handle, sendbuf, recvbuf, count, datatype, op, stream)

    def bcast(self, buff, count, datatype, root, stream):
        _lib.get().gp_nccl_bcast(self.handle, buff, count, datatype, root, stream)

    def reduce(self, sendbuf, recvbuf, count, datatype, op, root, stream):
        _lib.get().gp_nccl_reduce(self.handle, sendbuf, recvbuf, count, datatype, op, root, stream)

    def destroy(self):
        if self.handle:
            self._lib.gp_nccl_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
