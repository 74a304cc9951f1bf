"""ctypes binding of ``libgradpath.so`` (the C-ABI declared in ``include/gradpath.h``).

The reference reaches its kernels through CuPy's NVRTC JIT
(``chainer.cuda.raw`` / ``chainer.cuda.elementwise``:
``chainermn/communicators/_memory_utility.py:289-429``,
``pure_nccl_communicator.py:183-186``, ``chainer/optimizers/momentum_sgd.py:80-85``,
``chainer/optimizers/adam.py:313-327``) and NCCL through ``cupy.cuda.nccl``
(``chainermn/nccl.py:1-14``).  Here the host code stays Python and calls the
ahead-of-time compiled sm_100a kernels through this thin ctypes layer.

There is deliberately NO fallback: if the shared library is missing, `load()`
raises, and every product entry point needs it.
"""
import ctypes
import os

import numpy as np

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_int64 = ctypes.c_int64
c_double = ctypes.c_double
c_size_t = ctypes.c_size_t
c_char_p = ctypes.c_char_p

GP_F16, GP_F32, GP_F64, GP_BF16 = 6, 7, 8, 9
GP_SEG_VEC_OK = 1
GP_ADAM_AMSGRAD = 1
GP_ADAM_ADABOUND = 2
GP_NCCL_SUM = 0
GP_NCCL_UNIQUE_ID_BYTES = 128

#: numpy mirror of ``gp_seg_t`` (64 bytes)
SEG_DTYPE = np.dtype([
    ('ptr', np.uint64, (5,)),
    ('buf_off', np.int64),
    ('dtype0', np.int32),
    ('dtype1', np.int32),
    ('flags', np.uint32),
    ('reserved', np.uint32),
])
assert SEG_DTYPE.itemsize == 64



RULE_SGD, RULE_CORRECTED_MOMENTUM, RULE_NESTEROV_AG = 0, 1, 2


class GpHooks(ctypes.Structure):
    """ctypes mirror of ``gp_hooks_t`` (include/gradpath.h)."""
    _fields_ = [('clip_rate', ctypes.c_void_p), ('weight_decay', ctypes.c_double),
                ('loss_scale', ctypes.c_double)]


_P = ctypes.POINTER
# name -> (restype, argtypes); one entry per declaration of include/gradpath.h
PROTOTYPES = {
    'gp_last_error': (c_char_p, []),
    'gp_abi_version': (c_int, []),
    'gp_device_count': (c_int, [_P(c_int)]),
    'gp_set_device': (c_int, [c_int]),
    'gp_get_device': (c_int, [_P(c_int)]),
    'gp_device_synchronize': (c_int, []),
    'gp_device_sm_count': (c_int, [_P(c_int)]),
    'gp_malloc': (c_int, [_P(c_void_p), c_size_t]),
    'gp_free': (c_int, [c_void_p]),
    'gp_malloc_host': (c_int, [_P(c_void_p), c_size_t]),
    'gp_free_host': (c_int, [c_void_p]),
    'gp_memcpy_async': (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    'gp_memset_async': (c_int, [c_void_p, c_int, c_size_t, c_void_p]),
    'gp_stream_create': (c_int, [_P(c_void_p), c_int]),
    'gp_stream_destroy': (c_int, [c_void_p]),
    'gp_stream_synchronize': (c_int, [c_void_p]),
    'gp_stream_wait_event': (c_int, [c_void_p, c_void_p]),
    'gp_event_create': (c_int, [_P(c_void_p), c_int]),
    'gp_event_destroy': (c_int, [c_void_p]),
    'gp_event_record': (c_int, [c_void_p, c_void_p]),
    'gp_event_synchronize': (c_int, [c_void_p]),
    'gp_event_elapsed_ms': (c_int, [_P(ctypes.c_float), c_void_p, c_void_p]),
    'gp_table_create': (c_int, [_P(c_void_p)]),
    'gp_table_destroy': (c_int, [c_void_p]),
    'gp_table_upload': (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, _P(c_void_p)]),
    'gp_pack': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int64,
                        c_double, c_int, c_void_p]),
    'gp_unpack_scale': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                c_double, c_int, c_void_p]),
    'gp_unpack_momentum_sgd': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64,
                                       c_int64, c_double, c_double, c_double, c_int, c_int, c_void_p]),
    'gp_unpack_adam': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int64,
                               c_double, c_double, c_double, c_double, c_double, c_double,
                               c_double, c_double, c_double, c_int, c_int, c_int, c_void_p]),
    'gp_scale': (c_int, [c_void_p, c_int, c_int64, c_double, c_void_p]),
    'gp_check_finite': (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    'gp_bn_workspace_bytes': (c_size_t, [c_int64]),
    'gp_bn_workspace_layout': (c_int, [c_int64, _P(c_int64)]),
    'gp_bn_fwd_stats': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int,
                                c_void_p, c_void_p]),
    'gp_bn_fwd_mean_var': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int,
                                   c_void_p, c_void_p]),
    'gp_bn_bwd_stats': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    'gp_bn_finish_mean_var': (c_int, [c_void_p, c_int, c_int64, c_double, c_void_p, c_void_p]),
    'gp_bn_fwd_stats_allreduce': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64,
                                          c_void_p, c_void_p, c_void_p]),
    'gp_bn_bwd_stats_allreduce': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                          c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p,
                                          c_void_p, c_void_p]),
    'gp_bn_fwd_apply': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_double, c_double, c_void_p]),
    'gp_bn_bwd_apply': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int64,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_double,
                                c_void_p, c_void_p]),
    'gp_nccl_load': (c_int, [c_char_p]),
    'gp_nccl_version': (c_int, [_P(c_int)]),
    'gp_nccl_get_unique_id': (c_int, [c_char_p]),
    'gp_nccl_comm_init_rank': (c_int, [_P(c_void_p), c_int, c_char_p, c_int]),
    'gp_nccl_comm_destroy': (c_int, [c_void_p]),
    'gp_nccl_allreduce': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                  c_void_p]),
    'gp_nccl_bcast': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    'gp_nccl_reduce': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                               c_void_p]),
    'gp_nccl_group_start': (c_int, []),
    'gp_nccl_group_end': (c_int, []),
    'gp_nccl_mem_alloc': (c_int, [_P(c_void_p), c_size_t]),
    'gp_nccl_mem_free': (c_int, [c_void_p]),
    'gp_nccl_comm_register': (c_int, [c_void_p, c_void_p, c_size_t, _P(c_void_p)]),
    'gp_nccl_comm_deregister': (c_int, [c_void_p, c_void_p]),
    'gp_unpack_momentum_sgd_hooked': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                              c_double, c_double, c_double, c_int, c_int,
                                              c_void_p, c_void_p]),
    'gp_unpack_adam_hooked': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                      c_double, c_double, c_double, c_double, c_double, c_double,
                                      c_double, c_double, c_double, c_int, c_int, c_int,
                                      c_void_p, c_void_p]),
    'gp_unpack_momentum_sgd_master': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64,
                                              c_int64, c_double, c_double, c_double, c_int,
                                              c_void_p, c_void_p, c_void_p]),
    'gp_unpack_adam_master': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                      c_double, c_double, c_double, c_double, c_double, c_double,
                                      c_double, c_double, c_double, c_int, c_int, c_void_p,
                                      c_void_p, c_void_p]),
    'gp_unpack_sgd_family': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                     c_double, c_int, c_double, c_double, c_int, c_int, c_void_p,
                                     c_void_p]),
    'gp_sqnorm_workspace_bytes': (c_size_t, []),
    'gp_sqnorm': (c_int, [c_void_p, c_int, c_int64, c_double, c_int, c_double, c_void_p, c_void_p,
                          c_void_p]),
    'gp_scale_by_device': (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    'gp_weight_decay': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_double, c_void_p]),
    'gp_divide': (c_int, [c_void_p, c_int, c_int64, c_double, c_void_p]),
    'gp_mc_supported': (c_int, [_P(c_int)]),
    'gp_mc_create': (c_int, [_P(c_void_p), c_int, c_int, c_size_t]),
    'gp_mc_export_fd': (c_int, [c_void_p, _P(c_int)]),
    'gp_mc_import_fd': (c_int, [c_void_p, c_int]),
    'gp_mc_add_device': (c_int, [c_void_p]),
    'gp_mc_bind': (c_int, [c_void_p]),
    'gp_mc_pointers': (c_int, [c_void_p, _P(c_void_p), _P(c_void_p), _P(c_size_t)]),
    'gp_mc_destroy': (c_int, [c_void_p]),
    'gp_mc_allreduce': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p]),
    'gp_mc_set_tuning': (c_int, [c_int, c_int, c_int]),
    'gp_step_supported': (c_int, [c_int, c_int, c_int, c_double, c_int]),
    'gp_step_momentum_sgd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                     c_int64, c_double, c_double, c_double, c_int, c_int, c_void_p]),
    'gp_step_adam': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64,
                             c_double, c_double, c_double, c_double, c_double, c_double, c_double,
                             c_double, c_double, c_int, c_int, c_int, c_void_p]),
    'gp_step_words_bytes': (c_size_t, [c_int64]),
    'gp_step_tile_elems': (c_int, []),
    'gp_p2p_set_step_words': (c_int, [c_void_p, _P(c_void_p), c_int64, c_int64]),
    'gp_step_set_tuning': (c_int, [c_char_p, c_int]),
    'gp_nccl_comm_window_register': (c_int, [c_void_p, c_void_p, c_size_t, _P(c_void_p), c_int]),
    'gp_nccl_comm_window_deregister': (c_int, [c_void_p, c_void_p]),
    'gp_ipc_get_handle': (c_int, [c_void_p, c_char_p]),
    'gp_ipc_open_handle': (c_int, [c_char_p, _P(c_void_p)]),
    'gp_ipc_close_handle': (c_int, [c_void_p]),
    'gp_p2p_flag_bytes': (c_size_t, []),
    'gp_p2p_create': (c_int, [_P(c_void_p), c_int, c_int, _P(c_void_p), _P(c_void_p)]),
    'gp_p2p_set_buffers': (c_int, [c_void_p, _P(c_void_p)]),
    'gp_p2p_destroy': (c_int, [c_void_p]),
    'gp_p2p_allreduce': (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p]),
    'gp_p2p_set_tuning': (c_int, [c_int, c_int, c_int]),
    'gp_p2p_small_bytes': (c_size_t, [c_int, c_int64]),
    'gp_p2p_set_small': (c_int, [c_void_p, _P(c_void_p), _P(c_void_p), c_int64]),
    'gp_p2p_allreduce_small': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_double,
                                       c_void_p]),
    'gp_nvtx_push': (c_int, [c_char_p]),
    'gp_nvtx_pop': (c_int, []),
    'gp_set_tuning': (c_int, [c_char_p, c_int]),
    'gp_get_tuning': (c_int, [c_char_p, _P(c_int)]),
}

# entry points that launch exactly one of OUR kernels (counted in `launches`)
KERNEL_FUNCS = frozenset([
    'gp_pack', 'gp_unpack_scale', 'gp_unpack_momentum_sgd', 'gp_unpack_adam', 'gp_scale',
    'gp_check_finite', 'gp_bn_fwd_stats', 'gp_bn_fwd_mean_var', 'gp_bn_bwd_stats', 'gp_bn_finish_mean_var',
    'gp_p2p_allreduce', 'gp_p2p_allreduce_small', 'gp_mc_allreduce',
    'gp_unpack_momentum_sgd_hooked', 'gp_unpack_adam_hooked', 'gp_sqnorm', 'gp_scale_by_device',
    'gp_weight_decay', 'gp_divide', 'gp_unpack_sgd_family', 'gp_step_momentum_sgd', 'gp_step_adam',
    'gp_bn_fwd_stats_allreduce', 'gp_bn_bwd_stats_allreduce', 'gp_bn_fwd_apply', 'gp_bn_bwd_apply',
    'gp_unpack_momentum_sgd_master', 'gp_unpack_adam_master'])

# functions whose int return value is an error code
_NO_CHECK = {'gp_last_error', 'gp_abi_version', 'gp_nvtx_push', 'gp_nvtx_pop', 'gp_bn_workspace_bytes', 'gp_p2p_flag_bytes', 'gp_p2p_small_bytes',
             'gp_sqnorm_workspace_bytes', 'gp_step_supported', 'gp_step_words_bytes', 'gp_step_tile_elems'}


class GradpathError(RuntimeError):
    """A libgradpath call failed (CUDA / NCCL error or bad argument)."""

    def __init__(self, fn, code, message):
        super(GradpathError, self).__init__(
            '{} failed with code {}: {}'.format(fn, code, message))
        self.fn = fn
        self.code = code


def library_path():
    env = os.environ.get('CHAINER_B200_LIBGRADPATH')
    if env:
        return env
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libgradpath.so')


class _Lib(object):
    """The loaded library.  Each C function becomes a method that raises
    :class:`GradpathError` on a non-zero return code."""

    accepts_host_pointers = False  # real kernels dereference device pointers only

    def __init__(self, path):
        self.path = path
        self.launches = 0      # kernels of this library launched so far
        self.cdll = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
        for name, (restype, argtypes) in PROTOTYPES.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError:
                raise ImportError(
                    '{} does not export {} (stale build? run __graft_entry__.build())'.format(
                        path, name))
            fn.restype = restype
            fn.argtypes = argtypes
            if name in _NO_CHECK:
                setattr(self, name, fn)
            else:
                setattr(self, name, self._checked(name, fn))
        if self.gp_abi_version() != 1:
            raise ImportError('libgradpath ABI version mismatch')

    def _checked(self, name, fn):
        last_error = self.cdll.gp_last_error
        counted = name in KERNEL_FUNCS

        def call(*args):
            if counted:
                self.launches += 1
            rc = fn(*args)
            if rc != 0:
                last_error.restype = c_char_p
                msg = last_error()
                raise GradpathError(name, rc, msg.decode('utf-8', 'replace') if msg else '')
            return rc
        call.__name__ = name
        return call


_lib = None


def load():
    """Load ``libgradpath.so`` (once).  Raises ImportError when it has not been
    built -- there is no CPU or CuPy fallback on this path."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(
                'libgradpath.so not found at {}: build it with '
                '`python -c "import __graft_entry__ as g; g.build()"` or '
                '`make -C chainer_b200/csrc`.  chainer_b200 has no fallback path.'.format(path))
        _lib = _Lib(path)
        # wall-time bound of the in-kernel waits for peer GPUs (default 30 minutes; 0: none)
        t = os.environ.get('CHAINER_B200_PEER_TIMEOUT_S')
        if t:
            _lib.gp_set_tuning(b'peer_timeout_s', int(float(t)))
    return _lib


def get():
    return load()


def set_backend_for_testing(obj):
    """Replace the loaded library by `obj` (TESTS ONLY: lets the host logic be exercised on
    machines without a GPU against an oracle-backed double that lives under tests/).
    Returns the previous backend.  Refused unless the process is a pytest run or sets
    CHAINER_B200_TEST_BACKEND=1 explicitly: a product process can never end up on a double."""
    import sys
    if 'pytest' not in sys.modules and os.environ.get('CHAINER_B200_TEST_BACKEND') != '1':
        raise RuntimeError('set_backend_for_testing is test machinery: run under pytest or set '
                           'CHAINER_B200_TEST_BACKEND=1')
    global _lib
    prev = _lib
    _lib = obj
    return prev


class nvtx_range(object):
    """``with nvtx_range('pack'):`` -- an NVTX range around the enclosed launches when
    CHAINER_B200_NVTX=1 (a no-op context otherwise, and without a profiler attached)."""
    enabled = os.environ.get('CHAINER_B200_NVTX', '0') not in ('0', '', 'false', 'False')

    def __init__(self, name):
        self.name = name.encode() if isinstance(name, str) else name

    def __enter__(self):
        if nvtx_range.enabled:
            get().gp_nvtx_push(self.name)
        return self

    def __exit__(self, *exc):
        if nvtx_range.enabled:
            get().gp_nvtx_pop()
        return False


def find_libnccl():
    """Path of libnccl.so.2: $CHAINER_B200_LIBNCCL, the `nvidia-nccl` wheel that
    PyTorch also uses, or the loader's default."""
    env = os.environ.get('CHAINER_B200_LIBNCCL')
    if env:
        return env
    try:
        import nvidia.nccl
        for base in list(nvidia.nccl.__path__):
            cand = os.path.join(base, 'lib', 'libnccl.so.2')
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return 'libnccl.so.2'
