"""Optimizer hooks of the gradient path: ``WeightDecay`` and ``GradientClipping``
(``chainer/optimizer_hooks/weight_decay.py``, ``gradient_clipping.py``).

Registered with ``optimizer.add_hook(...)`` exactly as in the reference.  Behind
``create_multi_node_optimizer`` the hook lists ``[GradientClipping]``,
``[WeightDecay]`` and ``[GradientClipping, WeightDecay]`` are FUSED into the
update kernel (``csrc/gp_sgd_hooks.cu``, ``gp_adam_hooks.cu``; the global norm is
one reduction over the allreduced packed buffer, ``csrc/gp_hooks.cu``), so the
``__call__`` methods below only run on the unfused path: a stand-alone
``optimizer.update()``, other hook orders, or next to custom hooks.  They are
device kernels too -- there is no host arithmetic on gradients anywhere.
"""
import ctypes

from chainer_b200 import _lib
from chainer_b200 import device as _dev


class WeightDecay(object):
    """``g += rate * p`` for every parameter (``weight_decay.py:6-57``)."""
    name = 'WeightDecay'
    call_for_each_param = True
    timing = 'pre'

    def __init__(self, rate):
        self.rate = rate

    def __call__(self, rule, param):
        p, g = param.data, param.grad
        if p is None or g is None:
            return
        rate = self.rate
        loss_scale = getattr(param, '_loss_scale', None)
        if loss_scale is not None:
            rate *= loss_scale
        dt = _dev.array_dtype(g)
        if _dev.array_dtype(p) != dt:
            raise ValueError('WeightDecay: parameter and gradient dtypes differ')
        _lib.get().gp_weight_decay(_dev.device_ptr(g), _dev.device_ptr(p), _dev.dtype_id(dt),
                                   _dev.array_size(g), float(rate), 0)


class _NormScratch(object):
    """Device scratch of the norm reduction: workspace + the 16-byte result
    ``{double sqsum; float rate; float norm}``."""

    def __init__(self):
        lib = _lib.get()
        nbytes = lib.gp_sqnorm_workspace_bytes()
        self._ws = _dev._Allocation(nbytes)
        self._out = _dev._Allocation(16)
        lib.gp_memset_async(self._ws.ptr, 0, nbytes, 0)
        lib.gp_memset_async(self._out.ptr, 0, 16, 0)
        self.ws = self._ws.ptr
        self.out = self._out.ptr
        self.rate_ptr = self._out.ptr + 8


class GradientClipping(object):
    """Scales all gradients so that their global L2 norm is at most ``threshold``
    (``gradient_clipping.py:55-106``, GPU branch: ``rate = (threshold /
    norm).clip(None, 1)`` stays on the device, no synchronisation)."""
    name = 'GradientClipping'
    timing = 'pre'

    def __init__(self, threshold):
        self.threshold = threshold
        self._scratch = None

    def scratch(self):
        if self._scratch is None:
            self._scratch = _NormScratch()
        return self._scratch

    def __call__(self, opt):
        lib = _lib.get()
        sc = self.scratch()
        grads = [p.grad for p in opt.target.params(False) if p.grad is not None]
        if not grads:
            return
        for i, g in enumerate(grads):
            lib.gp_sqnorm(_dev.device_ptr(g), _dev.dtype_id(_dev.array_dtype(g)),
                          _dev.array_size(g), 1.0, 1 if i else 0, float(self.threshold),
                          sc.ws, sc.out, 0)
        for g in grads:
            lib.gp_scale_by_device(_dev.device_ptr(g), _dev.dtype_id(_dev.array_dtype(g)),
                                   _dev.array_size(g), sc.rate_ptr, 0)

    def last_norm(self):
        """L2 norm seen by the most recent call (reads 16 bytes back; diagnostics)."""
        sc = self.scratch()
        host = (ctypes.c_char * 16)()
        lib = _lib.get()
        lib.gp_memcpy_async(ctypes.addressof(host), sc.out, 16, 1, 0)
        lib.gp_stream_synchronize(0)
        import struct
        sqsum, rate, norm = struct.unpack('dff', bytes(host))
        return norm
