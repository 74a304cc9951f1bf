"""Host-memory double of ``libgradpath`` for the CPU test-suite (TEST
INFRASTRUCTURE).

It exposes the same ``gp_*`` methods as ``chainer_b200._lib._Lib`` but executes
them with the NumPy oracle on HOST pointers, and implements the ``gp_nccl_*``
calls on a ``torch.distributed`` gloo group.  Installing it with
``_lib.set_backend_for_testing`` lets the host logic of the product (parameter
ordering, table building, bucketing, the optimizer protocol, the control
plane) run without a GPU, including with world_size 2.  The product never
imports this module; on a GPU box the real library is the only backend.
"""
import ctypes

import numpy as np

from chainer_b200 import _lib
from oracle import gradpath as og

_ID2DT = {6: np.dtype(np.float16), 7: np.dtype(np.float32), 8: np.dtype(np.float64), 9: og.BF16}


def _view(ptr, count, dtype):
    """NumPy view of `count` elements of `dtype` at host address `ptr`."""
    if count == 0:
        return np.zeros(0, dtype=dtype)
    dtype = np.dtype(dtype)
    buf = (ctypes.c_ubyte * (count * dtype.itemsize)).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype, count=count)


def _buf_view(ptr, count, dtype_id):
    """View of a packed buffer; bfloat16 buffers are viewed as uint16."""
    if dtype_id == 9:
        return _view(ptr, count, np.uint16)
    return _view(ptr, count, _ID2DT[dtype_id])


def _buf_read(ptr, count, dtype_id):
    v = _buf_view(ptr, count, dtype_id)
    if dtype_id == 9:
        return (v.astype(np.uint32) << 16).view(np.float32)
    return v


def _buf_write(ptr, count, dtype_id, values):
    v = _buf_view(ptr, count, dtype_id)
    if dtype_id == 9:
        v[...] = (np.ascontiguousarray(values, dtype=np.float32).view(np.uint32) >> 16).astype(np.uint16)
    else:
        v[...] = values


class FakeLib(object):
    accepts_host_pointers = True
    path = '<oracle-backed host double>'

    def __init__(self):
        self._allocs = {}
        self._tables = {}
        self._next = 1
        self.calls = []           # (name, args) log for assertions
        self._nccl = {}
        self.tuning = {}

    # ---------------------------------------------------------------- misc --
    def gp_abi_version(self):
        return 1

    def gp_last_error(self):
        return b''

    def gp_nvtx_push(self, name):
        self.calls.append(('nvtx_push', (name,)))
        return 0

    def gp_nvtx_pop(self):
        self.calls.append(('nvtx_pop', ()))
        return 0

    def gp_set_tuning(self, key, value):
        self.tuning[key] = value
        return 0

    def gp_get_tuning(self, key, out):
        out._obj.value = self.tuning.get(key, 0)
        return 0

    def gp_device_count(self, out):
        out._obj.value = 1
        return 0

    def gp_set_device(self, d):
        return 0

    def gp_get_device(self, out):
        out._obj.value = 0
        return 0

    def gp_device_synchronize(self):
        return 0

    def gp_device_sm_count(self, out):
        out._obj.value = 148
        return 0

    # -------------------------------------------------------------- memory --
    def gp_malloc(self, out, nbytes):
        buf = ctypes.create_string_buffer(max(int(nbytes), 1) + 64)
        addr = (ctypes.addressof(buf) + 63) // 64 * 64
        self._allocs[addr] = buf
        out._obj.value = addr
        return 0

    def gp_free(self, ptr):
        self._allocs.pop(int(ptr) if ptr else 0, None)
        return 0

    gp_malloc_host = gp_malloc
    gp_free_host = gp_free

    def gp_memcpy_async(self, dst, src, nbytes, kind, stream):
        ctypes.memmove(int(dst), int(src), int(nbytes))
        return 0

    def gp_memset_async(self, dst, value, nbytes, stream):
        ctypes.memset(int(dst), value, int(nbytes))
        return 0

    def _handle(self, out):
        out._obj.value = self._next
        self._next += 1
        return 0

    def gp_stream_create(self, out, non_blocking):
        return self._handle(out)

    def gp_stream_destroy(self, s):
        return 0

    def gp_stream_synchronize(self, s):
        return 0

    def gp_stream_wait_event(self, s, e):
        self.calls.append(('wait_event', (s, e)))
        return 0

    def gp_event_create(self, out, timing):
        return self._handle(out)

    def gp_event_destroy(self, e):
        return 0

    def gp_event_record(self, e, s):
        self.calls.append(('record', (e, s)))
        return 0

    def gp_event_synchronize(self, e):
        return 0

    def gp_event_elapsed_ms(self, out, a, b):
        out._obj.value = 0.0
        return 0

    # --------------------------------------------------------------- table --
    def gp_table_create(self, out):
        self._handle(out)
        self._tables[out._obj.value] = []
        return 0

    def gp_table_destroy(self, t):
        self._tables.pop(t, None)
        return 0

    def gp_table_upload(self, t, host_src, nbytes, stream, out):
        buf = ctypes.create_string_buffer(int(nbytes) + 256)
        addr = (ctypes.addressof(buf) + 255) // 256 * 256
        ctypes.memmove(addr, int(host_src), int(nbytes))
        ring = self._tables.setdefault(t, [])
        ring.append(buf)
        if len(ring) > 4:
            ring.pop(0)
        out._obj.value = addr
        return 0

    # ------------------------------------------------------------- kernels --
    def _tables_of(self, d_csum, d_segs, n):
        csum = _view(d_csum, n + 1, np.int64)
        segs = _view(d_segs, n, _lib.SEG_DTYPE)
        return csum, segs

    def _pieces(self, csum, n, begin, end):
        """(j, e0, e1) for every segment intersecting work range [begin, end)."""
        for j in range(n):
            lo, hi = max(int(csum[j]), begin), min(int(csum[j + 1]), end)
            if lo < hi:
                yield j, lo - int(csum[j]), hi - int(csum[j])

    def gp_pack(self, buffer, buf_dtype, d_csum, d_segs, n, begin, end, scale, layout_hint, stream):
        self.calls.append(('gp_pack', (buf_dtype, n, begin, end, scale)))
        assert begin % 4 == 0
        csum, segs = self._tables_of(d_csum, d_segs, n)
        for j, e0, e1 in self._pieces(csum, n, begin, end):
            src = _view(int(segs['ptr'][j, 0]), int(csum[j + 1] - csum[j]), _ID2DT[int(segs['dtype0'][j])])
            vals = og.pack([src[e0:e1]], _ID2DT[buf_dtype], scale)
            off = int(segs['buf_off'][j])
            isz = 2 if buf_dtype in (6, 9) else (4 if buf_dtype == 7 else 8)
            _buf_write(int(buffer) + (off + e0) * isz, e1 - e0, buf_dtype, vals)
        return 0

    def _mean_grad(self, buffer, buf_dtype, off, e0, e1, scale, gdt, seg=None, size=None):
        if not buffer:        # stand-alone update: the gradient array is the source
            src = _view(int(seg['ptr'][0]), size, _ID2DT[buf_dtype])[e0:e1]
            return og.scale_buffer(np.array(src), _ID2DT[buf_dtype], scale).astype(gdt)
        isz = 2 if buf_dtype in (6, 9) else (4 if buf_dtype == 7 else 8)
        raw = _buf_read(int(buffer) + (off + e0) * isz, e1 - e0, buf_dtype)
        return og.scale_buffer(np.array(raw), _ID2DT[buf_dtype], scale).astype(gdt)

    def gp_unpack_scale(self, buffer, buf_dtype, d_csum, d_segs, n, begin, end, scale, layout_hint,
                        stream):
        self.calls.append(('gp_unpack_scale', (buf_dtype, n, begin, end, scale)))
        csum, segs = self._tables_of(d_csum, d_segs, n)
        for j, e0, e1 in self._pieces(csum, n, begin, end):
            gdt = _ID2DT[int(segs['dtype0'][j])]
            dst = _view(int(segs['ptr'][j, 0]), int(csum[j + 1] - csum[j]), gdt)
            dst[e0:e1] = self._mean_grad(buffer, buf_dtype, int(segs['buf_off'][j]), e0, e1, scale, gdt)
        return 0

    def gp_unpack_momentum_sgd(self, buffer, buf_dtype, d_csum, d_segs, n, begin, end, scale, lr,
                               momentum, write_grad, layout_hint, stream):
        self.calls.append(('gp_unpack_momentum_sgd', (buf_dtype, n, begin, end, scale, lr, momentum,
                                                      write_grad)))
        csum, segs = self._tables_of(d_csum, d_segs, n)
        for j, e0, e1 in self._pieces(csum, n, begin, end):
            pdt = _ID2DT[int(segs['dtype1'][j])]
            size = int(csum[j + 1] - csum[j])
            g = self._mean_grad(buffer, buf_dtype, int(segs['buf_off'][j]), e0, e1, scale, pdt,
                                segs[j], size)
            p = _view(int(segs['ptr'][j, 1]), size, pdt)[e0:e1]
            v = _view(int(segs['ptr'][j, 2]), size, pdt)[e0:e1]
            og.momentum_sgd_update(p, g, v, lr, momentum)
            if write_grad:
                _view(int(segs['ptr'][j, 0]), size, pdt)[e0:e1] = g
        return 0

    def gp_unpack_adam(self, buffer, buf_dtype, d_csum, d_segs, n, begin, end, scale, alpha_t,
                       omb1, omb2, eps, eta, wd, lower, upper, flags, write_grad, layout_hint, stream):
        self.calls.append(('gp_unpack_adam', (buf_dtype, n, begin, end, scale, alpha_t, flags,
                                              write_grad)))
        csum, segs = self._tables_of(d_csum, d_segs, n)
        for j, e0, e1 in self._pieces(csum, n, begin, end):
            pdt = _ID2DT[int(segs['dtype1'][j])]
            size = int(csum[j + 1] - csum[j])
            g = self._mean_grad(buffer, buf_dtype, int(segs['buf_off'][j]), e0, e1, scale, pdt,
                                segs[j], size)
            p = _view(int(segs['ptr'][j, 1]), size, pdt)[e0:e1]
            m = _view(int(segs['ptr'][j, 2]), size, pdt)[e0:e1]
            v = _view(int(segs['ptr'][j, 3]), size, pdt)[e0:e1]
            vh = _view(int(segs['ptr'][j, 4]), size, pdt)[e0:e1] if flags & 1 else None
            _adam_kernel(p, g, m, v, vh, alpha_t, omb1, omb2, eps, eta, wd, lower, upper, flags)
            if write_grad:
                _view(int(segs['ptr'][j, 0]), size, pdt)[e0:e1] = g
        return 0

    # -- one-launch step (csrc/gp_step.cu): same results as the separate launches -----
    def gp_step_supported(self, n_ranks, buf_dtype, layout_hint, scale, adam_flags):
        # the double covers the one-rank step only (no peer memory on the host)
        return 1 if (n_ranks == 1 and buf_dtype in (6, 7, 9) and layout_hint == 7 and
                     not (adam_flags & 1) and float(scale) == 1.0) else 0

    def gp_step_tile_elems(self):
        return 16384

    def gp_step_words_bytes(self, cap):
        return cap * 36

    def gp_step_set_tuning(self, key, value):
        self.tuning['step_' + (key.decode() if isinstance(key, bytes) else key)] = value
        return 0

    def gp_step_momentum_sgd(self, comm, mc_ptr, buffer, buf_dtype, d_csum, d_segs, n, n_elems,
                             scale, lr, momentum, write_grad, layout_hint, stream):
        assert comm is None and mc_ptr is None
        self.calls.append(('gp_step_momentum_sgd', (buf_dtype, n, n_elems, scale, lr, momentum,
                                                    write_grad)))
        log = self.calls
        self.calls = []
        try:
            self.gp_pack(buffer, buf_dtype, d_csum, d_segs, n, 0, n_elems, 1.0, layout_hint, stream)
            self.gp_unpack_momentum_sgd(buffer, buf_dtype, d_csum, d_segs, n, 0, n_elems, scale, lr,
                                        momentum, write_grad, layout_hint, stream)
        finally:
            self.calls = log
        return 0

    def gp_step_adam(self, comm, mc_ptr, buffer, buf_dtype, d_csum, d_segs, n, n_elems, scale,
                     alpha_t, omb1, omb2, eps, eta, wd, lower, upper, flags, write_grad,
                     layout_hint, stream):
        assert comm is None and mc_ptr is None
        self.calls.append(('gp_step_adam', (buf_dtype, n, n_elems, scale, alpha_t, flags,
                                            write_grad)))
        log = self.calls
        self.calls = []
        try:
            self.gp_pack(buffer, buf_dtype, d_csum, d_segs, n, 0, n_elems, 1.0, layout_hint, stream)
            self.gp_unpack_adam(buffer, buf_dtype, d_csum, d_segs, n, 0, n_elems, scale, alpha_t,
                                omb1, omb2, eps, eta, wd, lower, upper, flags, write_grad,
                                layout_hint, stream)
        finally:
            self.calls = log
        return 0

    # -- hooks (csrc/gp_hooks.cu, gp_sgd_hooks.cu, gp_adam_hooks.cu) --------------
    @staticmethod
    def _apply_hooks(g, p, hooks_addr, pdt):
        """g (a fresh array in the parameter dtype) after clip-rate, decay and
        loss-scale division, each rounded to the parameter dtype."""
        h = _lib.GpHooks.from_address(int(hooks_addr))
        if h.clip_rate:
            rate = _view(h.clip_rate, 1, np.float32)[0]
            g *= pdt.type(rate)
        if h.weight_decay != 0.0:
            og.weight_decay_hook(p, g, h.weight_decay)
        if h.loss_scale != 0.0:
            og.loss_scale_divide(g, h.loss_scale)
        return g

    def gp_unpack_momentum_sgd_hooked(self, buffer, buf_dtype, d_csum, d_segs, n, begin, end,
                                      scale, lr, momentum, write_grad, layout_hint, hooks, stream):
        self.calls.append(('gp_unpack_momentum_sgd_hooked', (buf_dtype, n, begin, end, scale, lr,
                                                             momentum, write_grad)))
        csum, segs = self._tables_of(d_csum, d_segs, n)
        for j, e0, e1 in self._pieces(csum, n, begin, end):
            pdt = _ID2DT[int(segs['dtype1'][j])]
            size = int(csum[j + 1] - csum[j])
            g = self._mean_grad(buffer, buf_dtype, int(segs['buf_off'][j]), e0, e1, scale, pdt,
                                segs[j], size)
            p = _view(int(segs['ptr'][j, 1]), size, pdt)[e0:e1]
            v = _view(int(segs['ptr'][j, 2]), size, pdt)[e0:e1]
            g = self._apply_hooks(np.array(g), p, hooks, pdt)
            og.momentum_sgd_update(p, g, v, lr, momentum)
            if write_grad:
                _view(int(segs['ptr'][j, 0]), size, pdt)[e0:e1] = g
        return 0

    def gp_unpack_adam_hooked(self, buffer, buf_dtype, d_csum, d_segs, n, begin, end, scale,
                              alpha_t, omb1, omb2, eps, eta, wd, lower, upper, flags, write_grad,
                              layout_hint, hooks, stream):
        self.calls.append(('gp_unpack_adam_hooked', (buf_dtype, n, begin, end, scale, alpha_t, flags,
                                                     write_grad)))
        csum, segs = self._tables_of(d_csum, d_segs, n)
        for j, e0, e1 in self._pieces(csum, n, begin, end):
            pdt = _ID2DT[int(segs['dtype1'][j])]
            size = int(csum[j + 1] - csum[j])
            g = self._mean_grad(buffer, buf_dtype, int(segs['buf_off'][j]), e0, e1, scale, pdt,
                                segs[j], size)
            p = _view(int(segs['ptr'][j, 1]), size, pdt)[e0:e1]
            m = _view(int(segs['ptr'][j, 2]), size, pdt)[e0:e1]
            v = _view(int(segs['ptr'][j, 3]), size, pdt)[e0:e1]
            vh = _view(int(segs['ptr'][j, 4]), size, pdt)[e0:e1] if flags & 1 else None
            g = self._apply_hooks(np.array(g), p, hooks, pdt)
            _adam_kernel(p, g, m, v, vh, alpha_t, omb1, omb2, eps, eta, wd, lower, upper, flags)
            if write_grad:
                _view(int(segs['ptr'][j, 0]), size, pdt)[e0:e1] = g
        return 0

    # -- float32 master weights (csrc/gp_master.cu) ------------------------------------
    def _master(self, kind, buffer, buf_dtype, d_csum, d_segs, n, begin, end, scale, write_grad,
                hooks, d_skip, update):
        csum, segs = self._tables_of(d_csum, d_segs, n)
        skip = bool(d_skip) and int(_view(d_skip, 1, np.int32)[0]) != 0
        h = _lib.GpHooks.from_address(int(hooks)) if hooks else None
        f16, f32 = np.dtype(np.float16), np.dtype(np.float32)
        for j, e0, e1 in self._pieces(csum, n, begin, end):
            assert int(segs['dtype0'][j]) == 6 and int(segs['dtype1'][j]) == 7
            size = int(csum[j + 1] - csum[j])
            g16 = self._mean_grad(buffer, buf_dtype, int(segs['buf_off'][j]), e0, e1, scale, f16,
                                  segs[j], size)
            g16 = np.array(g16)
            gout = _view(int(segs['ptr'][j, 0]), size, f16)[e0:e1]
            if skip:
                if write_grad:
                    gout[...] = g16
                continue
            p32 = _view(int(segs['ptr'][j, 1]), size, f32)[e0:e1]
            p16 = _view(int(segs['ptr'][j, 4]), size, f16)[e0:e1]
            if h is not None and h.clip_rate:
                g16 *= f16.type(_view(h.clip_rate, 1, np.float32)[0])
            if h is not None and h.weight_decay != 0.0:
                og.weight_decay_hook(p32.astype(f16), g16, h.weight_decay)
            g32 = g16.astype(f32)
            if h is not None and h.loss_scale != 0.0:
                og.loss_scale_divide(g32, h.loss_scale)
            update(segs, j, size, e0, e1, p32, g32)
            p16[...] = p32.astype(f16)
            if write_grad:
                gout[...] = g16

    def gp_unpack_momentum_sgd_master(self, buffer, buf_dtype, d_csum, d_segs, n, begin, end, scale,
                                      lr, momentum, write_grad, hooks, d_skip, stream):
        self.calls.append(('gp_unpack_momentum_sgd_master', (buf_dtype, n, begin, end, scale, lr,
                                                             momentum, write_grad)))

        def update(segs, j, size, e0, e1, p32, g32):
            v = _view(int(segs['ptr'][j, 2]), size, np.float32)[e0:e1]
            og.momentum_sgd_update(p32, g32, v, lr, momentum)
        self._master('sgd', buffer, buf_dtype, d_csum, d_segs, n, begin, end, scale, write_grad, hooks,
                     d_skip, update)
        return 0

    def gp_unpack_adam_master(self, buffer, buf_dtype, d_csum, d_segs, n, begin, end, scale, alpha_t,
                              omb1, omb2, eps, eta, wd, lower, upper, flags, write_grad, hooks,
                              d_skip, stream):
        self.calls.append(('gp_unpack_adam_master', (buf_dtype, n, begin, end, scale, alpha_t, flags,
                                                     write_grad)))

        def update(segs, j, size, e0, e1, p32, g32):
            m = _view(int(segs['ptr'][j, 2]), size, np.float32)[e0:e1]
            v = _view(int(segs['ptr'][j, 3]), size, np.float32)[e0:e1]
            _adam_kernel(p32, g32, m, v, None, alpha_t, omb1, omb2, eps, eta, wd, lower, upper, flags)
        self._master('adam', buffer, buf_dtype, d_csum, d_segs, n, begin, end, scale, write_grad, hooks,
                     d_skip, update)
        return 0

    def gp_unpack_sgd_family(self, buffer, buf_dtype, d_csum, d_segs, n, begin, end, scale, rule,
                             lr, momentum, write_grad, layout_hint, hooks, stream):
        self.calls.append(('gp_unpack_sgd_family', (buf_dtype, n, begin, end, scale, rule, lr,
                                                    momentum, write_grad)))
        csum, segs = self._tables_of(d_csum, d_segs, n)
        for j, e0, e1 in self._pieces(csum, n, begin, end):
            pdt = _ID2DT[int(segs['dtype1'][j])]
            size = int(csum[j + 1] - csum[j])
            g = self._mean_grad(buffer, buf_dtype, int(segs['buf_off'][j]), e0, e1, scale, pdt,
                                segs[j], size)
            p = _view(int(segs['ptr'][j, 1]), size, pdt)[e0:e1]
            g = np.array(g)
            if hooks:
                g = self._apply_hooks(g, p, hooks, pdt)
            if rule == 0:
                og.sgd_update(p, g, lr)
            else:
                v = _view(int(segs['ptr'][j, 2]), size, pdt)[e0:e1]
                if rule == 1:
                    og.corrected_momentum_sgd_update(p, g, v, lr, momentum)
                else:
                    og.nesterov_ag_update(p, g, v, lr, momentum)
            if write_grad:
                _view(int(segs['ptr'][j, 0]), size, pdt)[e0:e1] = g
        return 0

    def gp_sqnorm_workspace_bytes(self):
        return 64

    def gp_sqnorm(self, x, dtype, n, scale, accumulate, threshold, ws, out, stream):
        self.calls.append(('gp_sqnorm', (dtype, n, scale, accumulate, threshold)))
        vals = og.scale_buffer(np.array(_buf_read(x, n, dtype)), _ID2DT[dtype], scale)
        sq = float(np.sum(np.asarray(vals, dtype=np.float64) ** 2))
        o_d = _view(out, 1, np.float64)
        o_f = _view(int(out) + 8, 2, np.float32)
        total = (float(o_d[0]) if accumulate else 0.0) + sq
        o_d[0] = total
        norm = np.float32(np.sqrt(total))
        with np.errstate(divide='ignore'):
            o_f[0] = min(np.float32(threshold) / norm, np.float32(1)) if threshold > 0 else 1.0
        o_f[1] = norm
        return 0

    def gp_scale_by_device(self, x, dtype, n, d_factor, stream):
        self.calls.append(('gp_scale_by_device', (dtype, n)))
        v = _view(x, n, _ID2DT[dtype])
        v *= _ID2DT[dtype].type(_view(d_factor, 1, np.float32)[0])
        return 0

    def gp_weight_decay(self, grad, param, dtype, n, decay, stream):
        self.calls.append(('gp_weight_decay', (dtype, n, decay)))
        og.weight_decay_hook(_view(param, n, _ID2DT[dtype]), _view(grad, n, _ID2DT[dtype]), decay)
        return 0

    def gp_divide(self, x, dtype, n, divisor, stream):
        self.calls.append(('gp_divide', (dtype, n, divisor)))
        og.loss_scale_divide(_view(x, n, _ID2DT[dtype]), divisor)
        return 0

    def gp_scale(self, buffer, dtype, n, scale, stream):
        self.calls.append(('gp_scale', (dtype, n, scale)))
        vals = og.scale_buffer(np.array(_buf_read(buffer, n, dtype)), _ID2DT[dtype], scale)
        _buf_write(buffer, n, dtype, vals)
        return 0

    def gp_check_finite(self, buffer, dtype, n, d_flag, stream):
        self.calls.append(('gp_check_finite', (dtype, n)))
        if not np.isfinite(_buf_read(buffer, n, dtype)).all():
            _view(d_flag, 1, np.int32)[0] |= 1
        return 0

    # ------------------------------------------------------------------ BN --
    def gp_bn_workspace_bytes(self, C):
        return 64

    def gp_bn_workspace_layout(self, C, out4):
        out4[0], out4[1], out4[2], out4[3] = 0, 32, 32, 64
        return 0

    def gp_bn_fwd_stats(self, x, x_dtype, N, C, HW, out, out_dtype, ws, stream):
        self.calls.append(('gp_bn_fwd_stats', (N, C, HW)))
        xs = _view(x, N * C * HW, _ID2DT[x_dtype]).reshape(N, C, HW)
        _view(out, 2 * C, _ID2DT[out_dtype])[...] = og.bn_fwd_stats(xs, _ID2DT[out_dtype])
        return 0

    def gp_bn_fwd_mean_var(self, x, x_dtype, N, C, HW, out, out_dtype, ws, stream):
        self.gp_bn_fwd_stats(x, x_dtype, N, C, HW, out, out_dtype, ws, stream)
        return self.gp_bn_finish_mean_var(out, out_dtype, C, 1.0, int(out) + C * (2 if out_dtype == 6 else (4 if out_dtype == 7 else 8)), stream)

    def gp_bn_bwd_stats(self, gy, gy_dtype, xh, x_dtype, mean, inv_std, stat_dtype, N, C, HW, out,
                        out_dtype, ws, stream):
        self.calls.append(('gp_bn_bwd_stats', (N, C, HW)))
        g = _view(gy, N * C * HW, _ID2DT[gy_dtype]).reshape(N, C, HW)
        x = _view(xh, N * C * HW, _ID2DT[x_dtype]).reshape(N, C, HW)
        if mean and inv_std:
            x = og.x_hat(x, _view(mean, C, _ID2DT[stat_dtype]), _view(inv_std, C, _ID2DT[stat_dtype]))
        _view(out, 2 * C, _ID2DT[out_dtype])[...] = og.bn_bwd_stats(g, x, _ID2DT[out_dtype])
        return 0

    def gp_bn_finish_mean_var(self, buf, dtype, C, scale, out_var, stream):
        dt = _ID2DT[dtype]
        s = og.scale_buffer(np.array(_view(buf, 2 * C, dt)), dt, scale)
        var = s[C:] - np.square(s[:C])
        _view(buf, 2 * C, dt)[...] = s
        _view(out_var, C, dt)[...] = var
        return 0

    def gp_bn_fwd_apply(self, x, x_dtype, N, C, HW, mean, var, gamma, beta, stat_dtype, eps, y,
                        inv_std_out, running_mean, running_var, running_dtype, decay, adjust, stream):
        self.calls.append(('gp_bn_fwd_apply', (N, C, HW)))
        xdt, sdt = _ID2DT[x_dtype], _ID2DT[stat_dtype]
        xs = _view(x, N * C * HW, xdt).reshape(N, C, HW)
        m, v = _view(mean, C, sdt), _view(var, C, sdt)
        _view(y, N * C * HW, xdt).reshape(N, C, HW)[...] = og.bn_fwd_apply(
            xs, m, v, _view(gamma, C, sdt), _view(beta, C, sdt), eps)
        if inv_std_out:
            _view(inv_std_out, C, sdt)[...] = og.bn_inv_std(np.array(v), eps)
        if running_mean:
            rdt = _ID2DT[running_dtype]
            og.bn_running_update(_view(running_mean, C, rdt), _view(running_var, C, rdt),
                                 np.array(m), np.array(v), decay, adjust)
        return 0

    def gp_bn_bwd_apply(self, gy, gy_dtype, x, x_dtype, N, C, HW, mean, inv_std, gamma, ggamma, gbeta,
                        stat_dtype, inv_m, gx, stream):
        self.calls.append(('gp_bn_bwd_apply', (N, C, HW)))
        xdt, sdt = _ID2DT[x_dtype], _ID2DT[stat_dtype]
        g = _view(gy, N * C * HW, xdt).reshape(N, C, HW)
        xs = _view(x, N * C * HW, xdt).reshape(N, C, HW)
        _view(gx, N * C * HW, xdt).reshape(N, C, HW)[...] = og.bn_bwd_apply(
            g, xs, _view(mean, C, sdt), _view(inv_std, C, sdt), _view(gamma, C, sdt),
            _view(ggamma, C, sdt), _view(gbeta, C, sdt), inv_m)
        return 0

    # ---------------------------------------------------------------- NCCL --
    def gp_nccl_load(self, path):
        return 0

    def gp_nccl_version(self, out):
        out._obj.value = 22809
        return 0

    def gp_nccl_get_unique_id(self, buf):
        buf.raw = b'fake-nccl-id'.ljust(128, b'\0')
        return 0

    def gp_nccl_comm_init_rank(self, out, n, uid, rank):
        import torch.distributed as dist
        assert dist.is_initialized() and dist.get_world_size() == n and dist.get_rank() == rank
        return self._handle(out)

    def gp_nccl_comm_destroy(self, comm):
        return 0

    def gp_nccl_allreduce(self, comm, send, recv, count, dtype, op, stream):
        import torch
        import torch.distributed as dist
        self.calls.append(('gp_nccl_allreduce', (count, dtype)))
        vals = np.array(_buf_read(send, count, dtype))
        t = torch.from_numpy(vals)
        if dtype == 6:          # sum in float16 like NCCL would
            t = t.to(torch.float16)
        dist.all_reduce(t)
        _buf_write(recv, count, dtype, t.numpy() if dtype != 9 else og.bf16_round(t.numpy()))
        return 0

    def gp_nccl_bcast(self, comm, buf, count, dtype, root, stream):
        import torch
        import torch.distributed as dist
        self.calls.append(('gp_nccl_bcast', (count, dtype, root)))
        t = torch.from_numpy(np.array(_buf_read(buf, count, dtype)))
        dist.broadcast(t, src=root)
        _buf_write(buf, count, dtype, t.numpy())
        return 0

    def gp_nccl_reduce(self, comm, send, recv, count, dtype, op, root, stream):
        return self.gp_nccl_allreduce(comm, send, recv, count, dtype, op, stream)

    def gp_p2p_flag_bytes(self):
        return 96

    def gp_p2p_small_bytes(self, n, cap):
        return 2 * n * cap * 4

    def gp_nccl_group_start(self):
        return 0

    def gp_nccl_group_end(self):
        return 0


def _adam_kernel(p, g, m, v, vh, alpha_t, omb1, omb2, eps, eta, wd, lower, upper, flags):
    """The C-ABI takes alpha_t / bounds pre-computed: same arithmetic as
    oracle.adam_update_gpu with those scalars given."""
    P = p.dtype
    T = np.float32 if P == np.float16 else P.type
    g_, m_, v_ = g.astype(T), m.astype(T), v.astype(T)
    m_ = m_ + T(omb1) * (g_ - m_)
    v_ = v_ + T(omb2) * (g_ * g_ - v_)
    if flags & 1:
        vh_ = np.maximum(vh.astype(T), v_)
        vh[...] = vh_.astype(P)
        d_ = vh_
    else:
        d_ = v_
    m[...] = m_.astype(P)
    v[...] = v_.astype(P)
    denom = np.sqrt(d_) + T(eps)
    if flags & 2:
        step = np.maximum(np.minimum(T(alpha_t) / denom, T(upper)), T(lower)) * m_
    else:
        step = T(alpha_t) * m_ / denom
    p_ = p.astype(T)
    p[...] = (p_ - T(eta) * (step + T(wd) * p_)).astype(P)


def install():
    """Install the double; returns (fake, previous_backend)."""
    fake = FakeLib()
    prev = _lib.set_backend_for_testing(fake)
    return fake, prev


def uninstall(prev):
    _lib.set_backend_for_testing(prev)
