// gp_sgd_hooks.cu -- gp_unpack_momentum_sgd with the pre-update gradient hooks fused
// in (gradient clipping rate, weight decay, loss-scale division; gp_common.cuh
// HookArgs).  A separate translation unit so that the plain kernels of gp_sgd.cu
// are byte-for-byte what they were and both compile in parallel.
//
// Reference being replaced: the hook kernels of chainer/optimizer_hooks/
// weight_decay.py:44-57 and gradient_clipping.py:84-106 (one launch per parameter
// each, before the per-parameter update launches) and `grad /= loss_scale`
// (chainer/optimizer.py:289-291).
#include "gp_sgd_op.cuh"

extern "C" int gp_unpack_momentum_sgd_hooked(const void* buffer, int buf_dtype,
                                             const int64_t* d_csum, const gp_seg_t* d_segs,
                                             int n_segs, int64_t elem_begin, int64_t elem_end,
                                             double scale, double lr, double momentum,
                                             int write_grad, int layout_hint,
                                             const gp_hooks_t* hooks, void* stream) {
  if (!hooks) {
    gp_set_error("gp_unpack_momentum_sgd_hooked: hooks is NULL");
    return GP_EINVAL;
  }
  const HookArgs h = {hooks->clip_rate, hooks->weight_decay, hooks->loss_scale};
  return launch_sgd<true>(buffer, buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, scale,
                          lr, momentum, write_grad, layout_hint, h, stream,
                          "gp_unpack_momentum_sgd_hooked");
}
