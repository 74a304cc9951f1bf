// gp_adam_hooks.cu -- gp_unpack_adam with the pre-update gradient hooks fused in
// (see gp_sgd_hooks.cu and gp_common.cuh HookArgs).
#include "gp_adam_op.cuh"

extern "C" int gp_unpack_adam_hooked(const void* buffer, int buf_dtype, const int64_t* d_csum,
                                     const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                                     int64_t elem_end, double scale, double alpha_t,
                                     double one_minus_beta1, double one_minus_beta2, double eps,
                                     double eta, double weight_decay_rate, double lower,
                                     double upper, int adam_flags, int write_grad,
                                     int layout_hint, const gp_hooks_t* hooks, void* stream) {
  if (!hooks) {
    gp_set_error("gp_unpack_adam_hooked: hooks is NULL");
    return GP_EINVAL;
  }
  const HookArgs h = {hooks->clip_rate, hooks->weight_decay, hooks->loss_scale};
  return launch_adam<true>(buffer, buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, scale,
                           alpha_t, one_minus_beta1, one_minus_beta2, eps, eta, weight_decay_rate,
                           lower, upper, adam_flags, write_grad, layout_hint, h, stream,
                           "gp_unpack_adam_hooked");
}
