// gp_master.cu -- fused unpack + descale + update for float16 parameters with FLOAT32 MASTER
// WEIGHTS, with the pre-update hooks, the loss-scale division and the dynamic-loss-scaling
// skip fused in.
//
// Reference being replaced (chainer v7.8.1), per parameter and per step:
//   chainer/optimizer.py:262-282  fp32_param = param.astype(float32) (kept across steps),
//                                 fp32_param.grad = param.grad.astype(float32)
//   chainer/optimizer.py:286-291  fp32_param.grad /= loss_scale
//   chainer/optimizers/momentum_sgd.py:75-88 | adam.py:224-332   update_core_gpu on the
//                                 float32 copy (states are float32)
//   chainer/optimizer.py:297-305  param.array = fp32_param.array.astype(float16)
//   chainer/optimizer.py:763-779  dynamic loss scaling: check_nan_in_grads() reads every
//                                 gradient back to decide is_safe_to_update()
// (3 casts + the update per parameter, 4 n launches, one host synchronisation per parameter
// for the NaN check), after K3 div_by_size and K2 unpack have made two more passes.
//
// Here ONE launch covers the parameter list.  Segment table: ptr[0] the float16 gradient
// (written back when write_grad), ptr[1] the float32 master, ptr[2..3] the float32 states,
// ptr[4] the float16 parameter (dtype0 = float16, dtype1 = float32).  Per element:
//   g16 = (half)( (B)(x * 1/N) )            the mean gradient as multi_node_mean_grad leaves it
//   g16 = g16 * rate; g16 += decay * p16    optimizer-level hooks act on the float16 arrays
//   g32 = (float)g16 / loss_scale           float32, as UpdateRule.update does
//   update(p32, states, g32);  p16 = (half)p32
// `d_skip` (may be NULL): a device word written by gp_check_finite over the reduced buffer;
// non-zero means a non-finite gradient somewhere -> nothing is updated (only the mean gradient
// is written back), which is `is_safe_to_update()` without a host round trip.
// Algorithmic bytes per element: b (buffer) + 8 (master r/w) + 8 per state + 2 (p16) [+2 grad].
#include "gp_sgd_op.cuh"
#include "gp_adam_op.cuh"

namespace {

struct MasterHooks {
  const float* clip_rate;   // device pointer (gp_sqnorm output) or nullptr
  double decay;             // 0: off (already multiplied by the loss scale, like HookArgs)
  double loss_scale;        // 0: off
  const int32_t* skip;      // device pointer or nullptr
};

// g16 (a half value carried in float) after the optimizer-level hooks, in float16 arithmetic
// (chainer/optimizer_hooks: the hooks see param.grad / param.array, both float16)
struct HookState {
  float rate, decay, ls;
  bool use_rate, use_decay, use_ls, skip;
  __device__ __forceinline__ explicit HookState(const MasterHooks& h) {
    using A = Arith<__half>;
    use_rate = h.clip_rate != nullptr;
    rate = use_rate ? A::cst((double)__ldg(h.clip_rate)) : 1.f;
    use_decay = h.decay != 0.0;
    decay = A::cst(h.decay);
    use_ls = h.loss_scale != 0.0;
    ls = (float)h.loss_scale;
    skip = h.skip != nullptr && __ldg(h.skip) != 0;
  }
  __device__ __forceinline__ float hooks16(float g16, float p32) const {
    using A = Arith<__half>;
    if (use_rate) g16 = A::mul(g16, rate);
    if (use_decay) g16 = A::add(g16, A::mul(decay, A::r(p32)));   // p16 == (half)p32
    return g16;
  }
  __device__ __forceinline__ float to32(float g16) const {
    return use_ls ? __fdiv_rn(g16, ls) : g16;
  }
};

template <class P> __device__ __forceinline__ void st_half4(__half* p, const float (&x)[4]) {
  st4(p, pack4<__half, float>(x));
}

// ------------------------------------------------------------ MomentumSGD --
struct SgdMasterOp {
  static constexpr int kMaxUnroll = 2;
  static constexpr int kDefaultUnroll = 2;
  const void* buffer;
  ScaleArg s;
  double lr, momentum;
  int write_grad;
  MasterHooks hooks;

  static __device__ __forceinline__ int key(const gp_seg_t& g) { return g.dtype1; }
  template <class B>
  __device__ __forceinline__ const B* grad_src(const gp_seg_t& g, int64_t e) const {
    return buffer ? reinterpret_cast<const B*>(buffer) + g.buf_off + e
                  : reinterpret_cast<const B*>(g.ptr[0]) + e;
  }

  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void vec(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                      const bool (&act)[U]) const {
    using CB = typename Carrier<B>::type;
    Raw4<B> rb[U];
    Raw4<float> rp[U], rv[U];
    float *pp[U], *pv[U];
    __half *p16[U], *g16[U];
    const HookState hk(hooks);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      pp[u] = mptr<float>(seg[u]->ptr[1]) + e[u];
      pv[u] = mptr<float>(seg[u]->ptr[2]) + e[u];
      p16[u] = mptr<__half>(seg[u]->ptr[4]) + e[u];
      g16[u] = mptr<__half>(seg[u]->ptr[0]) + e[u];
      rb[u] = ld4_stream(grad_src<B>(*seg[u], e[u]));
      if (!hk.skip) {
        rp[u] = ld4(pp[u]);
        rv[u] = ld4(pv[u]);
      }
    }
    const float lr_ = (float)lr, mom_ = (float)momentum;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      CB xb[4];
      float g[4], p[4], v[4];
      unpack4(rb[u], xb);
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = gpw::mean_grad_value<B, __half, SM>(xb[i], s);
      if (hk.skip) {
        if (write_grad) st_half4<float>(g16[u], g);
        continue;
      }
      unpack4(rp[u], p);
      unpack4(rv[u], v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i] = hk.hooks16(g[i], p[i]);
        const float g32 = hk.to32(g[i]);
        SgdOp<false>::math<float>(g32, p[i], v[i], lr_, mom_);
        // what param.grad holds afterwards in the reference: the float16 array after the
        // hooks (the division happened on the float32 copy)
      }
      st4(pp[u], pack4<float, float>(p));
      st4(pv[u], pack4<float, float>(v));
      st_half4<float>(p16[u], p);
      if (write_grad) st_half4<float>(g16[u], g);
    }
  }

  template <class B, class P, int SM>
  __device__ __forceinline__ void one(const gp_seg_t& sg, int64_t e) const {
    const HookState hk(hooks);
    float g = gpw::mean_grad_value<B, __half, SM>(to_carrier(*grad_src<B>(sg, e)), s);
    __half* g16 = mptr<__half>(sg.ptr[0]) + e;
    if (hk.skip) {
      if (write_grad) *g16 = __float2half_rn(g);
      return;
    }
    float* pp = mptr<float>(sg.ptr[1]) + e;
    float* pv = mptr<float>(sg.ptr[2]) + e;
    float p = *pp, v = *pv;
    g = hk.hooks16(g, p);
    SgdOp<false>::math<float>(hk.to32(g), p, v, (float)lr, (float)momentum);
    *pp = p;
    *pv = v;
    mptr<__half>(sg.ptr[4])[e] = __float2half_rn(p);
    if (write_grad) *g16 = __float2half_rn(g);
  }
};

// ------------------------------------------------------------------- Adam --
struct AdamMasterOp {
  static constexpr int kMaxUnroll = 2;
  static constexpr int kDefaultUnroll = 2;
  const void* buffer;
  ScaleArg s;
  AdamOp<false, false> adam;   // the constants and math() of the plain op
  int write_grad;
  MasterHooks hooks;

  static __device__ __forceinline__ int key(const gp_seg_t& g) { return g.dtype1; }
  template <class B>
  __device__ __forceinline__ const B* grad_src(const gp_seg_t& g, int64_t e) const {
    return buffer ? reinterpret_cast<const B*>(buffer) + g.buf_off + e
                  : reinterpret_cast<const B*>(g.ptr[0]) + e;
  }

  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void vec(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                      const bool (&act)[U]) const {
    using CB = typename Carrier<B>::type;
    Raw4<B> rb[U];
    Raw4<float> rp[U], rm[U], rv[U];
    float *pp[U], *pm[U], *pv[U];
    __half *p16[U], *g16[U];
    const HookState hk(hooks);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      pp[u] = mptr<float>(seg[u]->ptr[1]) + e[u];
      pm[u] = mptr<float>(seg[u]->ptr[2]) + e[u];
      pv[u] = mptr<float>(seg[u]->ptr[3]) + e[u];
      p16[u] = mptr<__half>(seg[u]->ptr[4]) + e[u];
      g16[u] = mptr<__half>(seg[u]->ptr[0]) + e[u];
      rb[u] = ld4_stream(grad_src<B>(*seg[u], e[u]));
      if (!hk.skip) {
        rp[u] = ld4(pp[u]);
        rm[u] = ld4(pm[u]);
        rv[u] = ld4(pv[u]);
      }
    }
    const auto c = adam.consts<float>();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      CB xb[4];
      float g[4], p[4], m[4], v[4];
      unpack4(rb[u], xb);
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = gpw::mean_grad_value<B, __half, SM>(xb[i], s);
      if (hk.skip) {
        if (write_grad) st_half4<float>(g16[u], g);
        continue;
      }
      unpack4(rp[u], p);
      unpack4(rm[u], m);
      unpack4(rv[u], v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i] = hk.hooks16(g[i], p[i]);
        float vh = 0.f;
        adam.math<float, float>(hk.to32(g[i]), p[i], m[i], v[i], vh, c);
      }
      st4(pp[u], pack4<float, float>(p));
      st4(pm[u], pack4<float, float>(m));
      st4(pv[u], pack4<float, float>(v));
      st_half4<float>(p16[u], p);
      if (write_grad) st_half4<float>(g16[u], g);
    }
  }

  template <class B, class P, int SM>
  __device__ __forceinline__ void one(const gp_seg_t& sg, int64_t e) const {
    const HookState hk(hooks);
    float g = gpw::mean_grad_value<B, __half, SM>(to_carrier(*grad_src<B>(sg, e)), s);
    __half* g16 = mptr<__half>(sg.ptr[0]) + e;
    if (hk.skip) {
      if (write_grad) *g16 = __float2half_rn(g);
      return;
    }
    float* pp = mptr<float>(sg.ptr[1]) + e;
    float* pm = mptr<float>(sg.ptr[2]) + e;
    float* pv = mptr<float>(sg.ptr[3]) + e;
    float p = *pp, m = *pm, v = *pv, vh = 0.f;
    g = hk.hooks16(g, p);
    adam.math<float, float>(hk.to32(g), p, m, v, vh, adam.consts<float>());
    *pp = p;
    *pm = m;
    *pv = v;
    mptr<__half>(sg.ptr[4])[e] = __float2half_rn(p);
    if (write_grad) *g16 = __float2half_rn(g);
  }
};

MasterHooks master_hooks(const gp_hooks_t* h, const void* d_skip) {
  MasterHooks m = {nullptr, 0.0, 0.0, (const int32_t*)d_skip};
  if (h) {
    m.clip_rate = (const float*)h->clip_rate;
    m.decay = h->weight_decay;
    m.loss_scale = h->loss_scale;
  }
  return m;
}

template <class Op>
int launch_master(int buf_dtype, const int64_t* d_csum, const gp_seg_t* d_segs, int n_segs,
                  int64_t begin, int64_t end, const Op& op, void* stream, const char* what) {
  switch (buf_dtype) {
    case GP_F32: return gpw::launch_f32<Op, float>(d_csum, d_segs, n_segs, begin, end, op, stream, what);
    case GP_F16: return gpw::launch_f32<Op, __half>(d_csum, d_segs, n_segs, begin, end, op, stream, what);
    case GP_BF16:
      return gpw::launch_f32<Op, __nv_bfloat16>(d_csum, d_segs, n_segs, begin, end, op, stream, what);
    default:
      gp_set_error("%s: unsupported buffer dtype id %d (float32, float16, bfloat16)", what, buf_dtype);
      return GP_EINVAL;
  }
}

}  // namespace

extern "C" int gp_unpack_momentum_sgd_master(const void* buffer, int buf_dtype,
                                             const int64_t* d_csum, const gp_seg_t* d_segs,
                                             int n_segs, int64_t elem_begin, int64_t elem_end,
                                             double scale, double lr, double momentum,
                                             int write_grad, const gp_hooks_t* hooks,
                                             const void* d_skip, void* stream) {
  SgdMasterOp op;
  op.buffer = buffer;
  op.s = make_scale(scale);
  op.lr = lr;
  op.momentum = momentum;
  op.write_grad = write_grad;
  op.hooks = master_hooks(hooks, d_skip);
  return launch_master(buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, op, stream,
                       "gp_unpack_momentum_sgd_master");
}

extern "C" int gp_unpack_adam_master(const void* buffer, int buf_dtype, const int64_t* d_csum,
                                     const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                                     int64_t elem_end, double scale, double alpha_t,
                                     double one_minus_beta1, double one_minus_beta2, double eps,
                                     double eta, double weight_decay_rate, double lower,
                                     double upper, int adam_flags, int write_grad,
                                     const gp_hooks_t* hooks, const void* d_skip, void* stream) {
  if (adam_flags & GP_ADAM_AMSGRAD) {
    gp_set_error("gp_unpack_adam_master: AMSGrad is not covered (no table slot left for vhat)");
    return GP_EINVAL;
  }
  AdamMasterOp op;
  op.buffer = buffer;
  op.s = make_scale(scale);
  op.adam.buffer = nullptr;
  op.adam.s = op.s;
  op.adam.alpha_t = alpha_t; op.adam.omb1 = one_minus_beta1; op.adam.omb2 = one_minus_beta2;
  op.adam.eps = eps; op.adam.eta = eta; op.adam.wd = weight_decay_rate;
  op.adam.lower = lower; op.adam.upper = upper; op.adam.flags = adam_flags;
  op.adam.write_grad = 0;
  op.adam.hooks = {nullptr, 0.0, 0.0};
  op.write_grad = write_grad;
  op.hooks = master_hooks(hooks, d_skip);
  return launch_master(buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, op, stream,
                       "gp_unpack_adam_master");
}
