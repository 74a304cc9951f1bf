// gp_sgd_family.cu -- fused unpack + descale [+ hooks] + update for the other
// first-order rules of the reference that share MomentumSGD's shape:
//
//   GP_RULE_SGD                  chainer/optimizers/sgd.py:45-63
//        param -= lr * grad                                        (no state)
//   GP_RULE_CORRECTED_MOMENTUM   chainer/optimizers/corrected_momentum_sgd.py:61-89
//        v *= momentum; v -= grad; param += lr * v
//   GP_RULE_NESTEROV_AG          chainer/optimizers/nesterov_ag.py:60-84 (update_core_cpu)
//        v *= momentum; v -= lr * grad;
//        param += (momentum * momentum) * v; param -= ((1 + momentum) * lr) * grad
//
//   Arithmetic is in the parameter's type, one
// rounding per array operation, in the order of update_core_cpu -- the products of
// hyperparameters are formed in double on the host exactly as Python forms them.
// The reference's nesterov_ag GPU kernel fuses the two parameter updates into
// `param += m*m*v - (1+m)*lr*grad` (one rounding less); the CPU order is the one
// pinned by golden vectors (tests/golden/sgd_family.npz).
//
// Same walker, same table layout and the same optional hooks as gp_sgd_hooks.cu
// (ptr[0] grad, ptr[1] param, ptr[2] v).  Algorithmic HBM bytes per element:
// SGD b + 8, the momentum rules b + 16 (+4 with write_grad).
#include "gp_walk.cuh"

namespace {


template <int KIND>
struct FamilyOp {
  static constexpr int kMaxUnroll = 4;
  static constexpr int kDefaultUnroll = 2;
  static constexpr bool kHasV = KIND != GP_RULE_SGD;
  const void* buffer;
  ScaleArg s;
  double lr, momentum, mm, c;   // mm = momentum * momentum, c = (1 + momentum) * lr
  int write_grad;
  HookArgs hooks;

  static __device__ __forceinline__ int key(const gp_seg_t& g) { return g.dtype1; }

  template <class B>
  __device__ __forceinline__ const B* grad_src(const gp_seg_t& g, int64_t e) const {
    return buffer ? reinterpret_cast<const B*>(buffer) + g.buf_off + e
                  : reinterpret_cast<const B*>(g.ptr[0]) + e;
  }

  template <class P> struct Consts { typename Carrier<P>::type lr, mom, mm, c; };
  template <class P> __device__ __forceinline__ Consts<P> consts() const {
    Consts<P> k;
    k.lr = Arith<P>::cst(lr);
    k.mom = Arith<P>::cst(momentum);
    k.mm = Arith<P>::cst(mm);
    k.c = Arith<P>::cst(c);
    return k;
  }

  template <class P>
  static __device__ __forceinline__ void math(typename Carrier<P>::type g,
                                              typename Carrier<P>::type& p,
                                              typename Carrier<P>::type& v,
                                              const Consts<P>& k) {
    using A = Arith<P>;
    if constexpr (KIND == GP_RULE_SGD) {
      p = A::sub(p, A::mul(k.lr, g));
    } else if constexpr (KIND == GP_RULE_CORRECTED_MOMENTUM) {
      v = A::sub(A::mul(v, k.mom), g);
      p = A::add(p, A::mul(k.lr, v));
    } else {
      v = A::sub(A::mul(v, k.mom), A::mul(k.lr, g));
      p = A::add(p, A::mul(k.mm, v));
      p = A::sub(p, A::mul(k.c, g));
    }
  }

  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void vec(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                      const bool (&act)[U]) const {
    using CB = typename Carrier<B>::type;
    using CP = typename Carrier<P>::type;
    Raw4<B> rb[U];
    Raw4<P> rp[U], rv[kHasV ? U : 1];
    P *pp[U], *pv[kHasV ? U : 1], *pg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (act[u]) {
        pp[u] = mptr<P>(seg[u]->ptr[1]) + e[u];
        pg[u] = mptr<P>(seg[u]->ptr[0]) + e[u];
        rb[u] = ld4_stream(grad_src<B>(*seg[u], e[u]));
        rp[u] = ld4(pp[u]);
        if constexpr (kHasV) {
          pv[u] = mptr<P>(seg[u]->ptr[2]) + e[u];
          rv[u] = ld4(pv[u]);
        }
      }
    }
    const Consts<P> k = consts<P>();
    const HookRegs<P, true> hk(hooks);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      CB xb[4];
      CP g[4], p[4], v[4];
      unpack4(rb[u], xb);
      unpack4(rp[u], p);
      if constexpr (kHasV) unpack4(rv[u], v);
      else { v[0] = v[1] = v[2] = v[3] = (CP)0; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i] = hk.apply(gpw::mean_grad_value<B, P, SM>(xb[i], s), p[i]);
        math<P>(g[i], p[i], v[i], k);
      }
      st4(pp[u], pack4<P, CP>(p));
      if constexpr (kHasV) st4(pv[u], pack4<P, CP>(v));
      if (write_grad) st4(pg[u], pack4<P, CP>(g));
    }
  }

  template <class B, class P, int SM>
  __device__ __forceinline__ void one(const gp_seg_t& sg, int64_t e) const {
    using CP = typename Carrier<P>::type;
    const auto xb = to_carrier(*grad_src<B>(sg, e));
    P* pp = mptr<P>(sg.ptr[1]) + e;
    CP p = to_carrier(*pp), v = (CP)0;
    P* pv = nullptr;
    if constexpr (kHasV) {
      pv = mptr<P>(sg.ptr[2]) + e;
      v = to_carrier(*pv);
    }
    const HookRegs<P, true> hk(hooks);
    const CP g = hk.apply(gpw::mean_grad_value<B, P, SM>(xb, s), p);
    math<P>(g, p, v, consts<P>());
    *pp = from_carrier<P>(p);
    if constexpr (kHasV) *pv = from_carrier<P>(v);
    if (write_grad) mptr<P>(sg.ptr[0])[e] = from_carrier<P>(g);
  }
  template <class B, int SM>
  __device__ __forceinline__ void scalar(const gp_seg_t& sg, int64_t e) const {
    switch (sg.dtype1) {
      case GP_F32: one<B, float, SM>(sg, e); break;
      case GP_F16: one<B, __half, SM>(sg, e); break;
      case GP_F64: one<B, double, SM>(sg, e); break;
      default: break;
    }
  }
};

template <int KIND>
int launch_family(const void* buffer, int buf_dtype, const int64_t* d_csum, const gp_seg_t* d_segs,
                  int n_segs, int64_t begin, int64_t end, double scale, double lr, double momentum,
                  int write_grad, int layout_hint, const HookArgs& hooks, void* stream) {
  FamilyOp<KIND> op;
  op.buffer = buffer;
  op.s = make_scale(scale);
  op.lr = lr;
  op.momentum = momentum;
  op.mm = momentum * momentum;
  op.c = (1.0 + momentum) * lr;
  op.write_grad = write_grad;
  op.hooks = hooks;
  return gpw::launch_buf(buf_dtype, d_csum, d_segs, n_segs, begin, end, op, stream,
                         "gp_unpack_sgd_family", layout_hint == GP_F32);
}

}  // namespace

extern "C" int gp_unpack_sgd_family(const void* buffer, int buf_dtype, const int64_t* d_csum,
                                    const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                                    int64_t elem_end, double scale, int rule, double lr,
                                    double momentum, int write_grad, int layout_hint,
                                    const gp_hooks_t* hooks, void* stream) {
  HookArgs h = {nullptr, 0.0, 0.0};
  if (hooks) h = HookArgs{hooks->clip_rate, hooks->weight_decay, hooks->loss_scale};
  switch (rule) {
    case GP_RULE_SGD:
      return launch_family<GP_RULE_SGD>(buffer, buf_dtype, d_csum, d_segs, n_segs, elem_begin,
                                        elem_end, scale, lr, momentum, write_grad, layout_hint, h,
                                        stream);
    case GP_RULE_CORRECTED_MOMENTUM:
      return launch_family<GP_RULE_CORRECTED_MOMENTUM>(buffer, buf_dtype, d_csum, d_segs, n_segs,
                                                       elem_begin, elem_end, scale, lr, momentum,
                                                       write_grad, layout_hint, h, stream);
    case GP_RULE_NESTEROV_AG:
      return launch_family<GP_RULE_NESTEROV_AG>(buffer, buf_dtype, d_csum, d_segs, n_segs,
                                                elem_begin, elem_end, scale, lr, momentum,
                                                write_grad, layout_hint, h, stream);
    default:
      gp_set_error("gp_unpack_sgd_family: unknown rule id %d", rule);
      return GP_EINVAL;
  }
}
