// gp_update.cu -- fused unpack + descale + optimizer update.
//
// Reference being replaced (chainer v7.8.1):
//   K3 div_by_size              chainermn/communicators/pure_nccl_communicator.py:183-189
//   K2 batched unpack           chainermn/communicators/_memory_utility.py:361-429
//   K5 momentum_sgd             chainer/optimizers/momentum_sgd.py:75-88
//        v = momentum * v - lr * grad;  param += v;            (T = param dtype)
//   K6 adam / amsgrad / adabound / amsbound    chainer/optimizers/adam.py:237-332
//        T grad_ = grad; T m_ = m; T v_ = v;
//        m_ += one_minus_beta1 * (grad_ - m_);
//        v_ += one_minus_beta2 * (grad_ * grad_ - v_);
//        [vhat_ = max(vhat_, v_); vhat = vhat_;]
//        m = m_; v = v_;
//        param -= eta * (alpha_t * m_ / (sqrt(v_|vhat_) + eps) + weight_decay_rate * param);
//        [adabound: max(min(alpha_t / (sqrt(.) + eps), upper), lower) * m_ instead]
//   (one launch per parameter in the reference; K3 and K2 are two more full
//    passes over the gradient.)
//
// Here one launch covers the whole parameter list and each mean-gradient
// element is read from HBM once (from the allreduced packed buffer), never
// materialised unless write_grad asks for param.grad to stay observable, as
// reference callers may expect (tests/chainermn_tests/optimizer_tests/
// test_multi_node_optimizer.py:57-110).
//
// Algorithmic HBM bytes per element (fp32 params, buffer itemsize b):
//   MomentumSGD  b + 8 (param r/w) + 8 (v r/w)  [+4 write_grad]
//   Adam         b + 8 + 8 (m) + 8 (v)          [+4 write_grad] [+8 vhat]
#include "gp_walk.cuh"

namespace {

template <class P> __device__ __forceinline__ P* mptr(uint64_t p) {
  return reinterpret_cast<P*>(p);
}

// ------------------------------------------------------------ MomentumSGD --
struct SgdOp {
  static constexpr int kMaxUnroll = 4;
  const void* buffer;
  ScaleArg s;
  double lr, momentum;
  int write_grad;

  static __device__ __forceinline__ int key(const gp_seg_t& g) { return g.dtype1; }

  // one element, arithmetic in P exactly as update_core_cpu
  // (momentum_sgd.py:61-73: v *= momentum; v -= lr * grad; param += v)
  template <class P>
  static __device__ __forceinline__ void math(typename Carrier<P>::type g,
                                              typename Carrier<P>::type& p,
                                              typename Carrier<P>::type& v,
                                              typename Carrier<P>::type lr_,
                                              typename Carrier<P>::type mom_) {
    using A = Arith<P>;
    v = A::sub(A::mul(mom_, v), A::mul(lr_, g));
    p = A::add(p, v);
  }

  template <class B, class P, int U>
  __device__ __forceinline__ void vec(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                      const bool (&act)[U]) const {
    using CB = typename Carrier<B>::type;
    using CP = typename Carrier<P>::type;
    Raw4<B> rb[U];
    Raw4<P> rp[U], rv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (act[u]) {
        rb[u] = ld4_stream(reinterpret_cast<const B*>(buffer) + seg[u]->buf_off + e[u]);
        rp[u] = ld4(mptr<P>(seg[u]->ptr[1]) + e[u]);
        rv[u] = ld4(mptr<P>(seg[u]->ptr[2]) + e[u]);
      }
    }
    const CP lr_ = Arith<P>::cst(lr), mom_ = Arith<P>::cst(momentum);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      CB xb[4];
      CP g[4], p[4], v[4];
      unpack4(rb[u], xb);
      unpack4(rp[u], p);
      unpack4(rv[u], v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i] = gpw::mean_grad_value<B, P>(xb[i], s);
        math<P>(g[i], p[i], v[i], lr_, mom_);
      }
      st4(mptr<P>(seg[u]->ptr[1]) + e[u], pack4<P, CP>(p));
      st4(mptr<P>(seg[u]->ptr[2]) + e[u], pack4<P, CP>(v));
      if (write_grad) st4(mptr<P>(seg[u]->ptr[0]) + e[u], pack4<P, CP>(g));
    }
  }

  template <class B, class P>
  __device__ __forceinline__ void one(const gp_seg_t& sg, int64_t e) const {
    using CP = typename Carrier<P>::type;
    const auto xb = to_carrier(reinterpret_cast<const B*>(buffer)[sg.buf_off + e]);
    const CP g = gpw::mean_grad_value<B, P>(xb, s);
    P* pp = mptr<P>(sg.ptr[1]) + e;
    P* pv = mptr<P>(sg.ptr[2]) + e;
    CP p = to_carrier(*pp), v = to_carrier(*pv);
    math<P>(g, p, v, Arith<P>::cst(lr), Arith<P>::cst(momentum));
    *pp = from_carrier<P>(p);
    *pv = from_carrier<P>(v);
    if (write_grad) mptr<P>(sg.ptr[0])[e] = from_carrier<P>(g);
  }
  template <class B>
  __device__ __forceinline__ void scalar(const gp_seg_t& sg, int64_t e) const {
    switch (sg.dtype1) {
      case GP_F32: one<B, float>(sg, e); break;
      case GP_F16: one<B, __half>(sg, e); break;
      case GP_F64: one<B, double>(sg, e); break;
      default: break;
    }
  }
};

// ------------------------------------------------------------------- Adam --
template <class P> struct AdamT { using type = float; };
template <> struct AdamT<double> { using type = double; };

struct AdamOp {
  static constexpr int kMaxUnroll = 2;
  const void* buffer;
  ScaleArg s;
  double alpha_t, omb1, omb2, eps, eta, wd, lower, upper;
  int flags;
  int write_grad;

  static __device__ __forceinline__ int key(const gp_seg_t& g) { return g.dtype1; }

  template <class T> struct Consts { T alpha_t, omb1, omb2, eps, eta, wd, lower, upper; };
  template <class T> __device__ __forceinline__ Consts<T> consts() const {
    Consts<T> c;
    c.alpha_t = (T)alpha_t; c.omb1 = (T)omb1; c.omb2 = (T)omb2; c.eps = (T)eps;
    c.eta = (T)eta; c.wd = (T)wd; c.lower = (T)lower; c.upper = (T)upper;
    return c;
  }

  // One element.  g, p, m, v, vh hold P-representable values in T; on return
  // p, m, v, vh are the values to store (rounded to P by the caller's pack).
  template <class P, class T>
  __device__ __forceinline__ void math(T g, T& p, T& m, T& v, T& vh, const Consts<T>& c) const {
    using I = Inter<T>;
    T m_ = I::add(m, I::mul(c.omb1, I::sub(g, m)));
    T v_ = I::add(v, I::mul(c.omb2, I::sub(I::mul(g, g), v)));
    T d_ = v_;
    if (flags & GP_ADAM_AMSGRAD) {
      vh = I::max(vh, v_);
      d_ = vh;
    }
    const T denom = I::add(I::sqrt(d_), c.eps);
    T step;
    if (flags & GP_ADAM_ADABOUND) {
      step = I::mul(I::max(I::min(I::div(c.alpha_t, denom), c.upper), c.lower), m_);
    } else {
      step = I::div(I::mul(c.alpha_t, m_), denom);
    }
    const T upd = I::mul(c.eta, I::add(step, I::mul(c.wd, p)));
    p = I::sub(p, upd);
    m = m_;
    v = v_;
  }

  template <class B, class P, int U>
  __device__ __forceinline__ void vec(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                      const bool (&act)[U]) const {
    using CB = typename Carrier<B>::type;
    using T = typename AdamT<P>::type;  // == Carrier<P>::type
    const bool ams = flags & GP_ADAM_AMSGRAD;
    Raw4<B> rb[U];
    Raw4<P> rp[U], rm[U], rv[U], rh[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (act[u]) {
        rb[u] = ld4_stream(reinterpret_cast<const B*>(buffer) + seg[u]->buf_off + e[u]);
        rp[u] = ld4(mptr<P>(seg[u]->ptr[1]) + e[u]);
        rm[u] = ld4(mptr<P>(seg[u]->ptr[2]) + e[u]);
        rv[u] = ld4(mptr<P>(seg[u]->ptr[3]) + e[u]);
        if (ams) rh[u] = ld4(mptr<P>(seg[u]->ptr[4]) + e[u]);
      }
    }
    const Consts<T> c = consts<T>();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      CB xb[4];
      T g[4], p[4], m[4], v[4], vh[4];
      unpack4(rb[u], xb);
      unpack4(rp[u], p);
      unpack4(rm[u], m);
      unpack4(rv[u], v);
      if (ams) unpack4(rh[u], vh);
      else { vh[0] = vh[1] = vh[2] = vh[3] = (T)0; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i] = gpw::mean_grad_value<B, P>(xb[i], s);
        math<P, T>(g[i], p[i], m[i], v[i], vh[i], c);
      }
      st4(mptr<P>(seg[u]->ptr[1]) + e[u], pack4<P, T>(p));
      st4(mptr<P>(seg[u]->ptr[2]) + e[u], pack4<P, T>(m));
      st4(mptr<P>(seg[u]->ptr[3]) + e[u], pack4<P, T>(v));
      if (ams) st4(mptr<P>(seg[u]->ptr[4]) + e[u], pack4<P, T>(vh));
      if (write_grad) st4(mptr<P>(seg[u]->ptr[0]) + e[u], pack4<P, T>(g));
    }
  }

  template <class B, class P>
  __device__ __forceinline__ void one(const gp_seg_t& sg, int64_t e) const {
    using T = typename AdamT<P>::type;
    const bool ams = flags & GP_ADAM_AMSGRAD;
    const auto xb = to_carrier(reinterpret_cast<const B*>(buffer)[sg.buf_off + e]);
    const T g = gpw::mean_grad_value<B, P>(xb, s);
    P* pp = mptr<P>(sg.ptr[1]) + e;
    P* pm = mptr<P>(sg.ptr[2]) + e;
    P* pv = mptr<P>(sg.ptr[3]) + e;
    P* ph = mptr<P>(sg.ptr[4]) + e;
    T p = to_carrier(*pp), m = to_carrier(*pm), v = to_carrier(*pv);
    T vh = ams ? (T)to_carrier(*ph) : (T)0;
    math<P, T>(g, p, m, v, vh, consts<T>());
    *pp = from_carrier<P>(p);
    *pm = from_carrier<P>(m);
    *pv = from_carrier<P>(v);
    if (ams) *ph = from_carrier<P>(vh);
    if (write_grad) mptr<P>(sg.ptr[0])[e] = from_carrier<P>(g);
  }
  template <class B>
  __device__ __forceinline__ void scalar(const gp_seg_t& sg, int64_t e) const {
    switch (sg.dtype1) {
      case GP_F32: one<B, float>(sg, e); break;
      case GP_F16: one<B, __half>(sg, e); break;
      case GP_F64: one<B, double>(sg, e); break;
      default: break;
    }
  }
};

}  // namespace

extern "C" int gp_unpack_momentum_sgd(const void* buffer, int buf_dtype, const int64_t* d_csum,
                                      const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                                      int64_t elem_end, double scale, double lr, double momentum,
                                      int write_grad, void* stream) {
  SgdOp op;
  op.buffer = buffer;
  op.s = make_scale(scale);
  op.lr = lr;
  op.momentum = momentum;
  op.write_grad = write_grad;
  return gpw::launch_buf(buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, op, stream,
                         "gp_unpack_momentum_sgd");
}

extern "C" int gp_unpack_adam(const void* buffer, int buf_dtype, const int64_t* d_csum,
                              const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                              int64_t elem_end, double scale, double alpha_t,
                              double one_minus_beta1, double one_minus_beta2, double eps,
                              double eta, double weight_decay_rate, double lower, double upper,
                              int adam_flags, int write_grad, void* stream) {
  AdamOp op;
  op.buffer = buffer;
  op.s = make_scale(scale);
  op.alpha_t = alpha_t;
  op.omb1 = one_minus_beta1;
  op.omb2 = one_minus_beta2;
  op.eps = eps;
  op.eta = eta;
  op.wd = weight_decay_rate;
  op.lower = lower;
  op.upper = upper;
  op.flags = adam_flags;
  op.write_grad = write_grad;
  return gpw::launch_buf(buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, op, stream,
                         "gp_unpack_adam");
}
