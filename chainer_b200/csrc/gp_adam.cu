// gp_adam.cu -- fused unpack + descale + Adam-family update (see also gp_sgd.cu).
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
#include "gp_bulk.cuh"

namespace {

template <class P> __device__ __forceinline__ P* mptr(uint64_t p) {
  return reinterpret_cast<P*>(p);
}

// ------------------------------------------------------------------- Adam --
template <class P> struct AdamT { using type = float; };
template <> struct AdamT<double> { using type = double; };

template <bool AMS>
struct AdamOp {
  static constexpr int kMaxUnroll = 4;
  static constexpr int kDefaultUnroll = 2;
  const void* buffer;
  ScaleArg s;
  double alpha_t, omb1, omb2, eps, eta, wd, lower, upper;
  int flags;
  int write_grad;

  static __device__ __forceinline__ int key(const gp_seg_t& g) { return g.dtype1; }

  // where the (summed) gradient of element e comes from: the packed buffer, or --
  // stand-alone optimizer.update() without a communicator, buffer == NULL -- the
  // gradient array itself (then dtype0 == buffer dtype is required by the host)
  template <class B>
  __device__ __forceinline__ const B* grad_src(const gp_seg_t& g, int64_t e) const {
    return buffer ? reinterpret_cast<const B*>(buffer) + g.buf_off + e
                  : reinterpret_cast<const B*>(g.ptr[0]) + e;
  }

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
    if constexpr (AMS) {
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

  template <class B, class P, int U> struct Regs {
    Raw4<B> rb[U];
    Raw4<P> rp[U], rm[U], rv[U], rh[AMS ? U : 1];
    P *pp[U], *pm[U], *pv[U], *ph[AMS ? U : 1], *pg[U];  // resolved once, before any store
  };

  template <class B, class P, int U>
  __device__ __forceinline__ void load(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                       const bool (&act)[U], Regs<B, P, U>& r) const {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (act[u]) {
        r.pp[u] = mptr<P>(seg[u]->ptr[1]) + e[u];
        r.pm[u] = mptr<P>(seg[u]->ptr[2]) + e[u];
        r.pv[u] = mptr<P>(seg[u]->ptr[3]) + e[u];
        r.pg[u] = mptr<P>(seg[u]->ptr[0]) + e[u];
        if constexpr (AMS) r.ph[u] = mptr<P>(seg[u]->ptr[4]) + e[u];
        r.rb[u] = ld4_stream(grad_src<B>(*seg[u], e[u]));
        r.rp[u] = ld4(r.pp[u]);
        r.rm[u] = ld4(r.pm[u]);
        r.rv[u] = ld4(r.pv[u]);
        if constexpr (AMS) r.rh[u] = ld4(r.ph[u]);
      }
    }
  }
  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void finish(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                         const bool (&act)[U], const Regs<B, P, U>& r) const {
    using CB = typename Carrier<B>::type;
    using T = typename AdamT<P>::type;  // == Carrier<P>::type
    const Consts<T> c = consts<T>();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      CB xb[4];
      T g[4], p[4], m[4], v[4], vh[4];
      unpack4(r.rb[u], xb);
      unpack4(r.rp[u], p);
      unpack4(r.rm[u], m);
      unpack4(r.rv[u], v);
      if constexpr (AMS) unpack4(r.rh[u], vh);
      else { vh[0] = vh[1] = vh[2] = vh[3] = (T)0; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i] = gpw::mean_grad_value<B, P, SM>(xb[i], s);
        math<P, T>(g[i], p[i], m[i], v[i], vh[i], c);
      }
      st4(r.pp[u], pack4<P, T>(p));
      st4(r.pm[u], pack4<P, T>(m));
      st4(r.pv[u], pack4<P, T>(v));
      if constexpr (AMS) st4(r.ph[u], pack4<P, T>(vh));
      if (write_grad) st4(r.pg[u], pack4<P, T>(g));
    }
  }
  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void vec(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                      const bool (&act)[U]) const {
    Regs<B, P, U> r;
    load<B, P, U>(seg, e, act, r);
    finish<B, P, U, SM>(seg, e, act, r);
  }

  // TMA path: one tile, in place in shared memory (gp_bulk.cuh).
  // arrays: 0 buffer, 1 param, 2 m, 3 v, [4 vhat,] last: grad out
  template <class B, class P, int SM>
  static __device__ __forceinline__ void tile(const AdamOp& op, unsigned char* st,
                                              const gpb::BulkArgs& a, int n_vec) {
    using CB = typename Carrier<B>::type;
    using T = typename AdamT<P>::type;
    B* sb = reinterpret_cast<B*>(st + a.arr[0].smem_off);
    P* sp = reinterpret_cast<P*>(st + a.arr[1].smem_off);
    P* sm = reinterpret_cast<P*>(st + a.arr[2].smem_off);
    P* sv = reinterpret_cast<P*>(st + a.arr[3].smem_off);
    P* sh = reinterpret_cast<P*>(st + a.arr[AMS ? 4 : 3].smem_off);
    P* sg = reinterpret_cast<P*>(st + a.arr[AMS ? 5 : 4].smem_off);
    const Consts<T> c = op.template consts<T>();
    for (int v = threadIdx.x; v < n_vec; v += gpb::kConsumers) {
      const Raw4<B> rb = gpb::lds4(sb + 4 * v);
      const Raw4<P> rp = gpb::lds4(sp + 4 * v);
      const Raw4<P> rm = gpb::lds4(sm + 4 * v);
      const Raw4<P> rv = gpb::lds4(sv + 4 * v);
      Raw4<P> rh;
      if constexpr (AMS) rh = gpb::lds4(sh + 4 * v);
      CB xb[4];
      T g[4], p[4], m[4], vv[4], vh[4];
      unpack4(rb, xb);
      unpack4(rp, p);
      unpack4(rm, m);
      unpack4(rv, vv);
      if constexpr (AMS) unpack4(rh, vh);
      else { vh[0] = vh[1] = vh[2] = vh[3] = (T)0; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i] = gpw::mean_grad_value<B, P, SM>(xb[i], op.s);
        op.template math<P, T>(g[i], p[i], m[i], vv[i], vh[i], c);
      }
      gpb::sts4(sp + 4 * v, pack4<P, T>(p));
      gpb::sts4(sm + 4 * v, pack4<P, T>(m));
      gpb::sts4(sv + 4 * v, pack4<P, T>(vv));
      if constexpr (AMS) gpb::sts4(sh + 4 * v, pack4<P, T>(vh));
      if (op.write_grad) gpb::sts4(sg + 4 * v, pack4<P, T>(g));
    }
  }

  template <class B, class P, int SM>
  __device__ __forceinline__ void one(const gp_seg_t& sg, int64_t e) const {
    using T = typename AdamT<P>::type;
    constexpr bool ams = AMS;
    const auto xb = to_carrier(*grad_src<B>(sg, e));
    const T g = gpw::mean_grad_value<B, P, SM>(xb, s);
    P* pp = mptr<P>(sg.ptr[1]) + e;
    P* pm = mptr<P>(sg.ptr[2]) + e;
    P* pv = mptr<P>(sg.ptr[3]) + e;
    P* ph = mptr<P>(sg.ptr[4]) + e;
    T p = to_carrier(*pp), m = to_carrier(*pm), v = to_carrier(*pv);
    T vh = (T)0;
    if constexpr (ams) vh = (T)to_carrier(*ph);
    math<P, T>(g, p, m, v, vh, consts<T>());
    *pp = from_carrier<P>(p);
    *pm = from_carrier<P>(m);
    *pv = from_carrier<P>(v);
    if constexpr (ams) *ph = from_carrier<P>(vh);
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

template <class Op>
void fill_adam(Op& op, const void* buffer, double scale, double alpha_t, double omb1, double omb2,
               double eps, double eta, double wd, double lower, double upper, int flags,
               int write_grad) {
  op.buffer = buffer;
  op.s = make_scale(scale);
  op.alpha_t = alpha_t; op.omb1 = omb1; op.omb2 = omb2; op.eps = eps; op.eta = eta; op.wd = wd;
  op.lower = lower; op.upper = upper; op.flags = flags; op.write_grad = write_grad;
}

}  // namespace

extern "C" int gp_unpack_adam(const void* buffer, int buf_dtype, const int64_t* d_csum,
                              const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                              int64_t elem_end, double scale, double alpha_t,
                              double one_minus_beta1, double one_minus_beta2, double eps,
                              double eta, double weight_decay_rate, double lower, double upper,
                              int adam_flags, int write_grad, int layout_hint, void* stream) {
  gpb::BulkArgs a = {};
  if (layout_hint && n_segs > 0 && buffer) {
    a.csum = d_csum; a.segs = d_segs; a.n_segs = n_segs; a.begin = elem_begin; a.end = elem_end;
    a.buffer = buffer;
    const int ps = gp_itemsize(layout_hint);
    const bool ams = adam_flags & GP_ADAM_AMSGRAD;
    int q = 0;
    a.arr[q++] = {-1, 0, 1, 0, 0};          // packed buffer
    a.arr[q++] = {1, ps, 1, 1, 0};          // param
    a.arr[q++] = {2, ps, 1, 1, 0};          // m
    a.arr[q++] = {3, ps, 1, 1, 0};          // v
    if (ams) a.arr[q++] = {4, ps, 1, 1, 0}; // vhat
    a.arr[q] = {0, ps, 0, 1, 0};            // mean gradient written back
    a.n_arrays = write_grad ? q + 1 : q;
  }
  if (adam_flags & GP_ADAM_AMSGRAD) {
    AdamOp<true> op;
    fill_adam(op, buffer, scale, alpha_t, one_minus_beta1, one_minus_beta2, eps, eta,
              weight_decay_rate, lower, upper, adam_flags, write_grad);
    if (layout_hint && n_segs > 0 && buffer) {
      const int r = gpb::launch_bulk(buf_dtype, layout_hint, a, op, stream, "gp_unpack_adam");
      if (r <= 0) return r;
    }
    return gpw::launch_buf(buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, op, stream,
                           "gp_unpack_adam", layout_hint == GP_F32);
  }
  AdamOp<false> op;
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
  if (layout_hint && n_segs > 0 && buffer) {
    const int r = gpb::launch_bulk(buf_dtype, layout_hint, a, op, stream, "gp_unpack_adam");
    if (r <= 0) return r;
  }
  return gpw::launch_buf(buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, op, stream,
                         "gp_unpack_adam", layout_hint == GP_F32);
}
