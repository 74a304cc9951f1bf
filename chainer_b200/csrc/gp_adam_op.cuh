// gp_adam_op.cuh -- the fused Adam-family op (see gp_adam.cu for the reference map).
// Included by gp_adam.cu (H = false) and gp_adam_hooks.cu (H = true: pre-update hooks).
#pragma once
#include "gp_bulk.cuh"

namespace {


// ------------------------------------------------------------------- Adam --
template <class P> struct AdamT { using type = float; };
template <> struct AdamT<double> { using type = double; };

template <bool AMS, bool H>
struct AdamOp {
  static constexpr int kMaxUnroll = 4;
  static constexpr int kDefaultUnroll = 2;
  const void* buffer;
  ScaleArg s;
  double alpha_t, omb1, omb2, eps, eta, wd, lower, upper;
  int flags;
  int write_grad;
  HookArgs hooks;  // read only when H

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

  // WITH_BUF false: the caller supplies r.rb (one-launch step at one rank, gp_step.cu)
  template <class B, class P, int U, bool WITH_BUF = true>
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
        if constexpr (WITH_BUF) r.rb[u] = ld4_stream(grad_src<B>(*seg[u], e[u]));
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
    const HookRegs<P, H> hk(hooks);
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
        g[i] = hk.apply(gpw::mean_grad_value<B, P, SM>(xb[i], s), p[i]);
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
    const HookRegs<P, H> hk(op.hooks);
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
        g[i] = hk.apply(gpw::mean_grad_value<B, P, SM>(xb[i], op.s), p[i]);
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
    P* pp = mptr<P>(sg.ptr[1]) + e;
    P* pm = mptr<P>(sg.ptr[2]) + e;
    P* pv = mptr<P>(sg.ptr[3]) + e;
    P* ph = mptr<P>(sg.ptr[4]) + e;
    T p = to_carrier(*pp), m = to_carrier(*pm), v = to_carrier(*pv);
    const HookRegs<P, H> hk(hooks);
    const T g = hk.apply(gpw::mean_grad_value<B, P, SM>(xb, s), p);
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

// fill + launch, shared by the plain (gp_adam.cu) and hooked (gp_adam_hooks.cu) entry points
template <bool AMS, bool H>
int launch_adam_t(const void* buffer, int buf_dtype, const int64_t* d_csum, const gp_seg_t* d_segs,
                  int n_segs, int64_t elem_begin, int64_t elem_end, double scale, double alpha_t,
                  double omb1, double omb2, double eps, double eta, double wd, double lower,
                  double upper, int flags, int write_grad, int layout_hint, const HookArgs& hooks,
                  void* stream, const char* what) {
  AdamOp<AMS, H> op;
  op.buffer = buffer;
  op.s = make_scale(scale);
  op.alpha_t = alpha_t; op.omb1 = omb1; op.omb2 = omb2; op.eps = eps; op.eta = eta; op.wd = wd;
  op.lower = lower; op.upper = upper; op.flags = flags; op.write_grad = write_grad;
  op.hooks = hooks;
  if constexpr (!H) if (layout_hint && n_segs > 0 && buffer) {
    gpb::BulkArgs a = {};
    a.csum = d_csum; a.segs = d_segs; a.n_segs = n_segs; a.begin = elem_begin; a.end = elem_end;
    a.buffer = buffer;
    const int ps = gp_itemsize(layout_hint);
    int q = 0;
    a.arr[q++] = {-1, 0, 1, 0, 0};          // packed buffer
    a.arr[q++] = {1, ps, 1, 1, 0};          // param
    a.arr[q++] = {2, ps, 1, 1, 0};          // m
    a.arr[q++] = {3, ps, 1, 1, 0};          // v
    if (AMS) a.arr[q++] = {4, ps, 1, 1, 0}; // vhat
    a.arr[q] = {0, ps, 0, 1, 0};            // mean gradient written back
    a.n_arrays = write_grad ? q + 1 : q;
    const int r = gpb::launch_bulk(buf_dtype, layout_hint, a, op, stream, what);
    if (r <= 0) return r;
  }
  return gpw::launch_buf(buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, op, stream,
                         what, layout_hint == GP_F32);
}

template <bool H>
int launch_adam(const void* buffer, int buf_dtype, const int64_t* d_csum, const gp_seg_t* d_segs,
                int n_segs, int64_t elem_begin, int64_t elem_end, double scale, double alpha_t,
                double omb1, double omb2, double eps, double eta, double wd, double lower,
                double upper, int flags, int write_grad, int layout_hint, const HookArgs& hooks,
                void* stream, const char* what) {
  if (flags & GP_ADAM_AMSGRAD)
    return launch_adam_t<true, H>(buffer, buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end,
                                  scale, alpha_t, omb1, omb2, eps, eta, wd, lower, upper, flags,
                                  write_grad, layout_hint, hooks, stream, what);
  return launch_adam_t<false, H>(buffer, buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end,
                                 scale, alpha_t, omb1, omb2, eps, eta, wd, lower, upper, flags,
                                 write_grad, layout_hint, hooks, stream, what);
}

}  // namespace
