// gp_sgd_op.cuh -- the fused MomentumSGD op (see gp_sgd.cu for the reference map).
// Included by gp_sgd.cu (H = false) and gp_sgd_hooks.cu (H = true: pre-update hooks).
#pragma once
#include "gp_bulk.cuh"

namespace {


// ------------------------------------------------------------ MomentumSGD --
template <bool H>
struct SgdOp {
  static constexpr int kMaxUnroll = 4;
  static constexpr int kDefaultUnroll = 2;
  const void* buffer;
  ScaleArg s;
  double lr, momentum;
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

  template <class B, class P, int U> struct Regs {
    Raw4<B> rb[U];
    Raw4<P> rp[U], rv[U];
    P *pp[U], *pv[U], *pg[U];  // resolved once, before any store (no table re-reads)
  };

  // WITH_BUF false: the caller supplies r.rb (one-launch step at one rank, gp_step.cu)
  template <class B, class P, int U, bool WITH_BUF = true>
  __device__ __forceinline__ void load(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                       const bool (&act)[U], Regs<B, P, U>& r) const {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (act[u]) {
        r.pp[u] = mptr<P>(seg[u]->ptr[1]) + e[u];
        r.pv[u] = mptr<P>(seg[u]->ptr[2]) + e[u];
        r.pg[u] = mptr<P>(seg[u]->ptr[0]) + e[u];
        if constexpr (WITH_BUF) r.rb[u] = ld4_stream(grad_src<B>(*seg[u], e[u]));
        r.rp[u] = ld4(r.pp[u]);
        r.rv[u] = ld4(r.pv[u]);
      }
    }
  }
  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void finish(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                         const bool (&act)[U], const Regs<B, P, U>& r) const {
    using CB = typename Carrier<B>::type;
    using CP = typename Carrier<P>::type;
    const CP lr_ = Arith<P>::cst(lr), mom_ = Arith<P>::cst(momentum);
    const HookRegs<P, H> hk(hooks);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      CB xb[4];
      CP g[4], p[4], v[4];
      unpack4(r.rb[u], xb);
      unpack4(r.rp[u], p);
      unpack4(r.rv[u], v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i] = hk.apply(gpw::mean_grad_value<B, P, SM>(xb[i], s), p[i]);
        math<P>(g[i], p[i], v[i], lr_, mom_);
      }
      st4(r.pp[u], pack4<P, CP>(p));
      st4(r.pv[u], pack4<P, CP>(v));
      if (write_grad) st4(r.pg[u], pack4<P, CP>(g));
    }
  }
  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void vec(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                      const bool (&act)[U]) const {
    Regs<B, P, U> r;
    load<B, P, U>(seg, e, act, r);
    finish<B, P, U, SM>(seg, e, act, r);
  }

  // TMA path: one tile, in place in shared memory (gp_bulk.cuh)
  template <class B, class P, int SM>
  static __device__ __forceinline__ void tile(const SgdOp& op, unsigned char* st,
                                              const gpb::BulkArgs& a, int n_vec) {
    using CB = typename Carrier<B>::type;
    using CP = typename Carrier<P>::type;
    B* sb = reinterpret_cast<B*>(st + a.arr[0].smem_off);
    P* sp = reinterpret_cast<P*>(st + a.arr[1].smem_off);
    P* sv = reinterpret_cast<P*>(st + a.arr[2].smem_off);
    P* sg = reinterpret_cast<P*>(st + a.arr[3].smem_off);
    const CP lr_ = Arith<P>::cst(op.lr), mom_ = Arith<P>::cst(op.momentum);
    const HookRegs<P, H> hk(op.hooks);
    constexpr int UN = 2;
    for (int v0 = threadIdx.x; v0 < n_vec; v0 += gpb::kConsumers * UN) {
      Raw4<B> rb[UN];
      Raw4<P> rp[UN], rv[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int v = v0 + u * gpb::kConsumers;
        if (v < n_vec) {
          rb[u] = gpb::lds4(sb + 4 * v);
          rp[u] = gpb::lds4(sp + 4 * v);
          rv[u] = gpb::lds4(sv + 4 * v);
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int v = v0 + u * gpb::kConsumers;
        if (v >= n_vec) continue;
        CB xb[4];
        CP g[4], p[4], vv[4];
        unpack4(rb[u], xb);
        unpack4(rp[u], p);
        unpack4(rv[u], vv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          g[i] = hk.apply(gpw::mean_grad_value<B, P, SM>(xb[i], op.s), p[i]);
          math<P>(g[i], p[i], vv[i], lr_, mom_);
        }
        gpb::sts4(sp + 4 * v, pack4<P, CP>(p));
        gpb::sts4(sv + 4 * v, pack4<P, CP>(vv));
        if (op.write_grad) gpb::sts4(sg + 4 * v, pack4<P, CP>(g));
      }
    }
  }

  template <class B, class P, int SM>
  __device__ __forceinline__ void one(const gp_seg_t& sg, int64_t e) const {
    using CP = typename Carrier<P>::type;
    const auto xb = to_carrier(*grad_src<B>(sg, e));
    P* pp = mptr<P>(sg.ptr[1]) + e;
    P* pv = mptr<P>(sg.ptr[2]) + e;
    CP p = to_carrier(*pp), v = to_carrier(*pv);
    const HookRegs<P, H> hk(hooks);
    const CP g = hk.apply(gpw::mean_grad_value<B, P, SM>(xb, s), p);
    math<P>(g, p, v, Arith<P>::cst(lr), Arith<P>::cst(momentum));
    *pp = from_carrier<P>(p);
    *pv = from_carrier<P>(v);
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

// fill + launch, shared by the plain (gp_sgd.cu) and hooked (gp_sgd_hooks.cu) entry points
template <bool H>
int launch_sgd(const void* buffer, int buf_dtype, const int64_t* d_csum, const gp_seg_t* d_segs,
               int n_segs, int64_t elem_begin, int64_t elem_end, double scale, double lr,
               double momentum, int write_grad, int layout_hint, const HookArgs& hooks,
               void* stream, const char* what) {
  SgdOp<H> op;
  op.buffer = buffer;
  op.s = make_scale(scale);
  op.lr = lr;
  op.momentum = momentum;
  op.write_grad = write_grad;
  op.hooks = hooks;
  if constexpr (!H) if (layout_hint && n_segs > 0 && buffer) {
    gpb::BulkArgs a = {};
    a.csum = d_csum; a.segs = d_segs; a.n_segs = n_segs; a.begin = elem_begin; a.end = elem_end;
    a.buffer = buffer;
    const int ps = gp_itemsize(layout_hint);
    a.n_arrays = write_grad ? 4 : 3;
    a.arr[0] = {-1, 0, 1, 0, 0};           // packed buffer: load only
    a.arr[1] = {1, ps, 1, 1, 0};           // param: load + store
    a.arr[2] = {2, ps, 1, 1, 0};           // v
    a.arr[3] = {0, ps, 0, 1, 0};           // mean gradient written back
    const int r = gpb::launch_bulk(buf_dtype, layout_hint, a, op, stream, what);
    if (r <= 0) return r;
  }
  return gpw::launch_buf(buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, op, stream,
                         what, layout_hint == GP_F32);
}

}  // namespace
