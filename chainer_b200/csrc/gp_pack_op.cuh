// gp_pack_op.cuh -- the pack / unpack ops of the segmented walker (see gp_pack.cu for
// the reference map).  Included by gp_pack.cu and by the one-launch step kernels
// (gp_step.cu), which run PackOp and a fused update op inside the same kernel.
#pragma once
#include "gp_walk.cuh"

namespace {


// buffer[buf_off + k] = (B)(scale * ptr0[k])
struct PackOp {
  static constexpr int kMaxUnroll = 4;
  static constexpr int kDefaultUnroll = 4;
  void* buffer;
  ScaleArg s;

  static __device__ __forceinline__ int key(const gp_seg_t& g) { return g.dtype0; }

  // (B)(scale * x): product in double (or exactly in float for 2^k), ONE rounding
  template <class B, int SM, class CP>
  __device__ __forceinline__ Raw4<B> convert(const CP (&x)[4]) const {
    if constexpr (SM == 2 || (SM == 1 && sizeof(CP) == 8)) {
      double y[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = __dmul_rn((double)x[i], s.ds);
      return pack4<B, double>(y);
    } else if constexpr (SM == 1) {
      float y[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = __fmul_rn((float)x[i], s.fs);  // exact: factor is 2^k
      return pack4<B, float>(y);
    } else {
      return pack4<B, CP>(x);
    }
  }

  template <class B, class P, int U> struct Regs { Raw4<P> in[U]; B* dst[U]; };

  template <class B, class P, int U>
  __device__ __forceinline__ void load(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                       const bool (&act)[U], Regs<B, P, U>& r) const {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (act[u]) {
        r.dst[u] = reinterpret_cast<B*>(buffer) + seg[u]->buf_off + e[u];
        r.in[u] = ld4_stream(cptr<P>(seg[u]->ptr[0]) + e[u]);
      }
  }
  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void finish(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                         const bool (&act)[U], const Regs<B, P, U>& r) const {
    using CP = typename Carrier<P>::type;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      CP x[4];
      unpack4(r.in[u], x);
      st4(r.dst[u], convert<B, SM, CP>(x));
    }
  }
  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void vec(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                      const bool (&act)[U]) const {
    Regs<B, P, U> r;
    load<B, P, U>(seg, e, act, r);
    finish<B, P, U, SM>(seg, e, act, r);
  }

  template <class B, class P, int SM>
  __device__ __forceinline__ void one(const gp_seg_t& g, int64_t e) const {
    using CP = typename Carrier<P>::type;
    const CP x = to_carrier(cptr<P>(g.ptr[0])[e]);
    B out;
    if constexpr (SM == 2 || (SM == 1 && sizeof(CP) == 8)) out = from_d<B>(__dmul_rn((double)x, s.ds));
    else if constexpr (SM == 1) out = from_f<B>(__fmul_rn((float)x, s.fs));
    else out = from_carrier<B>(x);
    reinterpret_cast<B*>(buffer)[g.buf_off + e] = out;
  }
  template <class B, int SM>
  __device__ __forceinline__ void scalar(const gp_seg_t& g, int64_t e) const {
    switch (g.dtype0) {
      case GP_F32: one<B, float, SM>(g, e); break;
      case GP_F16: one<B, __half, SM>(g, e); break;
      case GP_F64: one<B, double, SM>(g, e); break;
      default: break;
    }
  }
};

// ptr0[k] = (P)( (B)(scale * buffer[buf_off + k]) )
struct UnpackOp {
  static constexpr int kMaxUnroll = 4;
  static constexpr int kDefaultUnroll = 4;
  const void* buffer;
  ScaleArg s;

  static __device__ __forceinline__ int key(const gp_seg_t& g) { return g.dtype0; }

  template <class B, class P, int U> struct Regs { Raw4<B> in[U]; P* dst[U]; };

  template <class B, class P, int U>
  __device__ __forceinline__ void load(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                       const bool (&act)[U], Regs<B, P, U>& r) const {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (act[u]) {
        r.dst[u] = mptr<P>(seg[u]->ptr[0]) + e[u];
        r.in[u] = ld4_stream(reinterpret_cast<const B*>(buffer) + seg[u]->buf_off + e[u]);
      }
  }
  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void finish(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                         const bool (&act)[U], const Regs<B, P, U>& r) const {
    using CB = typename Carrier<B>::type;
    using CP = typename Carrier<P>::type;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      CB x[4];
      CP g[4];
      unpack4(r.in[u], x);
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = gpw::mean_grad_value<B, P, SM>(x[i], s);
      st4(r.dst[u], pack4<P, CP>(g));
    }
  }
  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void vec(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                      const bool (&act)[U]) const {
    Regs<B, P, U> r;
    load<B, P, U>(seg, e, act, r);
    finish<B, P, U, SM>(seg, e, act, r);
  }

  template <class B, class P, int SM>
  __device__ __forceinline__ void one(const gp_seg_t& g, int64_t e) const {
    const auto x = to_carrier(reinterpret_cast<const B*>(buffer)[g.buf_off + e]);
    mptr<P>(g.ptr[0])[e] = from_carrier<P>(gpw::mean_grad_value<B, P, SM>(x, s));
  }
  template <class B, int SM>
  __device__ __forceinline__ void scalar(const gp_seg_t& g, int64_t e) const {
    switch (g.dtype0) {
      case GP_F32: one<B, float, SM>(g, e); break;
      case GP_F16: one<B, __half, SM>(g, e); break;
      case GP_F64: one<B, double, SM>(g, e); break;
      default: break;
    }
  }
};

}  // namespace
