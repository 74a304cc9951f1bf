// gp_common.cuh -- shared device helpers of libgradpath (sm_100a).
//
// Nothing on this path is a contraction, so there are no tensor-core
// instructions here: every kernel is an HBM-bound stream.  The helpers below
// are what such kernels need on B200: 4-element wide (up to 128-bit) loads and
// stores, streaming cache hints for read-once data, exact (non-contracted)
// IEEE arithmetic so that results are bit-identical to the NumPy restatement
// of the reference, and warp-shuffle reductions.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gradpath.h"

// table entries carry pointers as uint64
template <class P> __device__ __forceinline__ const P* cptr(uint64_t p) {
  return reinterpret_cast<const P*>(p);
}
template <class P> __device__ __forceinline__ P* mptr(uint64_t p) {
  return reinterpret_cast<P*>(p);
}

// ----------------------------------------------------------------- errors --
void gp_set_error(const char* fmt, ...);
int gp_cuda_fail(cudaError_t e, const char* what);
#define GP_CUDA(expr)                                 \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) return gp_cuda_fail(_e, #expr); \
  } while (0)

struct GpTuning {
  int threads;      // threads per CTA of the streaming kernels
  int unroll;       // 4-element vectors in flight per thread and array (1, 2, 4)
  int ctas_per_sm;  // persistent grid = SMs * ctas_per_sm
  int persistent;   // 1: balanced contiguous range per CTA; 0: one tile per CTA
  int bn_threads;
  int pipeline;     // (unused: the register-pipelined walker was measured slower and removed)
  int bn_ctas_per_sm;
};
extern GpTuning g_gp_tuning;
// wall-time bound of the in-kernel waits for peer GPUs, ns (0: unbounded); gp_set_tuning("peer_timeout_s")
extern unsigned long long g_gp_peer_timeout_ns;
int gp_sm_count_cached();

// ------------------------------------------------------------ dtype traits --
template <class T> struct Carrier { using type = float; };
template <> struct Carrier<double> { using type = double; };

template <class T> struct DtypeId;
template <> struct DtypeId<__half> { static constexpr int value = GP_F16; };
template <> struct DtypeId<__nv_bfloat16> { static constexpr int value = GP_BF16; };
template <> struct DtypeId<float> { static constexpr int value = GP_F32; };
template <> struct DtypeId<double> { static constexpr int value = GP_F64; };

__host__ __device__ __forceinline__ int gp_itemsize(int dtype) {
  return dtype == GP_F64 ? 8 : (dtype == GP_F32 ? 4 : 2);
}

// scalar conversions: storage type -> carrier (always exact) ...
__device__ __forceinline__ float to_carrier(__half x) { return __half2float(x); }
__device__ __forceinline__ float to_carrier(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float to_carrier(float x) { return x; }
__device__ __forceinline__ double to_carrier(double x) { return x; }

// ... and carrier -> storage type, round-to-nearest-even, ONE rounding.
template <class T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ double from_f<double>(float x) { return (double)x; }

template <class T> __device__ __forceinline__ T from_d(double x);
template <> __device__ __forceinline__ __half from_d<__half>(double x) { return __double2half(x); }
template <> __device__ __forceinline__ __nv_bfloat16 from_d<__nv_bfloat16>(double x) { return __double2bfloat16(x); }
template <> __device__ __forceinline__ float from_d<float>(double x) { return __double2float_rn(x); }
template <> __device__ __forceinline__ double from_d<double>(double x) { return x; }

template <class T> __device__ __forceinline__ T from_carrier(float x) { return from_f<T>(x); }
template <class T> __device__ __forceinline__ T from_carrier(double x) { return from_d<T>(x); }

// round a carrier value to storage type T and bring it back as T's carrier
template <class T, class C>
__device__ __forceinline__ typename Carrier<T>::type round_through(C x) {
  return to_carrier(from_carrier<T>(x));
}

// ------------------------------------------------------------- the scale --
// `x *= (1.0/size)` of the reference is evaluated in double and rounded once
// to the buffer type (pure_nccl_communicator.py:183-189).  When the factor is
// a power of two the same result is obtained with one float multiply
// (the product is exact before the single rounding), which keeps FP64 units
// out of the stream; mode 2 is the general double path.
struct ScaleArg {
  int mode;  // 0: identity, 1: power of two (float exact), 2: general (double)
  float fs;
  double ds;
};

inline ScaleArg make_scale(double s) {
  ScaleArg a;
  a.ds = s;
  a.fs = (float)s;
  if (s == 1.0) {
    a.mode = 0;
  } else {
    union { double d; uint64_t u; } b;
    b.d = s;
    const bool pow2 = (b.u & 0x000FFFFFFFFFFFFFull) == 0 && s > 0 && s >= 1e-30 && s <= 1e30;
    a.mode = pow2 ? 1 : 2;
  }
  return a;
}

// value read from a buffer of type B (as carrier) -> scaled, rounded to B.
// The mode is a compile-time parameter of the kernels: with a run-time switch
// the compiler evaluated the double path speculatively for every element.
template <class B, int SM>
__device__ __forceinline__ typename Carrier<B>::type descale(typename Carrier<B>::type x,
                                                             const ScaleArg& s) {
  if constexpr (SM == 0) {
    return x;
  } else if constexpr (SM == 1) {
    if constexpr (sizeof(typename Carrier<B>::type) == 8) {
      return __dmul_rn(x, s.ds);
    } else {
      return round_through<B>(__fmul_rn(x, s.fs));
    }
  } else {
    return round_through<B>(__dmul_rn((double)x, s.ds));
  }
}
// run-time mode (small flat kernels only)
template <class B>
__device__ __forceinline__ typename Carrier<B>::type descale_rt(typename Carrier<B>::type x,
                                                                const ScaleArg& s) {
  if (s.mode == 0) return descale<B, 0>(x, s);
  if (s.mode == 1) return descale<B, 1>(x, s);
  return descale<B, 2>(x, s);
}

// ------------------------------------------------------- 4-wide accessors --
// Every array is accessed 4 elements at a time: 8 B for f16/bf16, 16 B for f32,
// 2 x 16 B for f64.  A warp instruction therefore always covers a contiguous,
// fully used span (256 B / 512 B), i.e. whole 32 B sectors.
template <class T> struct Raw4;
template <> struct Raw4<__half> { uint2 r; };
template <> struct Raw4<__nv_bfloat16> { uint2 r; };
template <> struct Raw4<float> { float4 r; };
template <> struct Raw4<double> { double2 a, b; };

// streaming (read-once, never written by this kernel) load
template <class T> __device__ __forceinline__ Raw4<T> ld4_stream(const T* p);
template <> __device__ __forceinline__ Raw4<__half> ld4_stream(const __half* p) {
  Raw4<__half> v; v.r = __ldcs(reinterpret_cast<const uint2*>(p)); return v;
}
template <> __device__ __forceinline__ Raw4<__nv_bfloat16> ld4_stream(const __nv_bfloat16* p) {
  Raw4<__nv_bfloat16> v; v.r = __ldcs(reinterpret_cast<const uint2*>(p)); return v;
}
template <> __device__ __forceinline__ Raw4<float> ld4_stream(const float* p) {
  Raw4<float> v; v.r = __ldcs(reinterpret_cast<const float4*>(p)); return v;
}
template <> __device__ __forceinline__ Raw4<double> ld4_stream(const double* p) {
  Raw4<double> v;
  v.a = __ldcs(reinterpret_cast<const double2*>(p));
  v.b = __ldcs(reinterpret_cast<const double2*>(p) + 1);
  return v;
}
// read-modify-write load (param / optimizer state): plain ld.global
template <class T> __device__ __forceinline__ Raw4<T> ld4(const T* p);
template <> __device__ __forceinline__ Raw4<__half> ld4(const __half* p) {
  Raw4<__half> v; v.r = *reinterpret_cast<const uint2*>(p); return v;
}
template <> __device__ __forceinline__ Raw4<__nv_bfloat16> ld4(const __nv_bfloat16* p) {
  Raw4<__nv_bfloat16> v; v.r = *reinterpret_cast<const uint2*>(p); return v;
}
template <> __device__ __forceinline__ Raw4<float> ld4(const float* p) {
  Raw4<float> v; v.r = *reinterpret_cast<const float4*>(p); return v;
}
template <> __device__ __forceinline__ Raw4<double> ld4(const double* p) {
  Raw4<double> v;
  v.a = *reinterpret_cast<const double2*>(p);
  v.b = *(reinterpret_cast<const double2*>(p) + 1);
  return v;
}
template <class T> __device__ __forceinline__ void st4(T* p, const Raw4<T>& v);
template <> __device__ __forceinline__ void st4(__half* p, const Raw4<__half>& v) {
  __stcs(reinterpret_cast<uint2*>(p), v.r);
}
template <> __device__ __forceinline__ void st4(__nv_bfloat16* p, const Raw4<__nv_bfloat16>& v) {
  __stcs(reinterpret_cast<uint2*>(p), v.r);
}
template <> __device__ __forceinline__ void st4(float* p, const Raw4<float>& v) {
  __stcs(reinterpret_cast<float4*>(p), v.r);
}
template <> __device__ __forceinline__ void st4(double* p, const Raw4<double>& v) {
  __stcs(reinterpret_cast<double2*>(p), v.a);
  __stcs(reinterpret_cast<double2*>(p) + 1, v.b);
}

// raw <-> carrier[4]
__device__ __forceinline__ void unpack4(const Raw4<__half>& v, float (&o)[4]) {
  const __half2 a = *reinterpret_cast<const __half2*>(&v.r.x);
  const __half2 b = *reinterpret_cast<const __half2*>(&v.r.y);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  o[0] = fa.x; o[1] = fa.y; o[2] = fb.x; o[3] = fb.y;
}
__device__ __forceinline__ void unpack4(const Raw4<__nv_bfloat16>& v, float (&o)[4]) {
  // bf16 -> f32 is a 16-bit shift
  o[0] = __uint_as_float(v.r.x << 16);
  o[1] = __uint_as_float(v.r.x & 0xFFFF0000u);
  o[2] = __uint_as_float(v.r.y << 16);
  o[3] = __uint_as_float(v.r.y & 0xFFFF0000u);
}
__device__ __forceinline__ void unpack4(const Raw4<float>& v, float (&o)[4]) {
  o[0] = v.r.x; o[1] = v.r.y; o[2] = v.r.z; o[3] = v.r.w;
}
__device__ __forceinline__ void unpack4(const Raw4<double>& v, double (&o)[4]) {
  o[0] = v.a.x; o[1] = v.a.y; o[2] = v.b.x; o[3] = v.b.y;
}

template <class T, class C> __device__ __forceinline__ Raw4<T> pack4(const C (&x)[4]);
template <class T, class C> struct Pack4Impl;
template <class C> struct Pack4Impl<__half, C> {
  static __device__ __forceinline__ Raw4<__half> run(const C (&x)[4]) {
    Raw4<__half> v;
    __half h0 = from_carrier<__half>(x[0]), h1 = from_carrier<__half>(x[1]);
    __half h2 = from_carrier<__half>(x[2]), h3 = from_carrier<__half>(x[3]);
    v.r.x = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    v.r.y = (uint32_t)__half_as_ushort(h2) | ((uint32_t)__half_as_ushort(h3) << 16);
    return v;
  }
};
template <class C> struct Pack4Impl<__nv_bfloat16, C> {
  static __device__ __forceinline__ Raw4<__nv_bfloat16> run(const C (&x)[4]) {
    Raw4<__nv_bfloat16> v;
    __nv_bfloat16 h0 = from_carrier<__nv_bfloat16>(x[0]), h1 = from_carrier<__nv_bfloat16>(x[1]);
    __nv_bfloat16 h2 = from_carrier<__nv_bfloat16>(x[2]), h3 = from_carrier<__nv_bfloat16>(x[3]);
    v.r.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    v.r.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
    return v;
  }
};
template <class C> struct Pack4Impl<float, C> {
  static __device__ __forceinline__ Raw4<float> run(const C (&x)[4]) {
    Raw4<float> v;
    v.r = make_float4(from_carrier<float>(x[0]), from_carrier<float>(x[1]),
                      from_carrier<float>(x[2]), from_carrier<float>(x[3]));
    return v;
  }
};
template <class C> struct Pack4Impl<double, C> {
  static __device__ __forceinline__ Raw4<double> run(const C (&x)[4]) {
    Raw4<double> v;
    v.a = make_double2((double)x[0], (double)x[1]);
    v.b = make_double2((double)x[2], (double)x[3]);
    return v;
  }
};
template <class T, class C> __device__ __forceinline__ Raw4<T> pack4(const C (&x)[4]) {
  return Pack4Impl<T, C>::run(x);
}

// ------------------------------------------------ exact IEEE arithmetic --
// Arithmetic "in type P" as NumPy / the reference kernels do it: every
// operation is rounded to P.  For P = half the operation is carried out in
// float and rounded to half (exact for + - *: float has > 2*11+2 bits).  The
// explicit _rn intrinsics are never contracted into FMAs by nvcc, so results
// do not depend on compiler flags.
template <class P> struct Arith;
template <> struct Arith<float> {
  using C = float;
  static __device__ __forceinline__ C cst(double x) { return (float)x; }
  static __device__ __forceinline__ C mul(C a, C b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ C add(C a, C b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ C sub(C a, C b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ C div(C a, C b) { return __fdiv_rn(a, b); }
};
template <> struct Arith<__half> {
  using C = float;
  static __device__ __forceinline__ C r(C x) { return __half2float(__float2half_rn(x)); }
  static __device__ __forceinline__ C cst(double x) { return __half2float(__double2half(x)); }
  static __device__ __forceinline__ C mul(C a, C b) { return r(__fmul_rn(a, b)); }
  static __device__ __forceinline__ C add(C a, C b) { return r(__fadd_rn(a, b)); }
  static __device__ __forceinline__ C sub(C a, C b) { return r(__fsub_rn(a, b)); }
  // float has 24 >= 2*11+2 bits: rounding the float quotient to half is the
  // correctly rounded half quotient
  static __device__ __forceinline__ C div(C a, C b) { return r(__fdiv_rn(a, b)); }
};
template <> struct Arith<double> {
  using C = double;
  static __device__ __forceinline__ C cst(double x) { return x; }
  static __device__ __forceinline__ C mul(C a, C b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ C add(C a, C b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ C sub(C a, C b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ C div(C a, C b) { return __ddiv_rn(a, b); }
};

// ------------------------------------------- fused pre-update gradient hooks --
// What GradientMethod.update (chainer/optimizer.py:857-894) and UpdateRule.update
// (:286-291) do to a gradient between the mean and update_core when the optimizer
// carries the hooks [GradientClipping][, WeightDecay] and static loss scaling:
//   g *= rate            optimizer_hooks/gradient_clipping.py:84-106 (rate <= 1, device scalar)
//   g += decay * p       optimizer_hooks/weight_decay.py:44-57 (decay = rate * loss_scale)
//   g /= loss_scale      optimizer.py:289-291
// every operation rounded to the parameter's type, as the array expressions are.
struct HookArgs {
  const float* clip_rate;  // device pointer (gp_sqnorm output) or nullptr
  double decay;            // 0: off
  double loss_scale;       // 0: off
};
template <class P, bool H> struct HookRegs;
template <class P> struct HookRegs<P, false> {
  using C = typename Carrier<P>::type;
  __device__ __forceinline__ explicit HookRegs(const HookArgs&) {}
  __device__ __forceinline__ C apply(C g, C) const { return g; }
};
template <class P> struct HookRegs<P, true> {
  using C = typename Carrier<P>::type;
  C rate, decay, ls;
  bool use_rate, use_decay, use_ls;
  __device__ __forceinline__ explicit HookRegs(const HookArgs& h) {
    use_rate = h.clip_rate != nullptr;
    rate = use_rate ? Arith<P>::cst((double)__ldg(h.clip_rate)) : (C)1;
    use_decay = h.decay != 0.0;
    decay = Arith<P>::cst(h.decay);
    use_ls = h.loss_scale != 0.0;
    ls = Arith<P>::cst(h.loss_scale);
  }
  __device__ __forceinline__ C apply(C g, C p) const {
    using A = Arith<P>;
    if (use_rate) g = A::mul(g, rate);
    if (use_decay) g = A::add(g, A::mul(decay, p));
    if (use_ls) g = A::div(g, ls);
    return g;
  }
};

// intermediate type T of the Adam kernels (adam.py:57-63): float for
// float16/float32 parameters, double for float64.
template <class T> struct Inter;
template <> struct Inter<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
  static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float min(float a, float b) { return fminf(a, b); }
};
template <> struct Inter<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
  static __device__ __forceinline__ double max(double a, double b) { return fmax(a, b); }
  static __device__ __forceinline__ double min(double a, double b) { return fmin(a, b); }
};

// ------------------------------------------------------- warp reductions --
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
