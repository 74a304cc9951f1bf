// gp_grad.cu -- the gradient-path stream kernels of libgradpath (sm_100a):
//   gp_pack                  gather + cast (+ pre-scale) of the parameter list
//   gp_unpack_scale          descale + scatter + cast (mean_grad stand-alone)
//   gp_unpack_momentum_sgd   fused unpack + descale + MomentumSGD
//   gp_unpack_adam           fused unpack + descale + Adam/AdamW/AMSGrad/AdaBound
//   gp_scale, gp_check_finite
//
// Reference being replaced (chainer v7.8.1):
//   chainermn/communicators/_memory_utility.py:289-429  (pack / unpack RawKernels:
//       one thread per element, a binary search through global memory per element,
//       scalar accesses)
//   chainermn/communicators/pure_nccl_communicator.py:183-189 (div_by_size)
//   chainer/optimizers/momentum_sgd.py:75-88, chainer/optimizers/adam.py:224-332
//       (one ElementwiseKernel launch per parameter)
//
// Design (HBM-bound, no data reuse):
//   * ONE launch walks the whole flat element space of the parameter list.
//     The grid is persistent (SMs x ctas_per_sm CTAs); each CTA owns an equal
//     contiguous slice of the flat space, so there is no wave quantisation and
//     no tail.
//   * The cumulative-size table (int64[n+1]) is staged into shared memory once
//     per CTA; a warp finds the parameter of its first element with one binary
//     search in shared memory and then only walks forward.
//   * A warp tile is 32 lanes x U vectors x 4 elements.  If every vector of the
//     tile lies inside one 4-aligned parameter of one dtype, the tile runs in
//     vector mode: all loads of all U vectors are issued first (up to 128-bit
//     each, fully coalesced: one warp instruction covers 256/512 contiguous
//     bytes), then the arithmetic, then the stores.  Otherwise (ragged parameter
//     sizes, mixed dtypes, the tail) the tile runs in scalar mode with a
//     lane-coalesced element mapping.  Tiny parameters therefore cost nothing
//     extra: neighbouring lanes simply resolve to different table entries.
//   * Arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction)
//     in the operation order of the reference kernels.
#include "gp_common.cuh"

namespace {

constexpr int kMaxSmemSegs = 8191;  // csum staged in smem up to 64 KB

struct WalkArgs {
  const int64_t* csum;
  const gp_seg_t* segs;
  int n_segs;
  int use_smem;
  int64_t begin, end;
  int64_t per_cta;
};

__device__ __forceinline__ int seg_find(const int64_t* cs, int n, int64_t flat) {
  // largest j in [0, n) with cs[j] <= flat  (cs[0] <= flat < cs[n] is guaranteed)
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (cs[mid] <= flat) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__device__ __forceinline__ int seg_seek(const int64_t* cs, int n, int j, int64_t flat) {
  while (j + 1 < n && flat >= cs[j + 1]) ++j;
  return j;
}

// ------------------------------------------------------------------ ops --
// An Op supplies
//   key(seg)                       the dtype id a vector tile must agree on
//   vec<B, P, U>(...)              U vectors of one dtype, loads batched
//   scalar<B>(seg, e, bidx)        one element, any dtype
// B = buffer element type (per launch), P = parameter element type (per tile).

template <class B, class P>
__device__ __forceinline__ typename Carrier<P>::type buf_to_param(typename Carrier<B>::type x) {
  // (dtype0)(buffer value): cast of the unpack kernel, _memory_utility.py:392-425
  return round_through<P>(x);
}

struct PackOp {
  void* buffer;
  ScaleArg s;

  static __device__ __forceinline__ int key(const gp_seg_t& g) { return g.dtype0; }

  template <class B, class P, int U>
  __device__ __forceinline__ void vec(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                      const bool (&act)[U]) const {
    using CP = typename Carrier<P>::type;
    Raw4<P> in[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (act[u]) in[u] = ld4_stream(reinterpret_cast<const P*>(seg[u]->ptr[0]) + e[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      CP x[4];
      unpack4(in[u], x);
      if (s.mode != 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // pre-scale evaluated in double (or exactly in float for 2^-k), then ONE
          // rounding to the buffer type
          if (s.mode == 1 && sizeof(CP) == 4) x[i] = (CP)__fmul_rn((float)x[i], s.fs);
          else x[i] = (CP)0, x[i] = x[i];  // placeholder, replaced below
        }
      }
      st4(reinterpret_cast<B*>(buffer) + seg[u]->buf_off + e[u], pack4<B, CP>(x));
    }
  }
};

}  // namespace
