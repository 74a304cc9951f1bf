// gp_walk.cuh -- the segmented stream walker shared by the pack, unpack and
// fused-update kernels of libgradpath (sm_100a).
//
// Reference being replaced: the thread-per-element kernels
// `cupy_batched_pack_params` / `cupy_batched_unpack_params`
// (chainermn/communicators/_memory_utility.py:289-429), which do a binary
// search through a global-memory int32 table for EVERY element and move one
// scalar per thread, and the per-parameter ElementwiseKernel launches of the
// optimizers (chainer/optimizers/momentum_sgd.py:75-88, adam.py:224-332).
//
// Design (HBM-bound, no data reuse):
//   * ONE launch walks the flat element space [begin, end) of the parameter
//     list.  By default the grid is persistent (SMs x ctas_per_sm CTAs) and each
//     CTA owns an equal contiguous slice: no wave quantisation, no tail.
//   * The cumulative-size table (int64[n+1]) is staged into shared memory once
//     per CTA; a warp finds the parameter of its first element with one binary
//     search in shared memory and afterwards only walks forward.
//   * A warp tile is U rows of 32 lanes x 4 elements.  If every row of the tile
//     lies inside a 4-aligned parameter and all rows have one dtype (resolved
//     per ROW with warp-uniform lookups, not per lane),
//     the tile runs in VECTOR mode: the loads of all U vectors of all arrays are
//     issued first (8/16 B per lane, one warp instruction = 256/512 contiguous
//     bytes), then the arithmetic, then the stores.  Otherwise (ragged sizes,
//     mixed dtypes, the tail) the tile runs in SCALAR mode with a lane-coalesced
//     element mapping, where neighbouring lanes simply resolve to different
//     table entries (runs of tiny parameters).
#pragma once
#include "gp_common.cuh"

namespace gpw {

constexpr int kMaxSmemSegs = 6143;  // (n+1) * 8 B <= 48 KB of dynamic shared memory
constexpr int kMaxThreads = 512;

struct WalkArgs {
  const int64_t* csum;
  const gp_seg_t* segs;
  int n_segs;
  int use_smem;
  int64_t begin, end;
  int64_t per_cta;
};

// largest j in [0, n) with cs[j] <= flat   (cs[0] <= flat < cs[n] guaranteed)
__device__ __forceinline__ int seg_find(const int64_t* cs, int n, int64_t flat) {
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

// An Op supplies
//   static int key(const gp_seg_t&)          dtype id a vector tile must agree on
//   ScaleArg s                               (s.mode selects the SM instantiation)
//   template <class B, class P, int U, int SM> void vec(seg[U], e[U], act[U]) const
//   template <class B, int SM> void scalar(const gp_seg_t&, int64_t e) const
// B = buffer element type (per launch), P = parameter element type (per tile).
// The CTA's share of the walk: every warp of the CTA takes warp tiles of the flat
// range [lo, hi) (lo a multiple of 4; the kernels call this with whole-CTA control
// flow, but there is no barrier inside).  `cs` is the cumulative-size table (shared
// or global memory).
template <class Op, class B, int U, int SM, int PT>
__device__ __forceinline__ void walk_range(const int64_t* cs, const gp_seg_t* segs, int n,
                                           int64_t lo, int64_t hi, const Op& op) {
  constexpr int WT = 32 * U * 4;  // elements per warp tile
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t stride = (int64_t)(blockDim.x >> 5) * WT;
  int64_t base = lo + (int64_t)warp * WT;
  if (base >= hi) return;
  int j = seg_find(cs, n, base);

  for (; base < hi; base += stride) {
    j = seg_seek(cs, n, j, base);

    // Resolve the tile row by row (a row = 32 lanes x 4 elements = 128 flat
    // elements).  Everything here is warp-uniform: one table lookup per ROW, not
    // per lane.  A row is a vector row if it lies inside one 4-aligned
    // parameter; the tile runs in vector mode if all its U rows are vector rows
    // of one dtype (rows may belong to different parameters).
    const gp_seg_t* sg[U];
    int64_t e[U];
    bool act[U];
    bool ok = base + WT <= hi;
    int key0 = 0;
    int jr = j;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = base + u * 128;
      sg[u] = segs + jr;
      e[u] = 0;
      act[u] = true;
      if (ok) {
        jr = seg_seek(cs, n, jr, row);
        const gp_seg_t* g = segs + jr;
        const int k = Op::key(*g);
        if (u == 0) key0 = k;
        ok = (row + 128 <= cs[jr + 1]) && (g->flags & GP_SEG_VEC_OK) && k == key0 && k != GP_F64;
        sg[u] = g;
        e[u] = row - cs[jr] + (int64_t)lane * 4;
      }
    }

    if (ok) {
      if constexpr (PT == GP_F32) {
        // the caller promised float32 everywhere: no dtype dispatch in the kernel
        op.template vec<B, float, U, SM>(sg, e, act);
      } else {
        if (key0 == GP_F32) op.template vec<B, float, U, SM>(sg, e, act);
        else op.template vec<B, __half, U, SM>(sg, e, act);
      }
    } else {
      // ragged sizes, mixed dtypes, float64, tiny parameters, the tail:
      // lane-coalesced scalar elements
      int js = j;
#pragma unroll 1
      for (int k = 0; k < U * 4; ++k) {
        const int64_t flat = base + (int64_t)k * 32 + lane;
        if (flat < hi) {
          js = seg_seek(cs, n, js, flat);
          if constexpr (PT == GP_F32) op.template one<B, float, SM>(segs[js], flat - cs[js]);
          else op.template scalar<B, SM>(segs[js], flat - cs[js]);
        }
      }
    }
  }
}

template <class Op, class B, int U, int SM, int PT>
__global__ void __launch_bounds__(kMaxThreads) walk_kernel(const WalkArgs a, const Op op) {
  extern __shared__ int64_t s_csum[];
  const int n = a.n_segs;
  const int64_t lo = a.begin + (int64_t)blockIdx.x * a.per_cta;
  const int64_t hi = (lo + a.per_cta < a.end) ? lo + a.per_cta : a.end;
  if (lo >= hi) return;

  const int64_t* cs = a.csum;
  if (a.use_smem) {
    for (int i = threadIdx.x; i <= n; i += blockDim.x) s_csum[i] = a.csum[i];
    __syncthreads();
    cs = s_csum;
  }
  walk_range<Op, B, U, SM, PT>(cs, a.segs, n, lo, hi, op);
}

// resident CTAs per SM of one instantiation (cached: the occupancy query is a
// host-side calculation but not free)
template <class Op, class B, int U, int SM, int PT>
int resident_ctas(int threads, size_t smem) {
  static int c_threads = -1, c_occ = 1;
  static size_t c_smem = 0;
  if (threads != c_threads || smem != c_smem) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, walk_kernel<Op, B, U, SM, PT>, threads, smem) !=
            cudaSuccess || occ < 1) {
      (void)cudaGetLastError();
      occ = 1;
    }
    c_threads = threads;
    c_smem = smem;
    c_occ = occ;
  }
  return c_occ;
}

template <class Op, class B, int U, int SM, int PT>
int launch_u(WalkArgs a, const Op& op, int threads, size_t smem, cudaStream_t st, const char* what) {
  const GpTuning& t = g_gp_tuning;
  const int64_t total = a.end - a.begin;
  const int64_t cta_tile = (int64_t)threads * U * 4;
  int64_t grid;
  if (t.persistent) {
    // every CTA must be resident at once, otherwise the equal slices would run
    // in waves: cap the grid by the real occupancy of this instantiation.
    int per_sm = resident_ctas<Op, B, U, SM, PT>(threads, smem);
    if (t.ctas_per_sm > 0 && t.ctas_per_sm < per_sm) per_sm = t.ctas_per_sm;
    const int64_t max_grid = (int64_t)gp_sm_count_cached() * per_sm;
    grid = (total + cta_tile - 1) / cta_tile;
    if (grid > max_grid) grid = max_grid;
    int64_t per = (total + grid - 1) / grid;
    per = (per + 127) & ~(int64_t)127;
    a.per_cta = per;
    grid = (total + per - 1) / per;
  } else {
    a.per_cta = cta_tile;
    grid = (total + cta_tile - 1) / cta_tile;
  }
  if (grid > 0x7fffffff) {
    gp_set_error("%s: grid too large", what);
    return GP_EINVAL;
  }
  walk_kernel<Op, B, U, SM, PT><<<(unsigned)grid, threads, smem, st>>>(a, op);
  return gp_cuda_fail(cudaGetLastError(), what);
}

// Choose the grid and launch.  `what` only labels error messages.
template <class Op, class B>
int launch(const int64_t* d_csum, const gp_seg_t* d_segs, int n_segs, int64_t begin, int64_t end,
           const Op& op, void* stream, const char* what, bool f32 = false) {
  if (n_segs <= 0 || end <= begin) return 0;
  if (begin < 0 || (begin & 3)) {
    gp_set_error("%s: elem_begin (%lld) must be a non-negative multiple of 4", what,
                 (long long)begin);
    return GP_EINVAL;
  }
  const GpTuning& t = g_gp_tuning;
  int threads = t.threads;
  if (threads < 32) threads = 32;
  if (threads > kMaxThreads) threads = kMaxThreads;
  threads &= ~31;
  // unroll 0 = the op's tuned default
  int U = t.unroll <= 0 ? Op::kDefaultUnroll : (t.unroll >= 4 ? 4 : 2);
  if (sizeof(B) == 8 && U > 2) U = 2;
  if (U > Op::kMaxUnroll && !f32) U = Op::kMaxUnroll;  // the float32-only variants are leaner

  WalkArgs a;
  a.csum = d_csum;
  a.segs = d_segs;
  a.n_segs = n_segs;
  a.use_smem = n_segs <= kMaxSmemSegs;
  a.begin = begin;
  a.end = end;
  a.per_cta = 0;
  const size_t smem = a.use_smem ? (size_t)(n_segs + 1) * sizeof(int64_t) : 0;

  cudaStream_t st = (cudaStream_t)stream;
  const int mode = op.s.mode;
#define GP_LAUNCH(UU)                                                              \
  switch (mode) {                                                                  \
    case 0: return f32 ? launch_u<Op, B, UU, 0, GP_F32>(a, op, threads, smem, st, what)  \
                       : launch_u<Op, B, UU, 0, 0>(a, op, threads, smem, st, what);        \
    case 1: return f32 ? launch_u<Op, B, UU, 1, GP_F32>(a, op, threads, smem, st, what)  \
                       : launch_u<Op, B, UU, 1, 0>(a, op, threads, smem, st, what);        \
    default: return launch_u<Op, B, UU, 2, 0>(a, op, threads, smem, st, what);             \
  }
  if (U >= 4) { GP_LAUNCH(4) }
  GP_LAUNCH(2)
#undef GP_LAUNCH
}

// Ops that only exist for all-float32 parameter arrays (PT = GP_F32: no dtype dispatch,
// no generic scalar path): the three scale modes, the op's default unroll.
template <class Op, class B>
int launch_f32(const int64_t* d_csum, const gp_seg_t* d_segs, int n_segs, int64_t begin,
               int64_t end, const Op& op, void* stream, const char* what) {
  if (n_segs <= 0 || end <= begin) return 0;
  if (begin < 0 || (begin & 3)) {
    gp_set_error("%s: elem_begin (%lld) must be a non-negative multiple of 4", what,
                 (long long)begin);
    return GP_EINVAL;
  }
  const GpTuning& t = g_gp_tuning;
  int threads = t.threads < 32 ? 32 : (t.threads > kMaxThreads ? kMaxThreads : t.threads);
  threads &= ~31;
  WalkArgs a;
  a.csum = d_csum;
  a.segs = d_segs;
  a.n_segs = n_segs;
  a.use_smem = n_segs <= kMaxSmemSegs;
  a.begin = begin;
  a.end = end;
  a.per_cta = 0;
  const size_t smem = a.use_smem ? (size_t)(n_segs + 1) * sizeof(int64_t) : 0;
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int U = Op::kDefaultUnroll;
  switch (op.s.mode) {
    case 0: return launch_u<Op, B, U, 0, GP_F32>(a, op, threads, smem, st, what);
    case 1: return launch_u<Op, B, U, 1, GP_F32>(a, op, threads, smem, st, what);
    default: return launch_u<Op, B, U, 2, GP_F32>(a, op, threads, smem, st, what);
  }
}

// dispatch on the runtime buffer dtype
// f32: the caller promises that every segment's arrays are float32 (layout_hint)
template <class Op>
int launch_buf(int buf_dtype, const int64_t* d_csum, const gp_seg_t* d_segs, int n_segs,
               int64_t begin, int64_t end, const Op& op, void* stream, const char* what,
               bool f32 = false) {
  switch (buf_dtype) {
    case GP_F32: return launch<Op, float>(d_csum, d_segs, n_segs, begin, end, op, stream, what, f32);
    case GP_F16: return launch<Op, __half>(d_csum, d_segs, n_segs, begin, end, op, stream, what, f32);
    case GP_BF16: return launch<Op, __nv_bfloat16>(d_csum, d_segs, n_segs, begin, end, op, stream, what, f32);
    case GP_F64: return launch<Op, double>(d_csum, d_segs, n_segs, begin, end, op, stream, what);
    default:
      gp_set_error("%s: unsupported buffer dtype id %d", what, buf_dtype);
      return GP_EINVAL;
  }
}

// value of the packed buffer (type B, as carrier) -> mean gradient in the
// parameter's type P (as carrier): descale, round to B (the reference scales the
// buffer in place), then cast to P (unpack kernel, _memory_utility.py:392-425).
template <class B, class P, int SM>
__device__ __forceinline__ typename Carrier<P>::type mean_grad_value(
    typename Carrier<B>::type x, const ScaleArg& s) {
  return round_through<P>(descale<B, SM>(x, s));
}

}  // namespace gpw
