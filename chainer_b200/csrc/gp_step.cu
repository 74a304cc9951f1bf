// gp_step.cu -- the whole per-step gradient path as ONE launch per rank:
//   pack (+cast)  ->  sum over ranks  ->  unpack + 1/N + optimizer update
// flowing tile by tile through the packed buffer instead of kernel by kernel.
//
// Reference being replaced (chainer v7.8.1), one step of
// `_MultiNodeOptimizer.update` (chainermn/optimizers.py:17-33):
//   pack      chainermn/communicators/_memory_utility.py:253-268, 289-358
//   allreduce chainermn/communicators/pure_nccl_communicator.py:180-182
//   1/N       chainermn/communicators/pure_nccl_communicator.py:183-189
//   unpack    chainermn/communicators/_memory_utility.py:271-286, 361-429
//   update    chainer/optimizers/momentum_sgd.py:75-88, chainer/optimizers/adam.py:224-332
// (3 + n_params launches and a kernel boundary between every stage there).
//
// The packed buffer is cut into TILES of `tile_elems` elements (a multiple of 4096:
// 16-byte vectors for every buffer dtype, whole walker tiles).
//
// One rank (gp_step*, comm == NULL): pack and update are fused at the register level
// (PackUpd below): the packed buffer is written once and never read back.
//
// N ranks (2, 4, 8 on one NVSwitch box): tile t is OWNED by rank t % N.
//   worker CTAs   pack quarter-tile items in order; after each item one
//                 `st.release.sys` of the epoch into the owner's word packed[t][rank][q];
//                 then, per item, wait for the owner's flag[t] and run the fused
//                 update on it;
//   reducer CTAs  (the first `reducers` CTAs) take the rank's own tiles in order:
//                 wait until packed[t][0..N-1][0..3] show that all N ranks have packed it, sum the N
//                 copies -- `multimem.ld_reduce` + `multimem.st` through the NVSwitch
//                 (transport MC) or N peer loads added in rank order + N peer stores
//                 (transport P2P, bit-exact with the oracle) -- then publish
//                 flag[t] = epoch to every rank.
// The NVLink-bound reduction of tile t therefore overlaps the HBM-bound pack of later
// tiles and update of earlier ones inside ONE launch, with no kernel-level barrier
// (the allreduce kernels need two) and no chunk launches.  Every word carries the EPOCH
// of the step that wrote it (packed[t][r] == flag[t] == epoch once tile t is through), so a
// wait never depends on the history of a tile: the element count may change from step to
// step.  The words live in a per-rank block shared through CUDA IPC, and every wait polls
// LOCAL memory.
// Every CTA of the launch is resident at once (the grid is capped by the occupancy),
// which the cross-rank waits require.
#include <string.h>

#include "gp_pack_op.cuh"
#include "gp_sgd_op.cuh"
#include "gp_adam_op.cuh"

namespace {

#include "gp_p2p.cuh"

constexpr int kStepThreads = 256;
// Workers pack and update in ITEMS of 1 / kSub tile.  Measured (N = 2, profiles/
// r02_step_sweep_q_n2_*.json): quarter-tile items shorten both ends of the pipeline but pay
// a system-scope release per item -- 0.373 ms vs 0.227 ms with whole-tile items -- so kSub = 1.
constexpr int kSub = 1;

struct StepTables {
  const int64_t* csum;
  const gp_seg_t* segs;
  int n_segs;
  int use_smem;
  int64_t n_elems;
  int64_t tile_elems;
  int64_t n_tiles;
};

__device__ __forceinline__ const int64_t* stage_csum(const StepTables& a, int64_t* s_csum) {
  if (!a.use_smem) return a.csum;
  for (int i = threadIdx.x; i <= a.n_segs; i += blockDim.x) s_csum[i] = a.csum[i];
  __syncthreads();
  return s_csum;
}

// ------------------------------------------------------------------ one rank --
// Pack and update fused at the REGISTER level: a warp loads its gradient vectors, casts
// them to the buffer type, stores them into the packed buffer (it stays observable as
// `gpu_buffer_a`, bit-exact with gp_pack) and feeds the same registers to the fused
// update -- the packed buffer is written once and never read back.  An ordinary walker
// launch (gp_walk.cuh) with this combined op.
template <class Upd>
struct PackUpd {
  static constexpr int kMaxUnroll = Upd::kMaxUnroll;
  static constexpr int kDefaultUnroll = Upd::kDefaultUnroll;
  Upd up;       // up.buffer: the packed buffer; up.s: the 1/size scale
  ScaleArg s;   // == up.s (the launcher reads op.s.mode)

  static __device__ __forceinline__ int key(const gp_seg_t& g) { return Upd::key(g); }

  template <class B, class P, int U, int SM>
  __device__ __forceinline__ void vec(const gp_seg_t* const (&seg)[U], const int64_t (&e)[U],
                                      const bool (&act)[U]) const {
    using CP = typename Carrier<P>::type;
    typename Upd::template Regs<B, P, U> r;
    Raw4<P> rg[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (act[u]) rg[u] = ld4_stream(cptr<P>(seg[u]->ptr[0]) + e[u]);
    up.template load<B, P, U, false>(seg, e, act, r);
    B* buf = reinterpret_cast<B*>(const_cast<void*>(up.buffer));
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (act[u]) {
        CP x[4];
        unpack4(rg[u], x);
        r.rb[u] = pack4<B, CP>(x);                       // the cast of gp_pack (RN)
        st4(buf + seg[u]->buf_off + e[u], r.rb[u]);
      }
    up.template finish<B, P, U, SM>(seg, e, act, r);
  }
  template <class B, class P, int SM>
  __device__ __forceinline__ void one(const gp_seg_t& g, int64_t e) const {
    B* buf = reinterpret_cast<B*>(const_cast<void*>(up.buffer));
    buf[g.buf_off + e] = from_carrier<B>(to_carrier(cptr<P>(g.ptr[0])[e]));
    up.template one<B, P, SM>(g, e);                     // reads it back (same thread)
  }
};

// ------------------------------------------------------------------- N ranks --
struct StepPeers {
  uint32_t* words[kMaxRanks];  // every rank's [packed: tile_cap x kMaxRanks x kSub | flag: tile_cap]
  void* bufs[kMaxRanks];       // P2P: this process's mappings of the packed buffers
  char* mc_base;               // MC: multicast address of the packed buffer
  int64_t tile_cap;
  int rank, n;
  uint32_t epoch;
  int reducers;
  unsigned long long timeout_ns;
};

// 16-byte multimem accessors (see gp_mc.cu)
template <class T> struct Mm;
template <> struct Mm<float> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};
template <> struct Mm<__half> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};
template <> struct Mm<__nv_bfloat16> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};
__device__ __forceinline__ void mm_st(void* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Transport MC: the switch adds the N copies and replicates the sum.  Software
// pipelined: the loads of the next UN vectors are in flight while the current UN
// are stored, so both NVLink directions stay busy.
template <class B, int UN>
struct RedMc {
  static __device__ __forceinline__ void load(uint4 (&x)[UN], const char* base, int64_t i, int64_t v1) {
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t v = i + (int64_t)u * kStepThreads;
      if (v < v1) x[u] = Mm<B>::ld_reduce(base + v * 16);
    }
  }
  static __device__ __forceinline__ void store(const uint4 (&x)[UN], char* base, int64_t i, int64_t v1) {
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t v = i + (int64_t)u * kStepThreads;
      if (v < v1) mm_st(base + v * 16, x[u]);
    }
  }
  static __device__ __forceinline__ void run(const StepPeers& p, int64_t v0, int64_t v1) {
    constexpr int64_t adv = (int64_t)kStepThreads * UN;
    int64_t i = v0 + threadIdx.x;
    if (i >= v1) return;
    uint4 x[UN], y[UN];
    load(x, p.mc_base, i, v1);
    while (true) {
      load(y, p.mc_base, i + adv, v1);
      store(x, p.mc_base, i, v1);
      i += adv;
      if (i >= v1) break;
      load(x, p.mc_base, i + adv, v1);
      store(y, p.mc_base, i, v1);
      i += adv;
      if (i >= v1) break;
    }
  }
};

// Transport P2P: N peer loads, added in RANK ORDER (the oracle's order: identical bits
// on every rank and for every N), N peer stores.
template <class T> struct Vec16;
template <> struct Vec16<float> {
  static __device__ __forceinline__ void add(uint4& acc, const uint4& x) {
    float4& a = reinterpret_cast<float4&>(acc);
    const float4& b = reinterpret_cast<const float4&>(x);
    a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y);
    a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
  }
};
template <> struct Vec16<__half> {
  static __device__ __forceinline__ void add(uint4& acc, const uint4& x) {
    __half2* a = reinterpret_cast<__half2*>(&acc);
    const __half2* b = reinterpret_cast<const __half2*>(&x);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __hadd2_rn(a[i], b[i]);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static __device__ __forceinline__ void add(uint4& acc, const uint4& x) {
    __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&acc);
    const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __hadd2_rn(a[i], b[i]);
  }
};
__device__ __forceinline__ uint4 ld_sys(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.relaxed.sys.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(uint4* p, const uint4& v) {
  asm volatile("st.global.relaxed.sys.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w) : "memory");
}
template <class B, int N>
struct RedP2p {
  // N * UN vector loads in flight per thread and pipeline stage
  static constexpr int UN = N >= 8 ? 1 : 2;
  static __device__ __forceinline__ void load(uint4 (&x)[UN][N], const StepPeers& p, int64_t i,
                                              int64_t v1) {
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t v = i + (int64_t)u * kStepThreads;
      if (v < v1) {
#pragma unroll
        for (int k = 0; k < N; ++k) x[u][k] = ld_sys(reinterpret_cast<const uint4*>(p.bufs[k]) + v);
      }
    }
  }
  static __device__ __forceinline__ void store(const uint4 (&x)[UN][N], const StepPeers& p,
                                               int64_t i, int64_t v1) {
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t v = i + (int64_t)u * kStepThreads;
      if (v < v1) {
        uint4 acc = x[u][0];
#pragma unroll
        for (int k = 1; k < N; ++k) Vec16<B>::add(acc, x[u][k]);
#pragma unroll
        for (int k = 0; k < N; ++k) st_sys(reinterpret_cast<uint4*>(p.bufs[k]) + v, acc);
      }
    }
  }
  static __device__ __forceinline__ void run(const StepPeers& p, int64_t v0, int64_t v1) {
    constexpr int64_t adv = (int64_t)kStepThreads * UN;
    int64_t i = v0 + threadIdx.x;
    if (i >= v1) return;
    if constexpr (N >= 4) {
      // 4 / 8 peers: N loads per vector are already a deep queue; no second stage
      for (; i < v1; i += adv) {
        uint4 x[UN][N];
        load(x, p, i, v1);
        store(x, p, i, v1);
      }
    } else {
      uint4 x[UN][N], y[UN][N];
      load(x, p, i, v1);
      while (true) {
        load(y, p, i + adv, v1);
        store(x, p, i, v1);
        i += adv;
        if (i >= v1) break;
        load(x, p, i + adv, v1);
        store(y, p, i, v1);
        i += adv;
        if (i >= v1) break;
      }
    }
  }
};

template <class Upd, class B, class Red>
__global__ void __launch_bounds__(kStepThreads) stepn_kernel(const StepTables a, const StepPeers p,
                                                            const PackOp pk, const Upd up) {
  extern __shared__ int64_t s_csum[];
  if ((int)blockIdx.x < p.reducers) {
    // ------------------------------------------------------------- reducer ----
    const uint32_t* packed = p.words[p.rank];
    constexpr int E = 16 / (int)sizeof(B);
    for (int64_t t = p.rank + (int64_t)blockIdx.x * p.n; t < a.n_tiles; t += (int64_t)p.reducers * p.n) {
      // thread (r, q) waits for rank r's "quarter q of tile t packed" word
      if ((int)threadIdx.x < p.n * kSub)
        spin_until(packed + t * (kMaxRanks * kSub) + threadIdx.x, p.epoch, p.timeout_ns);
      __syncthreads();
      const int64_t lo = t * a.tile_elems;
      const int64_t hi = lo + a.tile_elems < a.n_elems ? lo + a.tile_elems : a.n_elems;
      // whole 16-byte vectors: the tail of the last tile is rounded up into the padding
      // of the buffer (never read back)
      Red::run(p, lo / E, (hi + E - 1) / E);
      __syncthreads();
      // release: cumulative over the CTA's stores (ordered before it by bar.sync)
      if ((int)threadIdx.x < p.n)
        st_release_sys(p.words[threadIdx.x] + p.tile_cap * (kMaxRanks * kSub) + t, p.epoch);
    }
    return;
  }
  // ---------------------------------------------------------------- worker ----
  const int64_t* cs = stage_csum(a, s_csum);
  const int64_t w = (int64_t)blockIdx.x - p.reducers;
  const int64_t nw = (int64_t)gridDim.x - p.reducers;
  const int64_t sub = a.tile_elems / kSub;
  const int64_t n_items = (a.n_elems + sub - 1) / sub;
  for (int64_t i = w; i < n_items; i += nw) {
    const int64_t lo = i * sub;
    const int64_t hi = lo + sub < a.n_elems ? lo + sub : a.n_elems;
    gpw::walk_range<PackOp, B, 4, 0, GP_F32>(cs, a.segs, a.n_segs, lo, hi, pk);
    __syncthreads();
    // release: cumulative over the CTA's stores (ordered before it by bar.sync)
    const int64_t t = i / kSub;
    if (threadIdx.x == 0)
      st_release_sys(p.words[t % p.n] + t * (kMaxRanks * kSub) + p.rank * kSub + (i - t * kSub), p.epoch);
  }
  // a short last tile: its missing quarters are signalled as packed too
  if (w == 0 && threadIdx.x == 0) {
    const int64_t t = a.n_tiles - 1;
    for (int64_t q = n_items - t * kSub; q < kSub; ++q)
      st_release_sys(p.words[t % p.n] + t * (kMaxRanks * kSub) + p.rank * kSub + q, p.epoch);
  }
  const uint32_t* flag = p.words[p.rank] + p.tile_cap * (kMaxRanks * kSub);
  for (int64_t i = w; i < n_items; i += nw) {
    // the acquire also drops this SM's L1 lines of the tile (written by pack earlier)
    if (threadIdx.x == 0) spin_until(flag + i / kSub, p.epoch, p.timeout_ns);
    __syncthreads();
    const int64_t lo = i * sub;
    const int64_t hi = lo + sub < a.n_elems ? lo + sub : a.n_elems;
    gpw::walk_range<Upd, B, Upd::kDefaultUnroll, 1, GP_F32>(cs, a.segs, a.n_segs, lo, hi, up);
  }
}

// ------------------------------------------------------------------ launching --
struct StepTuning {
  int tile_elems;   // multiple of 4096
  int reducers;     // reducer CTAs per rank (N ranks); 0: the transport's default below
  int unroll;       // MC transport: 16-byte vectors per thread and pipeline stage (4, 8)
  int ctas_per_sm;  // cap of resident CTAs per SM (0: the occupancy)
  int reducers_mc, reducers_p2p;
};
StepTuning g_step = {16384, 0, 8, 4, 256, 192};

template <class K>
int occupancy(K kernel, size_t smem) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kStepThreads, smem) != cudaSuccess ||
      occ < 1) {
    (void)cudaGetLastError();
    occ = 1;
  }
  return occ;
}

StepTables make_tables(const int64_t* d_csum, const gp_seg_t* d_segs, int n_segs, int64_t n_elems,
                       int64_t tile, size_t* smem) {
  StepTables a;
  a.csum = d_csum;
  a.segs = d_segs;
  a.n_segs = n_segs;
  a.use_smem = n_segs <= gpw::kMaxSmemSegs;
  a.n_elems = n_elems;
  a.tile_elems = tile;
  a.n_tiles = (n_elems + tile - 1) / tile;
  *smem = a.use_smem ? (size_t)(n_segs + 1) * sizeof(int64_t) : 0;
  return a;
}

template <class Upd, class B>
int launch_step1(const int64_t* d_csum, const gp_seg_t* d_segs, int n_segs, int64_t n_elems,
                 const Upd& up, cudaStream_t st, const char* what) {
  PackUpd<Upd> op;
  op.up = up;
  op.s = up.s;
  const GpTuning& t = g_gp_tuning;
  int threads = t.threads < 32 ? 32 : (t.threads > gpw::kMaxThreads ? gpw::kMaxThreads : t.threads);
  threads &= ~31;
  gpw::WalkArgs a;
  a.csum = d_csum;
  a.segs = d_segs;
  a.n_segs = n_segs;
  a.use_smem = n_segs <= gpw::kMaxSmemSegs;
  a.begin = 0;
  a.end = n_elems;
  a.per_cta = 0;
  const size_t smem = a.use_smem ? (size_t)(n_segs + 1) * sizeof(int64_t) : 0;
  const bool u4 = t.unroll >= 4;
  if (up.s.mode == 0) {
    if (u4) return gpw::launch_u<PackUpd<Upd>, B, 4, 0, GP_F32>(a, op, threads, smem, st, what);
    return gpw::launch_u<PackUpd<Upd>, B, 2, 0, GP_F32>(a, op, threads, smem, st, what);
  }
  if (u4) return gpw::launch_u<PackUpd<Upd>, B, 4, 1, GP_F32>(a, op, threads, smem, st, what);
  return gpw::launch_u<PackUpd<Upd>, B, 2, 1, GP_F32>(a, op, threads, smem, st, what);
}

template <class Upd, class B, class Red>
int launch_stepn_t(const StepTables& a, StepPeers p, size_t smem, const PackOp& pk, const Upd& up,
                   cudaStream_t st) {
  auto kernel = stepn_kernel<Upd, B, Red>;
  int occ = occupancy(kernel, smem);
  if (g_step.ctas_per_sm > 0 && g_step.ctas_per_sm < occ) occ = g_step.ctas_per_sm;
  const int64_t cap = (int64_t)gp_sm_count_cached() * occ;   // everything resident at once
  // defaults from the sweeps of tools/step_sweep.py (profiles/r02_step_sweep_n*.json):
  // multicast 256 / N reducer CTAs (N = 8: 32, N = 4: 64), peer memory 192
  int64_t reducers = g_step.reducers > 0 ? g_step.reducers
                                          : (p.mc_base ? g_step.reducers_mc / p.n : g_step.reducers_p2p);
  const int64_t own_tiles = (a.n_tiles + p.n - 1) / p.n;
  if (reducers > own_tiles) reducers = own_tiles;
  if (reducers < 1) reducers = 1;
  if (reducers > cap / 2) reducers = cap / 2 > 0 ? cap / 2 : 1;
  int64_t workers = a.n_tiles * kSub;
  if (workers > cap - reducers) workers = cap - reducers;
  if (workers < 1) workers = 1;
  p.reducers = (int)reducers;
  kernel<<<(unsigned)(reducers + workers), kStepThreads, smem, st>>>(a, p, pk, up);
  return gp_cuda_fail(cudaGetLastError(), "stepn_kernel launch");
}

template <class Upd, class B>
int launch_stepn(const StepTables& a, const StepPeers& p, size_t smem, const PackOp& pk,
                 const Upd& up, cudaStream_t st) {
  if (p.mc_base) {
    if (g_step.unroll >= 8) return launch_stepn_t<Upd, B, RedMc<B, 8>>(a, p, smem, pk, up, st);
    if (g_step.unroll <= 2) return launch_stepn_t<Upd, B, RedMc<B, 2>>(a, p, smem, pk, up, st);
    return launch_stepn_t<Upd, B, RedMc<B, 4>>(a, p, smem, pk, up, st);
  }
  switch (p.n) {
    case 2: return launch_stepn_t<Upd, B, RedP2p<B, 2>>(a, p, smem, pk, up, st);
    case 4: return launch_stepn_t<Upd, B, RedP2p<B, 4>>(a, p, smem, pk, up, st);
    default: return launch_stepn_t<Upd, B, RedP2p<B, 8>>(a, p, smem, pk, up, st);
  }
}

// what the step kernels cover; everything else stays on the pack / allreduce / update
// launches (the host checks with gp_step_supported first)
bool step_covers(int n_ranks, int buf_dtype, int layout_hint, double scale) {
  if (n_ranks != 1 && n_ranks != 2 && n_ranks != 4 && n_ranks != 8) return false;
  if (buf_dtype != GP_F32 && buf_dtype != GP_F16 && buf_dtype != GP_BF16) return false;
  if (layout_hint != GP_F32) return false;
  const ScaleArg s = make_scale(scale);
  if (n_ranks == 1) return s.mode != 2;
  return s.mode == 1;
}

template <class Upd>
int step_dispatch(void* p2p_comm, void* mc_ptr, void* buffer, int buf_dtype, const int64_t* d_csum,
                  const gp_seg_t* d_segs, int n_segs, int64_t n_elems, double scale, int layout_hint,
                  Upd up, void* stream, const char* what) {
  if (n_segs <= 0 || n_elems <= 0) return 0;
  P2PComm* c = (P2PComm*)p2p_comm;
  const int n_ranks = c ? c->n : 1;
  if (!step_covers(n_ranks, buf_dtype, layout_hint, scale)) {
    gp_set_error("%s: not covered by the one-launch step (ranks %d, buffer dtype %d, layout hint %d, "
                 "scale %g)", what, n_ranks, buf_dtype, layout_hint, scale);
    return GP_EINVAL;
  }
  PackOp pk;
  pk.buffer = buffer;
  pk.s = make_scale(1.0);
  up.buffer = buffer;
  up.s = make_scale(scale);
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = 0;
  if (!c) {
    switch (buf_dtype) {
      case GP_F32: return launch_step1<Upd, float>(d_csum, d_segs, n_segs, n_elems, up, st, what);
      case GP_F16: return launch_step1<Upd, __half>(d_csum, d_segs, n_segs, n_elems, up, st, what);
      default: return launch_step1<Upd, __nv_bfloat16>(d_csum, d_segs, n_segs, n_elems, up, st, what);
    }
  }
  if (c->step_tile_cap <= 0 || c->step_tile_elems <= 0) {
    gp_set_error("%s: per-tile words not set (gp_p2p_set_step_words)", what);
    return GP_EINVAL;
  }
  const StepTables a = make_tables(d_csum, d_segs, n_segs, n_elems, c->step_tile_elems, &smem);
  if (a.n_tiles > c->step_tile_cap) {
    gp_set_error("%s: %lld tiles exceed the capacity %lld of the per-tile words", what,
                 (long long)a.n_tiles, (long long)c->step_tile_cap);
    return GP_EINVAL;
  }
  StepPeers p;
  for (int k = 0; k < kMaxRanks; ++k) {
    p.words[k] = k < c->n ? c->step_words[k] : nullptr;
    p.bufs[k] = k < c->n ? c->bufs[k] : nullptr;
  }
  p.mc_base = (char*)mc_ptr;
  p.tile_cap = c->step_tile_cap;
  p.rank = c->rank;
  p.n = c->n;
  p.reducers = 0;
  p.timeout_ns = g_gp_peer_timeout_ns;
  if (!mc_ptr) {
    if (c->bufs[c->rank] != buffer) {
      gp_set_error("%s: buffer is not the rank's registered packed buffer (gp_p2p_set_buffers)", what);
      return GP_EINVAL;
    }
  }
  p.epoch = ++c->step_epoch;
  switch (buf_dtype) {
    case GP_F32: return launch_stepn<Upd, float>(a, p, smem, pk, up, st);
    case GP_F16: return launch_stepn<Upd, __half>(a, p, smem, pk, up, st);
    default: return launch_stepn<Upd, __nv_bfloat16>(a, p, smem, pk, up, st);
  }
}

}  // namespace

extern "C" {

int gp_step_supported(int n_ranks, int buf_dtype, int layout_hint, double scale, int adam_flags) {
  if (adam_flags & GP_ADAM_AMSGRAD) return 0;
  return step_covers(n_ranks, buf_dtype, layout_hint, scale) ? 1 : 0;
}

size_t gp_step_words_bytes(int64_t tile_cap) {
  return (size_t)tile_cap * (kMaxRanks * kSub + 1) * sizeof(uint32_t);
}

int gp_step_tile_elems(void) { return g_step.tile_elems; }

// Collective (host side): `blocks[k]` is this process's mapping of rank k's ZEROED word
// block of gp_step_words_bytes(tile_cap) bytes; restarts the step epoch.
int gp_p2p_set_step_words(void* comm, void* const* blocks, int64_t tile_cap, int64_t tile_elems) {
  P2PComm* c = (P2PComm*)comm;
  if (!c || tile_cap <= 0 || tile_elems < 4096 || (tile_elems % 4096)) {
    gp_set_error("gp_p2p_set_step_words: bad arguments (tile_cap %lld, tile_elems %lld)",
                 (long long)tile_cap, (long long)tile_elems);
    return GP_EINVAL;
  }
  for (int k = 0; k < c->n; ++k) c->step_words[k] = (uint32_t*)blocks[k];
  c->step_tile_cap = tile_cap;
  c->step_tile_elems = tile_elems;
  c->step_epoch = 0;
  return 0;
}

int gp_step_momentum_sgd(void* p2p_comm, void* mc_ptr, void* buffer, int buf_dtype,
                         const int64_t* d_csum, const gp_seg_t* d_segs, int n_segs, int64_t n_elems,
                         double scale, double lr, double momentum, int write_grad, int layout_hint,
                         void* stream) {
  SgdOp<false> up;
  up.lr = lr;
  up.momentum = momentum;
  up.write_grad = write_grad;
  up.hooks = {nullptr, 0.0, 0.0};
  return step_dispatch(p2p_comm, mc_ptr, buffer, buf_dtype, d_csum, d_segs, n_segs, n_elems, scale,
                       layout_hint, up, stream, "gp_step_momentum_sgd");
}

int gp_step_adam(void* p2p_comm, void* mc_ptr, void* buffer, int buf_dtype, const int64_t* d_csum,
                 const gp_seg_t* d_segs, int n_segs, int64_t n_elems, double scale, double alpha_t,
                 double one_minus_beta1, double one_minus_beta2, double eps, double eta,
                 double weight_decay_rate, double lower, double upper, int adam_flags,
                 int write_grad, int layout_hint, void* stream) {
  if (adam_flags & GP_ADAM_AMSGRAD) {
    gp_set_error("gp_step_adam: AMSGrad is not covered by the one-launch step");
    return GP_EINVAL;
  }
  AdamOp<false, false> up;
  up.alpha_t = alpha_t; up.omb1 = one_minus_beta1; up.omb2 = one_minus_beta2; up.eps = eps;
  up.eta = eta; up.wd = weight_decay_rate; up.lower = lower; up.upper = upper;
  up.flags = adam_flags; up.write_grad = write_grad;
  up.hooks = {nullptr, 0.0, 0.0};
  return step_dispatch(p2p_comm, mc_ptr, buffer, buf_dtype, d_csum, d_segs, n_segs, n_elems, scale,
                       layout_hint, up, stream, "gp_step_adam");
}

int gp_step_set_tuning(const char* key, int value) {
  if (!key) return GP_EINVAL;
  if (!strcmp(key, "tile_elems")) g_step.tile_elems = value < 4096 ? 4096 : value / 4096 * 4096;
  else if (!strcmp(key, "reducers")) g_step.reducers = value < 0 ? 0 : value;
  else if (!strcmp(key, "unroll")) g_step.unroll = value;
  else if (!strcmp(key, "ctas_per_sm")) g_step.ctas_per_sm = value;
  else {
    gp_set_error("gp_step_set_tuning: unknown key '%s'", key);
    return GP_EINVAL;
  }
  return 0;
}

}  // extern "C"
