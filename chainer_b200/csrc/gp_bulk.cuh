// gp_bulk.cuh -- TMA-staged variant of the fused unpack + update stream.
//
// The register-path walker (gp_walk.cuh) is limited by how many bytes its
// threads can keep in flight: every warp alternates "issue loads -> wait ->
// arithmetic -> stores", so memory latency and the arithmetic (Adam: IEEE sqrt
// and division per element) do not overlap.  This kernel decouples them the
// Blackwell way: one producer warp drives the TMA engine with 1-D bulk copies
// (`cp.async.bulk`, SASS UBLKCP) into a ring of shared-memory stages guarded by
// mbarriers; eight consumer warps do the arithmetic from shared memory in
// place; the producer writes finished stages back with bulk stores.  Bytes in
// flight per SM = (stages - 1) x stage size (~100 KB), independent of register
// pressure and occupancy.
//
//   stage s:  [ buffer tile (B) | param tile (P) | state tiles (P) ... | grad-out tile (P) ]
//   producer: wait done[j] -> bulk-store tile j -> wait its smem reads -> bulk-load tile j+S
//   consumer: wait full[k] -> LDS.128, math, STS.128 in place -> fence.proxy.async -> arrive done[k]
//
// A tile is a contiguous range of the flat element space; it may span several
// parameters: the producer issues one bulk copy per (array, piece), its 32 lanes
// taking one piece each, so runs of tiny parameters are issued in parallel.  The
// consumers never look at the segment table: in shared memory the tile is
// contiguous and the arithmetic is elementwise.
//
// Requirements (checked on the host, otherwise the register path is used):
// every pointer 16-byte aligned, every csum / buf_off / range boundary a multiple
// of A = 16 / min(sizeof(B), sizeof(P)) elements, one parameter dtype.
#pragma once
#include "gp_walk.cuh"

namespace gpb {

constexpr int kConsumerWarps = 8;
constexpr int kConsumers = kConsumerWarps * 32;
constexpr int kThreads = kConsumers + 32;
constexpr int kMaxArrays = 6;

// ------------------------------------------------------------- PTX helpers --
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared, completion counted in bytes on an mbarrier (TMA, 1-D)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes,
                                          uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read1() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// One array of a stage: which pointer of the segment it mirrors, element size,
// direction.
struct ArrayDesc {
  int ptr_index;  // -1: the packed buffer (indexed by buf_off), else gp_seg_t.ptr[ptr_index]
  int itemsize;
  int load;   // copied global -> shared before the arithmetic
  int store;  // copied shared -> global after it
  int smem_off;  // byte offset inside a stage
};

struct BulkArgs {
  const int64_t* csum;
  const gp_seg_t* segs;
  int n_segs;
  int64_t begin, end;
  int64_t per_cta;  // multiple of the tile
  const void* buffer;
  int tile;        // elements per tile
  int stages;
  int stage_bytes;
  int n_arrays;
  ArrayDesc arr[kMaxArrays];
  uint32_t load_bytes_per_elem;
  int chunk_bytes;  // bytes per bulk copy (multiple of 16)
  int debug;  // experiments only: 1 = skip the bulk stores, 2 = skip the arithmetic
};

// Op interface:  template <class B, class P, int SM> static void tile(const Op&, char* stage,
//                const BulkArgs&, int n_vec4)   -- consumers, in place in shared memory.
template <class Op, class B, class P, int SM>
__global__ void __launch_bounds__(kThreads, 2) bulk_kernel(const BulkArgs a, const Op op) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [csum int64[n+1]] [full[S]] [done[S]] [stages...]
  const int n = a.n_segs;
  int64_t* s_csum = reinterpret_cast<int64_t*>(smem_raw);
  const size_t csum_bytes = (((size_t)(n + 1) * 8) + 127) & ~(size_t)127;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + csum_bytes);
  uint64_t* done = full + a.stages;
  unsigned char* stage0 = smem_raw + csum_bytes + (((size_t)a.stages * 16) + 127 & ~(size_t)127);

  const int64_t lo = a.begin + (int64_t)blockIdx.x * a.per_cta;
  int64_t hi = lo + a.per_cta;
  if (hi > a.end) hi = a.end;
  if (lo >= hi) return;
  const int T = a.tile;
  const int K = (int)((hi - lo + T - 1) / T);
  const int S = a.stages;

  for (int i = threadIdx.x; i <= n; i += blockDim.x) s_csum[i] = a.csum[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(done + s, kConsumers);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kConsumerWarps) {
    // ------------------------------------------------------------ producer --
    int j_first = gpw::seg_find(s_csum, n, lo);
    int j_store = j_first;

    auto issue = [&](int k, bool is_load, int& jf) {
      const int64_t t0 = lo + (int64_t)k * T;
      int64_t t1 = t0 + T;
      if (t1 > hi) t1 = hi;
      unsigned char* st = stage0 + (size_t)(k % S) * a.stage_bytes;
      jf = gpw::seg_seek(s_csum, n, jf, t0);
      if (is_load && lane == 0) mbar_expect_tx(full + (k % S), (uint32_t)(t1 - t0) * a.load_bytes_per_elem);
      __syncwarp();
      // Work items = (piece, array, chunk); the 32 lanes take them round-robin.
      // Chunking matters: one bulk copy is serviced with limited parallelism, so
      // many concurrent copies of a few KB move more bytes than few large ones.
      int item = 0;
      for (int j = jf; j < n && s_csum[j] < t1; ++j) {
        const int64_t p0 = s_csum[j] > t0 ? s_csum[j] : t0;
        const int64_t p1 = s_csum[j + 1] < t1 ? s_csum[j + 1] : t1;
        if (p1 <= p0) continue;
        const gp_seg_t* g = a.segs + j;
        const int64_t e0 = p0 - s_csum[j];          // element offset inside the parameter
        const uint32_t cnt = (uint32_t)(p1 - p0);
        const uint32_t toff = (uint32_t)(p0 - t0);  // element offset inside the tile
#pragma unroll
        for (int q = 0; q < kMaxArrays; ++q) {
          if (q >= a.n_arrays) break;
          const ArrayDesc& d = a.arr[q];
          if (is_load ? !d.load : !d.store) continue;
          const uint32_t che = (uint32_t)a.chunk_bytes / (uint32_t)d.itemsize;  // elements per chunk
          for (uint32_t c0 = 0; c0 < cnt; c0 += che, ++item) {
            if ((item & 31) != lane) continue;
            const uint32_t c1 = c0 + che < cnt ? c0 + che : cnt;
            unsigned char* sp = st + d.smem_off + (size_t)(toff + c0) * d.itemsize;
            unsigned char* gp = d.ptr_index < 0
                                    ? (unsigned char*)a.buffer + (g->buf_off + e0 + c0) * d.itemsize
                                    : (unsigned char*)g->ptr[d.ptr_index] + (e0 + c0) * d.itemsize;
            if (is_load) bulk_load(sp, gp, (c1 - c0) * d.itemsize, full + (k % S));
            else bulk_store(gp, sp, (c1 - c0) * d.itemsize);
          }
        }
      }
    };

    const int pro = K < S ? K : S;
    for (int k = 0; k < pro; ++k) issue(k, true, j_first);
    for (int j = 0; j < K; ++j) {
      if (lane == 0) mbar_wait(done + (j % S), (uint32_t)((j / S) & 1));
      __syncwarp();
      if (!(a.debug & 1)) issue(j, false, j_store);
      bulk_commit();
      // Refill the stage of the PREVIOUS tile: its bulk store has had one whole
      // tile time to read shared memory, so this wait does not stall the
      // load-issuing chain (waiting on the store just issued costs ~1.4 us/tile).
      if (j >= 1 && (j - 1) + S < K) {
        bulk_wait_read1();
        __syncwarp();
        issue(j - 1 + S, true, j_first);
      }
    }
    if (K >= 1 && (K - 1) + S < K) { /* unreachable: the last tile never triggers a refill */ }
    bulk_wait0();  // all writes globally performed before the kernel ends
  } else {
    // ----------------------------------------------------------- consumers --
    for (int k = 0; k < K; ++k) {
      const int s = k % S;
      mbar_wait(full + s, (uint32_t)((k / S) & 1));
      const int64_t t0 = lo + (int64_t)k * T;
      const int n_el = (int)((hi - t0) < T ? (hi - t0) : T);
      if (!(a.debug & 2)) Op::template tile<B, P, SM>(op, stage0 + (size_t)s * a.stage_bytes, a, n_el >> 2);
      fence_proxy_async();  // make the in-place results visible to the bulk-store engine
      mbar_arrive(done + s);
    }
  }
}

// ---------------------------------------------- shared-memory 4-wide access --
template <class T> __device__ __forceinline__ Raw4<T> lds4(const T* p);
template <> __device__ __forceinline__ Raw4<float> lds4(const float* p) {
  Raw4<float> v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.r.x), "=f"(v.r.y), "=f"(v.r.z), "=f"(v.r.w)
               : "r"(smem_u32(p)));
  return v;
}
template <> __device__ __forceinline__ Raw4<__half> lds4(const __half* p) {
  Raw4<__half> v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.r.x), "=r"(v.r.y) : "r"(smem_u32(p)));
  return v;
}
template <> __device__ __forceinline__ Raw4<__nv_bfloat16> lds4(const __nv_bfloat16* p) {
  Raw4<__nv_bfloat16> v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.r.x), "=r"(v.r.y) : "r"(smem_u32(p)));
  return v;
}
template <class T> __device__ __forceinline__ void sts4(T* p, const Raw4<T>& v);
template <> __device__ __forceinline__ void sts4(float* p, const Raw4<float>& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "f"(v.r.x),
               "f"(v.r.y), "f"(v.r.z), "f"(v.r.w)
               : "memory");
}
template <> __device__ __forceinline__ void sts4(__half* p, const Raw4<__half>& v) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(smem_u32(p)), "r"(v.r.x), "r"(v.r.y)
               : "memory");
}

// ------------------------------------------------------------- host side --
struct BulkTuning {
  int enable, tile, stages, ctas, debug, chunk;
};
extern BulkTuning g_bulk_tuning;

// Fill the array table; returns the stage size in bytes.
inline int layout_stage(BulkArgs& a, int tile) {
  int off = 0;
  uint32_t load_b = 0;
  for (int q = 0; q < a.n_arrays; ++q) {
    a.arr[q].smem_off = off;
    off += ((tile * a.arr[q].itemsize) + 127) & ~127;
    if (a.arr[q].load) load_b += a.arr[q].itemsize;
  }
  a.load_bytes_per_elem = load_b;
  return off;
}

template <class Op, class B, class P, int SM>
int launch_bulk_t(BulkArgs a, const Op& op, cudaStream_t st, const char* what) {
  int tile = g_bulk_tuning.tile;
  int stages = g_bulk_tuning.stages;
  const size_t fixed = ((((size_t)a.n_segs + 1) * 8 + 127) & ~(size_t)127) + 256;
  int stage_bytes = layout_stage(a, tile);
  const int ctas = g_bulk_tuning.ctas >= 2 ? 2 : 1;
  const size_t cap = ctas == 2 ? 113 * 1024 : 227 * 1024;
  while (stages > 2 && fixed + (size_t)stages * stage_bytes > cap) --stages;
  while (tile > 512 && fixed + (size_t)stages * stage_bytes > cap) {
    tile >>= 1;
    stage_bytes = layout_stage(a, tile);
  }
  if (fixed + (size_t)stages * stage_bytes > cap) return 1;  // does not fit: use the register path
  a.tile = tile;
  a.debug = g_bulk_tuning.debug;
  a.chunk_bytes = g_bulk_tuning.chunk < 256 ? 256 : (g_bulk_tuning.chunk & ~255);
  a.stages = stages;
  a.stage_bytes = stage_bytes;
  const size_t smem = fixed + (size_t)stages * stage_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bulk_kernel<Op, B, P, SM>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return gp_cuda_fail(e, "cudaFuncSetAttribute(bulk_kernel)");
    attr_set = true;
  }
  const int64_t total = a.end - a.begin;
  int64_t grid = (int64_t)gp_sm_count_cached() * ctas;
  const int64_t n_tiles = (total + tile - 1) / tile;
  if (grid > n_tiles) grid = n_tiles;
  int64_t per = (total + grid - 1) / grid;
  per = (per + 7) & ~(int64_t)7;          // keep every CTA range 16-byte conformant
  a.per_cta = per;
  grid = (total + per - 1) / per;
  bulk_kernel<Op, B, P, SM><<<(unsigned)grid, kThreads, smem, st>>>(a, op);
  return gp_cuda_fail(cudaGetLastError(), what);
}

// returns 1 when the bulk path does not apply (caller falls back), 0 on launch, <0 on error
template <class Op, class B>
int launch_bulk_p(int hint, const BulkArgs& a, const Op& op, cudaStream_t st, const char* what) {
  const int mode = op.s.mode;
  if (hint == GP_F32) {
    if (mode == 0) return launch_bulk_t<Op, B, float, 0>(a, op, st, what);
    if (mode == 1) return launch_bulk_t<Op, B, float, 1>(a, op, st, what);
    return launch_bulk_t<Op, B, float, 2>(a, op, st, what);
  }
  return 1;
}

template <class Op>
int launch_bulk(int buf_dtype, int hint, BulkArgs a, const Op& op, void* stream, const char* what) {
  if (!g_bulk_tuning.enable || hint != GP_F32 || a.n_segs > gpw::kMaxSmemSegs) return 1;
  const int A = buf_dtype == GP_F32 ? 4 : 8;
  if (buf_dtype == GP_F64 || (a.begin % A) || ((a.end - a.begin) % A) || a.end <= a.begin) return 1;
  if (((uintptr_t)a.buffer & 15) != 0) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  a.arr[0].itemsize = gp_itemsize(buf_dtype);
  switch (buf_dtype) {
    case GP_F32: return launch_bulk_p<Op, float>(hint, a, op, st, what);
    case GP_F16: return launch_bulk_p<Op, __half>(hint, a, op, st, what);
    case GP_BF16: return launch_bulk_p<Op, __nv_bfloat16>(hint, a, op, st, what);
    default: return 1;
  }
}

}  // namespace gpb
