// gp_p2p.cu -- allreduce of the packed gradient buffer over NVLink PEER MEMORY,
// written as ONE kernel per rank (two-shot: reduce-scatter by peer loads +
// all-gather by peer stores), with the cross-GPU barriers inside the kernel.
//
// Reference being replaced: `nccl_comm.allReduce(sendbuf, recvbuf, n_elems,
// type_id, NCCL_SUM, stream.ptr)` in
// chainermn/communicators/pure_nccl_communicator.py:180-182 for the case the
// reference calls "intra node": all ranks on one NVSwitch box.  (NCCL stays the
// transport for anything else: gp_nccl.cu.)
//
// Why: on an NVSwitch box every GPU reaches every peer at full bandwidth, so
// the allreduce of an S-byte buffer needs only S(N-1)/N bytes in and out of
// each GPU.  Rank r owns the r-th 1/N of the flat buffer.  For each 16-byte
// vector of its shard it loads the N copies (its own from HBM, N-1 through
// NVLink), adds them in RANK ORDER (deterministic, identical bits on every rank,
// the same order as the oracle), and stores the sum into all N buffers.  No
// address is written by one rank and read by another inside the kernel, so the
// only synchronisation is a barrier before (every rank has packed) and after
// (every rank has stored), both done with flags in peer memory.
//
//   kernel(rank r):  signal+wait "ready"  ->  for v in shard r: sum_k buf_k[v] -> store to buf_0..N-1[v]
//                    ->  grid-wide completion -> signal+wait "done"
//
// Buffers and flag words are cudaMalloc allocations shared with
// cudaIpcGetMemHandle / cudaIpcOpenMemHandle (gp_ipc_*), exchanged by the host
// code over the control plane.
#include <string.h>

#include "gp_common.cuh"

namespace {

#include "gp_p2p.cuh"

template <class T> struct Vec16;  // 16 bytes of T
template <> struct Vec16<float> {
  static constexpr int kElems = 4;
  static __device__ __forceinline__ void add(uint4& acc, const uint4& x) {
    float4& a = reinterpret_cast<float4&>(acc);
    const float4& b = reinterpret_cast<const float4&>(x);
    a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y);
    a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
  }
  static __device__ __forceinline__ float add1(float a, float b) { return __fadd_rn(a, b); }
};
template <> struct Vec16<double> {
  static constexpr int kElems = 2;
  static __device__ __forceinline__ void add(uint4& acc, const uint4& x) {
    double2& a = reinterpret_cast<double2&>(acc);
    const double2& b = reinterpret_cast<const double2&>(x);
    a.x = __dadd_rn(a.x, b.x); a.y = __dadd_rn(a.y, b.y);
  }
  static __device__ __forceinline__ double add1(double a, double b) { return __dadd_rn(a, b); }
};
template <> struct Vec16<__half> {
  static constexpr int kElems = 8;
  // every partial sum is rounded to half, as a half-precision ring would do
  static __device__ __forceinline__ void add(uint4& acc, const uint4& x) {
    __half2* a = reinterpret_cast<__half2*>(&acc);
    const __half2* b = reinterpret_cast<const __half2*>(&x);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __hadd2_rn(a[i], b[i]);
  }
  static __device__ __forceinline__ __half add1(__half a, __half b) { return __hadd_rn(a, b); }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int kElems = 8;
  static __device__ __forceinline__ void add(uint4& acc, const uint4& x) {
    __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&acc);
    const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __hadd2_rn(a[i], b[i]);
  }
  static __device__ __forceinline__ __nv_bfloat16 add1(__nv_bfloat16 a, __nv_bfloat16 b) {
    return __hadd_rn(a, b);
  }
};

// Data accesses.  Every address is read once and written once per kernel and
// the barriers order them against the other ranks, so plain (weak) accesses are
// sufficient; MODE 1 keeps system-scope relaxed accesses for comparison.
template <int MODE>
__device__ __forceinline__ uint4 ld_peer(const uint4* p) {
  uint4 v;
  if constexpr (MODE == 1) {
    asm volatile("ld.global.relaxed.sys.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
  } else {
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
  }
  return v;
}
template <int MODE>
__device__ __forceinline__ void st_peer(uint4* p, const uint4& v) {
  if constexpr (MODE == 1) {
    asm volatile("st.global.relaxed.sys.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
  } else {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
  }
}

template <class T, int N, int MODE>
__global__ void __launch_bounds__(512) p2p_allreduce_kernel(const P2PArgs a) {
  constexpr int E = Vec16<T>::kElems;
  // ---- barrier 1: every rank has finished writing (packing) its buffer ------
  if (blockIdx.x == 0) signal_all(a, 0, a.epoch);
  wait_all(a, 0, a.epoch);
  __syncthreads();

  // ---- this rank's shard: sum of the N copies, in rank order, to all N copies
  const int64_t v_begin = a.begin / E;           // shard boundaries are 16-byte aligned
  const int64_t v_end = a.end / E;               // (the global tail is handled below)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int UN = N >= 4 ? 2 : 4;             // N * UN vector loads in flight per thread
  for (int64_t v0 = v_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < v_end;
       v0 += stride * UN) {
    uint4 x[UN][N];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < v_end) {
#pragma unroll
        for (int k = 0; k < N; ++k) x[u][k] = ld_peer<MODE>(reinterpret_cast<const uint4*>(a.bufs[k]) + v);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < v_end) {
        uint4 acc = x[u][0];
#pragma unroll
        for (int k = 1; k < N; ++k) Vec16<T>::add(acc, x[u][k]);
#pragma unroll
        for (int k = 0; k < N; ++k) st_peer<MODE>(reinterpret_cast<uint4*>(a.bufs[k]) + v, acc);
      }
    }
  }
  // scalar tail of the whole buffer (fewer than E elements), owned by the last rank
  {
    const int64_t t = v_end * E + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.end) {
      T acc = reinterpret_cast<const T*>(a.bufs[0])[t];
#pragma unroll
      for (int k = 1; k < N; ++k) acc = Vec16<T>::add1(acc, reinterpret_cast<const T*>(a.bufs[k])[t]);
#pragma unroll
      for (int k = 0; k < N; ++k) reinterpret_cast<T*>(a.bufs[k])[t] = acc;
    }
  }

  // ---- barrier 2: every rank's stores have landed everywhere ----------------
  finish_all(a);
}

// ---------------------------------------------------------------------------
// One-shot allreduce for SMALL float32 messages (MultiNodeBatchNormalization's
// 2C statistics, chainermn/functions/batch_normalization.py:57-60, 83-86): every
// rank stores its vector into slot [rank] of every peer's receive area, raises a
// flag, waits for the N flags, then adds the N slots in rank order, applies the
// 1/size scale and (forward statistics) forms var = sqmean - mean^2 -- the work of
// allReduce + div_by_size + the var kernel in ONE single-CTA launch whose latency
// is one NVLink store round.  Receive areas are double-buffered by call parity.
#include "gp_p2p_small.cuh"

__global__ void __launch_bounds__(1024) p2p_small_kernel(const SmallArgs a) { small_exchange(a); }

int g_p2p_ctas = 0;     // 0: default
int g_p2p_threads = 512;
int g_p2p_mode = 1;      // system-scope relaxed accesses measured faster (tools/p2p_bench.py)

template <class T>
int launch_n(const P2PArgs& a, int grid, int threads, cudaStream_t st) {
  const bool sys = g_p2p_mode == 1;
  switch (a.n) {
    case 2:
      if (sys) p2p_allreduce_kernel<T, 2, 1><<<grid, threads, 0, st>>>(a);
      else p2p_allreduce_kernel<T, 2, 0><<<grid, threads, 0, st>>>(a);
      break;
    case 4:
      if (sys) p2p_allreduce_kernel<T, 4, 1><<<grid, threads, 0, st>>>(a);
      else p2p_allreduce_kernel<T, 4, 0><<<grid, threads, 0, st>>>(a);
      break;
    case 8:
      if (sys) p2p_allreduce_kernel<T, 8, 1><<<grid, threads, 0, st>>>(a);
      else p2p_allreduce_kernel<T, 8, 0><<<grid, threads, 0, st>>>(a);
      break;
    default:
      gp_set_error("gp_p2p_allreduce: world size %d is not supported (2, 4, 8)", a.n);
      return GP_EINVAL;
  }
  return gp_cuda_fail(cudaGetLastError(), "p2p_allreduce_kernel launch");
}

}  // namespace

extern "C" {

int gp_ipc_get_handle(void* device_ptr, char* handle64) {
  cudaIpcMemHandle_t h;
  GP_CUDA(cudaIpcGetMemHandle(&h, device_ptr));
  static_assert(sizeof(h) == GP_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  memcpy(handle64, &h, sizeof(h));
  return 0;
}

int gp_ipc_open_handle(const char* handle64, void** device_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  GP_CUDA(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int gp_ipc_close_handle(void* device_ptr) {
  if (device_ptr) GP_CUDA(cudaIpcCloseMemHandle(device_ptr));
  return 0;
}

size_t gp_p2p_flag_bytes(void) { return (2 * kMaxRanks + 8) * sizeof(uint32_t); }

int gp_p2p_create(void** comm, int rank, int n_ranks, void* const* buffers, void* const* flags) {
  if (n_ranks != 2 && n_ranks != 4 && n_ranks != 8) {
    gp_set_error("gp_p2p_create: world size %d is not supported (2, 4, 8)", n_ranks);
    return GP_EINVAL;
  }
  if (rank < 0 || rank >= n_ranks) return GP_EINVAL;
  P2PComm* c = new P2PComm();
  c->rank = rank;
  c->n = n_ranks;
  c->epoch = 0;
  c->small_cap = 0;
  c->small_epoch = 0;
  c->step_tile_cap = 0;
  c->step_tile_elems = 0;
  c->step_epoch = 0;
  for (int k = 0; k < n_ranks; ++k) {
    c->bufs[k] = buffers[k];
    c->flags[k] = (uint32_t*)flags[k];
  }
  *comm = c;
  return 0;
}

int gp_p2p_set_buffers(void* comm, void* const* buffers) {
  P2PComm* c = (P2PComm*)comm;
  if (!c) return GP_EINVAL;
  for (int k = 0; k < c->n; ++k) c->bufs[k] = buffers[k];
  return 0;
}

int gp_p2p_destroy(void* comm) {
  delete (P2PComm*)comm;
  return 0;
}

int gp_p2p_allreduce(void* comm, int dtype, int64_t offset_elems, int64_t n_elems, void* stream) {
  P2PComm* c = (P2PComm*)comm;
  if (!c) return GP_EINVAL;
  if (n_elems <= 0) return 0;
  const int isz = gp_itemsize(dtype);
  const int E = 16 / isz;
  if (offset_elems % E) {
    gp_set_error("gp_p2p_allreduce: offset must be a multiple of %d elements", E);
    return GP_EINVAL;
  }
  P2PArgs a;
  a.rank = c->rank;
  a.n = c->n;
  a.epoch = ++c->epoch;
  a.timeout_ns = g_gp_peer_timeout_ns;
  for (int k = 0; k < c->n; ++k) {
    a.bufs[k] = (char*)c->bufs[k] + offset_elems * isz;
    a.flags[k] = c->flags[k];
    if ((uintptr_t)a.bufs[k] & 15) {
      gp_set_error("gp_p2p_allreduce: buffers must be 16-byte aligned");
      return GP_EINVAL;
    }
  }
  // shard boundaries: multiples of E elements; the last rank takes the remainder
  const int64_t n_vec = n_elems / E;
  const int64_t per = (n_vec + c->n - 1) / c->n;
  int64_t vb = per * c->rank, ve = per * (c->rank + 1);
  if (vb > n_vec) vb = n_vec;
  if (ve > n_vec) ve = n_vec;
  a.begin = vb * E;
  a.end = (c->rank == c->n - 1) ? n_elems : ve * E;
  if (c->rank == c->n - 1 && ve < n_vec) a.end = n_elems;
  const int threads = g_p2p_threads;
  int64_t grid = g_p2p_ctas > 0 ? g_p2p_ctas : 2 * gp_sm_count_cached();
  const int64_t need = ((ve - vb) + threads - 1) / threads;
  if (grid > need) grid = need > 0 ? need : 1;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case GP_F32: return launch_n<float>(a, (int)grid, threads, st);
    case GP_F16: return launch_n<__half>(a, (int)grid, threads, st);
    case GP_BF16: return launch_n<__nv_bfloat16>(a, (int)grid, threads, st);
    case GP_F64: return launch_n<double>(a, (int)grid, threads, st);
    default:
      gp_set_error("gp_p2p_allreduce: unsupported dtype id %d", dtype);
      return GP_EINVAL;
  }
}

size_t gp_p2p_small_bytes(int n_ranks, int64_t capacity_elems) {
  return (size_t)2 * n_ranks * capacity_elems * sizeof(float);
}

int gp_p2p_set_small(void* comm, void* const* recv_areas, void* const* flag_blocks,
                     int64_t capacity_elems) {
  P2PComm* c = (P2PComm*)comm;
  if (!c) return GP_EINVAL;
  for (int k = 0; k < c->n; ++k) {
    c->small_recv[k] = (float*)recv_areas[k];
    c->small_flags[k] = (uint32_t*)flag_blocks[k];
  }
  c->small_cap = capacity_elems;
  c->small_epoch = 0;
  return 0;
}

int gp_p2p_allreduce_small(void* comm, const void* in, void* out, int64_t n_elems, int64_t C,
                           double scale, void* stream) {
  P2PComm* c = (P2PComm*)comm;
  if (!c || c->small_cap <= 0) {
    gp_set_error("gp_p2p_allreduce_small: small-message area not set");
    return GP_EINVAL;
  }
  if (n_elems <= 0) return 0;
  if (n_elems > c->small_cap || (C > 0 && 2 * C != n_elems)) {
    gp_set_error("gp_p2p_allreduce_small: %lld elements exceed the capacity %lld (or C mismatch)",
                 (long long)n_elems, (long long)c->small_cap);
    return GP_EINVAL;
  }
  SmallArgs a;
  small_args_from(c, &a, scale);
  a.in = (const float*)in;
  a.out = (float*)out;
  a.n_elems = (int)n_elems;
  a.C = (int)C;
  int threads = 1024;
  while (threads > 64 && threads / 2 >= n_elems) threads >>= 1;
  p2p_small_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(a);
  return gp_cuda_fail(cudaGetLastError(), "p2p_small_kernel launch");
}

int gp_p2p_set_tuning(int ctas, int threads, int mode) {
  g_p2p_ctas = ctas;
  if (threads >= 32 && threads <= 512) g_p2p_threads = threads & ~31;
  g_p2p_mode = mode;
  return 0;
}

}  // extern "C"
