// gp_bn.cu -- fused batch-normalisation statistics for MultiNodeBatchNormalization.
//
// Reference being replaced (chainer v7.8.1):
//   chainermn/functions/batch_normalization.py:44-68  _NcclImpl.get_mean_and_var
//       x.mean(axis, out=buf[:C]); xp.square(x).mean(axis, out=buf[C:])
//       -> two CuPy reduction kernels + one elementwise kernel that materialises
//          square(x): x is read twice and |x| bytes are written and re-read.
//   chainermn/functions/batch_normalization.py:70-93  get_ggamma_and_gbeta
//       gy.sum(axis, out=buf[:C]); (gy * x_hat).sum(axis, out=buf[C:])
//   :65-67  var = sqmean - square(mean)      (after the allreduce + div_by_size)
//
// Here: ONE pass over x (or gy, x_hat) producing both statistics.  Grid =
// (C channels) x (S splits over the batch axis); each CTA streams its rows
// with 4-element vector loads (4 independent loads in flight per thread), does
// a warp-shuffle + shared-memory block reduction, and the last CTA of a channel
// (atomic ticket) adds the S partials in a fixed order, so the result is
// deterministic.  Algorithmic bytes: |x| * itemsize read once, 2C * 4 written.
//
// N ranks on one NVSwitch box (gp_bn_*_stats_allreduce): the CTA that completes the LAST
// channel runs the one-shot peer-memory allreduce of the 2C values (gp_p2p_small.cuh) in
// the same launch -- statistics, allReduce, div_by_size and var = sqmean - mean^2
// (chainermn/functions/batch_normalization.py:53-68, 79-93) are ONE kernel.
#include "gp_common.cuh"

namespace {

#include "gp_p2p.cuh"
#include "gp_p2p_small.cuh"

constexpr int kMaxSplits = 64;

template <class T> struct AccOf { using type = float; };
template <> struct AccOf<double> { using type = double; };

__device__ __forceinline__ double load_stat(const void* p, int dtype, int64_t i) {
  switch (dtype) {
    case GP_F16: return (double)__half2float(reinterpret_cast<const __half*>(p)[i]);
    case GP_F32: return (double)reinterpret_cast<const float*>(p)[i];
    default: return reinterpret_cast<const double*>(p)[i];
  }
}
__device__ __forceinline__ void store_stat(void* p, int dtype, int64_t i, double v) {
  switch (dtype) {
    case GP_F16: reinterpret_cast<__half*>(p)[i] = __double2half(v); break;
    case GP_F32: reinterpret_cast<float*>(p)[i] = __double2float_rn(v); break;
    default: reinterpret_cast<double*>(p)[i] = v; break;
  }
}

struct BnArgs {
  const void* x;   // fwd: x;  bwd: x_hat (MODE 1) or x (MODE 2)
  const void* gy;  // bwd only
  const void* mean;
  const void* inv_std;
  int stat_dtype;
  int64_t N, C, HW;
  int S, rows_per_split;
  float* partials;  // [C][S][2] (double-sized slots when accumulating in double)
  int* counters;    // [C]
  void* out;
  int out_dtype;
  double out_scale;  // fwd: 1 / (N * HW); bwd: 1
  int finish_var;    // fwd, one rank: write [mean | var] instead of [mean | sqmean]
  int use_xchg;      // N ranks: the CTA finishing the last channel exchanges `out` (xchg.in)
  int* done;         // channels completed so far (zero between launches)
  SmallArgs xchg;
};

// Channel c's reduction is complete in (t0, t1) of the calling thread (one per CTA):
// combine the S split partials (the last CTA of the channel does, in fixed order) and
// write the channel's two outputs.  Returns true when THIS thread wrote them.
__device__ __forceinline__ bool bn_finish_channel(const BnArgs& a, int64_t c, int sp, double t0,
                                                  double t1) {
  if (a.S > 1) {
    double* part = reinterpret_cast<double*>(a.partials) + ((int64_t)c * a.S) * 2;
    part[sp * 2 + 0] = t0;
    part[sp * 2 + 1] = t1;
    __threadfence();
    const int ticket = atomicAdd(a.counters + c, 1);
    if (ticket != a.S - 1) return false;
    __threadfence();
    t0 = 0;
    t1 = 0;
    for (int k = 0; k < a.S; ++k) {  // fixed order: deterministic
      t0 += __ldcg(part + k * 2 + 0);
      t1 += __ldcg(part + k * 2 + 1);
    }
    a.counters[c] = 0;  // leave the workspace zeroed for the next call
  }
  if (a.finish_var) {
    // single rank: no allreduce follows, so form var = sqmean - mean^2 here with
    // the arithmetic of bn_finish_kernel (operands rounded to the output dtype)
    if (a.out_dtype == GP_F32) {
      const float m = __double2float_rn(t0 * a.out_scale), q = __double2float_rn(t1 * a.out_scale);
      reinterpret_cast<float*>(a.out)[c] = m;
      reinterpret_cast<float*>(a.out)[a.C + c] = __fsub_rn(q, __fmul_rn(m, m));
    } else if (a.out_dtype == GP_F16) {
      const float m = __half2float(__double2half(t0 * a.out_scale));
      const float q = __half2float(__double2half(t1 * a.out_scale));
      reinterpret_cast<__half*>(a.out)[c] = __float2half_rn(m);
      reinterpret_cast<__half*>(a.out)[a.C + c] =
          __float2half_rn(__fsub_rn(q, __half2float(__float2half_rn(__fmul_rn(m, m)))));
    } else {
      const double m = t0 * a.out_scale, q = t1 * a.out_scale;
      reinterpret_cast<double*>(a.out)[c] = m;
      reinterpret_cast<double*>(a.out)[a.C + c] = __dsub_rn(q, __dmul_rn(m, m));
    }
    return true;
  }
  store_stat(a.out, a.out_dtype, c, t0 * a.out_scale);
  store_stat(a.out, a.out_dtype, a.C + c, t1 * a.out_scale);
  return true;
}

// MODE 0: a = sum x, b = sum x^2
// MODE 1: a = sum gy, b = sum gy * xhat           (x holds x_hat)
// MODE 2: a = sum gy, b = sum gy * (x - mean) * inv_std
template <class TX, class TG, int MODE, bool VEC>
__global__ void __launch_bounds__(512) bn_stats_kernel(const BnArgs a) {
  using Acc = typename AccOf<TX>::type;
  using CX = typename Carrier<TX>::type;
  using CG = typename Carrier<TG>::type;
  const int64_t c = blockIdx.x;
  const int sp = blockIdx.y;
  const int64_t n0 = (int64_t)sp * a.rows_per_split;
  int64_t n1 = n0 + a.rows_per_split;
  if (n1 > a.N) n1 = a.N;
  const int rows = (int)(n1 - n0);

  const TX* __restrict__ x = reinterpret_cast<const TX*>(a.x);
  const TG* __restrict__ gy = reinterpret_cast<const TG*>(a.gy);
  Acc mu = 0, is = 0;
  if (MODE == 2) {
    mu = (Acc)load_stat(a.mean, a.stat_dtype, c);
    is = (Acc)load_stat(a.inv_std, a.stat_dtype, c);
  }

  Acc s0 = 0, s1 = 0;
  const int64_t row_stride = a.C * a.HW;
  const int64_t ch_off = c * a.HW;

  if (VEC) {
    const int VR = (int)(a.HW >> 2);  // vectors per row
    const int total = rows * VR;
    constexpr int UN = 4;
    Acc p0[4] = {0, 0, 0, 0}, p1[4] = {0, 0, 0, 0};
    // (row, col) of this thread's next vector, advanced incrementally: an integer
    // division per 16-byte load made the forward kernel instruction-bound
    const int step = blockDim.x;
    const int dq = step / VR, dr = step - dq * VR;
    int row = threadIdx.x / VR, col = threadIdx.x - row * VR;
    for (int base = threadIdx.x; base < total; base += step * UN) {
      Raw4<TX> rx[UN];
      Raw4<TG> rg[UN];
      bool act[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int idx = base + u * step;
        act[u] = idx < total;
        if (act[u]) {
          const int64_t off = (n0 + row) * row_stride + ch_off + (int64_t)col * 4;
          rx[u] = ld4_stream(x + off);
          if (MODE != 0) rg[u] = ld4_stream(gy + off);
        }
        col += dr;
        row += dq;
        if (col >= VR) {
          col -= VR;
          ++row;
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (!act[u]) continue;
        CX vx[4];
        unpack4(rx[u], vx);
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const Acc v = (Acc)vx[i];
            p0[i] += v;
            p1[i] += v * v;
          }
        } else {
          CG vg[4];
          unpack4(rg[u], vg);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const Acc g = (Acc)vg[i];
            Acc xh = (Acc)vx[i];
            if (MODE == 2) xh = (xh - mu) * is;
            p0[i] += g;
            p1[i] += g * xh;
          }
        }
      }
    }
    s0 = (p0[0] + p0[1]) + (p0[2] + p0[3]);
    s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]);
  } else {
    const int HWi = (int)a.HW;
    const int total = rows * HWi;
    const int step = blockDim.x;
    const int dq = step / HWi, dr = step - dq * HWi;
    int row = threadIdx.x / HWi, col = threadIdx.x - row * HWi;
    for (int idx = threadIdx.x; idx < total; idx += step) {
      const int64_t off = (n0 + row) * row_stride + ch_off + col;
      const Acc v = (Acc)to_carrier(x[off]);
      if (MODE == 0) {
        s0 += v;
        s1 += v * v;
      } else {
        const Acc g = (Acc)to_carrier(gy[off]);
        Acc xh = v;
        if (MODE == 2) xh = (xh - mu) * is;
        s0 += g;
        s1 += g * xh;
      }
      col += dr;
      row += dq;
      if (col >= HWi) {
        col -= HWi;
        ++row;
      }
    }
  }

  // block reduction: warp shuffles, then one value per warp through smem
  __shared__ Acc red[2][16];
  __shared__ int s_exchange;
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) s_exchange = 0;
  if (lane == 0) {
    red[0][warp] = s0;
    red[1][warp] = s1;
  }
  __syncthreads();
  if (warp == 0) {
    s0 = lane < nw ? red[0][lane] : (Acc)0;
    s1 = lane < nw ? red[1][lane] : (Acc)0;
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0 && bn_finish_channel(a, c, sp, (double)s0, (double)s1) && a.use_xchg) {
      // N ranks: whoever completes the last channel exchanges the 2C values
      __threadfence();
      if (atomicAdd(a.done, 1) == (int)a.C - 1) {
        *a.done = 0;
        s_exchange = 1;
      }
    }
  }
  if (!a.use_xchg) return;
  __syncthreads();
  if (!s_exchange) return;
  __threadfence();            // the other CTAs' channel outputs (read through L2 below)
  small_exchange(a.xchg);
}

// buf[0:2C] *= scale (rounded to T); var[c] = sqmean[c] - mean[c]^2
template <class T>
__global__ void bn_finish_kernel(T* buf, int64_t C, ScaleArg s, T* var) {
  using A = Arith<T>;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const auto m = descale_rt<T>(to_carrier(buf[c]), s);
  const auto q = descale_rt<T>(to_carrier(buf[C + c]), s);
  buf[c] = from_carrier<T>(m);
  buf[C + c] = from_carrier<T>(q);
  var[c] = from_carrier<T>(A::sub(q, A::mul(m, m)));
}

template <class TX, class TG, int MODE>
int launch_mode(const BnArgs& a, bool vec, dim3 grid, int threads, cudaStream_t st) {
  if (vec) bn_stats_kernel<TX, TG, MODE, true><<<grid, threads, 0, st>>>(a);
  else bn_stats_kernel<TX, TG, MODE, false><<<grid, threads, 0, st>>>(a);
  return gp_cuda_fail(cudaGetLastError(), "bn_stats_kernel launch");
}

template <class TX, int MODE>
int launch_g(int gy_dtype, const BnArgs& a, bool vec, dim3 grid, int threads, cudaStream_t st) {
  if (MODE == 0) return launch_mode<TX, TX, MODE>(a, vec, grid, threads, st);
  switch (gy_dtype) {
    case GP_F32: return launch_mode<TX, float, MODE>(a, vec, grid, threads, st);
    case GP_F16: return launch_mode<TX, __half, MODE>(a, vec, grid, threads, st);
    case GP_F64: return launch_mode<TX, double, MODE>(a, vec, grid, threads, st);
    default:
      gp_set_error("gp_bn_bwd_stats: unsupported gy dtype id %d", gy_dtype);
      return GP_EINVAL;
  }
}

template <int MODE>
int launch_x(int x_dtype, int gy_dtype, const BnArgs& a, bool vec, dim3 grid, int threads,
             cudaStream_t st) {
  switch (x_dtype) {
    case GP_F32: return launch_g<float, MODE>(gy_dtype, a, vec, grid, threads, st);
    case GP_F16: return launch_g<__half, MODE>(gy_dtype, a, vec, grid, threads, st);
    case GP_F64: return launch_g<double, MODE>(gy_dtype, a, vec, grid, threads, st);
    default:
      gp_set_error("gp_bn_stats: unsupported x dtype id %d", x_dtype);
      return GP_EINVAL;
  }
}

// Workspace layout.  ONE workspace serves every BN layer of a communicator, whatever its C, so
// the words the kernels rely on being ZERO between launches -- the "channels done" counter and
// the per-channel tickets -- sit in a header whose offsets do not depend on C; the areas that
// are left holding junk (split partials, staged local statistics) come after it.  (With the
// C-dependent offsets of round 1 the partials of a C = 64 layer landed on the tickets of a
// C = 128 layer: those channels were never completed.)
//   [0, 256)                       "channels done" counter
//   [256, 256 + 4 * kTicketCap)    channel tickets (splits S > 1 only while C <= kTicketCap)
//   then                           split partials  [C][kMaxSplits][2] doubles
//   then                           local 2C float statistics staged for the exchange
constexpr int64_t kTicketCap = 16384;
constexpr size_t kWsDoneOff = 0;
constexpr size_t kWsTicketsOff = 256;
constexpr size_t kWsPartialsOff = kWsTicketsOff + (size_t)kTicketCap * sizeof(int);
size_t ws_partials_bytes(int64_t C) { return (size_t)(C * kMaxSplits * 2 * sizeof(double)); }
size_t ws_local_off(int64_t C) { return kWsPartialsOff + ws_partials_bytes(C); }
size_t ws_local_bytes(int64_t C) { return (size_t)(2 * C * sizeof(float) + 255) / 256 * 256; }

int bn_launch(int mode, const void* x, int x_dtype, const void* gy, int gy_dtype, const void* mean,
              const void* inv_std, int stat_dtype, int64_t N, int64_t C, int64_t HW, void* out,
              int out_dtype, void* workspace, double out_scale, void* stream, int finish_var = 0,
              P2PComm* comm = nullptr) {
  if (N <= 0 || C <= 0 || HW <= 0) return 0;
  if (out_dtype != GP_F16 && out_dtype != GP_F32 && out_dtype != GP_F64) {
    gp_set_error("gp_bn_stats: unsupported output dtype id %d", out_dtype);
    return GP_EINVAL;
  }
  if (N * HW >= (int64_t)1 << 31) {
    gp_set_error("gp_bn_stats: N * HW per channel must be < 2^31");
    return GP_EINVAL;
  }
  BnArgs a;
  a.x = x; a.gy = gy; a.mean = mean; a.inv_std = inv_std; a.stat_dtype = stat_dtype;
  a.N = N; a.C = C; a.HW = HW;
  a.out = out; a.out_dtype = out_dtype; a.out_scale = out_scale;
  a.finish_var = finish_var;
  a.use_xchg = 0;
  a.done = nullptr;
  if (comm) {
    // local statistics go to a staging area of the workspace; the exchange writes `out`
    if (!workspace || out_dtype != GP_F32 || 2 * C > comm->small_cap || comm->small_cap <= 0) {
      gp_set_error("gp_bn_*_stats_allreduce: needs a workspace, float32 statistics and 2C <= %lld",
                   (long long)comm->small_cap);
      return GP_EINVAL;
    }
    char* ws = reinterpret_cast<char*>(workspace);
    a.done = reinterpret_cast<int*>(ws + kWsDoneOff);
    float* local = reinterpret_cast<float*>(ws + ws_local_off(C));
    a.out = local;
    a.use_xchg = 1;
    small_args_from(comm, &a.xchg, 1.0 / comm->n);
    a.xchg.in = local;
    a.xchg.out = reinterpret_cast<float*>(out);
    a.xchg.n_elems = (int)(2 * C);
    a.xchg.C = mode == 0 ? (int)C : 0;
  }

  // splits over the batch axis: enough CTAs for ~16 per SM (several waves, so that
  // the uneven last wave is short) while keeping >= ~8 KB of rows per CTA.
  const int sms = gp_sm_count_cached();
  int64_t S = (g_gp_tuning.bn_ctas_per_sm * (int64_t)sms + C - 1) / C;
  const int64_t bytes_per_row = HW * gp_itemsize(x_dtype);
  int64_t max_by_size = (N * bytes_per_row) / 8192;
  if (max_by_size < 1) max_by_size = 1;
  if (S > max_by_size) S = max_by_size;
  if (S > N) S = N;
  if (S > kMaxSplits) S = kMaxSplits;
  if (S < 1) S = 1;
  if (workspace == nullptr || C > kTicketCap) S = 1;
  a.rows_per_split = (int)((N + S - 1) / S);
  a.S = (int)((N + a.rows_per_split - 1) / a.rows_per_split);
  a.counters = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + kWsTicketsOff);
  a.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kWsPartialsOff);

  const int xs = gp_itemsize(x_dtype), gs = gp_itemsize(gy_dtype);
  bool vec = (HW % 4 == 0) && ((uintptr_t)x % (xs == 2 ? 8 : 16) == 0);
  if (mode != 0) vec = vec && ((uintptr_t)gy % (gs == 2 ? 8 : 16) == 0);

  int threads = g_gp_tuning.bn_threads;
  const int64_t work = (int64_t)a.rows_per_split * (vec ? HW / 4 : HW);
  while (threads > 64 && work < threads * 2) threads >>= 1;
  if (threads > 512) threads = 512;
  threads &= ~31;
  if (threads < 32) threads = 32;
  dim3 grid((unsigned)C, (unsigned)a.S);
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case 0: return launch_x<0>(x_dtype, gy_dtype, a, vec, grid, threads, st);
    case 1: return launch_x<1>(x_dtype, gy_dtype, a, vec, grid, threads, st);
    default: return launch_x<2>(x_dtype, gy_dtype, a, vec, grid, threads, st);
  }
}

}  // namespace

// ["channels done" counter | channel tickets | split partials | local 2C statistics]
extern "C" size_t gp_bn_workspace_bytes(int64_t C) {
  if (C < 0) C = 0;
  return ws_local_off(C) + ws_local_bytes(C);
}

// Byte ranges of the workspace for a layer of C channels, for tests and integrators:
// out[0..1] the zero-between-launches header, out[2..3] the scratch that may hold junk.
extern "C" int gp_bn_workspace_layout(int64_t C, int64_t* out4) {
  if (!out4 || C < 0) {
    gp_set_error("gp_bn_workspace_layout: bad arguments");
    return GP_EINVAL;
  }
  out4[0] = 0;
  out4[1] = (int64_t)kWsPartialsOff;
  out4[2] = (int64_t)kWsPartialsOff;
  out4[3] = (int64_t)(ws_local_off(C) + ws_local_bytes(C));
  return 0;
}

extern "C" int gp_bn_fwd_stats_allreduce(void* p2p_comm, const void* x, int x_dtype, int64_t N,
                                         int64_t C, int64_t HW, void* out_mean_var,
                                         void* workspace, void* stream) {
  if (!p2p_comm) {
    gp_set_error("gp_bn_fwd_stats_allreduce: no peer-memory communicator");
    return GP_EINVAL;
  }
  const double inv = (N > 0 && HW > 0) ? 1.0 / ((double)N * (double)HW) : 0.0;
  return bn_launch(0, x, x_dtype, nullptr, x_dtype, nullptr, nullptr, GP_F32, N, C, HW,
                   out_mean_var, GP_F32, workspace, inv, stream, 0, (P2PComm*)p2p_comm);
}

extern "C" int gp_bn_bwd_stats_allreduce(void* p2p_comm, const void* gy, int gy_dtype,
                                         const void* xhat_or_x, int x_dtype, const void* mean,
                                         const void* inv_std, int stat_dtype, int64_t N, int64_t C,
                                         int64_t HW, void* out, void* workspace, void* stream) {
  if (!p2p_comm) {
    gp_set_error("gp_bn_bwd_stats_allreduce: no peer-memory communicator");
    return GP_EINVAL;
  }
  const int mode = (mean != nullptr && inv_std != nullptr) ? 2 : 1;
  return bn_launch(mode, xhat_or_x, x_dtype, gy, gy_dtype, mean, inv_std, stat_dtype, N, C, HW,
                   out, GP_F32, workspace, 1.0, stream, 0, (P2PComm*)p2p_comm);
}

extern "C" int gp_bn_fwd_stats(const void* x, int x_dtype, int64_t N, int64_t C, int64_t HW,
                               void* out, int out_dtype, void* workspace, void* stream) {
  const double inv = (N > 0 && HW > 0) ? 1.0 / ((double)N * (double)HW) : 0.0;
  return bn_launch(0, x, x_dtype, nullptr, x_dtype, nullptr, nullptr, GP_F32, N, C, HW, out,
                   out_dtype, workspace, inv, stream);
}

extern "C" int gp_bn_fwd_mean_var(const void* x, int x_dtype, int64_t N, int64_t C, int64_t HW,
                                  void* out, int out_dtype, void* workspace, void* stream) {
  const double inv = (N > 0 && HW > 0) ? 1.0 / ((double)N * (double)HW) : 0.0;
  return bn_launch(0, x, x_dtype, nullptr, x_dtype, nullptr, nullptr, GP_F32, N, C, HW, out,
                   out_dtype, workspace, inv, stream, 1);
}

extern "C" int gp_bn_bwd_stats(const void* gy, int gy_dtype, const void* xhat_or_x, int x_dtype,
                               const void* mean, const void* inv_std, int stat_dtype, int64_t N,
                               int64_t C, int64_t HW, void* out, int out_dtype, void* workspace,
                               void* stream) {
  const int mode = (mean != nullptr && inv_std != nullptr) ? 2 : 1;
  return bn_launch(mode, xhat_or_x, x_dtype, gy, gy_dtype, mean, inv_std, stat_dtype, N, C, HW,
                   out, out_dtype, workspace, 1.0, stream);
}

extern "C" int gp_bn_finish_mean_var(void* buf, int dtype, int64_t C, double scale, void* out_var,
                                     void* stream) {
  if (C <= 0) return 0;
  const ScaleArg s = make_scale(scale);
  const int threads = 128;
  const int grid = (int)((C + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case GP_F32: bn_finish_kernel<float><<<grid, threads, 0, st>>>((float*)buf, C, s, (float*)out_var); break;
    case GP_F16: bn_finish_kernel<__half><<<grid, threads, 0, st>>>((__half*)buf, C, s, (__half*)out_var); break;
    case GP_F64: bn_finish_kernel<double><<<grid, threads, 0, st>>>((double*)buf, C, s, (double*)out_var); break;
    default:
      gp_set_error("gp_bn_finish_mean_var: unsupported dtype id %d", dtype);
      return GP_EINVAL;
  }
  return gp_cuda_fail(cudaGetLastError(), "bn_finish_kernel launch");
}
