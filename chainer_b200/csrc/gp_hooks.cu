// gp_hooks.cu -- the optimizer-hook kernels that cannot live inside the fused
// update: the global gradient norm of GradientClipping, and the stand-alone
// (unfused) forms of the clipping and weight-decay hooks.
//
// Reference being replaced (chainer v7.8.1):
//   chainer/optimizer_hooks/gradient_clipping.py:9-52   _sum_sqnorm_grads: one
//       `g.ravel().dot(g)` per parameter (cuBLAS dot launches) summed on the device
//   :84-106  norm = sqrt(sqnorm); rate = threshold / norm; rate.clip(None, 1);
//       `grad *= rate` -- one launch per parameter
//   chainer/optimizer_hooks/weight_decay.py:44-57   `g += decay * p`, one launch
//       per parameter
//
// Here the squared norm of ALL mean gradients is ONE deterministic reduction over
// the allreduced packed buffer (the gradients are contiguous there), which also
// forms the clipping rate on the device; the fused update kernels read that rate
// (gp_sgd_hooks.cu / gp_adam_hooks.cu), so a clipped, decayed step is
//   pack -> allreduce -> gp_sqnorm -> fused update        (no host synchronisation).
// The reduction is deterministic for a given size (fixed grid, fixed order), so
// every rank derives the same rate from the same allreduced bytes.
#include "gp_common.cuh"

namespace {

constexpr int kSqThreads = 256;
constexpr int kSqMaxGrid = 1024;

struct SqnormOut {   // 16 bytes of device memory owned by the caller
  double sqsum;
  float rate;
  float norm;
};

struct SqWorkspace {
  double partial[kSqMaxGrid];
  unsigned int ticket;
};

template <class B>
__device__ __forceinline__ double sq_of(const B* x, int64_t i, const ScaleArg& s) {
  // the mean gradient as the update kernels see it: descaled, rounded to B
  const auto g = descale_rt<B>(to_carrier(x[i]), s);
  return (double)g * (double)g;
}

template <class B>
__global__ void __launch_bounds__(kSqThreads) sqnorm_kernel(const B* __restrict__ x, int64_t n,
                                                            const ScaleArg s, int accumulate,
                                                            double threshold, SqWorkspace* ws,
                                                            SqnormOut* out) {
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // four independent loads in flight per thread
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const double a0 = sq_of(x, i, s), a1 = sq_of(x, i + stride, s);
    const double a2 = sq_of(x, i + 2 * stride, s), a3 = sq_of(x, i + 3 * stride, s);
    acc += (a0 + a1) + (a2 + a3);
  }
  for (; i < n; i += stride) acc += sq_of(x, i, s);

  __shared__ double s_warp[kSqThreads / 32];
  __shared__ bool s_last;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kSqThreads / 32; ++w) t += s_warp[w];
    ws->partial[blockIdx.x] = t;
    __threadfence();
    s_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // last CTA: the partials in index order (lane-strided, then a fixed shuffle tree)
  __threadfence();
  if (threadIdx.x < 32) {
    double t = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) t += __ldcg(&ws->partial[b]);
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      const double total = (accumulate ? out->sqsum : 0.0) + t;
      const double norm = sqrt(total);
      out->sqsum = total;
      out->norm = (float)norm;
      // gradient_clipping.py:91-101: rate = threshold / norm, clipped to <= 1
      // (norm == 0 gives inf -> 1)
      out->rate = threshold > 0.0 ? fminf(__fdiv_rn((float)threshold, (float)norm), 1.0f) : 1.0f;
      ws->ticket = 0;
    }
  }
}

template <class B>
int launch_sqnorm(const void* x, int64_t n, const ScaleArg& s, int accumulate, double threshold,
                  void* ws, void* out, cudaStream_t st) {
  int64_t grid = (n + (int64_t)kSqThreads * 16 - 1) / ((int64_t)kSqThreads * 16);
  const int64_t cap = (int64_t)gp_sm_count_cached() * 4;
  if (grid > cap) grid = cap;
  if (grid > kSqMaxGrid) grid = kSqMaxGrid;
  if (grid < 1) grid = 1;
  sqnorm_kernel<B><<<(unsigned)grid, kSqThreads, 0, st>>>(
      reinterpret_cast<const B*>(x), n, s, accumulate, threshold,
      reinterpret_cast<SqWorkspace*>(ws), reinterpret_cast<SqnormOut*>(out));
  return gp_cuda_fail(cudaGetLastError(), "sqnorm_kernel launch");
}

// x *= *factor, in the array's type (gradient_clipping.py:103-106)
template <class B>
__global__ void __launch_bounds__(256) scale_by_kernel(B* __restrict__ x, int64_t n,
                                                       const float* __restrict__ factor) {
  using A = Arith<B>;
  const auto f = A::cst((double)__ldg(factor));
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    x[i] = from_carrier<B>(A::mul(to_carrier(x[i]), f));
}

// g += decay * p, in the arrays' type (weight_decay.py:55-57)
template <class B>
__global__ void __launch_bounds__(256) weight_decay_kernel(B* __restrict__ g,
                                                           const B* __restrict__ p, int64_t n,
                                                           double decay) {
  using A = Arith<B>;
  const auto d = A::cst(decay);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    g[i] = from_carrier<B>(A::add(to_carrier(g[i]), A::mul(d, to_carrier(p[i]))));
}

// x /= divisor, in the array's type (chainer/optimizer.py:289-291 `grad /= loss_scale`)
template <class B>
__global__ void __launch_bounds__(256) divide_kernel(B* __restrict__ x, int64_t n, double divisor) {
  using A = Arith<B>;
  const auto d = A::cst(divisor);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    x[i] = from_carrier<B>(A::div(to_carrier(x[i]), d));
}

int flat_grid(int64_t n) {
  int64_t grid = (n + 255) / 256;
  const int64_t cap = (int64_t)gp_sm_count_cached() * 8;
  if (grid > cap) grid = cap;
  return grid < 1 ? 1 : (int)grid;
}

}  // namespace

extern "C" {

size_t gp_sqnorm_workspace_bytes(void) { return sizeof(SqWorkspace); }

int gp_sqnorm(const void* x, int dtype, int64_t n_elems, double scale, int accumulate,
              double threshold, void* workspace, void* out, void* stream) {
  if (!workspace || !out) {
    gp_set_error("gp_sqnorm: workspace and out must be device pointers");
    return GP_EINVAL;
  }
  if (n_elems < 0) return GP_EINVAL;
  const ScaleArg s = make_scale(scale);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case GP_F32: return launch_sqnorm<float>(x, n_elems, s, accumulate, threshold, workspace, out, st);
    case GP_F16: return launch_sqnorm<__half>(x, n_elems, s, accumulate, threshold, workspace, out, st);
    case GP_BF16: return launch_sqnorm<__nv_bfloat16>(x, n_elems, s, accumulate, threshold, workspace, out, st);
    case GP_F64: return launch_sqnorm<double>(x, n_elems, s, accumulate, threshold, workspace, out, st);
    default:
      gp_set_error("gp_sqnorm: unsupported dtype id %d", dtype);
      return GP_EINVAL;
  }
}

int gp_scale_by_device(void* x, int dtype, int64_t n_elems, const void* d_factor, void* stream) {
  if (n_elems <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = flat_grid(n_elems);
  const float* f = reinterpret_cast<const float*>(d_factor);
  switch (dtype) {
    case GP_F32: scale_by_kernel<float><<<grid, 256, 0, st>>>((float*)x, n_elems, f); break;
    case GP_F16: scale_by_kernel<__half><<<grid, 256, 0, st>>>((__half*)x, n_elems, f); break;
    case GP_F64: scale_by_kernel<double><<<grid, 256, 0, st>>>((double*)x, n_elems, f); break;
    default:
      gp_set_error("gp_scale_by_device: unsupported dtype id %d", dtype);
      return GP_EINVAL;
  }
  return gp_cuda_fail(cudaGetLastError(), "scale_by_kernel launch");
}

int gp_weight_decay(void* grad, const void* param, int dtype, int64_t n_elems, double decay,
                    void* stream) {
  if (n_elems <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = flat_grid(n_elems);
  switch (dtype) {
    case GP_F32:
      weight_decay_kernel<float><<<grid, 256, 0, st>>>((float*)grad, (const float*)param, n_elems, decay);
      break;
    case GP_F16:
      weight_decay_kernel<__half><<<grid, 256, 0, st>>>((__half*)grad, (const __half*)param, n_elems, decay);
      break;
    case GP_F64:
      weight_decay_kernel<double><<<grid, 256, 0, st>>>((double*)grad, (const double*)param, n_elems, decay);
      break;
    default:
      gp_set_error("gp_weight_decay: unsupported dtype id %d", dtype);
      return GP_EINVAL;
  }
  return gp_cuda_fail(cudaGetLastError(), "weight_decay_kernel launch");
}

int gp_divide(void* x, int dtype, int64_t n_elems, double divisor, void* stream) {
  if (n_elems <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = flat_grid(n_elems);
  switch (dtype) {
    case GP_F32: divide_kernel<float><<<grid, 256, 0, st>>>((float*)x, n_elems, divisor); break;
    case GP_F16: divide_kernel<__half><<<grid, 256, 0, st>>>((__half*)x, n_elems, divisor); break;
    case GP_F64: divide_kernel<double><<<grid, 256, 0, st>>>((double*)x, n_elems, divisor); break;
    default:
      gp_set_error("gp_divide: unsupported dtype id %d", dtype);
      return GP_EINVAL;
  }
  return gp_cuda_fail(cudaGetLastError(), "divide_kernel launch");
}

}  // extern "C"
