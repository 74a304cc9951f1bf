// gp_bn_apply.cu -- the elementwise halves of batch normalisation around the statistics
// kernels of gp_bn.cu: everything that follows `get_mean_and_var` in the forward pass and
// `get_ggamma_and_gbeta` in the backward pass, one launch each.
//
// Reference being replaced (chainer v7.8.1,
// chainer/functions/normalization/batch_normalization.py):
//   forward  :40-77   inv_std = rsqrt(var + eps)                         (1 launch)
//                     y = gamma * (x - mean) * inv_std + beta  `bn_fwd`  (1 launch, :864-867)
//                     r_mean = r_mean * decay + mean * (1 - decay);
//                     r_var  = r_var  * decay + var  * (1 - decay) * adjust
//                                                     `update_mean_var`  (1 launch, :69-77)
//   backward :105-133 x_hat = (x - mean) * inv_std      materialised (|x| written + re-read)
//                     gx = (gamma * inv_std) * (gy - (x_hat * ggamma + gbeta) * inv_m)  `bn_bwd`
//
// Here: gp_bn_fwd_apply reads x once and writes y, and its first C threads also write
// inv_std and update the running statistics; gp_bn_bwd_apply forms x_hat on the fly from
// x, mean and inv_std (never materialised) and writes gx.  Both are HBM-bound streams:
// |x| read + |y| written (forward), |gy| + |x| read + |gx| written (backward), 4-element
// vectors per thread when a channel plane (H*W) is a multiple of 4 elements.
// Arithmetic: float (double for float64 activations), the operation order of the reference
// kernels with every operation rounded separately (the reference's NVRTC build may
// contract to FMA: parity is at the 1e-6 level the north star states, tests/test_bn_apply_gpu.py).
#include "gp_common.cuh"

namespace {

__device__ __forceinline__ double ld_stat(const void* p, int dtype, int64_t i) {
  switch (dtype) {
    case GP_F16: return (double)__half2float(reinterpret_cast<const __half*>(p)[i]);
    case GP_F32: return (double)reinterpret_cast<const float*>(p)[i];
    default: return reinterpret_cast<const double*>(p)[i];
  }
}
__device__ __forceinline__ void st_stat(void* p, int dtype, int64_t i, double v) {
  switch (dtype) {
    case GP_F16: reinterpret_cast<__half*>(p)[i] = __double2half(v); break;
    case GP_F32: reinterpret_cast<float*>(p)[i] = __double2float_rn(v); break;
    default: reinterpret_cast<double*>(p)[i] = v; break;
  }
}

template <class C> __device__ __forceinline__ C rsqrt_rn(C v);
template <> __device__ __forceinline__ float rsqrt_rn(float v) { return __frsqrt_rn(v); }
template <> __device__ __forceinline__ double rsqrt_rn(double v) { return 1.0 / sqrt(v); }

struct ApplyArgs {
  const void* x;
  const void* gy;
  void* out;          // y (forward) / gx (backward)
  const void* mean;
  const void* var;      // forward
  const void* inv_std;  // backward
  const void* gamma;
  const void* beta;     // forward
  const void* ggamma;   // backward
  const void* gbeta;    // backward
  int stat_dtype;
  int64_t N, C, HW;
  double eps;           // forward
  double inv_m;         // backward
  void* inv_std_out;    // forward, may be NULL
  void* running_mean;   // forward, may be NULL
  void* running_var;
  int running_dtype;
  double decay, adjust;
};

// Work decomposition: a UNIT is up to kUnit consecutive elements of one (n, c) plane and
// is handled by one warp (2 vectors / 2 scalars per lane), so the channel -- and the two
// integer divisions that find it -- is resolved once per unit, not per element; planes of
// any size (112 x 112 down to 7 x 7) keep every lane busy with coalesced accesses.
template <bool VEC> struct UnitOf { static constexpr int value = VEC ? 256 : 64; };

struct Units {
  uint32_t upp;      // units per plane
  uint32_t n_units;  // planes * upp
};

template <class CT> struct FwdCh { CT mean, inv_std, gamma, beta; };
template <class CT> struct BwdCh { CT mean, inv_std, gi, ggamma, gbeta, inv_m; };

template <class CT>
__device__ __forceinline__ CT fwd_value(CT x, const FwdCh<CT>& k) {
  using I = Inter<CT>;   // gamma * (x - mean) * inv_std + beta, every operation rounded
  return I::add(I::mul(I::mul(k.gamma, I::sub(x, k.mean)), k.inv_std), k.beta);
}
template <class CT>
__device__ __forceinline__ CT bwd_value(CT gy, CT x, const BwdCh<CT>& k) {
  using I = Inter<CT>;   // (gamma * inv_std) * (gy - (x_hat * ggamma + gbeta) * inv_m)
  const CT xh = I::mul(I::sub(x, k.mean), k.inv_std);
  return I::mul(k.gi, I::sub(gy, I::mul(I::add(I::mul(xh, k.ggamma), k.gbeta), k.inv_m)));
}

template <class T, bool VEC>
__global__ void __launch_bounds__(256) bn_fwd_apply_kernel(const ApplyArgs a, const Units u) {
  using CT = typename Carrier<T>::type;
  using I = Inter<CT>;
  const T* __restrict__ x = reinterpret_cast<const T*>(a.x);
  T* __restrict__ y = reinterpret_cast<T*>(a.out);
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;

  // per-channel outputs: inv_std and the running statistics (first C threads of the grid)
  for (int64_t c = gtid; c < a.C; c += gthreads) {
    const CT mean = (CT)ld_stat(a.mean, a.stat_dtype, c);
    const CT var = (CT)ld_stat(a.var, a.stat_dtype, c);
    if (a.inv_std_out)
      st_stat(a.inv_std_out, a.stat_dtype, c, (double)rsqrt_rn<CT>(I::add(var, (CT)a.eps)));
    if (a.running_mean) {
      const CT decay = (CT)a.decay, adjust = (CT)a.adjust;
      const CT rm = (CT)ld_stat(a.running_mean, a.running_dtype, c);
      const CT rv = (CT)ld_stat(a.running_var, a.running_dtype, c);
      // r_mean * decay + mean * (1 - decay);  r_var * decay + var * (1 - decay) * adjust
      const CT omd = I::sub((CT)1, decay);
      st_stat(a.running_mean, a.running_dtype, c, (double)I::add(I::mul(rm, decay), I::mul(mean, omd)));
      st_stat(a.running_var, a.running_dtype, c,
              (double)I::add(I::mul(rv, decay), I::mul(I::mul(var, omd), adjust)));
    }
  }

  constexpr int kUnit = UnitOf<VEC>::value;
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (uint32_t)(gthreads >> 5);
  for (uint32_t unit = (uint32_t)(gtid >> 5); unit < u.n_units; unit += warps) {
    const uint32_t plane = unit / u.upp;
    const uint32_t c = plane % (uint32_t)a.C;
    const int64_t e0 = (int64_t)(unit - plane * u.upp) * kUnit;     // inside the plane
    const int64_t base = (int64_t)plane * a.HW;
    FwdCh<CT> k;
    k.mean = (CT)ld_stat(a.mean, a.stat_dtype, c);
    k.inv_std = rsqrt_rn<CT>(I::add((CT)ld_stat(a.var, a.stat_dtype, c), (CT)a.eps));
    k.gamma = (CT)ld_stat(a.gamma, a.stat_dtype, c);
    k.beta = (CT)ld_stat(a.beta, a.stat_dtype, c);
    if (VEC) {
      Raw4<T> r[2];
      bool act[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int64_t e = e0 + (q * 32 + lane) * 4;
        act[q] = e < a.HW;
        if (act[q]) r[q] = ld4_stream(x + base + e);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (!act[q]) continue;
        CT v[4];
        unpack4(r[q], v);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = fwd_value(v[i], k);
        st4(y + base + e0 + (q * 32 + lane) * 4, pack4<T, CT>(v));
      }
    } else {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int64_t e = e0 + q * 32 + lane;
        if (e < a.HW) y[base + e] = from_carrier<T>(fwd_value((CT)to_carrier(x[base + e]), k));
      }
    }
  }
}

template <class T, bool VEC>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const ApplyArgs a, const Units u) {
  using CT = typename Carrier<T>::type;
  using I = Inter<CT>;
  const T* __restrict__ x = reinterpret_cast<const T*>(a.x);
  const T* __restrict__ gy = reinterpret_cast<const T*>(a.gy);
  T* __restrict__ gx = reinterpret_cast<T*>(a.out);
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
  constexpr int kUnit = UnitOf<VEC>::value;
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (uint32_t)(gthreads >> 5);
  for (uint32_t unit = (uint32_t)(gtid >> 5); unit < u.n_units; unit += warps) {
    const uint32_t plane = unit / u.upp;
    const uint32_t c = plane % (uint32_t)a.C;
    const int64_t e0 = (int64_t)(unit - plane * u.upp) * kUnit;
    const int64_t base = (int64_t)plane * a.HW;
    BwdCh<CT> k;
    k.mean = (CT)ld_stat(a.mean, a.stat_dtype, c);
    k.inv_std = (CT)ld_stat(a.inv_std, a.stat_dtype, c);
    k.gi = I::mul((CT)ld_stat(a.gamma, a.stat_dtype, c), k.inv_std);
    k.ggamma = (CT)ld_stat(a.ggamma, a.stat_dtype, c);
    k.gbeta = (CT)ld_stat(a.gbeta, a.stat_dtype, c);
    k.inv_m = (CT)a.inv_m;
    if (VEC) {
      Raw4<T> rx[2], rg[2];
      bool act[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int64_t e = e0 + (q * 32 + lane) * 4;
        act[q] = e < a.HW;
        if (act[q]) {
          rx[q] = ld4_stream(x + base + e);
          rg[q] = ld4_stream(gy + base + e);
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (!act[q]) continue;
        CT vx[4], vg[4];
        unpack4(rx[q], vx);
        unpack4(rg[q], vg);
#pragma unroll
        for (int i = 0; i < 4; ++i) vx[i] = bwd_value(vg[i], vx[i], k);
        st4(gx + base + e0 + (q * 32 + lane) * 4, pack4<T, CT>(vx));
      }
    } else {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int64_t e = e0 + q * 32 + lane;
        if (e < a.HW)
          gx[base + e] = from_carrier<T>(bwd_value((CT)to_carrier(gy[base + e]),
                                                   (CT)to_carrier(x[base + e]), k));
      }
    }
  }
}

// units of the launch, or an error for sizes beyond 32-bit unit counts
bool make_units(int64_t N, int64_t C, int64_t HW, bool vec, Units* u) {
  const int64_t per = vec ? 256 : 64;
  const int64_t upp = (HW + per - 1) / per;
  const int64_t n = N * C * upp;
  if (n >= ((int64_t)1 << 31) || N * C >= ((int64_t)1 << 31)) return false;
  u->upp = (uint32_t)upp;
  u->n_units = (uint32_t)n;
  return true;
}

int apply_grid(const Units& u) {
  int64_t g = ((int64_t)u.n_units + 7) / 8;           // 8 warps per CTA, one unit each
  const int64_t cap = (int64_t)gp_sm_count_cached() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

bool aligned4(const void* p, int dtype) {
  return ((uintptr_t)p % (gp_itemsize(dtype) == 2 ? 8 : 16)) == 0;
}

}  // namespace

extern "C" int gp_bn_fwd_apply(const void* x, int x_dtype, int64_t N, int64_t C, int64_t HW,
                               const void* mean, const void* var, const void* gamma,
                               const void* beta, int stat_dtype, double eps, void* y,
                               void* inv_std_out, void* running_mean, void* running_var,
                               int running_dtype, double decay, double adjust, void* stream) {
  if (N <= 0 || C <= 0 || HW <= 0) return 0;
  if ((running_mean == nullptr) != (running_var == nullptr)) {
    gp_set_error("gp_bn_fwd_apply: running_mean and running_var go together");
    return GP_EINVAL;
  }
  ApplyArgs a = {};
  a.x = x; a.out = y; a.mean = mean; a.var = var; a.gamma = gamma; a.beta = beta;
  a.stat_dtype = stat_dtype; a.N = N; a.C = C; a.HW = HW; a.eps = eps;
  a.inv_std_out = inv_std_out; a.running_mean = running_mean; a.running_var = running_var;
  a.running_dtype = running_dtype; a.decay = decay; a.adjust = adjust;
  const bool vec = (HW % 4 == 0) && aligned4(x, x_dtype) && aligned4(y, x_dtype) && x_dtype != GP_F64;
  Units u;
  if (!make_units(N, C, HW, vec, &u)) {
    gp_set_error("gp_bn_fwd_apply: activation too large");
    return GP_EINVAL;
  }
  const int grid = apply_grid(u);
  cudaStream_t st = (cudaStream_t)stream;
  switch (x_dtype) {
    case GP_F32:
      if (vec) bn_fwd_apply_kernel<float, true><<<grid, 256, 0, st>>>(a, u);
      else bn_fwd_apply_kernel<float, false><<<grid, 256, 0, st>>>(a, u);
      break;
    case GP_F16:
      if (vec) bn_fwd_apply_kernel<__half, true><<<grid, 256, 0, st>>>(a, u);
      else bn_fwd_apply_kernel<__half, false><<<grid, 256, 0, st>>>(a, u);
      break;
    case GP_F64: bn_fwd_apply_kernel<double, false><<<grid, 256, 0, st>>>(a, u); break;
    default:
      gp_set_error("gp_bn_fwd_apply: unsupported x dtype id %d", x_dtype);
      return GP_EINVAL;
  }
  return gp_cuda_fail(cudaGetLastError(), "bn_fwd_apply_kernel launch");
}

extern "C" int gp_bn_bwd_apply(const void* gy, int gy_dtype, const void* x, int x_dtype, int64_t N,
                               int64_t C, int64_t HW, const void* mean, const void* inv_std,
                               const void* gamma, const void* ggamma, const void* gbeta,
                               int stat_dtype, double inv_m, void* gx, void* stream) {
  if (N <= 0 || C <= 0 || HW <= 0) return 0;
  if (gy_dtype != x_dtype) {
    gp_set_error("gp_bn_bwd_apply: gy and x must have one dtype (got %d and %d)", gy_dtype, x_dtype);
    return GP_EINVAL;
  }
  ApplyArgs a = {};
  a.x = x; a.gy = gy; a.out = gx; a.mean = mean; a.inv_std = inv_std; a.gamma = gamma;
  a.ggamma = ggamma; a.gbeta = gbeta; a.stat_dtype = stat_dtype; a.N = N; a.C = C; a.HW = HW;
  a.inv_m = inv_m;
  const bool vec = (HW % 4 == 0) && aligned4(x, x_dtype) && aligned4(gy, x_dtype) &&
                   aligned4(gx, x_dtype) && x_dtype != GP_F64;
  Units u;
  if (!make_units(N, C, HW, vec, &u)) {
    gp_set_error("gp_bn_bwd_apply: activation too large");
    return GP_EINVAL;
  }
  const int grid = apply_grid(u);
  cudaStream_t st = (cudaStream_t)stream;
  switch (x_dtype) {
    case GP_F32:
      if (vec) bn_bwd_apply_kernel<float, true><<<grid, 256, 0, st>>>(a, u);
      else bn_bwd_apply_kernel<float, false><<<grid, 256, 0, st>>>(a, u);
      break;
    case GP_F16:
      if (vec) bn_bwd_apply_kernel<__half, true><<<grid, 256, 0, st>>>(a, u);
      else bn_bwd_apply_kernel<__half, false><<<grid, 256, 0, st>>>(a, u);
      break;
    case GP_F64: bn_bwd_apply_kernel<double, false><<<grid, 256, 0, st>>>(a, u); break;
    default:
      gp_set_error("gp_bn_bwd_apply: unsupported dtype id %d", x_dtype);
      return GP_EINVAL;
  }
  return gp_cuda_fail(cudaGetLastError(), "bn_bwd_apply_kernel launch");
}
