// gp_pack.cu -- gp_pack / gp_unpack_scale / gp_scale / gp_check_finite.
//
// Reference being replaced:
//   chainermn/communicators/_memory_utility.py:253-268, 289-358  (batched pack)
//   chainermn/communicators/_memory_utility.py:271-286, 361-429  (batched unpack)
//   chainermn/communicators/pure_nccl_communicator.py:183-189    (div_by_size)
//   chainermn/communicators/mpi_communicator_base.py:730-733     (_ensure_all_finite)
// Layout contract (bit-exact with the reference): element k of parameter j sits
// at flat index csum[j] + k of the packed buffer, parameters in the order the
// caller lists them (sorted(model.namedparams()), _memory_utility.py:154-165).
#include "gp_pack_op.cuh"

namespace {

// ------------------------------------------------- flat (single array) ops --
template <class B>
__global__ void __launch_bounds__(256) scale_kernel(B* __restrict__ buf, int64_t n, ScaleArg s) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    typename Carrier<B>::type x[4];
    unpack4(ld4(buf + 4 * i), x);
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = descale_rt<B>(x[k], s);
    st4(buf + 4 * i, pack4<B, typename Carrier<B>::type>(x));
  }
  // tail (< 4 elements)
  const int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) buf[t] = from_carrier<B>(descale_rt<B>(to_carrier(buf[t]), s));
}

__device__ __forceinline__ bool is_finite_c(float x) { return isfinite(x); }
__device__ __forceinline__ bool is_finite_c(double x) { return isfinite(x); }

template <class B>
__global__ void __launch_bounds__(256) finite_kernel(const B* __restrict__ buf, int64_t n,
                                                     int32_t* flag) {
  bool bad = false;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    typename Carrier<B>::type x[4];
    unpack4(ld4_stream(buf + 4 * i), x);
#pragma unroll
    for (int k = 0; k < 4; ++k) bad = bad || !is_finite_c(x[k]);
  }
  const int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) bad = bad || !is_finite_c(to_carrier(buf[t]));
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

int flat_grid(int64_t n) {
  int64_t g = (n / 4 + 255) / 256;
  const int64_t cap = (int64_t)gp_sm_count_cached() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" int gp_pack(void* buffer, int buf_dtype, const int64_t* d_csum, const gp_seg_t* d_segs,
                       int n_segs, int64_t elem_begin, int64_t elem_end, double scale,
                       int layout_hint, void* stream) {
  PackOp op;
  op.buffer = buffer;
  op.s = make_scale(scale);
  return gpw::launch_buf(buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, op, stream,
                         "gp_pack", layout_hint == GP_F32);
}

extern "C" int gp_unpack_scale(const void* buffer, int buf_dtype, const int64_t* d_csum,
                               const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                               int64_t elem_end, double scale, int layout_hint, void* stream) {
  UnpackOp op;
  op.buffer = buffer;
  op.s = make_scale(scale);
  return gpw::launch_buf(buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, op, stream,
                         "gp_unpack_scale", layout_hint == GP_F32);
}

extern "C" int gp_scale(void* buffer, int dtype, int64_t n, double scale, void* stream) {
  if (n <= 0) return 0;
  if (((uintptr_t)buffer) & (gp_itemsize(dtype) == 2 ? 7 : 15)) {
    gp_set_error("gp_scale: buffer must be aligned to 4 elements");
    return GP_EINVAL;
  }
  const ScaleArg s = make_scale(scale);
  if (s.mode == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = flat_grid(n);
  switch (dtype) {
    case GP_F32: scale_kernel<float><<<g, 256, 0, st>>>((float*)buffer, n, s); break;
    case GP_F16: scale_kernel<__half><<<g, 256, 0, st>>>((__half*)buffer, n, s); break;
    case GP_BF16: scale_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((__nv_bfloat16*)buffer, n, s); break;
    case GP_F64: scale_kernel<double><<<g, 256, 0, st>>>((double*)buffer, n, s); break;
    default:
      gp_set_error("gp_scale: unsupported dtype id %d", dtype);
      return GP_EINVAL;
  }
  return gp_cuda_fail(cudaGetLastError(), "gp_scale launch");
}

extern "C" int gp_check_finite(const void* buffer, int dtype, int64_t n, int32_t* d_flag,
                               void* stream) {
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = flat_grid(n);
  switch (dtype) {
    case GP_F32: finite_kernel<float><<<g, 256, 0, st>>>((const float*)buffer, n, d_flag); break;
    case GP_F16: finite_kernel<__half><<<g, 256, 0, st>>>((const __half*)buffer, n, d_flag); break;
    case GP_BF16:
      finite_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)buffer, n, d_flag);
      break;
    case GP_F64: finite_kernel<double><<<g, 256, 0, st>>>((const double*)buffer, n, d_flag); break;
    default:
      gp_set_error("gp_check_finite: unsupported dtype id %d", dtype);
      return GP_EINVAL;
  }
  return gp_cuda_fail(cudaGetLastError(), "gp_check_finite launch");
}
