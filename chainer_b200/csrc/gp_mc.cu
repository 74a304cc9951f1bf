// gp_mc.cu -- allreduce of the packed gradient buffer through NVSwitch MULTICAST
// (NVLS): ONE kernel per rank whose loads are reduced inside the switch
// (`multimem.ld_reduce`) and whose stores are replicated by it (`multimem.st`).
//
// Reference being replaced: `nccl_comm.allReduce(sendbuf, recvbuf, n_elems,
// type_id, NCCL_SUM, stream.ptr)` in
// chainermn/communicators/pure_nccl_communicator.py:180-182, for 4 and 8 ranks on
// one NVSwitch box.
//
// Why: the peer-memory kernel (gp_p2p.cu) moves S(N-1)/N bytes in and out of
// every GPU twice (gather the N copies of the own shard, scatter the sum to N
// buffers).  With a multicast mapping of the N packed buffers rank r issues, for
// every 16-byte vector of ITS shard, one `multimem.ld_reduce` (the switch pulls
// the vector from the N GPUs and returns the sum: S/N bytes into the GPU) and one
// `multimem.st` (S/N bytes out, replicated by the switch), so each NVLink
// direction carries about S(N+1)/N instead of 2S(N-1)/N -- 1.56x less at N = 8.
// The order in which the switch adds the N values is fixed by the hardware, not
// by rank, so results agree with the oracle to rounding (fp32: 1e-6 relative),
// and are identical on all ranks (one rank reduces each element, all receive
// the same bits).  fp16 / bf16 buffers are accumulated in fp32 (`.acc::f32`).
//
// The buffers are cuMemCreate allocations bound to one multicast object
// (cuMulticastCreate / AddDevice / BindMem) and mapped twice: a unicast address
// (the packed buffer the pack and update kernels use) and the multicast
// address.  The cross-GPU barriers are the flag words of gp_p2p.cu.
// Driver entry points are resolved at run time (cudaGetDriverEntryPoint): the
// library has no link-time dependency on libcuda.
#include <cuda.h>
#include <string.h>
#include <unistd.h>

#include "gp_common.cuh"

namespace {

#include "gp_p2p.cuh"

struct DriverApi {
  bool tried, ok;
  CUresult (*GetErrorString)(CUresult, const char**);
  CUresult (*DeviceGet)(CUdevice*, int);
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
};
DriverApi g_drv = {};

bool load_driver() {
  if (g_drv.tried) return g_drv.ok;
  g_drv.tried = true;
  if (cudaFree(0) != cudaSuccess) return false;   // primary context current
  bool ok = true;
#define GP_DRV(field, sym)                                                                  \
  do {                                                                                      \
    void* fn = nullptr;                                                                     \
    cudaDriverEntryPointQueryResult qr;                                                     \
    if (cudaGetDriverEntryPoint(sym, &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn ||  \
        qr != cudaDriverEntryPointSuccess) {                                                \
      ok = false;                                                                           \
      (void)cudaGetLastError();                                                             \
    }                                                                                       \
    *(void**)(&g_drv.field) = fn;                                                           \
  } while (0)
  GP_DRV(GetErrorString, "cuGetErrorString");
  GP_DRV(DeviceGet, "cuDeviceGet");
  GP_DRV(DeviceGetAttribute, "cuDeviceGetAttribute");
  GP_DRV(MulticastGetGranularity, "cuMulticastGetGranularity");
  GP_DRV(MulticastCreate, "cuMulticastCreate");
  GP_DRV(MulticastAddDevice, "cuMulticastAddDevice");
  GP_DRV(MulticastBindMem, "cuMulticastBindMem");
  GP_DRV(MulticastUnbind, "cuMulticastUnbind");
  GP_DRV(MemGetAllocationGranularity, "cuMemGetAllocationGranularity");
  GP_DRV(MemCreate, "cuMemCreate");
  GP_DRV(MemRelease, "cuMemRelease");
  GP_DRV(MemExportToShareableHandle, "cuMemExportToShareableHandle");
  GP_DRV(MemImportFromShareableHandle, "cuMemImportFromShareableHandle");
  GP_DRV(MemAddressReserve, "cuMemAddressReserve");
  GP_DRV(MemAddressFree, "cuMemAddressFree");
  GP_DRV(MemMap, "cuMemMap");
  GP_DRV(MemUnmap, "cuMemUnmap");
  GP_DRV(MemSetAccess, "cuMemSetAccess");
#undef GP_DRV
  g_drv.ok = ok;
  return ok;
}

int drv_fail(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return 0;
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  gp_set_error("CUDA driver error %d (%s) in %s", (int)r, s ? s : "?", what);
  return -(3000 + (int)r);
}
#define GP_DRVCALL(call, what)                  \
  do {                                          \
    int _e = drv_fail(g_drv.call, what);        \
    if (_e) return _e;                          \
  } while (0)

struct McBuffer {
  int rank, n;
  CUdevice dev;
  size_t size;                       // bytes, multiple of every granularity involved
  CUmemGenericAllocationHandle mc;   // the multicast object (created by rank 0, imported elsewhere)
  CUmemGenericAllocationHandle mem;  // this rank's physical memory
  CUdeviceptr uc_va, mc_va;
  bool have_mc, added, have_mem, bound, uc_mapped, mc_mapped;
};

CUmulticastObjectProp mc_prop(int n, size_t size) {
  CUmulticastObjectProp p;
  memset(&p, 0, sizeof(p));
  p.numDevices = (unsigned)n;
  p.size = size;
  p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  p.flags = 0;
  return p;
}

CUmemAllocationProp mem_prop(CUdevice dev) {
  CUmemAllocationProp p;
  memset(&p, 0, sizeof(p));
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = (int)dev;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

// ---------------------------------------------------------------------------
// multimem accessors: 16 bytes per instruction
template <class T> struct Mm;
template <> struct Mm<float> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
  }
};
template <> struct Mm<__half> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
  }
};
template <> struct Mm<__nv_bfloat16> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
  }
};
__device__ __forceinline__ void mm_st(void* p, const uint4& v) {
  // the element type of a multimem store only names the vector shape
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

struct McArgs {
  P2PArgs p;        // flags, rank, n, epoch; begin/end = this rank's shard in 16-byte vectors
  char* mc_base;    // multicast address of the first vector
};

template <class T, int UN>
__global__ void __launch_bounds__(512) mc_allreduce_kernel(const McArgs a) {
  // ---- barrier 1: every rank has finished writing (packing) its buffer ------
  if (blockIdx.x == 0) signal_all(a.p, 0, a.p.epoch);
  wait_all(a.p, 0, a.p.epoch);
  __syncthreads();

  const int64_t v_end = a.p.end;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v0 = a.p.begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < v_end;
       v0 += stride * UN) {
    uint4 x[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < v_end) x[u] = Mm<T>::ld_reduce(a.mc_base + v * 16);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < v_end) mm_st(a.mc_base + v * 16, x[u]);
    }
  }
  // ---- barrier 2: every rank's stores have landed everywhere ----------------
  finish_all(a.p);
}

// few, small CTAs keep the switch's reduction pipeline fed without over-subscribing
// it (tools/p2p_bench.py at N = 8: 32 x 256 threads 241 us, 296 x 512 304 us)
int g_mc_ctas = 32;
int g_mc_threads = 256;
int g_mc_unroll = 4;

template <class T>
int launch_mc(const McArgs& a, int grid, int threads, cudaStream_t st) {
  switch (g_mc_unroll) {
    case 1: mc_allreduce_kernel<T, 1><<<grid, threads, 0, st>>>(a); break;
    case 2: mc_allreduce_kernel<T, 2><<<grid, threads, 0, st>>>(a); break;
    case 8: mc_allreduce_kernel<T, 8><<<grid, threads, 0, st>>>(a); break;
    default: mc_allreduce_kernel<T, 4><<<grid, threads, 0, st>>>(a); break;
  }
  return gp_cuda_fail(cudaGetLastError(), "mc_allreduce_kernel launch");
}

}  // namespace

extern "C" {

int gp_mc_supported(int* supported) {
  *supported = 0;
  if (!load_driver()) return 0;
  int ordinal = 0;
  GP_CUDA(cudaGetDevice(&ordinal));
  CUdevice dev;
  if (g_drv.DeviceGet(&dev, ordinal) != CUDA_SUCCESS) return 0;
  int v = 0;
  if (g_drv.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS)
    return 0;
  *supported = v;
  return 0;
}

int gp_mc_create(void** out, int rank, int n_ranks, size_t nbytes) {
  if (!load_driver()) {
    gp_set_error("gp_mc_create: CUDA driver multicast entry points are unavailable");
    return GP_ENOSYS;
  }
  if (n_ranks < 2 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks || nbytes == 0) {
    gp_set_error("gp_mc_create: bad arguments (rank %d of %d, %zu bytes)", rank, n_ranks, nbytes);
    return GP_EINVAL;
  }
  int ordinal = 0;
  GP_CUDA(cudaGetDevice(&ordinal));
  McBuffer* b = new McBuffer();
  memset(b, 0, sizeof(*b));
  b->rank = rank;
  b->n = n_ranks;
  int e = drv_fail(g_drv.DeviceGet(&b->dev, ordinal), "cuDeviceGet");
  if (e) { delete b; return e; }
  // size: a multiple of the multicast granularity and of the allocation granularity
  CUmulticastObjectProp mp = mc_prop(n_ranks, nbytes);
  size_t g_mc = 0, g_mem = 0;
  e = drv_fail(g_drv.MulticastGetGranularity(&g_mc, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED),
               "cuMulticastGetGranularity");
  if (e) { delete b; return e; }
  CUmemAllocationProp ap = mem_prop(b->dev);
  e = drv_fail(g_drv.MemGetAllocationGranularity(&g_mem, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
               "cuMemGetAllocationGranularity");
  if (e) { delete b; return e; }
  size_t g = g_mc > g_mem ? g_mc : g_mem;
  if (g == 0 || (g % g_mem) || (g % g_mc)) g = g_mc * g_mem;   // not expected: powers of two
  b->size = (nbytes + g - 1) / g * g;
  if (rank == 0) {
    mp.size = b->size;
    e = drv_fail(g_drv.MulticastCreate(&b->mc, &mp), "cuMulticastCreate");
    if (e) { delete b; return e; }
    b->have_mc = true;
  }
  *out = b;
  return 0;
}

int gp_mc_export_fd(void* mc, int* fd) {
  McBuffer* b = (McBuffer*)mc;
  if (!b || !b->have_mc) return GP_EINVAL;
  int h = -1;
  GP_DRVCALL(MemExportToShareableHandle(&h, b->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
             "cuMemExportToShareableHandle");
  *fd = h;
  return 0;
}

int gp_mc_import_fd(void* mc, int fd) {
  McBuffer* b = (McBuffer*)mc;
  if (!b || b->have_mc) return GP_EINVAL;
  GP_DRVCALL(MemImportFromShareableHandle(&b->mc, (void*)(uintptr_t)fd,
                                          CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
             "cuMemImportFromShareableHandle");
  b->have_mc = true;
  return 0;
}

int gp_mc_add_device(void* mc) {
  McBuffer* b = (McBuffer*)mc;
  if (!b || !b->have_mc) return GP_EINVAL;
  GP_DRVCALL(MulticastAddDevice(b->mc, b->dev), "cuMulticastAddDevice");
  b->added = true;
  return 0;
}

// After EVERY rank has returned from gp_mc_add_device: allocate this rank's
// memory, bind it to the multicast object and map both views.
int gp_mc_bind(void* mc) {
  McBuffer* b = (McBuffer*)mc;
  if (!b || !b->added) return GP_EINVAL;
  CUmemAllocationProp ap = mem_prop(b->dev);
  GP_DRVCALL(MemCreate(&b->mem, b->size, &ap, 0), "cuMemCreate");
  b->have_mem = true;
  GP_DRVCALL(MulticastBindMem(b->mc, 0, b->mem, 0, b->size, 0), "cuMulticastBindMem");
  b->bound = true;
  CUmemAccessDesc ad;
  memset(&ad, 0, sizeof(ad));
  ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ad.location.id = (int)b->dev;
  ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  GP_DRVCALL(MemAddressReserve(&b->uc_va, b->size, 0, 0, 0), "cuMemAddressReserve (unicast)");
  GP_DRVCALL(MemMap(b->uc_va, b->size, 0, b->mem, 0), "cuMemMap (unicast)");
  b->uc_mapped = true;
  GP_DRVCALL(MemSetAccess(b->uc_va, b->size, &ad, 1), "cuMemSetAccess (unicast)");
  GP_DRVCALL(MemAddressReserve(&b->mc_va, b->size, 0, 0, 0), "cuMemAddressReserve (multicast)");
  GP_DRVCALL(MemMap(b->mc_va, b->size, 0, b->mc, 0), "cuMemMap (multicast)");
  b->mc_mapped = true;
  GP_DRVCALL(MemSetAccess(b->mc_va, b->size, &ad, 1), "cuMemSetAccess (multicast)");
  return 0;
}

int gp_mc_pointers(void* mc, void** unicast, void** multicast, size_t* nbytes) {
  McBuffer* b = (McBuffer*)mc;
  if (!b) return GP_EINVAL;
  if (unicast) *unicast = (void*)b->uc_va;
  if (multicast) *multicast = (void*)b->mc_va;
  if (nbytes) *nbytes = b->size;
  return 0;
}

int gp_mc_destroy(void* mc) {
  McBuffer* b = (McBuffer*)mc;
  if (!b) return 0;
  if (b->mc_mapped) g_drv.MemUnmap(b->mc_va, b->size);
  if (b->mc_va) g_drv.MemAddressFree(b->mc_va, b->size);
  if (b->bound) g_drv.MulticastUnbind(b->mc, b->dev, 0, b->size);
  if (b->uc_mapped) g_drv.MemUnmap(b->uc_va, b->size);
  if (b->uc_va) g_drv.MemAddressFree(b->uc_va, b->size);
  if (b->have_mem) g_drv.MemRelease(b->mem);
  if (b->have_mc) g_drv.MemRelease(b->mc);
  delete b;
  return 0;
}

int gp_mc_allreduce(void* p2p_comm, void* mc, int dtype, int64_t offset_elems, int64_t n_elems,
                    void* stream) {
  P2PComm* c = (P2PComm*)p2p_comm;
  McBuffer* b = (McBuffer*)mc;
  if (!c || !b || !b->mc_mapped || c->n != b->n || c->rank != b->rank) {
    gp_set_error("gp_mc_allreduce: communicator / multicast buffer mismatch");
    return GP_EINVAL;
  }
  if (n_elems <= 0) return 0;
  if (dtype != GP_F32 && dtype != GP_F16 && dtype != GP_BF16) {
    gp_set_error("gp_mc_allreduce: dtype id %d is not supported (float32, float16, bfloat16)", dtype);
    return GP_EINVAL;
  }
  const int isz = gp_itemsize(dtype);
  const int E = 16 / isz;
  if (offset_elems % E) {
    gp_set_error("gp_mc_allreduce: offset must be a multiple of %d elements", E);
    return GP_EINVAL;
  }
  // whole 16-byte vectors: a ragged tail is rounded up into the padding of the
  // allocation (its sum is never read)
  const int64_t n_vec = (n_elems + E - 1) / E;
  if ((size_t)((offset_elems / E + n_vec) * 16) > b->size) {
    gp_set_error("gp_mc_allreduce: range exceeds the multicast buffer");
    return GP_EINVAL;
  }
  McArgs a;
  a.p.rank = c->rank;
  a.p.n = c->n;
  a.p.epoch = ++c->epoch;
  a.p.timeout_ns = g_gp_peer_timeout_ns;
  for (int k = 0; k < c->n; ++k) {
    a.p.bufs[k] = nullptr;
    a.p.flags[k] = c->flags[k];
  }
  const int64_t per = (n_vec + c->n - 1) / c->n;
  int64_t vb = per * c->rank, ve = per * (c->rank + 1);
  if (vb > n_vec) vb = n_vec;
  if (ve > n_vec) ve = n_vec;
  a.p.begin = vb;
  a.p.end = ve;
  a.mc_base = (char*)b->mc_va + offset_elems * isz;
  const int threads = g_mc_threads;
  int64_t grid = g_mc_ctas > 0 ? g_mc_ctas : gp_sm_count_cached();
  const int64_t need = ((ve - vb) + threads - 1) / threads;
  if (grid > need) grid = need > 0 ? need : 1;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case GP_F32: return launch_mc<float>(a, (int)grid, threads, st);
    case GP_F16: return launch_mc<__half>(a, (int)grid, threads, st);
    default: return launch_mc<__nv_bfloat16>(a, (int)grid, threads, st);
  }
}

int gp_mc_set_tuning(int ctas, int threads, int unroll) {
  g_mc_ctas = ctas;
  if (threads >= 32 && threads <= 512) g_mc_threads = threads & ~31;
  if (unroll == 1 || unroll == 2 || unroll == 4 || unroll == 8) g_mc_unroll = unroll;
  return 0;
}

}  // extern "C"
