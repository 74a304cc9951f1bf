// gp_runtime.cu -- error reporting, tuning knobs and the device plumbing of
// libgradpath: memory, streams, events, and the sync-free table upload.
//
// Reference being replaced: the CuPy plumbing used by
// chainermn/communicators/_memory_utility.py:65-151 (DeviceMemory,
// HostPinnedMemory: cupy.cuda.alloc / alloc_pinned_memory /
// copy_from_device_async) and the three synchronous `cupy.asarray` uploads of
// ParamsData (_memory_utility.py:59-61).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>

#include "gp_bulk.cuh"

static thread_local char g_err[512] = "";

void gp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int gp_cuda_fail(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  gp_set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return -(1000 + (int)e);
}

// defaults: see DESIGN.md "tuning"
GpTuning g_gp_tuning = {/*threads*/ 256, /*unroll*/ 0, /*ctas_per_sm*/ 16, /*persistent*/ 0,
                        /*bn_threads*/ 256, /*pipeline*/ 0, /*bn_ctas_per_sm*/ 4};

// 30 minutes: a collective watchdog, not a skew limit (gp_p2p.cuh: spin_until)
unsigned long long g_gp_peer_timeout_ns = 1800ull * 1000000000ull;

gpb::BulkTuning gpb::g_bulk_tuning = {/*enable*/ 0, /*tile*/ 4096, /*stages*/ 4, /*ctas*/ 1, /*debug*/ 0, /*chunk*/ 2048};

int gp_sm_count_cached() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) {
      cached_sms = sms;
      cached_dev = dev;
    }
  }
  return cached_sms;
}

extern "C" {

const char* gp_last_error(void) { return g_err; }
int gp_abi_version(void) { return GP_ABI_VERSION; }

int gp_set_tuning(const char* key, int value) {
  if (!key) return GP_EINVAL;
  if (!strcmp(key, "threads")) g_gp_tuning.threads = value;
  else if (!strcmp(key, "unroll")) g_gp_tuning.unroll = value;
  else if (!strcmp(key, "ctas_per_sm")) g_gp_tuning.ctas_per_sm = value;
  else if (!strcmp(key, "persistent")) g_gp_tuning.persistent = value;
  else if (!strcmp(key, "bn_threads")) g_gp_tuning.bn_threads = value;
  else if (!strcmp(key, "pipeline")) g_gp_tuning.pipeline = value;
  else if (!strcmp(key, "bn_ctas_per_sm")) g_gp_tuning.bn_ctas_per_sm = value < 1 ? 1 : value;
  else if (!strcmp(key, "peer_timeout_s")) g_gp_peer_timeout_ns = value <= 0 ? 0ull : (unsigned long long)value * 1000000000ull;
  else if (!strcmp(key, "bulk")) gpb::g_bulk_tuning.enable = value;
  else if (!strcmp(key, "bulk_tile")) gpb::g_bulk_tuning.tile = value < 512 ? 512 : (value & ~511);
  else if (!strcmp(key, "bulk_stages")) gpb::g_bulk_tuning.stages = value < 2 ? 2 : value;
  else if (!strcmp(key, "bulk_ctas")) gpb::g_bulk_tuning.ctas = value;
  else if (!strcmp(key, "bulk_debug")) gpb::g_bulk_tuning.debug = value;
  else if (!strcmp(key, "bulk_chunk")) gpb::g_bulk_tuning.chunk = value;
  else {
    gp_set_error("gp_set_tuning: unknown key '%s'", key);
    return GP_EINVAL;
  }
  return 0;
}
int gp_get_tuning(const char* key, int* value) {
  if (!key || !value) return GP_EINVAL;
  if (!strcmp(key, "threads")) *value = g_gp_tuning.threads;
  else if (!strcmp(key, "unroll")) *value = g_gp_tuning.unroll;
  else if (!strcmp(key, "ctas_per_sm")) *value = g_gp_tuning.ctas_per_sm;
  else if (!strcmp(key, "persistent")) *value = g_gp_tuning.persistent;
  else if (!strcmp(key, "bn_threads")) *value = g_gp_tuning.bn_threads;
  else if (!strcmp(key, "pipeline")) *value = g_gp_tuning.pipeline;
  else if (!strcmp(key, "bn_ctas_per_sm")) *value = g_gp_tuning.bn_ctas_per_sm;
  else if (!strcmp(key, "peer_timeout_s")) *value = (int)(g_gp_peer_timeout_ns / 1000000000ull);
  else if (!strcmp(key, "bulk")) *value = gpb::g_bulk_tuning.enable;
  else if (!strcmp(key, "bulk_tile")) *value = gpb::g_bulk_tuning.tile;
  else if (!strcmp(key, "bulk_stages")) *value = gpb::g_bulk_tuning.stages;
  else if (!strcmp(key, "bulk_ctas")) *value = gpb::g_bulk_tuning.ctas;
  else {
    gp_set_error("gp_get_tuning: unknown key '%s'", key);
    return GP_EINVAL;
  }
  return 0;
}

// ------------------------------------------------------------ NVTX ---------
// Named ranges around the stages of a step (pack / allreduce / update, BN statistics) for
// Nsight Systems / ncu --nvtx; header-only NVTX v3: no link dependency, a no-op unless a
// tool is attached.  Reference analogue: chainer/function_hooks/cuda_profile.py:14-24.
int gp_nvtx_push(const char* name) { return nvtxRangePushA(name ? name : "gradpath"); }
int gp_nvtx_pop(void) { return nvtxRangePop(); }

// ------------------------------------------------------------ device ------
int gp_device_count(int* count) { GP_CUDA(cudaGetDeviceCount(count)); return 0; }
int gp_set_device(int device) { GP_CUDA(cudaSetDevice(device)); return 0; }
int gp_get_device(int* device) { GP_CUDA(cudaGetDevice(device)); return 0; }
int gp_device_synchronize(void) { GP_CUDA(cudaDeviceSynchronize()); return 0; }
int gp_device_sm_count(int* count) {
  int dev = 0;
  GP_CUDA(cudaGetDevice(&dev));
  GP_CUDA(cudaDeviceGetAttribute(count, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}
int gp_malloc(void** ptr, size_t nbytes) {
  *ptr = nullptr;
  if (nbytes == 0) return 0;
  GP_CUDA(cudaMalloc(ptr, nbytes));
  return 0;
}
int gp_free(void* ptr) {
  if (ptr) GP_CUDA(cudaFree(ptr));
  return 0;
}
int gp_malloc_host(void** ptr, size_t nbytes) {
  *ptr = nullptr;
  if (nbytes == 0) return 0;
  GP_CUDA(cudaHostAlloc(ptr, nbytes, cudaHostAllocDefault));
  return 0;
}
int gp_free_host(void* ptr) {
  if (ptr) GP_CUDA(cudaFreeHost(ptr));
  return 0;
}
int gp_memcpy_async(void* dst, const void* src, size_t nbytes, int kind, void* stream) {
  if (nbytes == 0) return 0;
  cudaMemcpyKind k;
  switch (kind) {
    case 0: k = cudaMemcpyHostToDevice; break;
    case 1: k = cudaMemcpyDeviceToHost; break;
    case 2: k = cudaMemcpyDeviceToDevice; break;
    default:
      gp_set_error("gp_memcpy_async: bad kind %d", kind);
      return GP_EINVAL;
  }
  GP_CUDA(cudaMemcpyAsync(dst, src, nbytes, k, (cudaStream_t)stream));
  return 0;
}
int gp_memset_async(void* dst, int value, size_t nbytes, void* stream) {
  if (nbytes == 0) return 0;
  GP_CUDA(cudaMemsetAsync(dst, value, nbytes, (cudaStream_t)stream));
  return 0;
}
int gp_stream_create(void** stream, int non_blocking) {
  cudaStream_t s;
  GP_CUDA(cudaStreamCreateWithFlags(&s, non_blocking ? cudaStreamNonBlocking : cudaStreamDefault));
  *stream = (void*)s;
  return 0;
}
int gp_stream_destroy(void* stream) {
  if (stream) GP_CUDA(cudaStreamDestroy((cudaStream_t)stream));
  return 0;
}
int gp_stream_synchronize(void* stream) {
  GP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
int gp_stream_wait_event(void* stream, void* event) {
  GP_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
  return 0;
}
int gp_event_create(void** event, int enable_timing) {
  cudaEvent_t e;
  GP_CUDA(cudaEventCreateWithFlags(&e, enable_timing ? cudaEventDefault : cudaEventDisableTiming));
  *event = (void*)e;
  return 0;
}
int gp_event_destroy(void* event) {
  if (event) GP_CUDA(cudaEventDestroy((cudaEvent_t)event));
  return 0;
}
int gp_event_record(void* event, void* stream) {
  GP_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
  return 0;
}
int gp_event_synchronize(void* event) {
  GP_CUDA(cudaEventSynchronize((cudaEvent_t)event));
  return 0;
}
int gp_event_elapsed_ms(float* ms, void* start, void* stop) {
  GP_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return 0;
}

// ------------------------------------------------------------ table -------
// A ring of kSlots (pinned staging, device copy, event) triples.  An upload
// memcpy's the host table into the next pinned slot, enqueues one async H2D
// copy on the caller's stream and records the slot's event; the slot is reused
// only after that event has completed, so the host never blocks in steady
// state and the gradient arrays may move every step (chainer reallocates
// param.grad after cleargrads(): chainer/link.py:574).
struct GpTable {
  static constexpr int kSlots = 4;
  void* pinned[kSlots];
  void* dev[kSlots];
  cudaEvent_t ev[kSlots];
  bool used[kSlots];
  size_t cap[kSlots];
  int next;
};

int gp_table_create(void** table) {
  GpTable* t = new GpTable();
  memset(t, 0, sizeof(*t));
  for (int i = 0; i < GpTable::kSlots; ++i) {
    cudaError_t e = cudaEventCreateWithFlags(&t->ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
      delete t;
      return gp_cuda_fail(e, "gp_table_create");
    }
  }
  *table = t;
  return 0;
}

int gp_table_destroy(void* table) {
  GpTable* t = (GpTable*)table;
  if (!t) return 0;
  for (int i = 0; i < GpTable::kSlots; ++i) {
    if (t->used[i]) cudaEventSynchronize(t->ev[i]);
    if (t->pinned[i]) cudaFreeHost(t->pinned[i]);
    if (t->dev[i]) cudaFree(t->dev[i]);
    cudaEventDestroy(t->ev[i]);
  }
  delete t;
  return 0;
}

int gp_table_upload(void* table, const void* host_src, size_t nbytes, void* stream,
                    void** device_ptr) {
  GpTable* t = (GpTable*)table;
  if (!t || !host_src || !device_ptr) {
    gp_set_error("gp_table_upload: null argument");
    return GP_EINVAL;
  }
  const int i = t->next;
  t->next = (t->next + 1) % GpTable::kSlots;
  if (t->used[i]) GP_CUDA(cudaEventSynchronize(t->ev[i]));
  if (t->cap[i] < nbytes) {
    size_t cap = 4096;
    while (cap < nbytes) cap <<= 1;
    if (t->pinned[i]) GP_CUDA(cudaFreeHost(t->pinned[i]));
    if (t->dev[i]) GP_CUDA(cudaFree(t->dev[i]));
    t->pinned[i] = t->dev[i] = nullptr;
    t->cap[i] = 0;
    GP_CUDA(cudaHostAlloc(&t->pinned[i], cap, cudaHostAllocDefault));
    GP_CUDA(cudaMalloc(&t->dev[i], cap));
    t->cap[i] = cap;
  }
  memcpy(t->pinned[i], host_src, nbytes);
  GP_CUDA(cudaMemcpyAsync(t->dev[i], t->pinned[i], nbytes, cudaMemcpyHostToDevice,
                          (cudaStream_t)stream));
  GP_CUDA(cudaEventRecord(t->ev[i], (cudaStream_t)stream));
  t->used[i] = true;
  *device_ptr = t->dev[i];
  return 0;
}

}  // extern "C"
