// gp_nccl.cu -- NCCL entry points of libgradpath, resolved from libnccl.so.2
// with dlopen so that the library itself loads on machines without NCCL.
//
// Reference being replaced: `cupy.cuda.nccl` as re-exported by
// chainermn/nccl.py:1-14 and used by
//   chainermn/communicators/_communication_utility.py:69-76 (init_nccl_comm:
//       get_unique_id on rank 0, bcast over MPI, NcclCommunicator(size, id, rank))
//   chainermn/communicators/pure_nccl_communicator.py:94-97 (bcast), :180-182
//       (allReduce(sendbuf, recvbuf, n_elems, type_id, NCCL_SUM, stream.ptr))
// The dtype ids of this ABI ARE NCCL's ncclDataType_t values (6 = half, 7 =
// float, 8 = double, 9 = bfloat16), as in the reference.
#include <dlfcn.h>
#include <string.h>

#include "gp_common.cuh"

namespace {

typedef struct { char internal[GP_NCCL_UNIQUE_ID_BYTES]; } nccl_uid_t;
typedef void* nccl_comm_t;

struct NcclApi {
  void* handle;
  int (*GetVersion)(int*);
  int (*GetUniqueId)(nccl_uid_t*);
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int);
  int (*CommDestroy)(nccl_comm_t);
  const char* (*GetErrorString)(int);
  const char* (*GetLastError)(nccl_comm_t);
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
  int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
  int (*Reduce)(const void*, void*, size_t, int, int, int, nccl_comm_t, cudaStream_t);
  int (*GroupStart)(void);
  int (*GroupEnd)(void);
  int (*MemAlloc)(void**, size_t);
  int (*MemFree)(void*);
  int (*CommRegister)(nccl_comm_t, void*, size_t, void**);
  int (*CommDeregister)(nccl_comm_t, void*);
  int (*CommWindowRegister)(nccl_comm_t, void*, size_t, void**, int);
  int (*CommWindowDeregister)(nccl_comm_t, void*);
};
NcclApi g_nccl = {};

int nccl_fail(int r, const char* what) {
  if (r == 0) return 0;
  const char* s = g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?";
  const char* d = g_nccl.GetLastError ? g_nccl.GetLastError(nullptr) : "";
  gp_set_error("NCCL error %d (%s) in %s: %s", r, s, what, d ? d : "");
  return -(2000 + r);
}

int need(const void* fn, const char* name) {
  if (!g_nccl.handle) {
    gp_set_error("NCCL is not loaded: call gp_nccl_load first (%s)", name);
    return GP_ENOSYS;
  }
  if (!fn) {
    gp_set_error("libnccl does not export %s", name);
    return GP_ENOSYS;
  }
  return 0;
}
#define GP_NEED(field, sym)                                     \
  do {                                                          \
    int _r = need((const void*)g_nccl.field, sym);              \
    if (_r) return _r;                                          \
  } while (0)

}  // namespace

extern "C" {

int gp_nccl_load(const char* path) {
  if (g_nccl.handle) return 0;
  const char* p = (path && path[0]) ? path : "libnccl.so.2";
  void* h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    gp_set_error("gp_nccl_load: dlopen(%s) failed: %s", p, dlerror());
    return GP_ENOSYS;
  }
  NcclApi a = {};
  a.handle = h;
#define GP_SYM(field, sym) *(void**)(&a.field) = dlsym(h, sym)
  GP_SYM(GetVersion, "ncclGetVersion");
  GP_SYM(GetUniqueId, "ncclGetUniqueId");
  GP_SYM(CommInitRank, "ncclCommInitRank");
  GP_SYM(CommDestroy, "ncclCommDestroy");
  GP_SYM(GetErrorString, "ncclGetErrorString");
  GP_SYM(GetLastError, "ncclGetLastError");
  GP_SYM(AllReduce, "ncclAllReduce");
  GP_SYM(Broadcast, "ncclBroadcast");
  GP_SYM(Reduce, "ncclReduce");
  GP_SYM(GroupStart, "ncclGroupStart");
  GP_SYM(GroupEnd, "ncclGroupEnd");
  GP_SYM(MemAlloc, "ncclMemAlloc");
  GP_SYM(MemFree, "ncclMemFree");
  GP_SYM(CommRegister, "ncclCommRegister");
  GP_SYM(CommDeregister, "ncclCommDeregister");
  GP_SYM(CommWindowRegister, "ncclCommWindowRegister");
  GP_SYM(CommWindowDeregister, "ncclCommWindowDeregister");
#undef GP_SYM
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy) {
    gp_set_error("gp_nccl_load: %s lacks the NCCL 2 core symbols", p);
    dlclose(h);
    return GP_ENOSYS;
  }
  g_nccl = a;
  return 0;
}

int gp_nccl_version(int* version) {
  GP_NEED(GetVersion, "ncclGetVersion");
  return nccl_fail(g_nccl.GetVersion(version), "ncclGetVersion");
}

int gp_nccl_get_unique_id(char* id128) {
  GP_NEED(GetUniqueId, "ncclGetUniqueId");
  nccl_uid_t id;
  int r = g_nccl.GetUniqueId(&id);
  if (r) return nccl_fail(r, "ncclGetUniqueId");
  memcpy(id128, id.internal, GP_NCCL_UNIQUE_ID_BYTES);
  return 0;
}

int gp_nccl_comm_init_rank(void** comm, int n_ranks, const char* id128, int rank) {
  GP_NEED(CommInitRank, "ncclCommInitRank");
  nccl_uid_t id;
  memcpy(id.internal, id128, GP_NCCL_UNIQUE_ID_BYTES);
  nccl_comm_t c = nullptr;
  int r = g_nccl.CommInitRank(&c, n_ranks, id, rank);
  if (r) return nccl_fail(r, "ncclCommInitRank");
  *comm = c;
  return 0;
}

int gp_nccl_comm_destroy(void* comm) {
  GP_NEED(CommDestroy, "ncclCommDestroy");
  if (!comm) return 0;
  return nccl_fail(g_nccl.CommDestroy((nccl_comm_t)comm), "ncclCommDestroy");
}

static int check_dtype(int dtype, const char* what) {
  if (dtype < GP_F16 || dtype > GP_BF16) {
    gp_set_error("%s: unsupported dtype id %d", what, dtype);
    return GP_EINVAL;
  }
  return 0;
}

int gp_nccl_allreduce(void* comm, const void* sendbuf, void* recvbuf, int64_t count, int dtype,
                      int op, void* stream) {
  GP_NEED(AllReduce, "ncclAllReduce");
  if (int r = check_dtype(dtype, "gp_nccl_allreduce")) return r;
  if (count <= 0) return 0;
  return nccl_fail(g_nccl.AllReduce(sendbuf, recvbuf, (size_t)count, dtype, op, (nccl_comm_t)comm,
                                    (cudaStream_t)stream),
                   "ncclAllReduce");
}

int gp_nccl_bcast(void* comm, void* buffer, int64_t count, int dtype, int root, void* stream) {
  GP_NEED(Broadcast, "ncclBroadcast");
  if (int r = check_dtype(dtype, "gp_nccl_bcast")) return r;
  if (count <= 0) return 0;
  return nccl_fail(g_nccl.Broadcast(buffer, buffer, (size_t)count, dtype, root, (nccl_comm_t)comm,
                                    (cudaStream_t)stream),
                   "ncclBroadcast");
}

int gp_nccl_reduce(void* comm, const void* sendbuf, void* recvbuf, int64_t count, int dtype,
                   int op, int root, void* stream) {
  GP_NEED(Reduce, "ncclReduce");
  if (int r = check_dtype(dtype, "gp_nccl_reduce")) return r;
  if (count <= 0) return 0;
  return nccl_fail(g_nccl.Reduce(sendbuf, recvbuf, (size_t)count, dtype, op, root,
                                 (nccl_comm_t)comm, (cudaStream_t)stream),
                   "ncclReduce");
}

int gp_nccl_group_start(void) {
  GP_NEED(GroupStart, "ncclGroupStart");
  return nccl_fail(g_nccl.GroupStart(), "ncclGroupStart");
}
int gp_nccl_group_end(void) {
  GP_NEED(GroupEnd, "ncclGroupEnd");
  return nccl_fail(g_nccl.GroupEnd(), "ncclGroupEnd");
}
int gp_nccl_mem_alloc(void** ptr, size_t nbytes) {
  GP_NEED(MemAlloc, "ncclMemAlloc");
  return nccl_fail(g_nccl.MemAlloc(ptr, nbytes), "ncclMemAlloc");
}
int gp_nccl_mem_free(void* ptr) {
  GP_NEED(MemFree, "ncclMemFree");
  return nccl_fail(g_nccl.MemFree(ptr), "ncclMemFree");
}
int gp_nccl_comm_register(void* comm, void* buffer, size_t nbytes, void** handle) {
  GP_NEED(CommRegister, "ncclCommRegister");
  return nccl_fail(g_nccl.CommRegister((nccl_comm_t)comm, buffer, nbytes, handle),
                   "ncclCommRegister");
}
int gp_nccl_comm_deregister(void* comm, void* handle) {
  GP_NEED(CommDeregister, "ncclCommDeregister");
  return nccl_fail(g_nccl.CommDeregister((nccl_comm_t)comm, handle), "ncclCommDeregister");
}

// NCCL >= 2.27 symmetric memory: a buffer from ncclMemAlloc registered on every
// rank as a collective-symmetric window lets ncclAllReduce use its symmetric
// (NVLS / low-latency) kernels.  flags: 1 = NCCL_WIN_COLL_SYMMETRIC.
int gp_nccl_comm_window_register(void* comm, void* buffer, size_t nbytes, void** window, int flags) {
  GP_NEED(CommWindowRegister, "ncclCommWindowRegister");
  return nccl_fail(g_nccl.CommWindowRegister((nccl_comm_t)comm, buffer, nbytes, window, flags),
                   "ncclCommWindowRegister");
}
int gp_nccl_comm_window_deregister(void* comm, void* window) {
  GP_NEED(CommWindowDeregister, "ncclCommWindowDeregister");
  return nccl_fail(g_nccl.CommWindowDeregister((nccl_comm_t)comm, window),
                   "ncclCommWindowDeregister");
}

}  // extern "C"
