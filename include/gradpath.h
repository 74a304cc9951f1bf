/*
 * gradpath.h -- C-ABI of libgradpath.so: the B200 (sm_100a) implementation of
 * ChainerMN's data-parallel gradient path.
 *
 * Every entry point replaces one piece of the reference's `pure_nccl` path; the
 * reference location is cited beside each declaration (paths are relative to the
 * chainer/chainer v7.8.1 tree).  The reference binds its kernels through CuPy's
 * NVRTC JIT (`chainer.cuda.raw` / `chainer.cuda.elementwise`) and NCCL through
 * `cupy.cuda.nccl`; this library is what a maintainer binds instead, with ctypes
 * (see INTEGRATION.md).
 *
 * Conventions
 *   - plain C types only: pointers, sizes, doubles.  No torch / cupy types.
 *   - every function returns 0 on success, a negative code on failure:
 *       -(1000 + cudaError_t)  for CUDA runtime failures
 *       -(2000 + ncclResult_t) for NCCL failures
 *       GP_EINVAL / GP_ENOSYS  for bad arguments / missing NCCL symbols
 *     gp_last_error() returns a thread-local human readable message.
 *   - the caller owns every data buffer.  The library owns only the handles it
 *     creates (gp_*_create) and frees them in gp_*_destroy.
 *   - all kernels are asynchronous on the `stream` argument (a cudaStream_t
 *     passed as void*; NULL is the legacy default stream, which is what
 *     `chainer.cuda.Stream.null.ptr` is in the reference).
 *   - dtype ids are NCCL's: 6 = float16, 7 = float32, 8 = float64, 9 = bfloat16
 *     (reference: chainermn/nccl.py:1-14, _communication_utility.py:177-186;
 *      bfloat16 is an extension with no reference counterpart).
 */
#ifndef GRADPATH_H_
#define GRADPATH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP_ABI_VERSION 1

#define GP_F16 6
#define GP_F32 7
#define GP_F64 8
#define GP_BF16 9

#define GP_EINVAL (-22)
#define GP_ENOSYS (-38)

/* flags of gp_seg_t.flags */
#define GP_SEG_VEC_OK 1u /* all pointers 4-element aligned and csum/buf_off % 4 == 0 */

/* flags of gp_unpack_adam */
#define GP_ADAM_AMSGRAD 1
#define GP_ADAM_ADABOUND 2

/*
 * One entry of the device-side segment table: one parameter of the model.
 * Replaces the three device arrays of `ParamsData`
 * (chainermn/communicators/_memory_utility.py:31-62: dptr int64[n],
 * dtype int32[n], size_csum int32[n+1]).  The cumulative sizes live in a
 * separate int64[n+1] array (`csum`) so that it can be staged in shared
 * memory compactly; int64 instead of the reference's int32 so that buffers
 * beyond 2^31 elements do not overflow.
 *
 *   ptr[0]  the array that is packed / unpacked (param.grad or param.data);
 *           in the fused update kernels: where the mean gradient is written
 *           back (may be 0 when write_grad == 0)
 *   ptr[1]  param.data                         (fused update only)
 *   ptr[2]  state 'v' (MomentumSGD) / 'm' (Adam)
 *   ptr[3]  state 'v' (Adam)
 *   ptr[4]  state 'vhat' (AMSGrad)
 *   buf_off element offset of this parameter inside the packed buffer
 *   dtype0  dtype id of ptr[0]; dtype1 dtype id of ptr[1..4]
 */
typedef struct gp_seg_t {
  uint64_t ptr[5];
  int64_t buf_off;
  int32_t dtype0;
  int32_t dtype1;
  uint32_t flags;
  uint32_t reserved;
} gp_seg_t; /* 64 bytes */

/* ---------------------------------------------------------------- errors -- */
const char* gp_last_error(void);
int gp_abi_version(void);

/* ------------------------------------------------------- device / memory -- */
/* Plumbing the reference gets from CuPy (cupy.cuda.alloc, Stream, Event,
 * MemoryPointer.copy_from_device_async: _memory_utility.py:65-151). */
int gp_device_count(int* count);
int gp_set_device(int device);
int gp_get_device(int* device);
int gp_device_synchronize(void);
int gp_device_sm_count(int* count);
int gp_malloc(void** ptr, size_t nbytes);
int gp_free(void* ptr);
int gp_malloc_host(void** ptr, size_t nbytes); /* pinned */
int gp_free_host(void* ptr);
/* kind: 0 = host->device, 1 = device->host, 2 = device->device */
int gp_memcpy_async(void* dst, const void* src, size_t nbytes, int kind, void* stream);
int gp_memset_async(void* dst, int value, size_t nbytes, void* stream);
int gp_stream_create(void** stream, int non_blocking);
int gp_stream_destroy(void* stream);
int gp_stream_synchronize(void* stream);
int gp_stream_wait_event(void* stream, void* event);
int gp_event_create(void** event, int enable_timing);
int gp_event_destroy(void* event);
int gp_event_record(void* event, void* stream);
int gp_event_synchronize(void* event);
int gp_event_elapsed_ms(float* ms, void* start, void* stop);

/* Upload of a host table (csum + segments) without a host sync.  The handle
 * owns a device arena and a ring of pinned staging slots; replaces the three
 * synchronous `cupy.asarray` calls of ParamsData (_memory_utility.py:59-61).
 * `*device_ptr` receives the device address of the uploaded copy (256-byte
 * aligned, valid until the 4th following upload on the same handle). */
int gp_table_create(void** table);
int gp_table_destroy(void* table);
int gp_table_upload(void* table, const void* host_src, size_t nbytes, void* stream,
                    void** device_ptr);

/* ------------------------------------------------------ gradient kernels -- */
/*
 * gp_pack: gather + cast (+ optional pre-scale) of n_segs arrays into the
 * contiguous packed buffer.  buffer[seg.buf_off + k] = (buf_dtype)(scale * ptr0[k]).
 * Replaces kernel `cupy_batched_pack_params` (_memory_utility.py:289-358,
 * launched at :253-268) and, with scale != 1, folds `div_by_size`
 * (pure_nccl_communicator.py:183-189) into the pack.
 *
 * d_csum[n_segs + 1] (int64, device) are the cumulative element counts in this
 * call's work space; d_segs[n_segs] the segment entries.  Only work elements in
 * [elem_begin, elem_end) are processed (bucketing); pass 0 and d_csum[n_segs]
 * for everything.
 */
int gp_pack(void* buffer, int buf_dtype, const int64_t* d_csum, const gp_seg_t* d_segs,
            int n_segs, int64_t elem_begin, int64_t elem_end, double scale, int layout_hint,
            void* stream);

/*
 * gp_unpack_scale: ptr0[k] = (dtype0)( (buf_dtype)(scale * buffer[buf_off + k]) ).
 * Replaces `div_by_size` + kernel `cupy_batched_unpack_params`
 * (pure_nccl_communicator.py:183-189, _memory_utility.py:361-429).  The
 * intermediate rounding to buf_dtype reproduces the reference, which scales the
 * receive buffer in place before unpacking it.  scale == 1.0 is a pure unpack
 * (used by bcast_data, pure_nccl_communicator.py:82-99).
 */
int gp_unpack_scale(const void* buffer, int buf_dtype, const int64_t* d_csum,
                    const gp_seg_t* d_segs, int n_segs, int64_t elem_begin, int64_t elem_end,
                    double scale, int layout_hint, void* stream);

/*
 * layout_hint (all four stream kernels): 0 = no promise; GP_F32 = the caller
 * guarantees that every segment has dtype0 == dtype1 == float32, every pointer
 * is 16-byte aligned and every csum / buf_off value is a multiple of 8 elements
 * (4 suffices with a 4-byte buffer dtype).  The library then runs its
 * float32-only kernels: the software-pipelined walker (two tiles in registers
 * per warp, no dtype dispatch) or, when enabled by gp_set_tuning("bulk", 1),
 * the TMA-staged kernel of gp_bulk.cuh.  Results are identical either way.
 *
 * buffer == NULL (fused update kernels): there is no packed buffer; the gradient
 * of element k is read from ptr0[k] itself (every dtype0 must equal buf_dtype).
 * This is `optimizer.update()` without a communicator as ONE multi-tensor launch
 * (chainer/optimizer.py:857-894: one kernel per parameter in the reference).
 *
 * gp_unpack_momentum_sgd: fused unpack + descale + MomentumSGD update.
 *   g = (dtype0)((buf_dtype)(scale * buffer[buf_off + k]))
 *   v = momentum * v - lr * g ;  param += v        (arithmetic in dtype1)
 *   if write_grad: ptr0[k] = g
 * Replaces unpack (above) + one `momentum_sgd` ElementwiseKernel launch per
 * parameter (chainer/optimizers/momentum_sgd.py:75-88).  Each gradient element
 * is read from HBM exactly once.  No fused multiply-add contraction is used:
 * the result is bit-identical to update_core_cpu (momentum_sgd.py:61-73).
 */
int gp_unpack_momentum_sgd(const void* buffer, int buf_dtype, const int64_t* d_csum,
                           const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                           int64_t elem_end, double scale, double lr, double momentum,
                           int write_grad, int layout_hint, void* stream);

/*
 * gp_unpack_adam: fused unpack + descale + Adam / AdamW / AMSGrad / AdaBound.
 *   m += (1-beta1)(g-m); v += (1-beta2)(g*g-v); [vhat = max(vhat, v)]
 *   param -= eta * (alpha_t * m / (sqrt(v) + eps) + weight_decay_rate * param)
 * with the exact operation order of the `adam` / `amsgrad` / `adabound` /
 * `amsbound` ElementwiseKernels (chainer/optimizers/adam.py:237-332);
 * intermediate type float for float16/float32 parameters, double for float64
 * (adam.py:57-63).  alpha_t (adam.py:47-54) and the AdaBound bounds
 * (adam.py:346-358) are computed by the caller in double.
 */
int gp_unpack_adam(const void* buffer, int buf_dtype, const int64_t* d_csum,
                   const gp_seg_t* d_segs, int n_segs, int64_t elem_begin, int64_t elem_end,
                   double scale, double alpha_t, double one_minus_beta1,
                   double one_minus_beta2, double eps, double eta, double weight_decay_rate,
                   double lower, double upper, int adam_flags, int write_grad, int layout_hint,
                   void* stream);

/* gp_scale: buffer[k] = (dtype)(buffer[k] * scale), in place.  Replaces the
 * `div_by_size` ElementwiseKernel (pure_nccl_communicator.py:183-189) where it
 * is used stand-alone (MNBN statistics, functions/batch_normalization.py:57-60). */
int gp_scale(void* buffer, int dtype, int64_t n_elems, double scale, void* stream);

/* gp_check_finite: *d_flag |= 1 if any element is NaN/Inf.  Replaces
 * `_ensure_all_finite` (mpi_communicator_base.py:730-733) in debug mode. */
int gp_check_finite(const void* buffer, int dtype, int64_t n_elems, int32_t* d_flag,
                    void* stream);

/* ------------------------------------------- batch-normalisation statistics -- */
/*
 * gp_bn_fwd_stats: one pass over x[N, C, HW] (C-contiguous, x_dtype) producing
 * out[0:C] = mean over (N, HW), out[C:2C] = mean of squares, in out_dtype.
 * Replaces `x.mean(axis)`, `xp.square(x).mean(axis)` of
 * _NcclImpl.get_mean_and_var (chainermn/functions/batch_normalization.py:53-56),
 * which reads x twice and materialises square(x).
 * workspace: device scratch of gp_bn_workspace_bytes(C) bytes, zero-initialised
 * once by the caller.  One workspace may serve layers of DIFFERENT C (size it for the
 * largest), one launch at a time: the words the kernels need zero between launches (the
 * "channels done" counter and the channel tickets) form a header at C-independent offsets
 * and are left zeroed; the scratch behind it (split partials, staged local statistics) may
 * hold junk.  gp_bn_workspace_layout reports both byte ranges: out4 = {header begin, header
 * end, scratch begin, scratch end for this C}.
 */
size_t gp_bn_workspace_bytes(int64_t C);
int gp_bn_workspace_layout(int64_t C, int64_t* out4);
int gp_bn_fwd_stats(const void* x, int x_dtype, int64_t N, int64_t C, int64_t HW, void* out,
                    int out_dtype, void* workspace, void* stream);
/* Same pass, single rank: out[0:C] = mean, out[C:2C] = var = sqmean - mean^2
 * (no allreduce follows, so the finish step is folded into the kernel). */
int gp_bn_fwd_mean_var(const void* x, int x_dtype, int64_t N, int64_t C, int64_t HW, void* out,
                       int out_dtype, void* workspace, void* stream);
/*
 * gp_bn_bwd_stats: out[0:C] = sum(gy), out[C:2C] = sum(gy * x_hat) over (N, HW).
 * Replaces `gy.sum(axis)`, `(gy * x_hat).sum(axis)` of
 * _NcclImpl.get_ggamma_and_gbeta (functions/batch_normalization.py:79-82).
 * If mean/inv_std are non-NULL, `xhat_or_x` holds x and x_hat is formed on the
 * fly as (x - mean[c]) * inv_std[c] (chainer/functions/normalization/
 * batch_normalization.py:107-108, `_x_hat`).
 */
int gp_bn_bwd_stats(const void* gy, int gy_dtype, const void* xhat_or_x, int x_dtype,
                    const void* mean, const void* inv_std, int stat_dtype, int64_t N, int64_t C,
                    int64_t HW, void* out, int out_dtype, void* workspace, void* stream);
/* var = sqmean - mean^2 in place on a [mean | sqmean] buffer after the
 * allreduce (functions/batch_normalization.py:65-67), fused with the 1/size
 * scale: buf[k] *= scale first.  out_var[C] may alias buf + C. */
int gp_bn_finish_mean_var(void* buf, int dtype, int64_t C, double scale, void* out_var,
                          void* stream);
/* Statistics + exchange in ONE kernel for N ranks on one NVSwitch box: the CTA that
 * completes the last channel runs the one-shot peer-memory allreduce of the 2C float32
 * values (allReduce + div_by_size [+ var = sqmean - mean^2]) inside the same launch.
 * Replaces the whole of _NcclImpl.get_mean_and_var / get_ggamma_and_gbeta
 * (chainermn/functions/batch_normalization.py:44-68, 70-93).  `p2p_comm`: a gp_p2p_create
 * handle whose small-message areas are set (gp_p2p_set_small), 2C <= their capacity;
 * `out`: float32 [2C] = [mean | var] over all ranks (forward) / the mean over ranks of
 * [sum gy | sum gy * x_hat] (backward).  Collective: every rank launches it. */
int gp_bn_fwd_stats_allreduce(void* p2p_comm, const void* x, int x_dtype, int64_t N, int64_t C,
                              int64_t HW, void* out_mean_var, void* workspace, void* stream);
int gp_bn_bwd_stats_allreduce(void* p2p_comm, const void* gy, int gy_dtype, const void* xhat_or_x,
                              int x_dtype, const void* mean, const void* inv_std, int stat_dtype,
                              int64_t N, int64_t C, int64_t HW, void* out, void* workspace,
                              void* stream);
/* The elementwise halves of batch normalisation (csrc/gp_bn_apply.cu), one launch each.
 * Forward (chainer/functions/normalization/batch_normalization.py:40-77, 864-867):
 *   inv_std = rsqrt(var + eps);  y = gamma * (x - mean) * inv_std + beta;
 *   r_mean = r_mean * decay + mean * (1 - decay);
 *   r_var  = r_var  * decay + var  * (1 - decay) * adjust        (running_* may be NULL)
 * replacing the rsqrt, `bn_fwd` and `update_mean_var` launches; inv_std_out [C] may be NULL.
 * Backward (:105-133): gx = (gamma * inv_std) * (gy - (x_hat * ggamma + gbeta) * inv_m)
 * with x_hat = (x - mean) * inv_std formed on the fly (the reference materialises it).
 * x, y, gy, gx: [N, C, HW] contiguous, one dtype; statistics / gamma / beta: [C] of
 * stat_dtype. */
int gp_bn_fwd_apply(const void* x, int x_dtype, int64_t N, int64_t C, int64_t HW, const void* mean,
                    const void* var, const void* gamma, const void* beta, int stat_dtype,
                    double eps, void* y, void* inv_std_out, void* running_mean, void* running_var,
                    int running_dtype, double decay, double adjust, void* stream);
int gp_bn_bwd_apply(const void* gy, int gy_dtype, const void* x, int x_dtype, int64_t N, int64_t C,
                    int64_t HW, const void* mean, const void* inv_std, const void* gamma,
                    const void* ggamma, const void* gbeta, int stat_dtype, double inv_m, void* gx,
                    void* stream);

/* ------------------------------------------------------------------ NCCL -- */
/* Thin wrappers over libnccl.so.2, resolved with dlopen at gp_nccl_load time;
 * replaces `cupy.cuda.nccl` (chainermn/nccl.py:1-14) and
 * `init_nccl_comm` (_communication_utility.py:69-76). */
#define GP_NCCL_UNIQUE_ID_BYTES 128
#define GP_NCCL_SUM 0
int gp_nccl_load(const char* libnccl_path); /* NULL: "libnccl.so.2" from the loader path */
int gp_nccl_version(int* version);
int gp_nccl_get_unique_id(char* id128);
int gp_nccl_comm_init_rank(void** comm, int n_ranks, const char* id128, int rank);
int gp_nccl_comm_destroy(void* comm);
int gp_nccl_allreduce(void* comm, const void* sendbuf, void* recvbuf, int64_t count, int dtype,
                      int op, void* stream);
int gp_nccl_bcast(void* comm, void* buffer, int64_t count, int dtype, int root, void* stream);
int gp_nccl_reduce(void* comm, const void* sendbuf, void* recvbuf, int64_t count, int dtype,
                   int op, int root, void* stream);
int gp_nccl_group_start(void);
int gp_nccl_group_end(void);
/* ncclMemAlloc / ncclCommRegister: user-buffer registration makes the packed
 * buffer eligible for NVLS (in-switch reduction) on NVSwitch systems. */
int gp_nccl_mem_alloc(void** ptr, size_t nbytes);
int gp_nccl_mem_free(void* ptr);
int gp_nccl_comm_register(void* comm, void* buffer, size_t nbytes, void** handle);
int gp_nccl_comm_deregister(void* comm, void* handle);
/* NCCL >= 2.27 symmetric windows (flags 1 = NCCL_WIN_COLL_SYMMETRIC; collective call) */
int gp_nccl_comm_window_register(void* comm, void* buffer, size_t nbytes, void** window, int flags);
int gp_nccl_comm_window_deregister(void* comm, void* window);

/* ------------------------------------------- peer-memory allreduce (NVLink) -- */
/*
 * Allreduce of the packed buffer written as ONE kernel per rank over NVLink
 * peer memory (csrc/gp_p2p.cu): rank r loads the N copies of its 1/N shard (its
 * own from HBM, N-1 through NVLink), adds them in rank order and stores the sum
 * into all N buffers; the "everyone has packed" / "everyone has stored" barriers
 * are flags in peer memory inside the kernel.  Replaces
 * `nccl_comm.allReduce(...)` (pure_nccl_communicator.py:180-182) when all ranks
 * share one NVSwitch box (n_ranks 2, 4 or 8).  Buffers and flag blocks are
 * cudaMalloc allocations exchanged with gp_ipc_* over the control plane.
 * Sums are formed in rank order with every partial sum rounded to the buffer
 * dtype, so the result is deterministic and identical on all ranks.
 */
#define GP_IPC_HANDLE_BYTES 64
int gp_ipc_get_handle(void* device_ptr, char* handle64);
int gp_ipc_open_handle(const char* handle64, void** device_ptr);
int gp_ipc_close_handle(void* device_ptr);
size_t gp_p2p_flag_bytes(void); /* size of one rank's (zero-initialised) flag block */
/* buffers[k] / flags[k]: this process's mapping of rank k's packed buffer / flag block */
int gp_p2p_create(void** comm, int rank, int n_ranks, void* const* buffers, void* const* flags);
int gp_p2p_set_buffers(void* comm, void* const* buffers);
int gp_p2p_destroy(void* comm);
/* in-place sum over ranks of elements [offset, offset + n_elems) of the buffers */
int gp_p2p_allreduce(void* comm, int dtype, int64_t offset_elems, int64_t n_elems, void* stream);
int gp_p2p_set_tuning(int ctas, int threads, int mode);
/*
 * One-shot allreduce for small float32 messages (MNBN's 2C statistics,
 * chainermn/functions/batch_normalization.py:57-60, 83-86): allReduce +
 * div_by_size (+ `var = sqmean - mean^2` when C > 0, :65-67) in ONE single-CTA
 * kernel over peer memory.  out[i] = scale * sum_ranks in[i] (rank order);
 * recv_areas[k] / flag_blocks[k]: mappings of rank k's zero-initialised receive
 * area (gp_p2p_small_bytes) and flag block (gp_p2p_flag_bytes).
 */
size_t gp_p2p_small_bytes(int n_ranks, int64_t capacity_elems);
int gp_p2p_set_small(void* comm, void* const* recv_areas, void* const* flag_blocks,
                     int64_t capacity_elems);
int gp_p2p_allreduce_small(void* comm, const void* in, void* out, int64_t n_elems, int64_t C,
                           double scale, void* stream);

/* ---------------------------------------------------------------- tuning -- */
/* ---------------------------------------------------------------------------
 * Optimizer hooks and loss scaling fused into the update.
 *
 * GradientMethod.update (chainer/optimizer.py:857-894) runs the optimizer-level
 * hooks over every parameter before the per-parameter updates, and UpdateRule.update
 * (:286-291) divides by the loss scale.  For the hook lists [GradientClipping],
 * [WeightDecay] and [GradientClipping, WeightDecay] (registration order) the
 * `_hooked` update kernels apply, per element, between the mean and update_core:
 *     g *= *clip_rate          optimizer_hooks/gradient_clipping.py:84-106
 *     g += weight_decay * p    optimizer_hooks/weight_decay.py:44-57 (pass rate * loss_scale)
 *     g /= loss_scale          optimizer.py:289-291
 * each operation rounded to the parameter's dtype; with write_grad, param.grad
 * receives the transformed gradient, as the in-place hooks leave it.
 */
typedef struct gp_hooks_t {
  const float* clip_rate;   /* DEVICE pointer: &((float*)out)[2] of gp_sqnorm; NULL: no clipping */
  double weight_decay;      /* 0: off */
  double loss_scale;        /* 0: off */
} gp_hooks_t;

int gp_unpack_momentum_sgd_hooked(const void* buffer, int buf_dtype, const int64_t* d_csum,
                                  const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                                  int64_t elem_end, double scale, double lr, double momentum,
                                  int write_grad, int layout_hint, const gp_hooks_t* hooks,
                                  void* stream);
int gp_unpack_adam_hooked(const void* buffer, int buf_dtype, const int64_t* d_csum,
                          const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                          int64_t elem_end, double scale, double alpha_t, double one_minus_beta1,
                          double one_minus_beta2, double eps, double eta,
                          double weight_decay_rate, double lower, double upper, int adam_flags,
                          int write_grad, int layout_hint, const gp_hooks_t* hooks, void* stream);

/* float16 parameters with FLOAT32 MASTER WEIGHTS (csrc/gp_master.cu): ONE launch replaces, per
 * parameter, `fp32_param.grad = param.grad.astype(float32)`, `grad /= loss_scale`,
 * update_core_gpu on the float32 copy and `param.array = fp32_param.array.astype(float16)`
 * (chainer/optimizer.py:262-305), with the optimizer-level hooks (acting on the float16
 * arrays, as the reference's do) and the dynamic-loss-scaling skip fused in.
 * Tables: ptr[0] float16 gradient (with write_grad it receives the mean gradient after the
 * hooks), ptr[1] float32 master, ptr[2..3] float32 states, ptr[4] float16 parameter;
 * dtype0 = GP_F16, dtype1 = GP_F32.  hooks may be NULL.  d_skip may be NULL; else a device
 * int32 (gp_check_finite over the reduced buffer): non-zero = a non-finite gradient, nothing
 * is updated (`is_safe_to_update()`, chainer/optimizer.py:763-779, without a host round trip
 * between the allreduce and the update).  AMSGrad is not covered. */
int gp_unpack_momentum_sgd_master(const void* buffer, int buf_dtype, const int64_t* d_csum,
                                  const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                                  int64_t elem_end, double scale, double lr, double momentum,
                                  int write_grad, const gp_hooks_t* hooks, const void* d_skip,
                                  void* stream);
int gp_unpack_adam_master(const void* buffer, int buf_dtype, const int64_t* d_csum,
                          const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                          int64_t elem_end, double scale, double alpha_t, double one_minus_beta1,
                          double one_minus_beta2, double eps, double eta,
                          double weight_decay_rate, double lower, double upper, int adam_flags,
                          int write_grad, const gp_hooks_t* hooks, const void* d_skip,
                          void* stream);

/* The other first-order rules with MomentumSGD's shape (csrc/gp_sgd_family.cu), same
 * tables (ptr[0] grad, ptr[1] param, ptr[2] v), hooks optional (NULL: none):
 *   GP_RULE_SGD                 param -= lr * grad                  chainer/optimizers/sgd.py:45-63
 *   GP_RULE_CORRECTED_MOMENTUM  v = momentum*v - grad; param += lr*v
 *                                                     chainer/optimizers/corrected_momentum_sgd.py:61-89
 *   GP_RULE_NESTEROV_AG         v = momentum*v - lr*grad; param += momentum^2 * v;
 *                               param -= (1+momentum)*lr * grad     chainer/optimizers/nesterov_ag.py:60-71
 * arithmetic in the parameter dtype in the order of update_core_cpu. */
#define GP_RULE_SGD 0
#define GP_RULE_CORRECTED_MOMENTUM 1
#define GP_RULE_NESTEROV_AG 2
int gp_unpack_sgd_family(const void* buffer, int buf_dtype, const int64_t* d_csum,
                         const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                         int64_t elem_end, double scale, int rule, double lr, double momentum,
                         int write_grad, int layout_hint, const gp_hooks_t* hooks, void* stream);

/* Squared L2 norm of `scale * x[0..n)` (x: the allreduced packed buffer, or one
 * gradient array) -- _sum_sqnorm_grads, gradient_clipping.py:9-52 -- and the
 * clipping rate min(threshold / sqrt(sum), 1) (:91-101), formed on the device.
 * out: 16 bytes of device memory {double sqsum; float rate; float norm}.
 * accumulate != 0 adds to out->sqsum (per-parameter use); rate and norm always
 * describe the running total.  Deterministic for a given n (fixed grid and order);
 * double accumulation.  workspace: gp_sqnorm_workspace_bytes() of ZEROED device
 * memory, reusable across calls on one stream. */
size_t gp_sqnorm_workspace_bytes(void);
int gp_sqnorm(const void* x, int dtype, int64_t n_elems, double scale, int accumulate,
              double threshold, void* workspace, void* out, void* stream);
/* x *= *d_factor (float on the device), in x's dtype: the unfused `grad *= rate`
 * of gradient_clipping.py:103-106 */
int gp_scale_by_device(void* x, int dtype, int64_t n_elems, const void* d_factor, void* stream);
/* grad += decay * param, in the arrays' dtype: weight_decay.py:55-57, unfused */
int gp_weight_decay(void* grad, const void* param, int dtype, int64_t n_elems, double decay,
                    void* stream);
/* x /= divisor, in x's dtype: `grad /= loss_scale` of chainer/optimizer.py:289-291, unfused */
int gp_divide(void* x, int dtype, int64_t n_elems, double divisor, void* stream);

/* ---------------------------------------------------------------------------
 * NVSwitch-multicast (NVLS) allreduce of the packed buffer, 4 / 8 ranks on one box
 * (csrc/gp_mc.cu).  Replaces nccl_comm.allReduce of
 * chainermn/communicators/pure_nccl_communicator.py:180-182: rank r reduces its
 * 1/N shard with multimem.ld_reduce (the switch adds the N copies) and
 * broadcasts it with multimem.st, both barriers inside the one kernel (the flag
 * words of the gp_p2p communicator).  float32 / float16 / bfloat16 (16-bit types
 * accumulate in fp32); summation order is the switch's: parity to rounding.
 *
 * Set-up is collective and staged (the host code synchronises between stages):
 *   gp_mc_create on every rank (rank 0 creates the multicast object)
 *   rank 0: gp_mc_export_fd -> POSIX fd, passed to the peers (SCM_RIGHTS)
 *   ranks != 0: gp_mc_import_fd
 *   gp_mc_add_device on every rank;            -- barrier --
 *   gp_mc_bind on every rank (cuMemCreate + cuMulticastBindMem + 2 mappings)
 *                                              -- barrier --
 *   gp_mc_pointers: the unicast address IS the packed buffer (gpu_buffer_a).
 */
int gp_mc_supported(int* supported);
int gp_mc_create(void** mc, int rank, int n_ranks, size_t nbytes);
int gp_mc_export_fd(void* mc, int* fd);
int gp_mc_import_fd(void* mc, int fd);
int gp_mc_add_device(void* mc);
int gp_mc_bind(void* mc);
int gp_mc_pointers(void* mc, void** unicast, void** multicast, size_t* nbytes);
int gp_mc_destroy(void* mc);
int gp_mc_allreduce(void* p2p_comm, void* mc, int dtype, int64_t offset_elems, int64_t n_elems,
                    void* stream);
int gp_mc_set_tuning(int ctas, int threads, int unroll);

/* ---------------------------------------------------------------- one-launch step
 * The WHOLE per-step path of `_MultiNodeOptimizer.update`
 * (chainermn/optimizers.py:17-33) as one kernel launch per rank: pack
 * (_memory_utility.py:253-268, 289-358) -> sum over ranks
 * (pure_nccl_communicator.py:180-182) -> 1/N (:183-189) -> unpack
 * (_memory_utility.py:271-286, 361-429) -> update_core_gpu
 * (chainer/optimizers/momentum_sgd.py:75-88, adam.py:224-332), flowing tile by tile
 * through the packed buffer (csrc/gp_step.cu).  Results are identical to gp_pack +
 * gp_p2p_allreduce / gp_mc_allreduce + gp_unpack_momentum_sgd / gp_unpack_adam.
 *
 *   p2p_comm  NULL: one rank (no exchange); else the gp_p2p_create handle of 2/4/8 ranks
 *             whose per-tile words were set with gp_p2p_set_step_words
 *   mc_ptr    NULL: peer-memory transport (the buffers of gp_p2p_set_buffers; sums in
 *             rank order, bit-exact on every rank); else the MULTICAST address of
 *             `buffer` (gp_mc_pointers): the NVSwitch adds the copies
 *   buffer    this rank's packed buffer (unicast address), n_elems elements of
 *             buf_dtype (float32 / float16 / bfloat16), capacity rounded up to 16 bytes
 *   tables    as gp_unpack_momentum_sgd / gp_unpack_adam (ptr[0] = gradient: packed
 *             from, and with write_grad the mean written back to); layout_hint must
 *             be GP_F32 (all arrays float32); scale = 1/size
 * gp_step_supported tells whether a configuration is covered (else use the separate
 * launches).  Collective at N > 1: every rank launches the same step.
 */
int gp_step_supported(int n_ranks, int buf_dtype, int layout_hint, double scale, int adam_flags);
int gp_step_momentum_sgd(void* p2p_comm, void* mc_ptr, void* buffer, int buf_dtype,
                         const int64_t* d_csum, const gp_seg_t* d_segs, int n_segs, int64_t n_elems,
                         double scale, double lr, double momentum, int write_grad, int layout_hint,
                         void* stream);
int gp_step_adam(void* p2p_comm, void* mc_ptr, void* buffer, int buf_dtype, const int64_t* d_csum,
                 const gp_seg_t* d_segs, int n_segs, int64_t n_elems, double scale, double alpha_t,
                 double one_minus_beta1, double one_minus_beta2, double eps, double eta,
                 double weight_decay_rate, double lower, double upper, int adam_flags,
                 int write_grad, int layout_hint, void* stream);
/* per-tile words of the N-rank step: [tile_cap x 8 "packed by rank r" words | tile_cap
 * "reduced" flags] (uint32, each holding the epoch of the step that wrote it) per rank,
 * zeroed, shared through gp_ipc_*; tile_elems (a multiple of 4096, gp_step_tile_elems()
 * is the tuned default) is fixed for the life of the words; the element count may
 * change from step to step */
size_t gp_step_words_bytes(int64_t tile_cap);
int gp_step_tile_elems(void);
int gp_p2p_set_step_words(void* p2p_comm, void* const* blocks, int64_t tile_cap, int64_t tile_elems);
/* keys: tile_elems, reducers, unroll, ctas_per_sm (N-rank step; the one-rank step is a walker
 * launch and follows gp_set_tuning) */
int gp_step_set_tuning(const char* key, int value);

/* key: "threads", "unroll", "ctas_per_sm", "persistent".  For benchmarking
 * sweeps; defaults are the tuned values recorded in DESIGN.md. */
/* NVTX ranges (Nsight Systems / `ncu --nvtx`): push returns the nesting level or a negative
 * value when no tool is attached.  The Python layer brackets pack / allreduce / update and
 * the BN statistics with them when CHAINER_B200_NVTX=1
 * (reference analogue: chainer/function_hooks/cuda_profile.py:14-24). */
int gp_nvtx_push(const char* name);
int gp_nvtx_pop(void);

int gp_set_tuning(const char* key, int value);
int gp_get_tuning(const char* key, int* value);

#ifdef __cplusplus
}
#endif
#endif /* GRADPATH_H_ */
