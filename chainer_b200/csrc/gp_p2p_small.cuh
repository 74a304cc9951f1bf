// gp_p2p_small.cuh -- the one-shot allreduce of SMALL float32 messages over NVLink peer
// memory as a device function, shared by the stand-alone kernel (gp_p2p.cu) and the fused
// BN statistics kernels (gp_bn.cu), whose last CTA runs it right after the reduction.
// Include inside the same anonymous namespace as gp_p2p.cuh.
//
// Reference being replaced: `comm._multi_node_mean_nccl(gpu_buffer_a, gpu_buffer_b, 2C, ...)`
// of MultiNodeBatchNormalization (chainermn/functions/batch_normalization.py:57-60, 83-86:
// ncclAllReduce + div_by_size) and `var = sqmean - mean^2` (:65-67).
//
// Every rank stores its vector into slot [rank] of every peer's receive area, raises a
// flag, waits for the N flags, then adds the N slots in RANK ORDER (bit-identical on all
// ranks), applies the 1/size scale (x * (1.0/N) in double, or exactly in float for 2^-k: one
// rounding) and -- forward statistics -- forms var = sqmean - mean^2.  Receive areas are
// double-buffered by call parity; the latency is one NVLink store round.
#pragma once

struct SmallArgs {
  float* recv[kMaxRanks];
  uint32_t* flags[kMaxRanks];
  int rank, n;
  int64_t cap;
  uint32_t* epoch_ctr;   // this rank's call counter, in DEVICE memory (see small_exchange)
  unsigned long long timeout_ns;
  const float* in;
  float* out;
  int n_elems;
  int C;         // > 0: out[C + c] = out[C + c] - out[c]^2 after scaling (mean | var)
  float scale_f;
  double scale_d;
  int scale_mode;
};

// Called by ALL threads of one CTA.  `in` may have been written by other CTAs of the same
// grid (the caller has fenced): it is read through L2.
//
// The epoch (call number) that tags the flags lives in device memory and is advanced by the
// kernel itself, not passed as an argument: a launch is therefore REPLAYABLE -- the BN
// statistics kernels (and with them a whole forward / backward) can be captured into a CUDA
// graph once and replayed every step.  Collectives of one communicator are issued in the same
// order on one stream by every rank, so all ranks count alike.
__device__ __forceinline__ void small_exchange(const SmallArgs& a) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) {
    const uint32_t e = __ldcg(a.epoch_ctr) + 1;      // through L2: the previous launch wrote it
    __stcg(a.epoch_ctr, e);
    s_epoch = e;
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int par = epoch & 1;
  // 1. my contribution into slot [rank] of every rank (own included)
  for (int i = threadIdx.x; i < a.n_elems; i += blockDim.x) {
    const float v = __ldcg(a.in + i);
    for (int k = 0; k < a.n; ++k)
      a.recv[k][((int64_t)par * a.n + a.rank) * a.cap + i] = v;
  }
  __syncthreads();
  if (threadIdx.x < a.n) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + a.rank, epoch);
    spin_until(a.flags[a.rank] + threadIdx.x, epoch, a.timeout_ns);
  }
  __syncthreads();
  // 2. rank-order sum of the N slots, scale, optional variance
  const float* mine = a.recv[a.rank] + (int64_t)par * a.n * a.cap;
  for (int i = threadIdx.x; i < a.n_elems; i += blockDim.x) {
    float acc = __ldcg(mine + i);
    for (int k = 1; k < a.n; ++k) acc = __fadd_rn(acc, __ldcg(mine + (int64_t)k * a.cap + i));
    if (a.scale_mode == 1) acc = __fmul_rn(acc, a.scale_f);
    else if (a.scale_mode == 2) acc = __double2float_rn(__dmul_rn((double)acc, a.scale_d));
    a.out[i] = acc;
  }
  if (a.C > 0) {
    __syncthreads();
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      const float m = a.out[c];
      a.out[a.C + c] = __fsub_rn(a.out[a.C + c], __fmul_rn(m, m));
    }
  }
}

// fills everything but in / out / n_elems / C / scale from the communicator
inline void small_args_from(P2PComm* c, SmallArgs* a, double scale) {
  for (int k = 0; k < kMaxRanks; ++k) {
    a->recv[k] = k < c->n ? c->small_recv[k] : nullptr;
    a->flags[k] = k < c->n ? c->small_flags[k] : nullptr;
  }
  a->rank = c->rank;
  a->n = c->n;
  a->cap = c->small_cap;
  // a spare word of this rank's own flag block (zeroed with it)
  a->epoch_ctr = c->small_flags[c->rank] + 2 * kMaxRanks + 1;
  a->timeout_ns = g_gp_peer_timeout_ns;
  const ScaleArg s = make_scale(scale);
  a->scale_f = s.fs;
  a->scale_d = s.ds;
  a->scale_mode = s.mode;
}
