// gp_p2p.cuh -- state and in-kernel cross-GPU barrier shared by the peer-memory
// allreduce (gp_p2p.cu) and the NVSwitch-multicast allreduce (gp_mc.cu).
// Include inside an anonymous namespace.
#pragma once
#include "gp_common.cuh"

constexpr int kMaxRanks = 8;

struct P2PComm {
  int rank, n;
  void* bufs[kMaxRanks];       // this process's mappings of every rank's packed buffer
  uint32_t* flags[kMaxRanks];  // every rank's flag block: [2 * kMaxRanks] words + grid counter
  uint32_t epoch;
  // small-message one-shot allreduce: every rank's receive area
  // [2 parities][n ranks][small_cap floats] and its flag words [kMaxRanks]
  float* small_recv[kMaxRanks];
  uint32_t* small_flags[kMaxRanks];
  int64_t small_cap;
  uint32_t small_epoch;
};

struct P2PArgs {
  void* bufs[kMaxRanks];
  uint32_t* flags[kMaxRanks];
  int rank, n;
  int64_t begin, end;  // this rank's shard, in elements
  uint32_t epoch;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all ranks have reached `value` in slot set `which` (0: ready, 1: done)
__device__ __forceinline__ void signal_all(const P2PArgs& a, int which, uint32_t value) {
  // thread t < n tells rank t
  if (threadIdx.x < a.n)
    st_release_sys(a.flags[threadIdx.x] + which * kMaxRanks + a.rank, value);
}
__device__ __forceinline__ void wait_all(const P2PArgs& a, int which, uint32_t value) {
  if (threadIdx.x < a.n) {
    const uint32_t* f = a.flags[a.rank] + which * kMaxRanks + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - value) < 0) {
      // a peer that never arrives (crashed process) must not hang the GPU:
      // give up after ~2e10 cycles (about 10 s) with a launch failure
      if (clock64() - t0 > 20000000000LL) __trap();
    }
  }
}


// Grid-wide completion followed by the "done" barrier: returns after every
// rank's stores of this launch have landed everywhere (last CTA only waits).
__device__ __forceinline__ void finish_all(const P2PArgs& a) {
  // bar.sync orders the CTA's stores before thread 0's system-scope fence, which
  // is cumulative: one fence per CTA instead of one per thread.
  __syncthreads();
  __shared__ bool last;
  uint32_t* counter = a.flags[a.rank] + 2 * kMaxRanks;
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *counter = 0;
  __threadfence_system();
  signal_all(a, 1, a.epoch);
  wait_all(a, 1, a.epoch);
}
