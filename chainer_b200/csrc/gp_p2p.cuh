// gp_p2p.cuh -- state and in-kernel cross-GPU synchronisation shared by the peer-memory
// allreduce (gp_p2p.cu), the NVSwitch-multicast allreduce (gp_mc.cu) and the one-launch
// step kernels (gp_step.cu).  Include inside an anonymous namespace.
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
  // one-launch step (gp_step.cu): every rank's per-tile words
  // [tile_cap "packed" counters | tile_cap "reduced" flags], the tile size they
  // were zeroed for, and the step epoch since then
  uint32_t* step_words[kMaxRanks];
  int64_t step_tile_cap;
  int64_t step_tile_elems;
  uint32_t step_epoch;
};

struct P2PArgs {
  void* bufs[kMaxRanks];
  uint32_t* flags[kMaxRanks];
  int rank, n;
  int64_t begin, end;  // this rank's shard, in elements
  uint32_t epoch;
  unsigned long long timeout_ns;  // 0: wait for ever (as NCCL does)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Wait until the word at `p` (in THIS GPU's memory; peers write it over NVLink) has
// reached `value`.  A peer may legitimately be late by seconds or minutes (a rank-0-only
// evaluation or snapshot, a data-loader stall, first-iteration autotuning): the wait
// backs off with nanosleep so that it does not hold issue slots, and it is bounded only
// by `timeout_ns` of WALL time (host-configurable, default 30 minutes like a collective
// watchdog; 0 = unbounded, as ncclAllReduce behaves).  Only when that bound expires --
// a peer process has died -- the kernel gives up with a launch failure instead of
// hanging the GPU for ever.
__device__ __forceinline__ void spin_until(const uint32_t* p, uint32_t value,
                                           unsigned long long timeout_ns) {
  unsigned ns = 32;
  unsigned long long t0 = 0;
  for (unsigned it = 1;; ++it) {
    if ((int32_t)(ld_acquire_sys(p) - value) >= 0) return;
    if (it > 32) {
      __nanosleep(ns);
      if (ns < 2048) ns <<= 1;
    }
    if ((it & 255u) == 0 && timeout_ns) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > timeout_ns) __trap();
    }
  }
}

// all ranks have reached `value` in slot set `which` (0: ready, 1: done)
__device__ __forceinline__ void signal_all(const P2PArgs& a, int which, uint32_t value) {
  // thread t < n tells rank t
  if (threadIdx.x < a.n)
    st_release_sys(a.flags[threadIdx.x] + which * kMaxRanks + a.rank, value);
}
__device__ __forceinline__ void wait_all(const P2PArgs& a, int which, uint32_t value) {
  if (threadIdx.x < a.n)
    spin_until(a.flags[a.rank] + which * kMaxRanks + threadIdx.x, value, a.timeout_ns);
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
