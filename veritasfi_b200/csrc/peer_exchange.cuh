// peer_exchange.cuh — K3p: the multi-GPU exchange and the global merge as ONE kernel over NVLink peer memory.
//
// after the per-shard search every rank holds its local top-k as (score, global id); the
// global top-k is the top-k of the union.  The NCCL route is all-gather + merge_kernel (three kernels and a
// collective launch).  Here every rank owns a receive window that its peers have mapped (CUDA IPC), and a
// single launch per rank does:
//   push   — the CTA that owns query q packs its k results into 64-bit keys and stores them straight into
//            slot [parity][my rank][q] of EVERY rank's window (remote stores over NVLink, 2 KB coalesced
//            bursts), then publishes them with a system-scope release store of (epoch << 1 | fail) into
//            flag [parity][my rank][q] of that rank — `fail` says that this rank's exactness certificate
//            flagged some query of the batch, i.e. the rows it is pushing may still be repaired;
//   wait   — the same CTA acquires flag [parity][r][q] of its own window for every rank r;
//   merge  — and selects the k_out best of the world*k keys now sitting in local HBM/L2; the OR of the fail
//            bits of all ranks goes to *any_fail, the same value on every rank, so the ranks agree on whether
//            the batch has to be exchanged again after the repair (pipelined sharded search, sharded.py).
// No CTA ever waits on another CTA of its own grid, and the grid is sized to be fully resident, so the
// only dependency is "the peer has reached the same exchange", exactly the dependency of a collective.
// Windows are double-buffered by epoch parity: a rank can only start epoch e+2 after its own epoch e+1
// completed, which needed every peer's epoch e+1 pushes, which those peers issue after finishing their
// epoch-e reads (stream order) — so overwriting parity e&1 at e+2 never races with a reader.
// Fused with the compute (round 2): the rescoring kernel K2 (dense_support.cuh: rescore_finalize_kernel) sends — the CTA that has
// just certified and ordered query q stores that row into every peer's window (push_row_data) while the other queries of the
// batch are still being rescored, and does not wait for the NVLink round trip; exchange_wait_merge_kernel, next in the stream,
// first publishes the flags of this rank's rows (with the batch's certificate state as the fail bit), then waits for the peers'
// and merges.  (First version: K2 published the flags itself behind a system-scope fence — the fence held every K2 CTA for the
// round trip, +45 us on the kernel for 1024 queries, as much as the separate push phase it replaced.)  The stream order
// push(e) < merge(e) < push(e+1) on every rank keeps the double-buffer argument above intact.
// The reference has no counterpart (its workers are replicas, experiments/retriever/step3_mul.py:405-446).
#pragma once
#include "select.cuh"
#include "topk_common.cuh"

namespace vfi {

constexpr int kMaxPeers = 16;

// where one rank's rows of one epoch go: a by-value kernel parameter of the pushing kernel (world == 0: no push)
struct PushTarget {
  uint64_t* win[kMaxPeers];     // window base of every rank as mapped in THIS process (win[rank] is local)
  uint32_t* flags[kMaxPeers];   // flag base of every rank
  int rank, world;
  int64_t max_nq;               // window geometry: [2][world][max_nq][max_k] keys, flags [2][world][max_nq]
  int max_k;
  uint32_t epoch;               // 1, 2, 3 ... the same sequence on every rank
};

struct ExchangeParams {
  uint64_t* win[kMaxPeers];
  uint32_t* flags[kMaxPeers];
  int rank, world;
  int nq, k, k_out;
  int64_t max_nq;
  int max_k;
  uint32_t epoch;
  const float* scores;          // this rank's results [nq][k]
  const int64_t* ids;           // global ids, -1 = padding
  float* out_scores;            // [nq][k_out]
  int64_t* out_ids;
  const int* fail_a;            // device: > 0 when this rank's local results are not final yet (may be null)
  const int* fail_b;
  int* any_fail;                // device: set to 1 when any rank reported fail (may be null)
  int* fail_acc;                // device, zero between launches: OR of the fail bits seen by this launch's CTAs (may be null)
  int* done_ctas;               // device, zero between launches: finished CTAs
  int* host_any_fail;           // mapped host memory: the last CTA stores the OR here — no copy operation in the stream
  unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// The calling CTA stores row q of this rank (k keys; key_at(j) -> the j-th key with its GLOBAL id, kKeyNone = padding) into
// slot [parity][rank][q] of every rank's window.  Plain remote stores, nothing waited for: the row is published later by
// publish_row_flags from a kernel that runs AFTER this one in the stream (its system-scope release store is cumulative over
// everything that happened before it, and a kernel boundary orders this kernel's stores before it), so a compute kernel can
// send its results without stalling on the NVLink round trip.
template <class KeyAt>
__device__ __forceinline__ void push_row_data(const PushTarget& t, int q, int k, KeyAt key_at) {
  const uint32_t parity = t.epoch & 1u;
  const int64_t slot_keys = t.max_nq * t.max_k;
  const int64_t par_keys = static_cast<int64_t>(t.world) * slot_keys;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const uint64_t key = key_at(j);
    const int64_t dst = parity * par_keys + t.rank * slot_keys + static_cast<int64_t>(q) * t.max_k + j;
#pragma unroll 1
    for (int r = 0; r < t.world; ++r) t.win[(t.rank + r) % t.world][dst] = key;   // start with the local copy, then fan out
  }
}
// flag [parity][rank][q] = epoch << 1 | fail in every rank's flag array; threads 0..world-1 of the CTA store one each
__device__ __forceinline__ void publish_row_flags(const PushTarget& t, int q, uint32_t fail) {
  const uint32_t parity = t.epoch & 1u;
  const int64_t par_flags = static_cast<int64_t>(t.world) * t.max_nq;
  if (threadIdx.x < static_cast<uint32_t>(t.world)) {
    const int r = (t.rank + threadIdx.x) % t.world;
    st_release_sys_u32(t.flags[r] + parity * par_flags + static_cast<int64_t>(t.rank) * t.max_nq + q, (t.epoch << 1) | (fail & 1u));
  }
}
// both in one kernel: the writers order their own stores before the flag.  All threads of the CTA must call (one barrier
// inside); blockDim.x >= world.
template <class KeyAt>
__device__ __forceinline__ void push_row(const PushTarget& t, int q, int k, uint32_t fail, KeyAt key_at) {
  push_row_data(t, q, k, key_at);
  __threadfence_system();
  __syncthreads();
  publish_row_flags(t, q, fail);
}

struct WindowSrc {
  const uint64_t* base;   // local window, parity already applied: [world][max_nq][max_k]
  int world, k, q;
  int64_t max_nq;
  int max_k;
  template <class F>
  __device__ void for_each(F f) const {
    const int n = world * k;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int r = i / k, j = i - r * k;
      // ordered after the acquire of the flags by the barrier that follows it
      const uint64_t key = __ldcg(base + (static_cast<int64_t>(r) * max_nq + q) * max_k + j);   // L2: peers wrote it
      if (key != kKeyNone) f(key);
    }
  }
};

__device__ __forceinline__ PushTarget push_target_of(const ExchangeParams& p) {
  PushTarget t;
  for (int r = 0; r < kMaxPeers; ++r) { t.win[r] = p.win[r]; t.flags[r] = p.flags[r]; }
  t.rank = p.rank; t.world = p.world; t.max_nq = p.max_nq; t.max_k = p.max_k; t.epoch = p.epoch;
  return t;
}

// push phase: all queries of this CTA first, so the stores of every query are in flight together
__device__ __forceinline__ void exchange_push_phase(const ExchangeParams& p) {
  const uint32_t my_fail = ((p.fail_a != nullptr && *p.fail_a > 0) || (p.fail_b != nullptr && *p.fail_b > 0)) ? 1u : 0u;
  const PushTarget t = push_target_of(p);
  for (int q = blockIdx.x; q < p.nq; q += gridDim.x) {
    push_row(t, q, p.k, my_fail, [&](int j) -> uint64_t {
      const int64_t o = static_cast<int64_t>(q) * p.k + j;
      const int64_t id = p.ids[o];
      return (id >= 0) ? make_key(p.scores[o], static_cast<uint32_t>(id)) : kKeyNone;
    });
  }
}

// wait + merge phase, then the OR of the fail bits for the host
__device__ __forceinline__ void exchange_publish_phase(const ExchangeParams& p) {
  const uint32_t my_fail = ((p.fail_a != nullptr && *p.fail_a > 0) || (p.fail_b != nullptr && *p.fail_b > 0)) ? 1u : 0u;
  const PushTarget t = push_target_of(p);
  __threadfence_system();
  for (int q = blockIdx.x; q < p.nq; q += gridDim.x) publish_row_flags(t, q, my_fail);
}

__device__ __forceinline__ void exchange_wait_merge_phase(const ExchangeParams& p, SelectSmem* sm) {
  const uint32_t parity = p.epoch & 1u;
  const int64_t slot_keys = p.max_nq * p.max_k;
  const int64_t par_keys = static_cast<int64_t>(p.world) * slot_keys;
  const int64_t par_flags = static_cast<int64_t>(p.world) * p.max_nq;
  __shared__ uint32_t s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();
  const uint32_t* my_flags = p.flags[p.rank] + parity * par_flags;
  const uint64_t* my_win = p.win[p.rank] + parity * par_keys;
  for (int q = blockIdx.x; q < p.nq; q += gridDim.x) {
    if (threadIdx.x < static_cast<uint32_t>(p.world)) {
      const uint32_t* f = my_flags + static_cast<int64_t>(threadIdx.x) * p.max_nq + q;
      const unsigned long long t0 = global_timer_ns();
      uint32_t spins = 0, v;
      while (((v = ld_acquire_sys_u32(f)) >> 1) != p.epoch) {
        // a peer that never arrives (crashed rank, mismatched call sequence) must not hang the GPU
        if ((++spins & 0x3FFu) == 0u && global_timer_ns() - t0 > p.timeout_ns) __trap();
      }
      if (v & 1u) s_fail = 1;
    }
    __syncthreads();
    const int n_all = p.world * p.k;
    bool merged = false;
    if (n_all <= static_cast<int>(kSortCap)) {
      for (int i = threadIdx.x; i < n_all; i += blockDim.x) {
        const int r = i / p.k, j = i - r * p.k;
        sm->keys[i] = __ldcg(my_win + (static_cast<int64_t>(r) * p.max_nq + q) * p.max_k + j);   // L2: peers wrote it
      }
      __syncthreads();
      merged = merge_sorted_rows(sm->keys, p.world, p.k, p.k_out, q, p.out_scores, p.out_ids);
    }
    if (!merged) {
      WindowSrc src{my_win, p.world, p.k, q, p.max_nq, p.max_k};
      const uint32_t n = block_topk(src, static_cast<uint32_t>(p.world * p.k), static_cast<uint32_t>(p.k_out), sm);
      for (int i = threadIdx.x; i < p.k_out; i += blockDim.x) {
        const bool has = static_cast<uint32_t>(i) < n;
        const uint64_t key = has ? sm->keys[i] : 0ull;
        p.out_scores[static_cast<int64_t>(q) * p.k_out + i] = has ? key_score(key) : -3.402823466e+38f;
        p.out_ids[static_cast<int64_t>(q) * p.k_out + i] = has ? static_cast<int64_t>(key_id(key)) : -1;
      }
    }
    __syncthreads();   // sm is reused by the next query
  }
  if (threadIdx.x == 0) {
    if (s_fail != 0) {
      if (p.any_fail != nullptr) *p.any_fail = 1;
      if (p.fail_acc != nullptr) atomicOr(p.fail_acc, 1);
    }
    if (p.done_ctas != nullptr) {
      __threadfence();
      if (atomicAdd(p.done_ctas, 1) == static_cast<int>(gridDim.x) - 1) {
        __threadfence();
        *reinterpret_cast<volatile int*>(p.host_any_fail) = *reinterpret_cast<volatile int*>(p.fail_acc);
        *p.fail_acc = 0;          // ready for the launch that reuses this slot (stream-ordered behind this one)
        *p.done_ctas = 0;
        __threadfence_system();
      }
    }
  }
}

// The three kernels built from these phases (push + wait + merge, push, wait + merge) are defined in api_exchange.cu.

}  // namespace vfi
