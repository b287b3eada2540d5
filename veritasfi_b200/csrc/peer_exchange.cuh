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
// The reference has no counterpart (its workers are replicas, experiments/retriever/step3_mul.py:405-446).
#pragma once
#include "select.cuh"
#include "topk_common.cuh"

namespace vfi {

constexpr int kMaxPeers = 16;

struct ExchangeParams {
  uint64_t* win[kMaxPeers];     // window base of every rank as mapped in THIS process (win[rank] is local)
  uint32_t* flags[kMaxPeers];   // flag base of every rank
  int rank, world;
  int nq, k, k_out;
  int64_t max_nq;               // window geometry: [2][world][max_nq][max_k] keys, flags [2][world][max_nq]
  int max_k;
  uint32_t epoch;               // 1, 2, 3 ... the same sequence on every rank
  const float* scores;          // this rank's results [nq][k]
  const int64_t* ids;           // global ids, -1 = padding
  float* out_scores;            // [nq][k_out]
  int64_t* out_ids;
  const int* fail_a;            // device: > 0 when this rank's local results are not final yet (may be null)
  const int* fail_b;
  int* any_fail;                // device: set to 1 when any rank reported fail (may be null)
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

__global__ void __launch_bounds__(256) exchange_merge_kernel(const ExchangeParams p) {
  extern __shared__ uint8_t smem_raw[];
  SelectSmem* sm = reinterpret_cast<SelectSmem*>(smem_raw);
  const uint32_t parity = p.epoch & 1u;
  const int64_t slot_keys = p.max_nq * p.max_k;                     // keys per (parity, source rank)
  const int64_t par_keys = static_cast<int64_t>(p.world) * slot_keys;
  const int64_t par_flags = static_cast<int64_t>(p.world) * p.max_nq;

  const uint32_t my_fail = ((p.fail_a != nullptr && *p.fail_a > 0) || (p.fail_b != nullptr && *p.fail_b > 0)) ? 1u : 0u;
  const uint32_t stamp = (p.epoch << 1) | my_fail;
  __shared__ uint32_t s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  // ---- push: all queries of this CTA first, so the stores of every query are in flight together
  for (int q = blockIdx.x; q < p.nq; q += gridDim.x) {
    for (int j = threadIdx.x; j < p.k; j += blockDim.x) {
      const int64_t o = static_cast<int64_t>(q) * p.k + j;
      const int64_t id = p.ids[o];
      const uint64_t key = (id >= 0) ? make_key(p.scores[o], static_cast<uint32_t>(id)) : kKeyNone;
      const int64_t dst = parity * par_keys + p.rank * slot_keys + static_cast<int64_t>(q) * p.max_k + j;
#pragma unroll 1
      for (int r = 0; r < p.world; ++r) p.win[(p.rank + r) % p.world][dst] = key;   // start with the local copy, then fan out
    }
    __threadfence_system();   // each writer orders its own stores before the flag
    __syncthreads();
    if (threadIdx.x < static_cast<uint32_t>(p.world)) {
      const int r = (p.rank + threadIdx.x) % p.world;
      st_release_sys_u32(p.flags[r] + parity * par_flags + static_cast<int64_t>(p.rank) * p.max_nq + q, stamp);
    }
  }
  // ---- wait + merge
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
  if (threadIdx.x == 0 && s_fail != 0 && p.any_fail != nullptr) *p.any_fail = 1;
}

}  // namespace vfi
