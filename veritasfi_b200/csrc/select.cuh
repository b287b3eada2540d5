// select.cuh — block-level exact top-k over 64-bit keys (see topk_common.cuh for the order).
// Used by: candidate reduction after the fused GEMM (K1c), the final ordering after exact
// rescoring (K2), the multi-GPU merge (K3), the BM25 reduction (K4) and rank fusion (K5).
#pragma once
#include "topk_common.cuh"

namespace vfi {

constexpr uint32_t kSortCap = 4096;  // keys sortable in shared memory by one CTA (32 KB)

// scratch of the block-wide MSD radix walk (block_kth_key)
struct RadixScratch {
  uint32_t hist[256];
  uint32_t digit, before, cnt, seen;
};

template <uint32_t CAP>
struct SelectSmemT {
  static constexpr uint32_t kCap = CAP;   // keys sortable in shared memory (multiple of 256)
  uint64_t keys[CAP];
  RadixScratch rs;
  uint32_t ctr;
};
using SelectSmem = SelectSmemT<kSortCap>;

__device__ __forceinline__ uint32_t next_pow2(uint32_t x) {
  uint32_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

// warp 0: find the bin holding the `remaining`-th largest key given the 256-bin histogram
__device__ __forceinline__ void hist_find_bin(const uint32_t* hist, uint32_t remaining,
                                              uint32_t lane, uint32_t* o_digit,
                                              uint32_t* o_before, uint32_t* o_cnt, uint32_t* o_seen) {
  uint32_t c[8];
  uint32_t mine = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i] = hist[lane * 8 + i]; mine += c[i]; }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_down_sync(0xFFFFFFFFu, incl, o);
    if (lane + o < 32) incl += t;
  }
  const uint32_t above = incl - mine;
  if (lane == 0) *o_seen = incl;   // keys counted in this pass
  if (above < remaining && remaining <= above + mine) {
    uint32_t run = above;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      if (run < remaining && remaining <= run + c[i]) {
        *o_digit = lane * 8 + i;
        *o_before = run;
        *o_cnt = c[i];
      }
      run += c[i];
    }
  }
}

// Selection among the n keys already staged in sm->keys[0..n): leaves the min(n,k) best sorted
// descending at the front and returns that count.  blockDim.x must be 256, n <= kSortCap.
// Above 1024 keys a full sort is wasteful: the k-th largest of the 256 per-thread maxima is a lower
// bound of the k-th largest key, so only keys at or above it (about k of them) need sorting.
template <class SM>
__device__ __forceinline__ uint32_t block_topk_smem(SM* sm, uint32_t n, uint32_t k) {
  const uint32_t tid = threadIdx.x;
  constexpr uint32_t kCap = SM::kCap;
  constexpr uint32_t kPer = kCap / 256;
  if (n > 1024 && k < n && k <= 256) {
    uint64_t mine[kPer];
    uint64_t mx = 0ull;
#pragma unroll
    for (uint32_t j = 0; j < kPer; ++j) {
      const uint32_t i = tid + j * 256;
      mine[j] = (i < n) ? sm->keys[i] : 0ull;
      mx = mine[j] > mx ? mine[j] : mx;
    }
    __syncthreads();
    sm->keys[tid] = mx;
    block_bitonic_desc(sm->keys, 256);
    const uint64_t t0 = sm->keys[k - 1];
    if (tid == 0) sm->ctr = 0;
    __syncthreads();
#pragma unroll
    for (uint32_t j = 0; j < kPer; ++j) {
      if (mine[j] >= t0 && mine[j] != 0ull) sm->keys[atomicAdd(&sm->ctr, 1u)] = mine[j];
    }
    __syncthreads();
    n = sm->ctr;
  }
  const uint32_t np = max(next_pow2(n), 2u);
  for (uint32_t i = n + tid; i < np; i += blockDim.x) sm->keys[i] = 0ull;
  block_bitonic_desc(sm->keys, np);
  return min(n, k);
}

// Merge of `rows` result rows of k keys each, staged in shared memory (row r at keys + r*k), into the k_out best of result
// row q.  Every producer in this library emits its rows in result order (score desc, id asc = key descending, padding 0
// last), so the global rank of a key is its place in its own row plus, for every other row, the number of keys above it
// there (binary search; keys of different rows never compare equal: ids are unique) — no sort, ~log2(k) steps x rows per
// key.  Returns false without writing anything when some row is NOT in order (the caller then selects generically).
// All threads of the block must call; ids are written as they are in the keys (global ids).
__device__ inline bool merge_sorted_rows(const uint64_t* keys, int rows, int k, int k_out, int64_t q,
                                         float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  __shared__ uint32_t s_unsorted, s_valid;
  const int n_all = rows * k;
  if (threadIdx.x == 0) { s_unsorted = 0; s_valid = 0; }
  __syncthreads();
  bool bad = false;
  uint32_t valid_mine = 0;
  for (int i = threadIdx.x; i < n_all; i += blockDim.x) {
    const uint64_t key = keys[i];
    valid_mine += key != kKeyNone;
    if (i % k > 0 && (key > keys[i - 1] || (key == keys[i - 1] && key != kKeyNone))) bad = true;
  }
  if (bad) s_unsorted = 1;
  if (valid_mine) atomicAdd(&s_valid, valid_mine);
  __syncthreads();
  const bool ok = s_unsorted == 0;
  if (ok) {
    for (int i = threadIdx.x; i < n_all; i += blockDim.x) {
      const uint64_t key = keys[i];
      if (key == kKeyNone) continue;
      const int r = i / k;
      int rank = i - r * k;
      for (int r2 = 0; r2 < rows; ++r2) {
        if (r2 == r) continue;
        const uint64_t* row = keys + r2 * k;
        int lo = 0, hi = k;                           // first position whose key is below `key`
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (row[mid] > key) lo = mid + 1; else hi = mid;
        }
        rank += lo;
      }
      if (rank < k_out) {
        out_scores[q * k_out + rank] = key_score(key);
        out_ids[q * k_out + rank] = static_cast<int64_t>(key_id(key));
      }
    }
    for (int i = threadIdx.x + static_cast<int>(s_valid); i < k_out; i += blockDim.x) {     // fewer than k_out real keys
      out_scores[q * k_out + i] = -3.402823466e+38f;
      out_ids[q * k_out + i] = -1;
    }
  }
  __syncthreads();
  return ok;
}

// Selection WITHOUT ordering among the n distinct non-zero keys staged in sm->keys[0..n), n <= SM::kCap: returns the k-th
// largest key (the smallest of the k best) — the keys >= it are exactly the k best — or 0 when n < k (every key is kept).
// MSD radix walk over 8-bit digits of the 64-bit keys, histogram in shared memory with warp-aggregated atomics (the keys
// of one query share their leading bytes, so plain atomics would serialise on one bin).  Three barriers per pass and at
// most 8 passes (typically 3-4: the walk stops as soon as a bin holds exactly the keys still wanted), against the 55
// barrier-separated stages of a 1024-key bitonic sort.  blockDim.x must be a multiple of 32, >= 64.  All threads must call.
__device__ __forceinline__ uint64_t block_kth_key(const uint64_t* keys, uint32_t n, uint32_t k, RadixScratch* rs) {
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  if (n < k) return 0ull;
  uint64_t prefix = 0;
  uint32_t remaining = k;
  for (int shift = 56; shift >= 0 && n > k; shift -= 8) {
    for (uint32_t i = tid; i < 256; i += blockDim.x) rs->hist[i] = 0;
    __syncthreads();
    const uint64_t himask = (shift == 56) ? 0ull : (~0ull << (shift + 8));
    for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {     // warp-uniform trip count: every lane takes part in the match
      const uint32_t i = i0 + tid;
      uint32_t digit = 0xFFFFFFFFu;
      if (i < n) {
        const uint64_t key = keys[i];
        if ((key & himask) == prefix) digit = static_cast<uint32_t>(key >> shift) & 0xFFu;
      }
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
      if (digit != 0xFFFFFFFFu && lane == static_cast<uint32_t>(__ffs(peers) - 1)) atomicAdd(&rs->hist[digit], __popc(peers));
    }
    __syncthreads();
    if (tid < 32) hist_find_bin(rs->hist, remaining, tid, &rs->digit, &rs->before, &rs->cnt, &rs->seen);
    __syncthreads();
    prefix |= static_cast<uint64_t>(rs->digit) << shift;
    remaining -= rs->before;
    const bool stop = (rs->cnt == remaining) || shift == 0;
    __syncthreads();
    if (stop) break;
  }
  // every key >= prefix is kept; the k-th largest is the smallest of them
  __shared__ unsigned long long s_min;
  if (tid == 0) s_min = ~0ull;
  __syncthreads();
  unsigned long long mn = ~0ull;
  for (uint32_t i = tid; i < n; i += blockDim.x) {
    const uint64_t key = keys[i];
    if (key >= prefix && key < mn) mn = key;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xFFFFFFFFu, mn, o);
    mn = t < mn ? t : mn;
  }
  if (lane == 0 && mn != ~0ull) atomicMin(&s_min, mn);
  __syncthreads();
  const uint64_t out = s_min;
  __syncthreads();     // s_min may be re-initialised by the next call
  return out;
}
template <class SM>
__device__ __forceinline__ uint64_t block_kth_key_smem(SM* sm, uint32_t n, uint32_t k) {
  return block_kth_key(sm->keys, n, k, &sm->rs);
}

// Src: struct with   template<class F> __device__ void for_each(F f) const
// calling f(key) for the keys assigned to this thread (each key visited by exactly one thread).
// Leaves the min(total,k) best keys sorted descending in sm->keys[0..); returns that count.
// Keys must be distinct and != 0.  k <= SM::kCap.  All 256 threads of the block must call.
template <class Src, class SM>
__device__ uint32_t block_topk(const Src& src, uint32_t total, uint32_t k, SM* sm) {
  const uint32_t tid = threadIdx.x;
  constexpr uint32_t kSortCap = SM::kCap;   // shadows the namespace constant: every bound below is this buffer's
  if (total <= kSortCap) {
    // stage once, select in shared memory
    if (tid == 0) sm->ctr = 0;
    __syncthreads();
    src.for_each([&](uint64_t key) {
      if (key != 0ull) {
        const uint32_t pos = atomicAdd(&sm->ctr, 1u);
        if (pos < kSortCap) sm->keys[pos] = key;
      }
    });
    __syncthreads();
    return block_topk_smem(sm, min(sm->ctr, kSortCap), k);
  }
  uint64_t thresh = 0;  // keep keys >= thresh
  bool have_thresh = false;
  if (k <= 256) {
    // streaming variant of the per-thread-maximum bound: two passes over the source
    uint64_t mx = 0ull;
    src.for_each([&](uint64_t key) { mx = key > mx ? key : mx; });
    sm->keys[tid] = mx;
    block_bitonic_desc(sm->keys, 256);
    const uint64_t t0 = sm->keys[k - 1];
    if (tid == 0) sm->ctr = 0;
    __syncthreads();
    src.for_each([&](uint64_t key) {
      if (key >= t0 && key != 0ull) {
        const uint32_t pos = atomicAdd(&sm->ctr, 1u);
        if (pos < kSortCap) sm->keys[pos] = key;
      }
    });
    __syncthreads();
    if (sm->ctr <= kSortCap) return block_topk_smem(sm, sm->ctr, k);
    __syncthreads();   // masses of ties above the bound: fall through to the exact radix walk
  }
  {
    // MSD radix walk over the 64-bit keys for the exact k-th largest key (<= 8 passes)
    uint64_t prefix = 0;
    uint32_t remaining = k;
    for (int shift = 56; shift >= 0; shift -= 8) {
      if (tid < 256) sm->rs.hist[tid] = 0;
      __syncthreads();
      const uint64_t himask = (shift == 56) ? 0ull : (~0ull << (shift + 8));
      src.for_each([&](uint64_t key) {
        if ((key & himask) == prefix && key != 0ull) atomicAdd(&sm->rs.hist[(key >> shift) & 0xFF], 1u);
      });
      __syncthreads();
      if (tid < 32) hist_find_bin(sm->rs.hist, remaining, tid, &sm->rs.digit, &sm->rs.before, &sm->rs.cnt, &sm->rs.seen);
      __syncthreads();
      if (sm->rs.seen <= remaining) {   // fewer real keys than asked for (padding in the source): keep them all
        __syncthreads();
        prefix = 0;
        break;
      }
      prefix |= static_cast<uint64_t>(sm->rs.digit) << shift;
      remaining -= sm->rs.before;
      const bool stop = (sm->rs.cnt == remaining) || shift == 0;
      __syncthreads();
      if (stop) break;
    }
    thresh = prefix;
    have_thresh = true;
  }
  (void)have_thresh;
  if (tid == 0) sm->ctr = 0;
  __syncthreads();
  src.for_each([&](uint64_t key) {
    if (key >= thresh && key != 0ull) {
      const uint32_t pos = atomicAdd(&sm->ctr, 1u);
      if (pos < kSortCap) sm->keys[pos] = key;
    }
  });
  __syncthreads();
  return block_topk_smem(sm, min(sm->ctr, kSortCap), k);
}

}  // namespace vfi
