// reduce.cuh — the two per-query reduction kernels shared by the dense and the sparse paths:
//   K1c cand_reduce   union of the per-group candidate buffers of one query -> the k' best keys
//   K2b finalize      final (score desc, id asc) order of exact keys + exactness certificate
// Templates only, so every translation unit of libvfi.so can include this header.
#pragma once
#include "select.cuh"
#include "topk_common.cuh"

namespace vfi {

// ------------------------------------------------------------------------------------------
// K1c: union of the per-group candidate buffers of one query -> the k' best keys, in no particular order (every consumer
// rescoring or ordering them itself) unless the union exceeds the shared-memory buffer (then sorted descending).
// bound[q] = an upper bound on the tensor-core score of every row that is NOT in the output:
//   the k'-th key's score when at least k' rows were admitted, else the admission hint (rows at
//   or below the hint were never admitted), else -inf (every row of the shard is a candidate).
// ------------------------------------------------------------------------------------------
struct GroupBufSrc {
  const uint64_t* cand;
  const uint32_t* pref;   // shared memory: exclusive prefix of the per-group counts, [n_groups + 1]
  int n_groups, nq_pad, cap, q;
  template <class F>
  __device__ void for_each(F f) const {
    // flat index -> (group, offset) by binary search in the prefix array; four independent loads in flight per thread
    const uint32_t total = pref[n_groups];
    auto locate = [&](uint32_t i) -> const uint64_t* {
      int lo = 0, hi = n_groups;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pref[mid] <= i) lo = mid; else hi = mid;
      }
      return cand + (static_cast<size_t>(lo) * nq_pad + q) * cap + (i - pref[lo]);
    };
    uint32_t i = threadIdx.x;
    const uint32_t stride = blockDim.x;
    for (; i + 3 * stride < total; i += 4 * stride) {
      uint64_t v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = *locate(i + j * stride);
#pragma unroll
      for (int j = 0; j < 4; ++j) f(v[j]);
    }
    for (; i < total; i += stride) f(*locate(i));
  }
};

using CandSmem = SelectSmemT<1024>;   // 8 KB of keys: seven CTAs per SM, so 1024 queries are one wave (k' <= 256 only)
template <class SM>
__global__ void __launch_bounds__(256, (sizeof(SM) <= 16384 ? 7 : 4)) cand_reduce_kernel(const uint64_t* __restrict__ cand,
                                                          const uint32_t* __restrict__ cnt,
                                                          int n_groups, int nq_pad, int cap,
                                                          int keep, const float* __restrict__ tau_init,
                                                          uint64_t* __restrict__ out_keys,
                                                          uint32_t* __restrict__ out_n,
                                                          float* __restrict__ bound) {
  extern __shared__ uint8_t smem_raw[];
  SM* sm = reinterpret_cast<SM*>(smem_raw);
  const int q = blockIdx.x;
  __shared__ uint32_t s_pref[1025];
  // counts of all group buffers in one coalesced sweep, then an exclusive scan (n_groups <= 1024)
  for (int g = threadIdx.x; g < n_groups; g += blockDim.x)
    s_pref[g + 1] = min(cnt[static_cast<size_t>(g) * nq_pad + q], static_cast<uint32_t>(cap));
  if (threadIdx.x == 0) s_pref[0] = 0;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t carry = 0;
    for (int base = 0; base < n_groups; base += 32) {
      const int g = base + threadIdx.x;
      uint32_t v = (g < n_groups) ? s_pref[g + 1] : 0u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (static_cast<int>(threadIdx.x) >= o) v += t;
      }
      if (g < n_groups) s_pref[g + 1] = v + carry;
      carry += __shfl_sync(0xFFFFFFFFu, v, 31);
    }
  }
  __syncthreads();
  const uint32_t total = s_pref[n_groups];
  GroupBufSrc src{cand, s_pref, n_groups, nq_pad, cap, q};
  if (total <= SM::kCap) {
    // the consumers (K2 rescoring, K2a) take the k' best in any order: stage the keys once, find the k'-th largest by a
    // radix walk and emit what is at or above it — no sort (round 2: 58 -> ~25 us for 1024 queries of ~700 keys)
    if (threadIdx.x == 0) sm->ctr = 0;
    __syncthreads();
    src.for_each([&](uint64_t key) {
      if (key != 0ull) sm->keys[atomicAdd(&sm->ctr, 1u)] = key;
    });
    __syncthreads();
    const uint32_t n_keys = sm->ctr;
    const uint64_t kth = block_kth_key_smem(sm, n_keys, static_cast<uint32_t>(keep));
    if (threadIdx.x == 0) sm->ctr = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_keys; i += blockDim.x) {
      const uint64_t key = sm->keys[i];
      if (key >= kth) out_keys[static_cast<size_t>(q) * keep + atomicAdd(&sm->ctr, 1u)] = key;
    }
    __syncthreads();
    const uint32_t n = sm->ctr;                     // min(n_keys, keep)
    for (uint32_t i = n + threadIdx.x; i < static_cast<uint32_t>(keep); i += blockDim.x)
      out_keys[static_cast<size_t>(q) * keep + i] = kKeyNone;
    if (threadIdx.x == 0) {
      out_n[q] = n;
      float b;
      if (n_keys >= static_cast<uint32_t>(keep)) b = key_score(kth);
      else b = (tau_init != nullptr) ? tau_init[q] : -INFINITY;
      bound[q] = b;
    }
    return;
  }
  const uint32_t n = block_topk(src, total, static_cast<uint32_t>(keep), sm);
  for (uint32_t i = threadIdx.x; i < static_cast<uint32_t>(keep); i += blockDim.x)
    out_keys[static_cast<size_t>(q) * keep + i] = (i < n) ? sm->keys[i] : kKeyNone;
  if (threadIdx.x == 0) {
    out_n[q] = n;
    float b;
    if (total >= static_cast<uint32_t>(keep)) b = key_score(sm->keys[keep - 1]);
    else b = (tau_init != nullptr) ? tau_init[q] : -INFINITY;
    bound[q] = b;
  }
}

// ------------------------------------------------------------------------------------------
// K2b: final order + certificate.  One CTA per selected query.
// The result is provably the exact top-k iff every excluded row r satisfies
//   exact(r) <= tc(r) + eps <= bound + eps < exact k-th of the candidates.
// Queries failing the test are appended to `flagged` for the exhaustive pass.
// ------------------------------------------------------------------------------------------
struct KeyArraySrc {
  const uint64_t* keys;
  int64_t n;
  template <class F>
  __device__ void for_each(F f) const {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint64_t k = keys[i];
      if (k != kKeyNone) f(k);
    }
  }
};

template <int CHECK>
__global__ void __launch_bounds__(256) finalize_kernel(
    const uint64_t* __restrict__ keys2, int64_t keys2_pitch, int64_t n_slots,
    const int* __restrict__ qsel, const uint32_t* __restrict__ n_cand, int k, int64_t id_offset,
    const float* __restrict__ bound, const float* __restrict__ eps,
    float* __restrict__ out_scores, int64_t* __restrict__ out_ids, int* __restrict__ flagged,
    int* __restrict__ n_flagged, int* __restrict__ done_ctas, int* __restrict__ host_n_flagged) {
  extern __shared__ uint8_t smem_raw[];
  SelectSmem* sm = reinterpret_cast<SelectSmem*>(smem_raw);
  const int qslot = blockIdx.x;
  const int q = (qsel != nullptr) ? qsel[qslot] : qslot;
  const uint32_t total = (n_cand != nullptr) ? n_cand[q] : static_cast<uint32_t>(n_slots);
  KeyArraySrc src{keys2 + static_cast<int64_t>(qslot) * keys2_pitch, n_slots};
  const uint32_t n = block_topk(src, total, static_cast<uint32_t>(k), sm);
  bool ok = true;
  if (CHECK) {
    const float b = bound[q];
    if (b != -INFINITY) {
      if (n < static_cast<uint32_t>(k)) ok = false;
      else ok = key_score(sm->keys[k - 1]) > b + eps[q];
    }
  }
  if (ok) {
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      const bool has = static_cast<uint32_t>(i) < n;
      const uint64_t key = has ? sm->keys[i] : 0ull;
      out_scores[static_cast<int64_t>(q) * k + i] = has ? key_score(key) : -3.402823466e+38f;
      out_ids[static_cast<int64_t>(q) * k + i] = has ? static_cast<int64_t>(key_id(key)) + id_offset : -1;
    }
  } else if (threadIdx.x == 0) {
    flagged[atomicAdd(n_flagged, 1)] = q;
  }
  if (CHECK) publish_flag_count(n_flagged, done_ctas, host_n_flagged);
}

}  // namespace vfi
