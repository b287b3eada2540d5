// sparse_fuse.cuh — the bandwidth-bound kernels of the sparse and fusion paths:
//   K4 bm25_kernel   BM25 scoring over token-major (CSC) postings with a fused top-k'
//   K3 merge_kernel  global top-k after the all-gather of per-shard (score, id) lists
//   K5 rrf_kernel / union_kernel   rank fusion across retrieval paths
//
// K4 restates bm25s.BM25.retrieve called at /root/reference/src/utils/bm25Retriever.py:75-79:
// scores[doc] starts at +0 and receives the precomputed impact of every query token, in query
// token order, in fp32 (np.add.at order).
#pragma once
#include <limits.h>

#include "ptx.cuh"
#include "select.cuh"
#include "topk_common.cuh"

namespace vfi {

// ---- K4: BM25 over doc ranges ------------------------------------------------------------
// Work item = (query, doc segment); a CTA sweeps the segment in ranges of 8192 docs whose fp32
// accumulator lives in shared memory.  Inside a range the query's tokens are applied ONE AFTER THE
// OTHER (a block barrier between tokens), each token's postings streamed by all threads with
// several loads in flight per thread: a doc occurs once per posting list, so the adds of one token
// never collide, and the barrier fixes the per-doc order to query-token order — np.add.at's fp32
// result bit for bit, without atomics.  All posting offsets of the segment (token x range boundary)
// are found by parallel binary searches once per work item, so the sweep itself has no searches,
// no speculation and no per-tile bookkeeping.
constexpr int kBmRange = 8192;    // docs per range (32 KB accumulator)
constexpr int kBmThreads = 256;
constexpr int kBmMaxTok = 64;     // tokens applied per chunk (longer queries take several chunks per range)
constexpr int kBmMaxRanges = 8;   // ranges per segment
constexpr int kBmScan = 256;      // head room of the key buffer above k' (a 512-key buffer for k' = 224: five CTAs per SM)
constexpr int kBmUnroll = 8;      // postings per thread in flight while a heavy token streams
constexpr int kBmLight = 8;       // light postings per thread prefetched per range (all light tokens at once)

struct Bm25Params {
  const int64_t* indptr;
  const int32_t* indices;   // doc ids (ascending per token): only the binary searches for range boundaries read them
  const uint2* packed;      // per posting {(doc % kBmRange) * 4 = byte offset in the range accumulator, impact bits}:
                            // ONE 8-byte load per posting and no address arithmetic in the inner loop
  int64_t n_docs;
  int n_seg;              // doc segments per query (work item = query x segment)
  int64_t seg_docs;       // docs per segment (multiple of kBmRange, at most kBmMaxRanges ranges)
  const int32_t* q_tokens;
  const int64_t* q_indptr;
  int nq, nq_pad;
  int keep;               // k'
  int cap;                // shared key buffer size, power of two >= keep + kBmScan
  int all_positive;       // every impact > 0: untouched docs score exactly 0 and can be skipped
  uint64_t* cand;         // [n_seg][nq_pad][keep]
  uint32_t* cand_count;   // [n_seg][nq_pad]
  uint32_t* work_counter;
  unsigned long long* qtau;  // [nq] best published threshold key per query, shared by its work items (zeroed per search)
  float* dump;            // != nullptr: write all scores of query 0 here instead of selecting
};

struct Bm25Smem {
  float acc[kBmRange];
  int64_t offs[kBmMaxTok][kBmMaxRanges + 1];   // posting offset of token t at range boundary r
  uint32_t cnt[kBmMaxTok];                     // postings of token t in the current range
  uint8_t kind[kBmMaxTok];                     // 0 none here, 1 light (prefetched, one posting per thread), 2 streamed
  uint8_t ltok[kBmLight + 1];                  // token index of the u-th light token (ascending), padded with T
  int n_light;
  uint32_t total;
  uint32_t overflow;
  uint32_t count;
  uint64_t tau;
  uint32_t work;
  RadixScratch rs;                             // k'-th largest key of the buffer by a radix walk (compaction without a sort)
};

__device__ __forceinline__ int64_t lower_bound_i32(const int32_t* a, int64_t lo, int64_t hi, int64_t v) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (static_cast<int64_t>(a[mid]) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kBmThreads, 5) bm25_kernel(const Bm25Params p) {
  extern __shared__ uint8_t smem_raw[];
  Bm25Smem* sm = reinterpret_cast<Bm25Smem*>(smem_raw);
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw + ((sizeof(Bm25Smem) + 15) & ~size_t(15)));
  const uint32_t tid = threadIdx.x;
  const uint32_t n_work = static_cast<uint32_t>(p.nq) * p.n_seg;

  // Keep the k' best of the c > k' keys in skeys[0..c), in any order, at the front; returns the k'-th largest key.  A radix
  // walk finds that key and the survivors are moved up in chunks of 2048 positions held in registers (the write cursor never
  // passes the chunk being read) — round 2: the full bitonic sort of the buffer this replaces made the 2048-key buffer 9 %
  // slower than a 1024-key one on the C4 shard.  Sets sm->count = k'.  All threads must call.
  auto compact = [&](uint32_t c) -> uint64_t {
    const uint64_t kth = block_kth_key(skeys, c, static_cast<uint32_t>(p.keep), &sm->rs);
    if (tid == 0) sm->count = 0;
    __syncthreads();
    for (uint32_t base = 0; base < c; base += 8 * kBmThreads) {
      uint64_t mine[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t i = base + tid + j * kBmThreads;
        const uint64_t key = (i < c) ? skeys[i] : 0ull;
        mine[j] = (key >= kth) ? key : 0ull;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (mine[j] != 0ull) skeys[atomicAdd(&sm->count, 1u)] = mine[j];
      __syncthreads();
    }
    return kth;
  };

  for (;;) {
    __syncthreads();
    if (tid == 0) sm->work = atomicAdd(p.work_counter, 1u);
    __syncthreads();
    const uint32_t w = sm->work;
    if (w >= n_work) break;
    // segment-major order: by the time a query's next segment starts, an earlier one has usually
    // published a threshold, so the warm-up (everything admitted, repeated compactions) is paid once
    const int seg = w / p.nq;
    const int q = w % p.nq;
    const int64_t d0 = static_cast<int64_t>(seg) * p.seg_docs;
    const int64_t d1 = min(d0 + p.seg_docs, p.n_docs);
    const int R = static_cast<int>((d1 - d0 + kBmRange - 1) / kBmRange);
    const int64_t tq0 = p.q_indptr[q];
    const int T_all = static_cast<int>(p.q_indptr[q + 1] - tq0);
    // Queries of more than kBmMaxTok tokens are applied in chunks of kBmMaxTok tokens, chunk after chunk inside every
    // range (the per-doc add order stays query-token order); their posting offsets are searched per (range, chunk)
    // instead of once per work item.  bm25s has no token limit (the reference sends paragraph-length rewritten queries).
    const int n_chunks = max(1, (T_all + kBmMaxTok - 1) / kBmMaxTok);
    if (n_chunks == 1) {
      // every (token, range boundary) offset by an independent binary search
      for (int i = tid; i < T_all * (R + 1); i += kBmThreads) {
        const int t = i / (R + 1), r = i % (R + 1);
        const int32_t tok = p.q_tokens[tq0 + t];
        const int64_t bound = min(d0 + static_cast<int64_t>(r) * kBmRange, d1);
        sm->offs[t][r] = lower_bound_i32(p.indices, p.indptr[tok], p.indptr[tok + 1], bound);
      }
    }
    if (tid == 0) { sm->count = 0; sm->tau = (p.qtau != nullptr) ? p.qtau[q] : kKeyNone; }
    for (int i = tid; i < kBmRange; i += kBmThreads) sm->acc[i] = 0.f;
    __syncthreads();

    for (int r = 0; r < R; ++r) {
      const int64_t rs = d0 + static_cast<int64_t>(r) * kBmRange;
      const int64_t re = min(rs + kBmRange, d1);
      const int range_docs = static_cast<int>(re - rs);
      uint8_t* acc_bytes = reinterpret_cast<uint8_t*>(sm->acc);   // packed postings carry byte offsets into it
      bool touched = false;
      for (int tc = 0; tc < n_chunks; ++tc) {
        const int tb = tc * kBmMaxTok;
        const int T = min(kBmMaxTok, T_all - tb);
        if (n_chunks > 1) {
          for (int i = tid; i < T * 2; i += kBmThreads) {
            const int t = i >> 1, e = i & 1;
            const int32_t tok = p.q_tokens[tq0 + tb + t];
            sm->offs[t][r + e] = lower_bound_i32(p.indices, p.indptr[tok], p.indptr[tok + 1], e ? re : rs);
          }
          __syncthreads();
        }
        // Per-range bookkeeping by warp 0 (T <= 64: two tokens per lane): posting counts, light/streamed classification
        // by ballot, the first kBmLight light tokens in token order.  A light token has at most kBmThreads postings in the
        // range: thread i prefetches its i-th posting, all light tokens at once, so their latencies overlap; heavier
        // tokens (and light ones beyond kBmLight) are streamed when their turn comes.
        if (tid < 32) {
          uint32_t sum = 0;
          uint32_t light_before = 0;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int t = half * 32 + static_cast<int>(tid);
            const uint32_t c = (t < T) ? static_cast<uint32_t>(sm->offs[t][r + 1] - sm->offs[t][r]) : 0u;
            const bool light = c > 0 && c <= static_cast<uint32_t>(kBmThreads);
            const uint32_t lmask = __ballot_sync(0xFFFFFFFFu, light);
            const uint32_t ord = light_before + __popc(lmask & ((1u << tid) - 1u));
            if (t < T) {
              sm->cnt[t] = c;
              sm->kind[t] = (c == 0) ? 0 : ((light && ord < static_cast<uint32_t>(kBmLight)) ? 1 : 2);
              if (light && ord < static_cast<uint32_t>(kBmLight)) sm->ltok[ord] = static_cast<uint8_t>(t);
            }
            light_before += __popc(lmask);
            sum += c;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
          if (tid == 0) {
            sm->total = sum;
            sm->n_light = static_cast<int>(min(light_before, static_cast<uint32_t>(kBmLight)));
            sm->overflow = 0;
            if (p.qtau != nullptr && tc == 0) {     // adopt a better threshold published by another segment of this query
              const uint64_t g = *reinterpret_cast<volatile unsigned long long*>(p.qtau + q);
              if (g > sm->tau) sm->tau = g;
            }
          }
        }
        __syncthreads();
        if (sm->total == 0) {                          // no token of this chunk touches the range
          __syncthreads();                             // (everyone has read total before it is rewritten)
          continue;
        }
        touched = true;
        const int n_light = sm->n_light;
        int32_t ldoc[kBmLight];
        float lval[kBmLight];
#pragma unroll
        for (int u = 0; u < kBmLight; ++u) {
          ldoc[u] = -1;
          lval[u] = 0.f;
          if (u < n_light) {
            const int t = sm->ltok[u];
            if (tid < sm->cnt[t]) {
              const uint2 pv = __ldg(p.packed + sm->offs[t][r] + tid);
              ldoc[u] = static_cast<int32_t>(pv.x);
              lval[u] = __uint_as_float(pv.y);
            }
          }
        }
        // Stream one token: coalesced 4-byte loads, kBmUnroll postings per thread in flight; adjacent lanes hold ADJACENT
        // postings, i.e. ascending (for heavy tokens consecutive) docs, so the read-modify-write of the accumulator is
        // free of bank conflicts.  Inside one token every doc occurs once: the adds of different threads never collide.
        auto stream = [&](int t) {
          const uint32_t n = sm->cnt[t];
          const uint2* pp = p.packed + sm->offs[t][r] + tid;
          uint32_t i = tid;
          for (; i + (kBmUnroll - 1) * kBmThreads < n; i += kBmUnroll * kBmThreads) {
            uint2 pv[kBmUnroll];
#pragma unroll
            for (int u = 0; u < kBmUnroll; ++u) pv[u] = __ldg(pp + u * kBmThreads);
#pragma unroll
            for (int u = 0; u < kBmUnroll; ++u) *reinterpret_cast<float*>(acc_bytes + pv[u].x) += __uint_as_float(pv[u].y);
            pp += kBmUnroll * kBmThreads;
          }
          // the remainder in steps of two postings per thread, then one: a list of a few hundred postings in this range costs
          // a few instructions per thread instead of a predicated pass over kBmUnroll slots (ncu, C4 shard: 25 thread
          // instructions per posting overall, most of them predicated-off slots and per-range bookkeeping)
          for (; i + kBmThreads < n; i += 2 * kBmThreads) {
            const uint2 a = __ldg(pp), b = __ldg(pp + kBmThreads);
            *reinterpret_cast<float*>(acc_bytes + a.x) += __uint_as_float(a.y);
            *reinterpret_cast<float*>(acc_bytes + b.x) += __uint_as_float(b.y);
            pp += 2 * kBmThreads;
          }
          if (i < n) {
            const uint2 a = __ldg(pp);
            *reinterpret_cast<float*>(acc_bytes + a.x) += __uint_as_float(a.y);
          }
        };
        // tokens in query order, a block barrier after each (token t fully applied before token t+1 touches the same
        // docs): the prefetched light tokens are walked by a compile-time index, streamed tokens in between
        {
          int t = 0;
#pragma unroll
          for (int u = 0; u <= kBmLight; ++u) {
            const int tl = (u < n_light) ? static_cast<int>(sm->ltok[u < kBmLight ? u : 0]) : T;
            for (; t < tl; ++t) {
              if (sm->kind[t] == 2) { stream(t); __syncthreads(); }
            }
            if (u < kBmLight && u < n_light) {
              if (ldoc[u] >= 0) *reinterpret_cast<float*>(acc_bytes + ldoc[u]) += lval[u];
              __syncthreads();
              t = tl + 1;
            }
          }
        }
      }
      if (!touched && p.all_positive && p.dump == nullptr) continue;   // all-positive impacts: untouched docs score 0

      if (p.dump != nullptr) {
        for (int i = tid; i < kBmRange; i += kBmThreads) {
          if (i < range_docs) p.dump[rs + i] = sm->acc[i];
          sm->acc[i] = 0.f;
        }
      } else {
        // filter the range against the threshold; a full key buffer is compacted and the range rescanned
        for (;;) {
          const uint64_t tau = sm->tau;
          const float tau_f = (tau == kKeyNone) ? -INFINITY : key_score(tau);
          const float4* acc4 = reinterpret_cast<const float4*>(sm->acc);
          for (int i4 = tid; i4 < kBmRange / 4; i4 += kBmThreads) {
            const float4 v = acc4[i4];
            const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
            if (mx >= tau_f && (!p.all_positive || mx > 0.f)) {
              const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = i4 * 4 + e;
                if (i < range_docs && vv[e] >= tau_f && (!p.all_positive || vv[e] > 0.f)) {
                  const uint64_t key = make_key(vv[e], static_cast<uint32_t>(rs + i));
                  if (key > tau) {
                    const uint32_t pos = atomicAdd(&sm->count, 1u);
                    if (pos < static_cast<uint32_t>(p.cap)) skeys[pos] = key;
                    else sm->overflow = 1;
                  }
                }
              }
            }
          }
          __syncthreads();
          const bool over = sm->overflow != 0;
          const uint32_t c = min(sm->count, static_cast<uint32_t>(p.cap));
          __syncthreads();
          if (!over && c + 256 <= static_cast<uint32_t>(p.cap)) break;
          // compact to the k' best (every stored key is a valid candidate, so the k'-th largest stored key is a valid
          // lower bound for the threshold)
          uint64_t kth = 0ull;
          if (c >= static_cast<uint32_t>(p.keep)) kth = (c > static_cast<uint32_t>(p.keep)) ? compact(c) : block_kth_key(skeys, c, c, &sm->rs);
          const uint32_t kept = min(c, static_cast<uint32_t>(p.keep));
          if (!over) {
            if (tid == 0) {
              sm->count = kept;
              if (c >= static_cast<uint32_t>(p.keep)) {
                sm->tau = kth;
                if (p.qtau != nullptr) atomicMax(p.qtau + q, static_cast<unsigned long long>(kth));
              }
            }
            __syncthreads();
            break;
          }
          // overflow: some keys of this range were dropped.  Raise the threshold to just below the k'-th
          // stored key, drop the survivors that belong to this range (the rescan re-admits them) and rescan.
          uint64_t mine[8];                       // kept <= 2048 = 8 per thread
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t i = tid + j * kBmThreads;
            mine[j] = (i < kept) ? skeys[i] : 0ull;
          }
          const uint64_t new_tau = (c >= static_cast<uint32_t>(p.keep)) ? kth - 1ull : sm->tau;
          __syncthreads();
          if (tid == 0) {
            sm->count = 0;
            sm->tau = new_tau;
            sm->overflow = 0;
            if (p.qtau != nullptr) atomicMax(p.qtau + q, static_cast<unsigned long long>(new_tau));
          }
          __syncthreads();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (mine[j] != 0ull) {
              const int64_t id = key_id(mine[j]);
              if (id < rs || id >= re) skeys[atomicAdd(&sm->count, 1u)] = mine[j];
            }
          }
          __syncthreads();
        }
        float4* z4 = reinterpret_cast<float4*>(sm->acc);
        for (int i4 = tid; i4 < kBmRange / 4; i4 += kBmThreads) z4[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
    }

    if (p.dump == nullptr) {
      __syncthreads();
      const uint32_t c = sm->count;
      uint32_t n = c;
      if (c > static_cast<uint32_t>(p.keep)) {         // more than k' survivors: keep the best k' (in any order)
        const uint64_t kth = compact(c);
        n = p.keep;
        if (tid == 0 && p.qtau != nullptr) atomicMax(p.qtau + q, static_cast<unsigned long long>(kth));
      }
      const size_t slot = static_cast<size_t>(seg) * p.nq_pad + q;
      for (uint32_t i = tid; i < n; i += kBmThreads) p.cand[slot * p.keep + i] = skeys[i];
      if (tid == 0) p.cand_count[slot] = n;
    }
  }
}

// When fewer than k docs matched (all-positive impacts), the remaining places go to the lowest ids among the docs
// with score exactly 0.  One warp per query (4 per CTA), k <= 1024: the lanes count the matched places together (rows are
// padded with id -1 after the last match); the missing places can only be ids below k, so the matched ids below k are
// marked in a 1024-bit map (one word per lane) and every lane emits the free ids of its word at the scanned position.
__global__ void __launch_bounds__(128) bm25_zero_fill_kernel(float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                                                             int nq, int k, int64_t n_docs, int64_t id_offset) {
  __shared__ uint32_t bitmap[4][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + w;
  if (q >= nq) return;
  float* sc = out_scores + static_cast<int64_t>(q) * k;
  int64_t* id = out_ids + static_cast<int64_t>(q) * k;
  int matched = 0;
  for (int base = 0; base < k; base += 32) {
    const int i = base + lane;
    const bool has = i < k && id[i] >= 0;
    matched += __popc(__ballot_sync(0xFFFFFFFFu, has));
  }
  if (matched >= k) return;
  bitmap[w][lane] = 0u;
  __syncwarp();
  for (int i = lane; i < matched; i += 32) {
    const int64_t v = id[i] - id_offset;
    if (v >= 0 && v < k) atomicOr(&bitmap[w][v >> 5], 1u << (v & 31));
  }
  __syncwarp();
  const int64_t limit = min(static_cast<int64_t>(k), n_docs);      // candidate ids 0 .. limit-1
  const int64_t first = static_cast<int64_t>(lane) * 32;
  uint32_t valid = 0u;
  if (first < limit) valid = (limit - first >= 32) ? 0xFFFFFFFFu : ((1u << (limit - first)) - 1u);
  uint32_t free_bits = ~bitmap[w][lane] & valid;
  int incl = __popc(free_bits);
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  int pos = matched + incl - __popc(free_bits);
  while (free_bits != 0u && pos < k) {
    const int b = __ffs(free_bits) - 1;
    free_bits &= free_bits - 1u;
    id[pos] = first + b + id_offset;
    sc[pos] = 0.f;
    ++pos;
  }
}

// ------------------------------------------------------------------------------------------
// K3: merge.  scores [g][nq][k_in], ids int64 [g][nq][k_in] (-1 = padding) -> top k_out.
// ------------------------------------------------------------------------------------------
struct MergeSrc {
  const float* scores;
  const int64_t* ids;
  int g;
  int64_t nq;
  int k_in, q;
  template <class F>
  __device__ void for_each(F f) const {
    const int n = g * k_in;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int gi = i / k_in, r = i % k_in;
      const int64_t o = (static_cast<int64_t>(gi) * nq + q) * k_in + r;
      const int64_t id = ids[o];
      if (id >= 0) f(make_key(scores[o], static_cast<uint32_t>(id)));
    }
  }
};

__global__ void __launch_bounds__(256) merge_kernel(const float* __restrict__ scores,
                                                    const int64_t* __restrict__ ids, int g,
                                                    int64_t nq, int k_in, int k_out,
                                                    float* __restrict__ out_scores,
                                                    int64_t* __restrict__ out_ids) {
  extern __shared__ uint8_t smem_raw[];
  SelectSmem* sm = reinterpret_cast<SelectSmem*>(smem_raw);
  const int q = blockIdx.x;
  if (g * k_in <= static_cast<int>(kSortCap)) {     // rows in result order (every searcher here emits them so): rank by search
    for (int i = threadIdx.x; i < g * k_in; i += blockDim.x) {
      const int gi = i / k_in, r = i % k_in;
      const int64_t o = (static_cast<int64_t>(gi) * nq + q) * k_in + r;
      const int64_t id = ids[o];
      sm->keys[i] = (id >= 0) ? make_key(scores[o], static_cast<uint32_t>(id)) : kKeyNone;
    }
    __syncthreads();
    if (merge_sorted_rows(sm->keys, g, k_in, k_out, q, out_scores, out_ids)) return;
  }
  MergeSrc src{scores, ids, g, nq, k_in, q};
  const uint32_t n = block_topk(src, static_cast<uint32_t>(g) * k_in, static_cast<uint32_t>(k_out), sm);
  for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
    const bool has = static_cast<uint32_t>(i) < n;
    const uint64_t key = has ? sm->keys[i] : 0ull;
    out_scores[static_cast<int64_t>(q) * k_out + i] = has ? key_score(key) : -3.402823466e+38f;
    out_ids[static_cast<int64_t>(q) * k_out + i] = has ? static_cast<int64_t>(key_id(key)) : -1;
  }
}

// ------------------------------------------------------------------------------------------
// K5a: reciprocal-rank fusion.  One CTA per query; n_paths*depth <= kSortCap.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rrf_kernel(const int64_t* __restrict__ ids, int n_paths,
                                                  int depth, float k_rrf, int k,
                                                  float* __restrict__ out_scores,
                                                  int64_t* __restrict__ out_ids) {
  extern __shared__ uint8_t smem_raw[];
  SelectSmem* sm = reinterpret_cast<SelectSmem*>(smem_raw);
  uint64_t* keys = sm->keys;
  const int q = blockIdx.x;
  const int n_in = n_paths * depth;
  const uint32_t np = max(next_pow2(static_cast<uint32_t>(n_in)), 2u);
  const int64_t* src = ids + static_cast<int64_t>(q) * n_in;
  // sort entries by (id asc, position asc): descending sort of the complemented key
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
    uint64_t key = 0ull;
    if (i < static_cast<uint32_t>(n_in) && src[i] >= 0)
      key = ~((static_cast<uint64_t>(static_cast<uint32_t>(src[i])) << 32) | i);
    keys[i] = key;
  }
  block_bitonic_desc(keys, np);
  if (threadIdx.x == 0) sm->ctr = 0;
  __syncthreads();
  // run heads sum their run sequentially = in (path, rank) order
  uint64_t fused_key[ (kSortCap + 255) / 256 ];
  int n_mine = 0;
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
    const uint64_t key = keys[i];
    uint64_t out = 0ull;
    if (key != 0ull) {
      const uint64_t v = ~key;
      const uint32_t id = static_cast<uint32_t>(v >> 32);
      const bool head = (i == 0) || (static_cast<uint32_t>((~keys[i - 1]) >> 32) != id) || keys[i - 1] == 0ull;
      if (head) {
        float s = 0.f;
        for (uint32_t r = i; r < np && keys[r] != 0ull && static_cast<uint32_t>((~keys[r]) >> 32) == id; ++r) {
          const uint32_t pos = static_cast<uint32_t>(~keys[r]);
          const float rank = static_cast<float>(pos % depth + 1);
          s = __fadd_rn(s, __fdiv_rn(1.0f, __fadd_rn(k_rrf, rank)));
        }
        out = make_key(s, id);
      }
    }
    fused_key[n_mine++] = out;
  }
  __syncthreads();
  n_mine = 0;
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) keys[i] = fused_key[n_mine++];
  block_bitonic_desc(keys, np);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const uint64_t key = (static_cast<uint32_t>(i) < np) ? keys[i] : 0ull;
    const bool has = key != 0ull;
    out_scores[static_cast<int64_t>(q) * k + i] = has ? key_score(key) : -3.402823466e+38f;
    out_ids[static_cast<int64_t>(q) * k + i] = has ? static_cast<int64_t>(key_id(key)) : -1;
  }
}

// ------------------------------------------------------------------------------------------
// K5b: priority-ordered de-duplicated union (ensembleRetriever.py seen_ids semantics at the id
// level): walk paths in order, ranks in order, keep the first occurrence of every id.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) union_kernel(const int64_t* __restrict__ ids,
                                                    const float* __restrict__ scores, int n_paths,
                                                    int depth, int64_t* __restrict__ out_ids,
                                                    float* __restrict__ out_scores,
                                                    int32_t* __restrict__ out_path,
                                                    int32_t* __restrict__ out_count) {
  extern __shared__ uint8_t smem_raw[];
  SelectSmem* sm = reinterpret_cast<SelectSmem*>(smem_raw);
  uint64_t* keys = sm->keys;
  const int q = blockIdx.x;
  const int n_in = n_paths * depth;
  const uint32_t np = max(next_pow2(static_cast<uint32_t>(n_in)), 2u);
  const int64_t* src = ids + static_cast<int64_t>(q) * n_in;
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
    uint64_t key = 0ull;
    if (i < static_cast<uint32_t>(n_in) && src[i] >= 0)
      key = ~((static_cast<uint64_t>(static_cast<uint32_t>(src[i])) << 32) | i);
    keys[i] = key;
  }
  block_bitonic_desc(keys, np);
  // keep run heads (first occurrence), re-key by position
  uint64_t pos_key[(kSortCap + 255) / 256];
  int n_mine = 0;
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
    const uint64_t key = keys[i];
    uint64_t out = 0ull;
    if (key != 0ull) {
      const uint32_t id = static_cast<uint32_t>((~key) >> 32);
      const bool head = (i == 0) || keys[i - 1] == 0ull || (static_cast<uint32_t>((~keys[i - 1]) >> 32) != id);
      if (head) out = ~static_cast<uint64_t>(static_cast<uint32_t>(~key));  // ~position
    }
    pos_key[n_mine++] = out;
  }
  __syncthreads();
  n_mine = 0;
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) keys[i] = pos_key[n_mine++];
  block_bitonic_desc(keys, np);   // descending ~position = ascending position
  if (threadIdx.x == 0) sm->ctr = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n_in; i += blockDim.x) {
    const uint64_t key = (static_cast<uint32_t>(i) < np) ? keys[i] : 0ull;
    const int64_t o = static_cast<int64_t>(q) * n_in + i;
    if (key != 0ull) {
      const uint32_t pos = static_cast<uint32_t>(~key);
      out_ids[o] = src[pos];
      out_scores[o] = scores[static_cast<int64_t>(q) * n_in + pos];
      out_path[o] = static_cast<int32_t>(pos / depth);
      atomicAdd(&sm->ctr, 1u);
    } else {
      out_ids[o] = -1;
      out_scores[o] = -3.402823466e+38f;
      out_path[o] = -1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out_count[q] = static_cast<int32_t>(sm->ctr);
}

// ------------------------------------------------------------------------------------------
// K5c: the fusion step of the hybrid retriever in ONE launch (one CTA per query):
//   title path   ids are rows of the title-summary corpus: mapped to the chunk each title stands for (the stand-in for
//                the title -> chunks lookup of /root/reference/src/utils/ensembleRetriever.py:143-145); several titles
//                can stand for one chunk: the first (best-ranked) occurrence counts and the ranks behind it close up
//   sparse path  entries with score <= 0 are dropped: with all-positive impacts they are the zero-score filler docs
//                bm25s' top-k pads with when a query matches fewer than `depth` docs, and must not earn 1/(60+rank)
//   then RRF     fused(d) = sum over paths (in path order) of 1/(k_rrf + rank_p(d)) in fp32; top k by (fused desc, id asc).
// Lists are addressed as ids[p * path_stride + q * query_stride + r], so both [B,P,L] and the path-major [P,B,L] layout
// the sharded exchange produces are read in place.  n_paths * depth <= kSortCap.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hybrid_fuse_kernel(const int64_t* __restrict__ ids, const float* __restrict__ scores,
                                                          int n_paths, int depth, int64_t path_stride, int64_t query_stride,
                                                          const int64_t* __restrict__ title_to_chunk, int64_t n_titles,
                                                          int title_path, int sparse_path, float k_rrf, int k,
                                                          float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  extern __shared__ uint8_t smem_raw[];
  SelectSmem* sm = reinterpret_cast<SelectSmem*>(smem_raw);
  uint64_t* keys = sm->keys;
  uint32_t* dropped = sm->rs.hist;                    // bit per list position (kSortCap bits = 128 words)
  const int q = blockIdx.x;
  const int n_in = n_paths * depth;
  const uint32_t np = max(next_pow2(static_cast<uint32_t>(n_in)), 2u);
  for (int i = threadIdx.x; i < 128; i += blockDim.x) dropped[i] = 0;
  // sort entries by (id asc, position asc): descending sort of the complemented key
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
    uint64_t key = 0ull;
    if (i < static_cast<uint32_t>(n_in)) {
      const int p = i / depth, r = i - p * depth;
      const int64_t o = p * path_stride + q * query_stride + r;
      int64_t id = ids[o];
      if (id >= 0 && p == title_path) id = (id < n_titles) ? title_to_chunk[id] : -1;
      if (id >= 0 && p == sparse_path && !(scores[o] > 0.f)) id = -1;
      if (id >= 0) key = ~((static_cast<uint64_t>(static_cast<uint32_t>(id)) << 32) | i);
    }
    keys[i] = key;
  }
  block_bitonic_desc(keys, np);
  // a later entry with the same id in the same path is a duplicate (same-path entries of an id are adjacent)
  for (uint32_t i = threadIdx.x + 1; i < np; i += blockDim.x) {
    const uint64_t a = keys[i - 1], b = keys[i];
    if (b != 0ull && a != 0ull && static_cast<uint32_t>((~a) >> 32) == static_cast<uint32_t>((~b) >> 32)) {
      const uint32_t pa = static_cast<uint32_t>(~a), pb = static_cast<uint32_t>(~b);
      if (pa / depth == pb / depth) atomicOr(&dropped[pb >> 5], 1u << (pb & 31));
    }
  }
  __syncthreads();
  // run heads sum their run sequentially = in (path, rank) order; rank = place among the kept entries of the path
  uint64_t fused_key[(kSortCap + 255) / 256];
  int n_mine = 0;
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
    const uint64_t key = keys[i];
    uint64_t out = 0ull;
    if (key != 0ull) {
      const uint32_t id = static_cast<uint32_t>((~key) >> 32);
      const bool head = (i == 0) || keys[i - 1] == 0ull || (static_cast<uint32_t>((~keys[i - 1]) >> 32) != id);
      if (head) {
        float s = 0.f;
        for (uint32_t r = i; r < np && keys[r] != 0ull && static_cast<uint32_t>((~keys[r]) >> 32) == id; ++r) {
          const uint32_t pos = static_cast<uint32_t>(~keys[r]);
          if (dropped[pos >> 5] & (1u << (pos & 31))) continue;
          const uint32_t p0 = (pos / depth) * depth;           // first position of this entry's path
          uint32_t gone = 0;                                    // dropped entries of the path before this one
          for (uint32_t wd = p0 >> 5; wd <= (pos >> 5); ++wd) {
            uint32_t m = dropped[wd];
            if (wd == (p0 >> 5)) m &= ~0u << (p0 & 31);
            if (wd == (pos >> 5)) m &= (1u << (pos & 31)) - 1u;
            gone += __popc(m);
          }
          const float rank = static_cast<float>(pos - p0 - gone + 1);
          s = __fadd_rn(s, __fdiv_rn(1.0f, __fadd_rn(k_rrf, rank)));
        }
        out = make_key(s, id);
      }
    }
    fused_key[n_mine++] = out;
  }
  __syncthreads();
  n_mine = 0;
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) keys[i] = fused_key[n_mine++];
  block_bitonic_desc(keys, np);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const uint64_t key = (static_cast<uint32_t>(i) < np) ? keys[i] : 0ull;
    const bool has = key != 0ull;
    out_scores[static_cast<int64_t>(q) * k + i] = has ? key_score(key) : -3.402823466e+38f;
    out_ids[static_cast<int64_t>(q) * k + i] = has ? static_cast<int64_t>(key_id(key)) : -1;
  }
}

// ------------------------------------------------------------------------------------------
// posting arrays checked on the device at vfi_bm25_create: flags[0] doc id out of range, flags[1] a posting list not
// strictly ascending, flags[2] an impact <= 0 (or NaN) present
// ------------------------------------------------------------------------------------------
__global__ void bm25_validate_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                     const float* __restrict__ data, int64_t n_vocab, int64_t nnz, int64_t n_docs,
                                     uint32_t* __restrict__ flags) {
  uint32_t bad_range = 0, bad_order = 0, non_pos = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nnz; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t d = indices[i];
    if (d < 0 || d >= n_docs) bad_range = 1;
    if (!(data[i] > 0.f)) non_pos = 1;
    if (i > 0 && indices[i - 1] >= d) {
      // descending neighbours are fine only across a list boundary: is i the first posting of some token?
      int64_t lo = 0, hi = n_vocab;             // largest t with indptr[t] <= i
      while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (indptr[mid] <= i) lo = mid; else hi = mid - 1;
      }
      if (indptr[lo] != i) bad_order = 1;
    }
  }
  if (bad_range) flags[0] = 1;
  if (bad_order) flags[1] = 1;
  if (non_pos) flags[2] = 1;
}

// postings as the scoring kernel reads them (built once at vfi_bm25_create)
__global__ void bm25_pack_kernel(const int32_t* __restrict__ indices, const float* __restrict__ data, int64_t nnz,
                                 uint2* __restrict__ packed) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nnz; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    packed[i] = make_uint2((static_cast<uint32_t>(indices[i]) & static_cast<uint32_t>(kBmRange - 1)) << 2, __float_as_uint(data[i]));
}

// rank-all helpers: (score, id) -> sortable pair, and back
__global__ void rank_keys_kernel(const float* __restrict__ s, int64_t n, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = ~float_orderable(s[i]);
    vals[i] = static_cast<uint32_t>(i);
  }
}
__global__ void ranked_out_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n, int64_t id_offset,
                                  float* __restrict__ scores, int64_t* __restrict__ ids) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const float sc = orderable_float(~keys[i]);
    const int64_t id = static_cast<int64_t>(vals[i]) + id_offset;
    scores[i] = sc;
    ids[i] = id;
  }
}

}  // namespace vfi
