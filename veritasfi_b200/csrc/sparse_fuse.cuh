// sparse_fuse.cuh — the bandwidth-bound kernels of the sparse and fusion paths:
//   K4 bm25_kernel   BM25 scoring over token-major (CSC) postings with a fused top-k'
//   K3 merge_kernel  global top-k after the all-gather of per-shard (score, id) lists
//   K5 rrf_kernel / union_kernel   rank fusion across retrieval paths
//
// K4 restates bm25s.BM25.retrieve called at /root/reference/src/utils/bm25Retriever.py:75-79:
// scores[doc] starts at +0 and receives the precomputed impact of every query token, in query
// token order, in fp32 (np.add.at order).  To reproduce that order without atomics each CTA
// sweeps a doc range tile by tile; inside a tile every active token scatters its impacts into
// its own shared-memory plane (a doc occurs at most once per posting list => one writer per
// cell), and the planes are then summed per doc in token order.  x + 0.0f == x, so tokens that
// do not touch a doc can be skipped without changing a bit.
#pragma once
#include <limits.h>

#include "ptx.cuh"
#include "select.cuh"
#include "topk_common.cuh"

namespace vfi {

constexpr int kBmTile = 2048;     // docs per tile
constexpr int kBmPlanes = 8;      // token planes resident per pass
constexpr int kBmThreads = 512;
constexpr int kBmMaxTok = 64;     // tokens per query handled by the kernel
constexpr int kBmPiece = 256;     // postings per warp work piece
constexpr int kBmScan = 1024;     // docs scanned between buffer-compaction checks

struct Bm25Params {
  const int64_t* indptr;
  const int32_t* indices;
  const float* data;
  int64_t n_docs;
  int n_seg;              // doc segments per query (work item = query x segment)
  int64_t seg_docs;       // docs per segment (multiple of kBmTile)
  const int32_t* q_tokens;
  const int64_t* q_indptr;
  int nq, nq_pad;
  int keep;               // k'
  int cap;                // shared key buffer size, power of two >= keep + kBmScan
  int all_positive;       // every impact > 0: untouched docs score exactly 0 and can be skipped
  uint64_t* cand;         // [n_seg][nq_pad][keep]
  uint32_t* cand_count;   // [n_seg][nq_pad]
  uint32_t* work_counter;
  float* dump;            // != nullptr: write all scores of query 0 here instead of selecting
};

struct Bm25Smem {
  float planes[kBmPlanes][kBmTile];
  float acc[kBmTile];
  int64_t cur[kBmMaxTok];
  int64_t end[kBmMaxTok];
  float dens[kBmMaxTok];
  int next_doc[kBmMaxTok];
  int act[kBmMaxTok];
  int64_t win[kBmPlanes];        // window length per plane token this round
  int piece_start[kBmPlanes + 1];
  uint32_t adv[kBmPlanes];
  int n_act;
  int min_next;
  int more;
  uint32_t count;
  uint64_t tau;
  uint32_t work;
};

__device__ __forceinline__ int64_t lower_bound_i32(const int32_t* a, int64_t lo, int64_t hi, int64_t v) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (static_cast<int64_t>(a[mid]) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kBmThreads, 2) bm25_kernel(const Bm25Params p) {
  extern __shared__ uint8_t smem_raw[];
  Bm25Smem* sm = reinterpret_cast<Bm25Smem*>(smem_raw);
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw + ((sizeof(Bm25Smem) + 15) & ~size_t(15)));
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kWarps = kBmThreads / 32;
  const uint32_t n_work = static_cast<uint32_t>(p.nq) * p.n_seg;

  for (;;) {
    __syncthreads();
    if (tid == 0) sm->work = atomicAdd(p.work_counter, 1u);
    __syncthreads();
    const uint32_t w = sm->work;
    if (w >= n_work) break;
    const int q = w / p.n_seg;
    const int seg = w % p.n_seg;
    const int64_t d0 = static_cast<int64_t>(seg) * p.seg_docs;
    const int64_t d1 = min(d0 + p.seg_docs, p.n_docs);
    const int64_t tq0 = p.q_indptr[q];
    const int T = static_cast<int>(p.q_indptr[q + 1] - tq0);
    // posting range of every token inside this doc segment
    if (tid < T) {
      const int32_t tok = p.q_tokens[tq0 + tid];
      const int64_t lo = p.indptr[tok], hi = p.indptr[tok + 1];
      const int64_t s = lower_bound_i32(p.indices, lo, hi, d0);
      const int64_t e = lower_bound_i32(p.indices, s, hi, d1);
      sm->cur[tid] = s;
      sm->end[tid] = e;
      sm->dens[tid] = static_cast<float>(e - s) / static_cast<float>((d1 - d0) > 0 ? (d1 - d0) : 1);
    }
    if (tid == 0) { sm->count = 0; sm->tau = kKeyNone; }
    __syncthreads();

    int64_t tile_start = d0;
    while (tile_start < d1) {
      // next doc of every token; jump over tiles no token touches
      if (tid == 0) sm->min_next = INT_MAX;
      __syncthreads();
      if (tid < T) {
        const int nd = (sm->cur[tid] < sm->end[tid]) ? p.indices[sm->cur[tid]] : INT_MAX;
        sm->next_doc[tid] = nd;
        if (nd != INT_MAX) atomicMin(&sm->min_next, nd);
      }
      __syncthreads();
      if (p.all_positive && p.dump == nullptr) {
        if (sm->min_next == INT_MAX) break;
        const int64_t jump = (static_cast<int64_t>(sm->min_next) / kBmTile) * kBmTile;
        tile_start = max(tile_start, jump);
      }
      const int64_t tile_end = min(tile_start + kBmTile, d1);
      if (tid == 0) {
        int n = 0;
        for (int t = 0; t < T; ++t)
          if (static_cast<int64_t>(sm->next_doc[t]) < tile_end) sm->act[n++] = t;
        sm->n_act = n;
      }
      for (int i = tid; i < kBmTile; i += kBmThreads) sm->acc[i] = 0.f;
      __syncthreads();
      const int n_act = sm->n_act;

      for (int pass0 = 0; pass0 < n_act; pass0 += kBmPlanes) {
        const int np = min(kBmPlanes, n_act - pass0);
        for (int i = tid; i < np * kBmTile; i += kBmThreads) (&sm->planes[0][0])[i] = 0.f;
        if (tid == 0) sm->more = 1;
        __syncthreads();
        while (sm->more) {
          // size one speculative window per plane token and cut the windows into warp pieces
          if (tid == 0) {
            int ps = 0;
            for (int j = 0; j < np; ++j) {
              const int t = sm->act[pass0 + j];
              const int64_t left = sm->end[t] - sm->cur[t];
              int64_t wlen = 0;
              if (left > 0 && static_cast<int64_t>(p.indices[sm->cur[t]]) < tile_end) {
                const int64_t est = static_cast<int64_t>(sm->dens[t] * static_cast<float>(tile_end - tile_start) * 1.25f) + 64;
                wlen = min(left, est);
              }
              sm->win[j] = wlen;
              sm->piece_start[j] = ps;
              ps += static_cast<int>((wlen + kBmPiece - 1) / kBmPiece);
              sm->adv[j] = 0;
            }
            sm->piece_start[np] = ps;
          }
          __syncthreads();
          const int n_pieces = sm->piece_start[np];
          for (int piece = warp; piece < n_pieces; piece += kWarps) {
            int j = 0;
            while (piece >= sm->piece_start[j + 1]) ++j;
            const int t = sm->act[pass0 + j];
            const int64_t off = static_cast<int64_t>(piece - sm->piece_start[j]) * kBmPiece;
            const int64_t base = sm->cur[t] + off;
            const int n = static_cast<int>(min(static_cast<int64_t>(kBmPiece), sm->win[j] - off));
            int32_t doc[kBmPiece / 32];
            float val[kBmPiece / 32];
#pragma unroll
            for (int i = 0; i < kBmPiece / 32; ++i) {
              const int o = i * 32 + lane;
              doc[i] = (o < n) ? __ldg(p.indices + base + o) : INT_MAX;
              val[i] = (o < n) ? __ldg(p.data + base + o) : 0.f;
            }
            uint32_t cnt = 0;
            float* plane = sm->planes[j];
#pragma unroll
            for (int i = 0; i < kBmPiece / 32; ++i) {
              if (static_cast<int64_t>(doc[i]) < tile_end) {
                plane[doc[i] - tile_start] = val[i];
                ++cnt;
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
            if (lane == 0 && cnt) atomicAdd(&sm->adv[j], cnt);
          }
          __syncthreads();
          if (tid == 0) {
            int more = 0;
            for (int j = 0; j < np; ++j) {
              const int t = sm->act[pass0 + j];
              const bool exhausted_window = (sm->adv[j] == static_cast<uint32_t>(sm->win[j])) && sm->win[j] > 0;
              sm->cur[t] += sm->adv[j];
              if (exhausted_window && sm->cur[t] < sm->end[t]) more = 1;  // may continue in this tile
            }
            sm->more = more;
          }
          __syncthreads();
        }
        // fold the planes into the accumulator in token order
        for (int i = tid; i < kBmTile; i += kBmThreads) {
          float s = sm->acc[i];
          for (int j = 0; j < np; ++j) s += sm->planes[j][i];
          sm->acc[i] = s;
        }
        __syncthreads();
      }

      const int tile_docs = static_cast<int>(tile_end - tile_start);
      if (p.dump != nullptr) {
        for (int i = tid; i < tile_docs; i += kBmThreads) p.dump[tile_start + i] = sm->acc[i];
      } else {
        for (int s0 = 0; s0 < tile_docs; s0 += kBmScan) {
          for (int i = s0 + tid; i < min(s0 + kBmScan, tile_docs); i += kBmThreads) {
            const float v = sm->acc[i];
            if (!p.all_positive || v > 0.f) {
              const uint64_t key = make_key(v, static_cast<uint32_t>(tile_start + i));
              if (key > sm->tau) skeys[atomicAdd(&sm->count, 1u)] = key;
            }
          }
          __syncthreads();
          const uint32_t c = sm->count;
          if (c + kBmScan > static_cast<uint32_t>(p.cap)) {
            for (uint32_t i = c + tid; i < static_cast<uint32_t>(p.cap); i += kBmThreads) skeys[i] = 0ull;
            block_bitonic_desc(skeys, p.cap);
            if (tid == 0) { sm->count = p.keep; sm->tau = skeys[p.keep - 1]; }
            __syncthreads();
          }
        }
      }
      tile_start = tile_end;
    }

    if (p.dump == nullptr) {
      __syncthreads();
      const uint32_t c = sm->count;
      for (uint32_t i = c + tid; i < static_cast<uint32_t>(p.cap); i += kBmThreads) skeys[i] = 0ull;
      block_bitonic_desc(skeys, p.cap);
      const uint32_t n = min(c, static_cast<uint32_t>(p.keep));
      const size_t slot = static_cast<size_t>(seg) * p.nq_pad + q;
      for (uint32_t i = tid; i < n; i += kBmThreads) p.cand[slot * p.keep + i] = skeys[i];
      if (tid == 0) p.cand_count[slot] = n;
    }
  }
}

// When fewer than k docs matched (all-positive impacts), the remaining places go to the lowest
// ids among the docs with score exactly 0.  One thread per query; k is small.
__global__ void bm25_zero_fill_kernel(float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                                      int nq, int k, int64_t n_docs, int64_t id_offset) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  float* sc = out_scores + static_cast<int64_t>(q) * k;
  int64_t* id = out_ids + static_cast<int64_t>(q) * k;
  int n = 0;
  while (n < k && id[n] >= 0) ++n;
  const int matched = n;
  int64_t cand = 0;
  while (n < k && cand < n_docs) {
    bool used = false;
    for (int i = 0; i < matched; ++i) used |= (id[i] - id_offset == cand);
    if (!used) { id[n] = cand + id_offset; sc[n] = 0.f; ++n; }
    ++cand;
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

}  // namespace vfi
