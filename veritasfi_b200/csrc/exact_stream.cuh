// exact_stream.cuh — the exact streaming scorer: canonical scores of EVERY row of the shard for a small group of
// queries in one pass over the corpus, then an exact top-k over those scores.  No tensor cores, no candidates, no
// certificate: the scores are already the parity contract's (sequential fp64 fused multiply-add over j = 0..d-1,
// rounded once to fp32; oracle/vfi_oracle.c:vfo_canon_dot).
//
// This is the path of the reference's own online call — faiss.IndexFlatIP.search with k = 2048 for 1..4 query
// strings (/root/reference/src/utils/ensembleRetriever.py:64-66 -> src/utils/faissRetriever.py:37) — whose k' exceeds
// what the candidate buffers of K1/K1b hold, of tiny shards, and of the repair of queries whose certificate failed.
//
//   exact_scores_kernel   one thread per row, NQ independent fp64 chains per thread (one per query of the group), the
//                         row fetched in 256-byte pieces by one bulk copy per lane into that lane's own padded
//                         shared-memory row (the K2 scheme, dense_support.cuh); the queries sit in shared memory as
//                         fp64.  Bound: HBM (the corpus is read once per group of up to 8 queries); per row the SM
//                         spends d/16 cycles on F2F.F64 conversions and NQ*d/64 cycles on DFMA.
//                         (A two-stage ring of 128-byte pieces — the next piece in flight while one is summed — was measured in
//                         round 2: 1.04 ms instead of 0.75 ms for one query over 1M x 1024 fp32 rows.  The size of the
//                         contiguous piece per request matters more than the overlap, as the gather micro-benchmark of
//                         round 1 had found for random rows; one 256-byte stage per warp stays.)
//   radix_hist_kernel     multi-CTA MSD radix select over the 64-bit keys (score, ~id) of the score array:
//                         11+11+10 bits of the score, then 11+11+10 bits of the id; the last CTA of a query to
//                         finish a pass picks the digit (ticket counter), so a pass is one launch.
//   radix_gather_kernel   keys >= the threshold key (exactly k of them) are collected and the last CTA of the
//                         query sorts them and writes the result rows.
//   exact_small_kernel    shards up to 32768 rows: one CTA per query selects straight from the score array.
#pragma once
#include "ptx.cuh"
#include "select.cuh"
#include "topk_common.cuh"

namespace vfi {

constexpr int kExThreads = 512;                 // 16 warps, one row per thread per step (the base configuration)
constexpr int kExMaxQ = 8;                      // queries per pass over the corpus
constexpr int kExPiece = 256;                   // bytes per bulk copy
constexpr int kExPitch = kExPiece + 16;         // conflict-free 128-bit reads of 32 different rows
constexpr int kExTileBytes = (kExThreads / 32) * 32 * kExPitch;
constexpr int kExSmemBudget = 227 * 1024 - 1024;
__host__ __device__ inline size_t exact_smem_bytes(int dp, int nq, int threads = kExThreads) {
  return static_cast<size_t>(nq) * dp * 8 + (threads / 32) * 8 + static_cast<size_t>(threads / 32) * 32 * kExPitch;
}
// largest query group whose fp64 copies fit beside the row tiles
inline int exact_max_group(int dp) {
  const int64_t room = kExSmemBudget - kExTileBytes - 256;
  const int g = static_cast<int>(room / (static_cast<int64_t>(dp) * 8));
  return g < 1 ? 0 : (g > kExMaxQ ? kExMaxQ : g);
}
// The kernel is bound by how many rows an SM has in flight (one 272-byte stage per row, 16 warps: fp64 / XU / LSU pipes each
// ~30 % busy, ncu).  Small groups need few registers and little shared memory for their queries, so they run with more
// warps per SM when everything still fits: 24 warps for 1-2 queries, 20 for 3-4.
__host__ __device__ constexpr int exact_pref_threads(int nq) { return nq <= 2 ? 768 : (nq <= 4 ? 640 : 512); }

template <typename RowT, int NQ, int THREADS = kExThreads>
__global__ void __launch_bounds__(THREADS, 1) exact_scores_kernel(
    const RowT* __restrict__ rows, int64_t row_pitch, int dp, int64_t n_rows, const float* __restrict__ qcanon,
    const int* __restrict__ qsel /* [NQ] rows of qcanon, or nullptr = q0 .. q0+NQ-1 */, int q0,
    float* __restrict__ scores_out /* [NQ][ld] */, int64_t ld) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  double* qd = reinterpret_cast<double*>(smem_raw);                       // [NQ][dp]
  uint64_t* bars = reinterpret_cast<uint64_t*>(qd + static_cast<size_t>(NQ) * dp);   // one per warp
  uint8_t* tiles = reinterpret_cast<uint8_t*>(bars + THREADS / 32);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int j = 0; j < NQ; ++j) {
    const int q = (qsel != nullptr) ? qsel[q0 + j] : q0 + j;
    const float* qv = qcanon + static_cast<int64_t>(q) * dp;
    for (int e = tid; e < dp; e += THREADS) qd[j * dp + e] = static_cast<double>(qv[e]);
  }
  uint64_t* bar = &bars[warp];
  if (lane == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int row_len = dp * static_cast<int>(sizeof(RowT));
  const int n_pieces = (row_len + kExPiece - 1) / kExPiece;
  constexpr int kPieceElems = kExPiece / static_cast<int>(sizeof(RowT));
  uint8_t* my_tile = tiles + (static_cast<size_t>(warp) * 32 + lane) * kExPitch;
  const int64_t n_groups = (n_rows + 31) / 32;
  uint32_t phase = 0;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * (THREADS / 32) + warp; g < n_groups;
       g += static_cast<int64_t>(gridDim.x) * (THREADS / 32)) {
    const int64_t row = g * 32 + lane;
    const bool valid = row < n_rows;
    const uint32_t n_valid = __popc(__ballot_sync(0xFFFFFFFFu, valid));
    const uint8_t* my_row = reinterpret_cast<const uint8_t*>(rows) + row * row_pitch * static_cast<int64_t>(sizeof(RowT));
    double acc[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) acc[j] = 0.0;
    for (int c = 0; c < n_pieces; ++c) {
      const uint32_t bytes = static_cast<uint32_t>(min(kExPiece, row_len - c * kExPiece));
      if (lane == 0) ptx::mbar_expect_tx(bar, bytes * n_valid);
      __syncwarp();
      if (valid) ptx::bulk_copy_g2s(my_tile, my_row + static_cast<int64_t>(c) * kExPiece, bytes, bar);
      ptx::mbar_wait(bar, phase);
      phase ^= 1u;
      const uint4* mine = reinterpret_cast<const uint4*>(my_tile);
      const int nvec = static_cast<int>(bytes) / 16;
      const double* qbase = qd + c * kPieceElems;
#pragma unroll 2
      for (int i = 0; i < nvec; ++i) {
        const uint4 u = mine[i];
        if (sizeof(RowT) == 2) {
          double x[8];
          x[0] = static_cast<double>(__uint_as_float(u.x << 16));
          x[1] = static_cast<double>(__uint_as_float(u.x & 0xFFFF0000u));
          x[2] = static_cast<double>(__uint_as_float(u.y << 16));
          x[3] = static_cast<double>(__uint_as_float(u.y & 0xFFFF0000u));
          x[4] = static_cast<double>(__uint_as_float(u.z << 16));
          x[5] = static_cast<double>(__uint_as_float(u.z & 0xFFFF0000u));
          x[6] = static_cast<double>(__uint_as_float(u.w << 16));
          x[7] = static_cast<double>(__uint_as_float(u.w & 0xFFFF0000u));
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            const double2* qq = reinterpret_cast<const double2*>(qbase + j * dp + 8 * i);
            const double2 q0v = qq[0], q1v = qq[1], q2v = qq[2], q3v = qq[3];
            acc[j] = fma(q0v.x, x[0], acc[j]); acc[j] = fma(q0v.y, x[1], acc[j]);
            acc[j] = fma(q1v.x, x[2], acc[j]); acc[j] = fma(q1v.y, x[3], acc[j]);
            acc[j] = fma(q2v.x, x[4], acc[j]); acc[j] = fma(q2v.y, x[5], acc[j]);
            acc[j] = fma(q3v.x, x[6], acc[j]); acc[j] = fma(q3v.y, x[7], acc[j]);
          }
        } else {
          double x[4];
          x[0] = static_cast<double>(__uint_as_float(u.x));
          x[1] = static_cast<double>(__uint_as_float(u.y));
          x[2] = static_cast<double>(__uint_as_float(u.z));
          x[3] = static_cast<double>(__uint_as_float(u.w));
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            const double2* qq = reinterpret_cast<const double2*>(qbase + j * dp + 4 * i);
            const double2 q0v = qq[0], q1v = qq[1];
            acc[j] = fma(q0v.x, x[0], acc[j]); acc[j] = fma(q0v.y, x[1], acc[j]);
            acc[j] = fma(q1v.x, x[2], acc[j]); acc[j] = fma(q1v.y, x[3], acc[j]);
          }
        }
      }
      __syncwarp();                                   // every lane is done with the tile before it is refilled
    }
    if (valid) {
#pragma unroll
      for (int j = 0; j < NQ; ++j) scores_out[static_cast<int64_t>(j) * ld + row] = static_cast<float>(acc[j]);
    }
  }
}

// ---- exact top-k over a score array --------------------------------------------------------------------------
struct ScoreKeySrc {
  const float* s;
  int64_t n;
  template <class F>
  __device__ void for_each(F f) const {
    const int64_t stride = blockDim.x;
    int64_t i = threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = s[i + j * stride];
#pragma unroll
      for (int j = 0; j < 4; ++j) f(make_key(v[j], static_cast<uint32_t>(i + j * stride)));
    }
    for (; i < n; i += stride) f(make_key(s[i], static_cast<uint32_t>(i)));
  }
};

// write one result row: keys[0..n) sorted descending -> out_scores/out_ids[row][0..k), padded like faiss
__device__ __forceinline__ void emit_result_row(const uint64_t* keys, uint32_t n, int k, int64_t id_offset, int64_t row,
                                                float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const bool has = static_cast<uint32_t>(i) < n;
    const uint64_t key = has ? keys[i] : 0ull;
    out_scores[row * k + i] = has ? key_score(key) : -3.402823466e+38f;
    out_ids[row * k + i] = has ? static_cast<int64_t>(key_id(key)) + id_offset : -1;
  }
}

// small shards: one CTA per query of the group
__global__ void __launch_bounds__(256) exact_small_kernel(const float* __restrict__ scores, int64_t ld, int64_t n, int k,
                                                          const int* __restrict__ qsel, int q0, int64_t id_offset,
                                                          float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  extern __shared__ uint8_t smem_raw[];
  SelectSmem* sm = reinterpret_cast<SelectSmem*>(smem_raw);
  const int j = blockIdx.x;
  const int q = (qsel != nullptr) ? qsel[q0 + j] : q0 + j;
  ScoreKeySrc src{scores + static_cast<int64_t>(j) * ld, n};
  const uint32_t got = block_topk(src, static_cast<uint32_t>(n), static_cast<uint32_t>(k), sm);
  emit_result_row(sm->keys, got, k, id_offset, q, out_scores, out_ids);
}

// ---- multi-CTA radix select -----------------------------------------------------------------------------------
constexpr int kRxBins = 2048;
constexpr int kRxPasses = 6;
struct RadixState {              // one per query of the group, zeroed before the first pass
  unsigned long long prefix;     // bits of the threshold key fixed so far
  uint32_t remaining;            // rank of the threshold among the keys matching the prefix
  uint32_t done;                 // threshold final: keys >= prefix are exactly the k best
  uint32_t ticket;               // CTAs of this query that finished the current pass
  uint32_t gathered;             // keys collected by the gather pass
  uint32_t gather_ticket;
  uint32_t pad;
};
__host__ __device__ inline int radix_shift(int pass) {
  return pass == 0 ? 53 : pass == 1 ? 42 : pass == 2 ? 32 : pass == 3 ? 21 : pass == 4 ? 10 : 0;
}
__host__ __device__ inline int radix_bits(int pass) { return (pass == 2 || pass == 5) ? 10 : 11; }

// grid (ctas_per_query, nq_group).  hist [nq_group][kRxBins] zeroed before pass 0 (the picker re-zeroes it).
__global__ void __launch_bounds__(256) radix_hist_kernel(const float* __restrict__ scores, int64_t ld, int64_t n, int k, int pass,
                                                         RadixState* __restrict__ state, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kRxBins];
  __shared__ uint32_t s_last;
  const int j = blockIdx.y;
  RadixState* st = state + j;
  if (pass > 0 && *reinterpret_cast<volatile uint32_t*>(&st->done)) return;   // uniform per CTA: written by an earlier launch
  const unsigned long long prefix = (pass == 0) ? 0ull : st->prefix;
  const int shift = radix_shift(pass), bits = radix_bits(pass);
  const uint32_t mask = (1u << bits) - 1u;
  const int hi = shift + bits;                                    // bits above this digit must equal the prefix
  for (int i = threadIdx.x; i < kRxBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const float* s = scores + static_cast<int64_t>(j) * ld;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = per * blockIdx.x, hi_row = min(n, lo + per);
  for (int64_t i = lo + threadIdx.x; i < hi_row; i += blockDim.x) {
    const uint64_t key = make_key(s[i], static_cast<uint32_t>(i));
    if (hi >= 64 || (key >> hi) == (prefix >> hi)) atomicAdd(&sh[(key >> shift) & mask], 1u);
  }
  __syncthreads();
  uint32_t* gh = hist + static_cast<size_t>(j) * kRxBins;
  for (int i = threadIdx.x; i < kRxBins; i += blockDim.x)
    if (sh[i]) atomicAdd(&gh[i], sh[i]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  // the last CTA of this query picks the digit: walk the bins from the top
  __threadfence();
  for (int i = threadIdx.x; i < kRxBins; i += blockDim.x) sh[i] = __ldcg(&gh[i]);
  __syncthreads();
  if (threadIdx.x < 32) {
    const uint32_t lane = threadIdx.x;
    const uint32_t remaining = (pass == 0) ? static_cast<uint32_t>(k) : st->remaining;
    constexpr int kPer = kRxBins / 32;                             // 64 bins per lane, lane 31 owns the top ones
    uint32_t mine = 0;
    for (int i = 0; i < kPer; ++i) mine += sh[lane * kPer + i];
    uint32_t incl = mine;                                          // suffix sum over lanes (higher lanes = larger keys)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_down_sync(0xFFFFFFFFu, incl, o);
      if (lane + o < 32) incl += t;
    }
    const uint32_t above = incl - mine;
    if (above < remaining && remaining <= above + mine) {
      uint32_t run = above;
      for (int i = kPer - 1; i >= 0; --i) {
        const uint32_t c = sh[lane * kPer + i];
        if (run < remaining && remaining <= run + c) {
          const unsigned long long np = prefix | (static_cast<unsigned long long>(lane * kPer + i) << shift);
          st->prefix = np;
          st->remaining = remaining - run;
          // every key of the bin is kept, or no bits are left: "key >= prefix" selects exactly the k best
          st->done = (c == remaining - run || pass == kRxPasses - 1) ? 1u : 0u;
          break;
        }
        run += c;
      }
    }
    if (lane == 0) st->ticket = 0;
  }
  for (int i = threadIdx.x; i < kRxBins; i += blockDim.x) gh[i] = 0;
}

// grid (ctas_per_query, nq_group).  keys [nq_group][k_cap] scratch.  The last CTA of a query sorts and emits.
__global__ void __launch_bounds__(256) radix_gather_kernel(const float* __restrict__ scores, int64_t ld, int64_t n, int k, int k_cap,
                                                           RadixState* __restrict__ state, uint64_t* __restrict__ keys,
                                                           const int* __restrict__ qsel, int q0, int64_t id_offset,
                                                           float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  extern __shared__ uint8_t smem_raw[];
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw);          // [next_pow2(k)]
  __shared__ uint32_t s_last;
  const int j = blockIdx.y;
  RadixState* st = state + j;
  const unsigned long long thr = st->prefix;
  const float* s = scores + static_cast<int64_t>(j) * ld;
  uint64_t* kq = keys + static_cast<size_t>(j) * k_cap;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = per * blockIdx.x, hi_row = min(n, lo + per);
  for (int64_t i = lo + threadIdx.x; i < hi_row; i += blockDim.x) {
    const uint64_t key = make_key(s[i], static_cast<uint32_t>(i));
    if (key >= thr) {
      const uint32_t pos = atomicAdd(&st->gathered, 1u);
      if (pos < static_cast<uint32_t>(k_cap)) kq[pos] = key;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->gather_ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const uint32_t got = min(*reinterpret_cast<volatile uint32_t*>(&st->gathered), static_cast<uint32_t>(k_cap));
  const uint32_t np = max(next_pow2(got), 2u);
  for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) skeys[i] = (i < got) ? __ldcg(&kq[i]) : 0ull;
  block_bitonic_desc(skeys, np);
  const int q = (qsel != nullptr) ? qsel[q0 + j] : q0 + j;
  emit_result_row(skeys, min(got, static_cast<uint32_t>(k)), k, id_offset, q, out_scores, out_ids);
}

}  // namespace vfi
