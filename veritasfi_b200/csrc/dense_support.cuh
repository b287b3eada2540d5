// dense_support.cuh — the bandwidth-bound kernels around the tensor-core pass:
//   K6  prep_rows / prep_queries / normalize_l2   (cast, 3-term bf16 split, norms, epsilon)
//   K1c cand_reduce                               (per-query union of the group buffers -> k' best)
//   K2  canon_score + finalize                    (exact rescoring in the canonical order, final
//                                                  (score desc, id asc) order, exactness certificate)
//   K1b gemv_topk                                 (small query batches: pure HBM stream, no tensor cores)
//
// Canonical score (the parity contract, oracle/vfi_oracle.c:vfo_canon_dot): sequential fp64
// fused multiply-add over j = 0..d-1, rounded once to fp32.  Products of two fp32 values are exact
// in fp64, so fma and mul+add agree and the host restatement is bit-identical.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"
#include "select.cuh"
#include "topk_common.cuh"

namespace vfi {

__device__ __forceinline__ uint16_t f32_to_bf16_rn(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ float bf16_to_f32(uint16_t h) {
  return __uint_as_float(static_cast<uint32_t>(h) << 16);
}

// butterfly (xor 16,8,4,2,1) sum in fp64 — the fixed order the oracle restates for norms
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// K6a: rows fp32 [n][d] -> gemm operand rows bf16 [n][kp] (+ fp32 master [n][dp] in F32 mode).
// BF16 mode: kp = dp, row = bf16(x).  F32 mode: kp = 3*dp, row = [hi | lo | hi].
// One warp per row.  xnorm_max: running max of the row L2 norms (of the stored values).
// ------------------------------------------------------------------------------------------
template <bool SPLIT, typename InT>
__global__ void prep_rows_kernel(const InT* __restrict__ x, int64_t n, int d, int dp,
                                 uint16_t* __restrict__ g, int64_t kp, float* __restrict__ master,
                                 uint32_t* __restrict__ xnorm_max_bits) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (row >= n) return;
  const InT* xr = x + row * d;
  uint16_t* gr = g + row * kp;
  double ss = 0.0;
  for (int j = lane; j < dp; j += 32) {
    float v = 0.f;
    if (j < d) {
      if (sizeof(InT) == 2) v = bf16_to_f32(static_cast<uint16_t>(xr[j]));
      else v = static_cast<float>(xr[j]);
    }
    const uint16_t hi = f32_to_bf16_rn(v);
    if (SPLIT) {
      const float lo_f = v - bf16_to_f32(hi);
      const uint16_t lo = f32_to_bf16_rn(lo_f);
      gr[j] = hi;
      gr[dp + j] = lo;
      gr[2 * dp + j] = hi;
      master[row * dp + j] = v;
      ss += static_cast<double>(v) * static_cast<double>(v);
    } else {
      gr[j] = hi;
      const float s = bf16_to_f32(hi);
      ss += static_cast<double>(s) * static_cast<double>(s);
    }
  }
  ss = warp_sum_f64(ss);
  if (lane == 0) atomicMax(xnorm_max_bits, __float_as_uint(static_cast<float>(sqrt(ss)) * 1.0000002f));
}

// ------------------------------------------------------------------------------------------
// K6b: queries fp32 [nq][d] -> canonical fp32 [nq][dp], gemm operand bf16 [nq_pad][kp], eps[nq].
// BF16 mode: canon = bf16(q) as fp32, gemm = bf16(q).  F32: canon = q, gemm = [hi | hi | lo].
// eps = acc_c * kp * 2^-24 * |q| * xnorm_max  (+ split_c * |q| * xnorm_max in F32 mode)
// ------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void prep_queries_kernel(const float* __restrict__ q, int nq, int d, int dp,
                                    float* __restrict__ canon, uint16_t* __restrict__ g, int64_t kp,
                                    const uint32_t* __restrict__ xnorm_max_bits,
                                    float* __restrict__ eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (row >= nq) return;
  const float* qr = q + static_cast<int64_t>(row) * d;
  uint16_t* gr = g + static_cast<int64_t>(row) * kp;
  float* cr = canon + static_cast<int64_t>(row) * dp;
  double ss = 0.0;
  for (int j = lane; j < dp; j += 32) {
    const float v = (j < d) ? qr[j] : 0.f;
    const uint16_t hi = f32_to_bf16_rn(v);
    if (SPLIT) {
      const uint16_t lo = f32_to_bf16_rn(v - bf16_to_f32(hi));
      gr[j] = hi;
      gr[dp + j] = hi;
      gr[2 * dp + j] = lo;
      cr[j] = v;
      ss += static_cast<double>(v) * static_cast<double>(v);
    } else {
      gr[j] = hi;
      const float s = bf16_to_f32(hi);
      cr[j] = s;
      ss += static_cast<double>(s) * static_cast<double>(s);
    }
  }
  ss = warp_sum_f64(ss);
  if (lane == 0) {
    const float qn = static_cast<float>(sqrt(ss)) * 1.0000002f;
    const float xn = __uint_as_float(*xnorm_max_bits);
    float e = 2.0f * static_cast<float>(kp) * 5.9604645e-8f * qn * xn;
    if (SPLIT) e += 2.0e-5f * qn * xn;
    eps[row] = e;
  }
}

// ------------------------------------------------------------------------------------------
// K6c: faiss.normalize_L2 — x[i,:] *= 1/sqrtf(sum x^2), zero rows untouched.  One warp per row;
// sum of squares: lane-strided sequential fp64 partials, butterfly-combined, rounded to fp32.
// ------------------------------------------------------------------------------------------
__global__ void normalize_l2_kernel(float* __restrict__ x, int64_t n, int d) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (row >= n) return;
  float* xr = x + row * d;
  double ss = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double v = static_cast<double>(xr[j]);
    ss = fma(v, v, ss);
  }
  ss = warp_sum_f64(ss);
  const float nrm2 = static_cast<float>(ss);
  if (nrm2 > 0.f) {
    const float inv = 1.0f / sqrtf(nrm2);
    for (int j = lane; j < d; j += 32) xr[j] = xr[j] * inv;
  }
}

// ------------------------------------------------------------------------------------------
// K1c: union of the per-group candidate buffers of one query -> the k' best keys, sorted.
// bound[q] = an upper bound on the tensor-core score of every row that is NOT in the output:
//   the k'-th key's score when at least k' rows were admitted, else the admission hint (rows at
//   or below the hint were never admitted), else -inf (every row of the shard is a candidate).
// ------------------------------------------------------------------------------------------
struct GroupBufSrc {
  const uint64_t* cand;
  const uint32_t* cnt;
  int n_groups, nq_pad, cap, q;
  template <class F>
  __device__ void for_each(F f) const {
    for (int g = 0; g < n_groups; ++g) {
      const size_t slot = static_cast<size_t>(g) * nq_pad + q;
      const uint32_t c = min(cnt[slot], static_cast<uint32_t>(cap));
      const uint64_t* b = cand + slot * cap;
      for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) f(b[i]);
    }
  }
};

__global__ void __launch_bounds__(256) cand_reduce_kernel(const uint64_t* __restrict__ cand,
                                                          const uint32_t* __restrict__ cnt,
                                                          int n_groups, int nq_pad, int cap,
                                                          int keep, const float* __restrict__ tau_init,
                                                          uint64_t* __restrict__ out_keys,
                                                          uint32_t* __restrict__ out_n,
                                                          float* __restrict__ bound) {
  extern __shared__ uint8_t smem_raw[];
  SelectSmem* sm = reinterpret_cast<SelectSmem*>(smem_raw);
  const int q = blockIdx.x;
  __shared__ uint32_t s_total;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  uint32_t part = 0;
  for (int g = threadIdx.x; g < n_groups; g += blockDim.x)
    part += min(cnt[static_cast<size_t>(g) * nq_pad + q], static_cast<uint32_t>(cap));
  if (part) atomicAdd(&s_total, part);
  __syncthreads();
  const uint32_t total = s_total;
  GroupBufSrc src{cand, cnt, n_groups, nq_pad, cap, q};
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

// tau[q] = the m-th best score among the sampled rows' raw tensor-core scores (one CTA per query)
struct ScoreRowSrc {
  const float* s;
  int n;
  template <class F>
  __device__ void for_each(F f) const {
    for (int i = threadIdx.x; i < n; i += blockDim.x) f(make_key(s[i], static_cast<uint32_t>(i)));
  }
};
__global__ void __launch_bounds__(256) tau_from_scores_kernel(const float* __restrict__ scores, int64_t ld, int n_valid,
                                                              int m, float* __restrict__ tau, int debug_inf) {
  extern __shared__ uint8_t smem_raw[];
  SelectSmem* sm = reinterpret_cast<SelectSmem*>(smem_raw);
  const int q = blockIdx.x;
  ScoreRowSrc src{scores + static_cast<int64_t>(q) * ld, n_valid};
  const uint32_t n = block_topk(src, static_cast<uint32_t>(n_valid), static_cast<uint32_t>(m), sm);
  if (threadIdx.x == 0) {
    float t = (n >= static_cast<uint32_t>(m)) ? key_score(sm->keys[m - 1]) : -INFINITY;
    if (debug_inf) t = INFINITY;
    tau[q] = t;
  }
}

// ------------------------------------------------------------------------------------------
// K2a: canonical scores.  One warp = 32 (query, row) pairs of the same query, lane = row.
// The 32 rows are staged through shared memory in 64-element chunks with coalesced 16-byte
// loads, then every lane walks its own row sequentially in fp64.
//   cand_keys != nullptr : rows are the ids in cand_keys[q][0..n_cand[q])   (rescoring)
//   cand_keys == nullptr : rows are all_lo .. all_lo+all_n of the shard       (exhaustive)
// Output keys2[q][i] = make_key(exact fp32 score, id); pairs beyond the count get kKeyNone.
// ------------------------------------------------------------------------------------------
template <typename RowT>
__global__ void __launch_bounds__(128) canon_score_kernel(
    const RowT* __restrict__ rows, int64_t row_pitch, int dp, const float* __restrict__ qcanon,
    const int* __restrict__ qsel, const uint64_t* __restrict__ cand_keys,
    const uint32_t* __restrict__ n_cand, int keep, int64_t all_n, uint64_t* __restrict__ keys2,
    int64_t keys2_pitch, uint32_t* __restrict__ max_err_bits) {
  constexpr int kWarps = 4;
  constexpr int kWordsPerRow = (sizeof(RowT) == 2) ? 32 : 64;  // 64 elements per chunk
  constexpr int kPitch = kWordsPerRow + 1;
  __shared__ uint32_t tile[kWarps][32 * kPitch];
  __shared__ float qs[kWarps][64];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qslot = blockIdx.y;                       // index into the selected-query list
  const int q = (qsel != nullptr) ? qsel[qslot] : qslot;
  const int64_t grp = static_cast<int64_t>(blockIdx.x) * kWarps + warp;  // group of 32 rows
  const int64_t first = grp * 32;
  const int64_t limit = (cand_keys != nullptr) ? static_cast<int64_t>(n_cand[q]) : all_n;
  const int64_t slots = (cand_keys != nullptr) ? keep : all_n;
  if (first >= slots) return;
  // row id of this lane
  const int64_t slot = first + lane;
  uint32_t id = 0;
  float gemm_score = 0.f;
  bool valid = slot < limit;
  if (valid) {
    if (cand_keys != nullptr) {
      const uint64_t k = cand_keys[static_cast<int64_t>(q) * keep + slot];
      id = key_id(k);
      gemm_score = key_score(k);
    } else {
      id = static_cast<uint32_t>(slot);
    }
  }
  const float* qv = qcanon + static_cast<int64_t>(q) * dp;
  double acc = 0.0;
  uint32_t* my_tile = tile[warp];
  for (int c0 = 0; c0 < dp; c0 += 64) {
    // stage: every 16-byte piece of the 32 row chunks, coalesced per row
    constexpr int kVecPerRow = kWordsPerRow / 4;           // uint4 per row chunk (8 or 16)
    constexpr int kRowsPerLoad = 32 / kVecPerRow;          // rows covered by one warp load (4 or 2)
#pragma unroll
    for (int it = 0; it < 32 / kRowsPerLoad; ++it) {
      const int r = it * kRowsPerLoad + lane / kVecPerRow;
      const int v = lane % kVecPerRow;
      const uint32_t rid = __shfl_sync(0xFFFFFFFFu, id, r);
      const bool rvalid = __shfl_sync(0xFFFFFFFFu, valid ? 1 : 0, r) != 0;
      uint4 w = make_uint4(0, 0, 0, 0);
      if (rvalid) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(rows + static_cast<int64_t>(rid) * row_pitch + c0);
        w = ptx::ld_nc_u4(src + v * 16);
      }
      uint32_t* dst = my_tile + r * kPitch + v * 4;
      dst[0] = w.x; dst[1] = w.y; dst[2] = w.z; dst[3] = w.w;
    }
    qs[warp][lane] = qv[c0 + lane];
    qs[warp][lane + 32] = qv[c0 + lane + 32];
    __syncwarp();
    const uint32_t* mine = my_tile + lane * kPitch;
    if (sizeof(RowT) == 2) {
#pragma unroll 8
      for (int w = 0; w < 32; ++w) {
        const uint32_t u = mine[w];
        const float x0 = __uint_as_float(u << 16);
        const float x1 = __uint_as_float(u & 0xFFFF0000u);
        acc = fma(static_cast<double>(qs[warp][2 * w]), static_cast<double>(x0), acc);
        acc = fma(static_cast<double>(qs[warp][2 * w + 1]), static_cast<double>(x1), acc);
      }
    } else {
#pragma unroll 8
      for (int w = 0; w < 64; ++w) {
        acc = fma(static_cast<double>(qs[warp][w]), static_cast<double>(__uint_as_float(mine[w])), acc);
      }
    }
    __syncwarp();
  }
  const float s = static_cast<float>(acc);
  if (slot < slots) {
    keys2[static_cast<int64_t>(qslot) * keys2_pitch + slot] = valid ? make_key(s, id) : kKeyNone;
  }
  if (valid && cand_keys != nullptr && max_err_bits != nullptr) {
    const float e = fabsf(s - gemm_score);
    atomicMax(max_err_bits, __float_as_uint(e));
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

__global__ void __launch_bounds__(256) finalize_kernel(
    const uint64_t* __restrict__ keys2, int64_t keys2_pitch, int64_t n_slots,
    const int* __restrict__ qsel, const uint32_t* __restrict__ n_cand, int k, int64_t id_offset,
    const float* __restrict__ bound, const float* __restrict__ eps, int check,
    float* __restrict__ out_scores, int64_t* __restrict__ out_ids, int* __restrict__ flagged,
    int* __restrict__ n_flagged) {
  extern __shared__ uint8_t smem_raw[];
  SelectSmem* sm = reinterpret_cast<SelectSmem*>(smem_raw);
  const int qslot = blockIdx.x;
  const int q = (qsel != nullptr) ? qsel[qslot] : qslot;
  const uint32_t total = (n_cand != nullptr) ? n_cand[q] : static_cast<uint32_t>(n_slots);
  KeyArraySrc src{keys2 + static_cast<int64_t>(qslot) * keys2_pitch, n_slots};
  const uint32_t n = block_topk(src, total, static_cast<uint32_t>(k), sm);
  bool ok = true;
  if (check) {
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
}

// ------------------------------------------------------------------------------------------
// K1b: streaming scorer for small query batches (latency mode).  No tensor cores: the corpus is
// read once with 16-byte loads, one warp per row, fp32 accumulate; each CTA keeps a per-query key
// buffer in shared memory that is compacted (block bitonic sort) between rounds.  Emits at most
// k' keys per (CTA, query) in the same layout the fused kernel uses, so K1c/K2 are shared.
// ------------------------------------------------------------------------------------------
constexpr int kGemvThreads = 512;               // 16 warps
constexpr int kGemvRowsPerRound = 16 * 8;       // each warp scores 8 rows per round
constexpr int kGemvMaxQ = 8;

template <typename RowT>
__global__ void __launch_bounds__(kGemvThreads) gemv_topk_kernel(
    const RowT* __restrict__ rows, int64_t row_pitch, int dp, int64_t n_rows,
    const float* __restrict__ qcanon, int nq, int keep, int cap /*pow2 >= keep + rows/round*/,
    uint64_t* __restrict__ cand, uint32_t* __restrict__ cand_count, int nq_pad, int cand_cap) {
  extern __shared__ uint8_t smem_raw[];
  // layout: q fp32 [nq][dp] | keys [nq][cap] | counts [nq] | tau [nq]
  float* sq = reinterpret_cast<float*>(smem_raw);
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw + ((static_cast<size_t>(nq) * dp * 4 + 15) & ~size_t(15)));
  uint32_t* scount = reinterpret_cast<uint32_t*>(skeys + static_cast<size_t>(nq) * cap);
  uint64_t* stau = reinterpret_cast<uint64_t*>(scount + 2 * kGemvMaxQ);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < nq * dp; i += blockDim.x) sq[i] = qcanon[i];
  if (threadIdx.x < kGemvMaxQ) { scount[threadIdx.x] = 0; stau[threadIdx.x] = kKeyNone; }
  __syncthreads();

  constexpr int kElemsPerVec = 16 / sizeof(RowT);  // 8 bf16 or 4 fp32 per 16-byte load
  const int vecs = dp / kElemsPerVec;               // dp is a multiple of 64
  const int64_t n_rounds = (n_rows + kGemvRowsPerRound - 1) / kGemvRowsPerRound;
  for (int64_t round = blockIdx.x; round < n_rounds; round += gridDim.x) {
    const int64_t base = round * kGemvRowsPerRound + warp * 8;
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
      const int64_t row = base + r;
      if (row >= n_rows) break;
      float acc[kGemvMaxQ];
#pragma unroll
      for (int j = 0; j < kGemvMaxQ; ++j) acc[j] = 0.f;
      const uint8_t* rp = reinterpret_cast<const uint8_t*>(rows + row * row_pitch);
      for (int v = lane; v < vecs; v += 32) {
        const uint4 w = ptx::ld_nc_u4(rp + static_cast<size_t>(v) * 16);
        float x[8];
        if (sizeof(RowT) == 2) {
          x[0] = __uint_as_float(w.x << 16); x[1] = __uint_as_float(w.x & 0xFFFF0000u);
          x[2] = __uint_as_float(w.y << 16); x[3] = __uint_as_float(w.y & 0xFFFF0000u);
          x[4] = __uint_as_float(w.z << 16); x[5] = __uint_as_float(w.z & 0xFFFF0000u);
          x[6] = __uint_as_float(w.w << 16); x[7] = __uint_as_float(w.w & 0xFFFF0000u);
        } else {
          x[0] = __uint_as_float(w.x); x[1] = __uint_as_float(w.y);
          x[2] = __uint_as_float(w.z); x[3] = __uint_as_float(w.w);
        }
#pragma unroll
        for (int j = 0; j < kGemvMaxQ; ++j) {
          if (j < nq) {
            const float* qj = sq + j * dp + v * kElemsPerVec;
#pragma unroll
            for (int e = 0; e < kElemsPerVec; ++e) acc[j] = fmaf(qj[e], x[e], acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kGemvMaxQ; ++j) {
        if (j < nq) {
          float s = acc[j];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
          if (lane == 0) {
            const uint64_t key = make_key(s, static_cast<uint32_t>(row));
            if (key > stau[j]) {
              const uint32_t pos = atomicAdd(&scount[j], 1u);
              skeys[static_cast<size_t>(j) * cap + pos] = key;   // pos < cap by construction
            }
          }
        }
      }
    }
    __syncthreads();
    // compact any buffer that could overflow in the next round
    for (int j = 0; j < nq; ++j) {
      const uint32_t c = scount[j];
      if (c + kGemvRowsPerRound > static_cast<uint32_t>(cap)) {
        uint64_t* kj = skeys + static_cast<size_t>(j) * cap;
        for (uint32_t i = c + threadIdx.x; i < static_cast<uint32_t>(cap); i += blockDim.x) kj[i] = 0ull;
        block_bitonic_desc(kj, cap);
        if (threadIdx.x == 0) { scount[j] = keep; stau[j] = kj[keep - 1]; }
        __syncthreads();
      }
    }
  }
  // emit: sort what is left and write the best k' of each query
  __syncthreads();
  for (int j = 0; j < nq; ++j) {
    const uint32_t c = scount[j];
    uint64_t* kj = skeys + static_cast<size_t>(j) * cap;
    for (uint32_t i = c + threadIdx.x; i < static_cast<uint32_t>(cap); i += blockDim.x) kj[i] = 0ull;
    block_bitonic_desc(kj, cap);
    const uint32_t n = min(c, static_cast<uint32_t>(keep));
    const size_t slot = static_cast<size_t>(blockIdx.x) * nq_pad + j;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) cand[slot * cand_cap + i] = kj[i];
    if (threadIdx.x == 0) cand_count[slot] = n;
    __syncthreads();
  }
}

}  // namespace vfi
