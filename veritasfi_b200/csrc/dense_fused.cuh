// dense_fused.cuh — K1: S = Q·Dᵀ on the 5th-gen tensor cores with the top-k' selection fused into
// the TMEM epilogue, so the score matrix never reaches HBM.
//
// Replaces the hot loop of faiss.IndexFlatIP.search (sgemm + heap) called at
// /root/reference/src/utils/faissRetriever.py:37.
//
// Shape of the computation
//   M = queries (128 per CTA, one TMEM lane = one query), N = corpus rows (256 per MMA, one TMEM
//   column = one row), K = embedding dimension streamed in 64-element (128-byte) blocks.
//   A = Q tile [128 x 64] bf16, B = D tile [256 x 64] bf16, both K-major, 128B-swizzled, staged by
//   TMA through a 4-deep mbarrier ring.  D accumulates in TMEM (fp32, 2 x 256 columns, double
//   buffered) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Work decomposition
//   CTA c handles query tile m = c % n_mtiles for the corpus tiles t = g, g+G, g+2G .. of its
//   group g = c / n_mtiles.  The n_mtiles CTAs of a group stream the same corpus tile at the same
//   time, so each corpus tile leaves HBM once and is shared through L2.
//
// Selection (MODE_TOPK)
//   Each epilogue thread owns one query for the whole kernel: its admission threshold lives in
//   registers, admitted (score,id) keys are appended to a per-(group,query) buffer in global
//   memory (L2 resident, written rarely) and when a buffer fills the warp compacts it to the
//   exact k' best keys (warp_select_compact) and raises the threshold.  Only rows that can still
//   enter the top-k' ever leave the SM.
#pragma once
#include "ptx.cuh"
#include "topk_common.cuh"

namespace vfi {

constexpr int kBM = 128;
constexpr int kBN = 256;
constexpr int kBK = 64;
constexpr int kStages = 4;
constexpr int kABytes = kBM * kBK * 2;   // 16 KB
constexpr int kBBytes = kBN * kBK * 2;   // 32 KB
constexpr int kDenseThreads = 384;       // warps 0-3: TMA / MMA / TMEM-alloc / spare, 4-7 and 8-11: two epilogue sets
constexpr int kTmemCols = 512;
constexpr int kDenseSmemBytes =
    1024 /*align slack*/ + kStages * (kABytes + kBBytes) + 256 /*barriers*/ + 8 * 256 * 4 /*hist*/;

enum { MODE_TOPK = 0, MODE_STORE = 1, MODE_CHUNKMAX = 2 };   // CHUNKMAX: one value per (query, 32-row chunk), its largest score

struct DenseParams {
  int nq;                 // queries in this launch
  int nq_pad;             // n_mtiles * 128
  int n_rows;             // corpus rows
  int n_kblocks;          // K' / 64
  int n_mtiles;
  int n_groups;
  int n_tiles;            // ceil(n_rows / 256)
  int keep;               // k'
  int cap;                // keys per (group, query) buffer; cap >= keep + 64
  uint64_t* cand;         // [2*n_groups][nq_pad][cap]   (one buffer per epilogue set)
  uint32_t* cand_count;   // [2*n_groups][nq_pad]
  const float* tau_init;  // [nq] admission hints (rows with score <= hint are ignored) or nullptr
  float* scores_out;      // MODE_STORE: [nq_pad][ld_scores] scores; MODE_CHUNKMAX: [nq_pad][ld_scores] chunk maxima (ld = 8 per tile)
  int64_t ld_scores;
};

// One 32-column chunk of one query's scores against its admission threshold: admitted (score, id) keys are
// appended to the query's buffer; a buffer that cannot take another 32 keys is compacted to the exact k' best by
// its warp (warp_select_compact) and the threshold rises.  All 32 lanes must call.
__device__ __forceinline__ void admit_chunk(const float (&v)[32], uint32_t id0, uint32_t n_rows, bool valid_q, float& tau_f,
                                            uint64_t& tau_key, uint32_t& count, uint64_t* buf, int cap, int keep,
                                            uint32_t* my_hist, uint32_t lane) {
  // group maxima (4 columns each) feed both the chunk test and the per-group tests
  float m4[8];
#pragma unroll
  for (int g = 0; g < 8; ++g)
    m4[g] = fmaxf(fmaxf(v[4 * g], v[4 * g + 1]), fmaxf(v[4 * g + 2], v[4 * g + 3]));
  const float mx = fmaxf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])),
                         fmaxf(fmaxf(m4[4], m4[5]), fmaxf(m4[6], m4[7])));
  if (valid_q && mx >= tau_f) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (m4[g] >= tau_f) {
#pragma unroll
        for (int j = 4 * g; j < 4 * g + 4; ++j) {
          if (v[j] >= tau_f && id0 + j < n_rows) {
            const uint64_t key = make_key(v[j], id0 + j);
            if (key > tau_key) buf[count++] = key;
          }
        }
      }
    }
  }
  // a buffer that cannot take another 32 keys is compacted now (warp-cooperative)
  uint32_t need = __ballot_sync(0xFFFFFFFFu, count + 32 > static_cast<uint32_t>(cap));
  while (need) {
    const uint32_t src = __ffs(need) - 1;
    need &= need - 1;
    uint64_t* sbuf = reinterpret_cast<uint64_t*>(
        __shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(buf), src));
    const uint32_t sn = __shfl_sync(0xFFFFFFFFu, count, src);
    __syncwarp();
    const uint64_t kth = warp_select_compact(sbuf, sn, keep, my_hist, lane);
    if (lane == src) {
      count = keep;
      tau_key = kth;
      tau_f = key_score(kth);
    }
    __syncwarp();
  }
}

// largest of the 32 scores of a chunk, rows at or beyond n_rows excluded (TMA fills them with zeros)
__device__ __forceinline__ float chunk_max(const float (&v)[32], uint32_t row_base, uint32_t n_rows) {
  float mx = -INFINITY;
  if (row_base + 32 <= n_rows) {
#pragma unroll
    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (row_base + j < n_rows) mx = fmaxf(mx, v[j]);
  }
  return mx;
}

// ===================== epilogue: TMEM -> registers -> threshold filter =====================
// Run by warps 4..11 of a CTA.  Two warp sets: set 0 (warps 4-7) drains accumulator stage 0 = the even tiles of
// this CTA, set 1 (warps 8-11) stage 1 = the odd tiles, so two tiles are filtered concurrently.  Each set keeps
// its own threshold and key buffer per query.  arrive_tempty(acc) hands the accumulator stage back to the MMA
// issuer (a local arrive for the single-CTA kernel, a remote one on the leader for the CTA-pair kernel).
template <int MODE, class ArriveFn>
__device__ __forceinline__ void dense_epilogue(const DenseParams& p, uint32_t tmem_base, uint32_t warp, uint32_t lane,
                                               int m_tile, int group, uint64_t* tfull_bar, uint32_t* hist,
                                               ArriveFn arrive_tempty) {
  const uint32_t set = (warp - 4) >> 2;
  const uint32_t wq = (warp - 4) & 3;            // TMEM lane quadrant of this warp (= warp % 4)
  const int qi = m_tile * kBM + wq * 32 + lane;  // the query this thread owns
  const bool valid_q = qi < p.nq;
  uint32_t* my_hist = hist + (warp - 4) * 256;

  uint64_t* buf = nullptr;
  uint32_t count = 0;
  uint64_t tau_key = kKeyNone;
  float tau_f = -INFINITY;
  const size_t slot = (static_cast<size_t>(group) * 2 + set) * p.nq_pad + qi;
  if (MODE == MODE_TOPK) {
    buf = p.cand + slot * p.cap;
    if (valid_q && p.tau_init != nullptr) {
      tau_f = p.tau_init[qi];
      tau_key = make_key(tau_f, 0u);   // admits score > hint only
    }
  }
  const uint32_t acc = set;
  uint32_t acc_phase = 0;
  int local_tile = 0;
  for (int t = group; t < p.n_tiles; t += p.n_groups, ++local_tile) {
    if ((local_tile & 1) != static_cast<int>(set)) continue;
    ptx::mbar_wait(&tfull_bar[acc], acc_phase);
    acc_phase ^= 1;
    ptx::tc_fence_after();
    const uint32_t row0 = static_cast<uint32_t>(t) * kBN;
    const uint32_t taddr = tmem_base + ((wq * 32u) << 16) + acc * kBN;
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      float v[32];
      ptx::tmem_ld_32x32(taddr + c * 32, v);
      ptx::tmem_ld_wait();
      if (MODE == MODE_STORE) {
        if (valid_q) {
          float4* dst = reinterpret_cast<float4*>(p.scores_out + static_cast<size_t>(qi) * p.ld_scores +
                                                  row0 + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      } else if (MODE == MODE_CHUNKMAX) {
        const float mx = chunk_max(v, row0 + c * 32, static_cast<uint32_t>(p.n_rows));
        if (valid_q) p.scores_out[static_cast<size_t>(qi) * p.ld_scores + (row0 >> 5) + c] = mx;
      } else {
        admit_chunk(v, row0 + c * 32, static_cast<uint32_t>(p.n_rows), valid_q, tau_f, tau_key, count, buf, p.cap, p.keep,
                    my_hist, lane);
      }
    }
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) arrive_tempty(acc);
  }
  if (MODE == MODE_TOPK) {
    p.cand_count[slot] = valid_q ? count : 0u;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kDenseThreads, 1)
dense_fused_kernel(const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ CUtensorMap tmap_d, const DenseParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kStages * kBBytes);
  uint64_t* full_bar = bars;                   // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;        // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;    // [2]        MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]     epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint32_t* hist = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_d);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tile = static_cast<int>(blockIdx.x) % p.n_mtiles;
  const int group = static_cast<int>(blockIdx.x) / p.n_mtiles;
  const bool active = group < p.n_groups;
  const int nkb = p.n_kblocks;

  if (active && warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0;
    const uint64_t d_hint = (p.n_mtiles > 1) ? ptx::kEvictNormal : ptx::kEvictFirst;
    for (int t = group; t < p.n_tiles; t += p.n_groups) {
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        ptx::mbar_expect_tx(&full_bar[stage], kABytes + kBBytes);
        ptx::tma_load_2d(sA + stage * kABytes, &tmap_q, &full_bar[stage], kb * kBK,
                         m_tile * kBM, ptx::kEvictLast);
        ptx::tma_load_2d(sB + stage * kBBytes, &tmap_d, &full_bar[stage], kb * kBK, t * kBN, d_hint);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (active && warp == 1 && lane == 0) {
    // ===================== MMA issuer (one thread) =====================
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(kBM, kBN);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int t = group; t < p.n_tiles; t += p.n_groups) {
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kBN;
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(sA + stage * kABytes);
        const uint32_t b_addr = ptx::smem_u32(sB + stage * kBBytes);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          ptx::umma_bf16_ss(d_tmem, ptx::umma_desc_k128(a_addr + k * 32),
                            ptx::umma_desc_k128(b_addr + k * 32), idesc,
                            (kb | k) != 0 ? 1u : 0u);
        }
        ptx::tc_commit(&empty_bar[stage]);   // frees the smem stage when MMAs retire
        if (kb == nkb - 1) ptx::tc_commit(&tfull_bar[acc]);  // accumulator complete
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (active && warp >= 4) {
    dense_epilogue<MODE>(p, tmem_base, warp, lane, m_tile, group, tfull_bar, hist,
                         [&](uint32_t a) { ptx::mbar_arrive(&tempty_bar[a]); });
  }

  // teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ==================================================================================================
// K1 CTA-pair variant (tcgen05 cta_group::2).  Two CTAs on the two SMs of a TPC own two adjacent query tiles
// (M = 256 over the pair) and work on the SAME corpus tile: each CTA stages only its own 128-query A block and
// ONE HALF (128 rows) of the 256-row corpus tile; the leader's single MMA thread issues
// tcgen05.mma.cta_group::2 and the tensor cores of both SMs read the other half of B through the pair link.
// Against the single-CTA kernel this cuts the L2->SM fill per FLOP by a third (512 KB instead of 768 KB per
// CTA and corpus tile) and the shared-memory operand reads per MMA from 12 KB to 8 KB.  Measured reason: under
// the 1 kW power cap the single-CTA kernel drops the SM clock to ~960-1140 MHz where cuBLAS holds ~1335 MHz on the
// same box (profiles/): the fill traffic, not the MMA, was the power hog.
//
// Work decomposition: pair c owns the corpus tiles c, c+P, c+2P .. (P pairs) and, for each of them, walks ALL query
// tile pairs (mp = 0 .. n_mtiles/2-1) before moving on.  A corpus tile is therefore requested by one TPC only: it
// leaves HBM once and its re-reads hit the L2 of that TPC's die.  (With several pairs per tile, requests from the
// two dies arrived far enough apart that every tile was fetched from HBM twice — ncu: 41 GB for a 20.5 GB corpus.)
// Every SM pair of the chip is busy for any batch size.  Per-query selection state (count, threshold) lives in
// shared memory between the visits of a query tile.
//
// Measured and not kept (round 2, profiles/r2_ncu_summary.md): a TMA L2 prefetch (cp.async.bulk.prefetch.tensor) of the pair's
// next corpus tile during the last visit of the current one — 1-2 % slower at C3, 6-10 % slower at C2 (the prefetches compete
// with the demand loads of an HBM-co-limited launch); a ring of 5 / 4 / 3 stages instead of 6 — 5 / 5 / 8 % slower on a 1/8
// shard of C3 at full clocks (no difference under the power cap at C3): the kernel is fill-latency-bound at 1.97 GHz.
//
// Barriers: full[s] lives in the leader and counts the TMA bytes of BOTH CTAs; empty[s] and tfull[a] exist in both
// CTAs and are signalled by one multicast tcgen05.commit; tempty[a] lives in the leader and collects the 8 epilogue
// warps of the pair (the peer's arrive remotely).
// ==================================================================================================
constexpr int kPairStages = 6;
constexpr int kPairBBytes = (kBN / 2) * kBK * 2;   // 16 KB: this CTA's half of the corpus tile
constexpr int kPairMaxMp = 4;                       // query tile pairs per launch (1024 queries)
struct PairSelState {                               // [epilogue set][query tile pair][thread of the set]
  uint64_t tau_key[2][kPairMaxMp][128];
  float tau_f[2][kPairMaxMp][128];
  uint32_t count[2][kPairMaxMp][128];
};
constexpr int kPairSmemBytes = 1024 /*align slack*/ + kPairStages * (kABytes + kPairBBytes) + 256 /*barriers*/ +
                               8 * 256 * 4 /*hist*/ + static_cast<int>(sizeof(PairSelState));

// buffers per query in MODE_TOPK: one per pair when every query tile pair is always drained by the same epilogue set
// (even number of tile pairs), else one per (pair, set)
__host__ __device__ inline int pair_sets_per_query(int n_mtiles) { return ((n_mtiles / 2) % 2 == 0) ? 1 : 2; }

template <int MODE>
__global__ void __launch_bounds__(kDenseThreads, 1)
dense_fused_pair_kernel(const __grid_constant__ CUtensorMap tmap_q,
                        const __grid_constant__ CUtensorMap tmap_d /* box = 128 rows */, const DenseParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kPairStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kPairStages * kPairBBytes);
  uint64_t* full_bar = bars;                          // [kPairStages]  TMA (both CTAs) -> MMA     (used in the leader)
  uint64_t* empty_bar = bars + kPairStages;           // [kPairStages]  MMA -> TMA                 (both CTAs)
  uint64_t* tfull_bar = bars + 2 * kPairStages;       // [2]            MMA -> epilogue            (both CTAs)
  uint64_t* tempty_bar = bars + 2 * kPairStages + 2;  // [2]            epilogues of the pair -> MMA (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPairStages + 4);
  uint32_t* hist = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 256);
  PairSelState* sel = reinterpret_cast<PairSelState*>(reinterpret_cast<uint8_t*>(hist) + 8 * 256 * 4);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();      // 0 = leader
  const bool leader = crank == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_d);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kPairStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 8);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {                                    // the same warp of BOTH CTAs: a pair-wide allocation
    ptx::tmem_alloc_pair(tmem_slot, kTmemCols);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                            // the peer's barriers and TMEM exist before anything is signalled
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int pair = static_cast<int>(blockIdx.x) >> 1;
  const int n_pairs = p.n_groups;                     // pairs that own corpus tiles (every launched pair)
  const int n_mp = p.n_mtiles >> 1;
  const int nkb = p.n_kblocks;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (one per CTA) =====================
    uint32_t stage = 0, phase = 0;
    for (int t = pair; t < p.n_tiles; t += n_pairs) {
      for (int mp = 0; mp < n_mp; ++mp) {
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0u);
          if (leader) ptx::mbar_expect_tx(&full_bar[stage], 2 * (kABytes + kPairBBytes));
          ptx::tma_load_2d_pair(sA + stage * kABytes, &tmap_q, full_leader, kb * kBK, (2 * mp + static_cast<int>(crank)) * kBM,
                                ptx::kEvictLast);
          // the first visit of a tile streams it from HBM, the other n_mp-1 find it in this die's L2
          ptx::tma_load_2d_pair(sB + stage * kPairBBytes, &tmap_d, full_leader, kb * kBK,
                                t * kBN + static_cast<int>(crank) * (kBN / 2), mp == n_mp - 1 ? ptx::kEvictFirst : ptx::kEvictNormal);
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (leader && warp == 1 && lane == 0) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * kBM, kBN);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int t = pair; t < p.n_tiles; t += n_pairs) {
      for (int mp = 0; mp < n_mp; ++mp) {
        ptx::mbar_wait_cluster_acquire(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBN;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(sA + stage * kABytes);
          const uint32_t b_addr = ptx::smem_u32(sB + stage * kPairBBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            ptx::umma_bf16_ss_pair(d_tmem, ptx::umma_desc_k128(a_addr + k * 32), ptx::umma_desc_k128(b_addr + k * 32),
                                   idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::tc_commit_pair_mcast(&empty_bar[stage], 3);            // frees the stage in both CTAs when the MMAs retire
          if (kb == nkb - 1) ptx::tc_commit_pair_mcast(&tfull_bar[acc], 3);   // accumulator complete, both epilogues
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: set 0 (warps 4-7) drains accumulator 0 = the even work items (tile, mp) of this
    // pair, set 1 (warps 8-11) accumulator 1 = the odd ones =====================
    const uint32_t set = (warp - 4) >> 2;
    const uint32_t wq = (warp - 4) & 3;               // TMEM lane quadrant of this warp (= warp % 4)
    const uint32_t tq = wq * 32 + lane;               // thread of the set = TMEM lane = query within the tile
    uint32_t* my_hist = hist + (warp - 4) * 256;
    const uint32_t tempty_leader = ptx::mapa_u32(ptx::smem_u32(&tempty_bar[set]), 0u);
    const int nspq = pair_sets_per_query(p.n_mtiles);
    if (MODE == MODE_TOPK) {
      for (int mp = 0; mp < n_mp; ++mp) {
        const int qi = (2 * mp + static_cast<int>(crank)) * kBM + static_cast<int>(tq);
        float tf = -INFINITY;
        uint64_t tk = kKeyNone;
        if (qi < p.nq && p.tau_init != nullptr) {
          tf = p.tau_init[qi];
          tk = make_key(tf, 0u);                      // admits score > hint only
        }
        sel->tau_f[set][mp][tq] = tf;
        sel->tau_key[set][mp][tq] = tk;
        sel->count[set][mp][tq] = 0u;
      }
    }
    uint32_t acc_phase = 0;
    int item = 0;
    for (int t = pair; t < p.n_tiles; t += n_pairs) {
      for (int mp = 0; mp < n_mp; ++mp, ++item) {
        if ((item & 1) != static_cast<int>(set)) continue;
        const int qi = (2 * mp + static_cast<int>(crank)) * kBM + static_cast<int>(tq);
        const bool valid_q = qi < p.nq;
        uint64_t* buf = nullptr;
        uint32_t count = 0;
        uint64_t tau_key = kKeyNone;
        float tau_f = -INFINITY;
        if (MODE == MODE_TOPK) {
          const size_t slot = (static_cast<size_t>(pair) * nspq + (nspq == 2 ? set : 0u)) * p.nq_pad + qi;
          buf = p.cand + slot * p.cap;
          count = sel->count[set][mp][tq];
          tau_key = sel->tau_key[set][mp][tq];
          tau_f = sel->tau_f[set][mp][tq];
        }
        ptx::mbar_wait(&tfull_bar[set], acc_phase);
        acc_phase ^= 1;
        ptx::tc_fence_after();
        const uint32_t row0 = static_cast<uint32_t>(t) * kBN;
        const uint32_t taddr = tmem_base + ((wq * 32u) << 16) + set * kBN;
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
          float v[32];
          ptx::tmem_ld_32x32(taddr + c * 32, v);
          ptx::tmem_ld_wait();
          if (MODE == MODE_STORE) {
            if (valid_q) {
              float4* dst = reinterpret_cast<float4*>(p.scores_out + static_cast<size_t>(qi) * p.ld_scores + row0 + c * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          } else if (MODE == MODE_CHUNKMAX) {
            const float mx = chunk_max(v, row0 + c * 32, static_cast<uint32_t>(p.n_rows));
            if (valid_q) p.scores_out[static_cast<size_t>(qi) * p.ld_scores + (row0 >> 5) + c] = mx;
          } else {
            admit_chunk(v, row0 + c * 32, static_cast<uint32_t>(p.n_rows), valid_q, tau_f, tau_key, count, buf, p.cap, p.keep,
                        my_hist, lane);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader);
        if (MODE == MODE_TOPK) {
          sel->count[set][mp][tq] = count;
          sel->tau_key[set][mp][tq] = tau_key;
          sel->tau_f[set][mp][tq] = tau_f;
        }
      }
    }
    if (MODE == MODE_TOPK) {
      // publish the counts of the buffers this set owns (untouched ones hold 0)
      for (int mp = 0; mp < n_mp; ++mp) {
        if (nspq == 1 && (mp & 1) != static_cast<int>(set)) continue;
        const int qi = (2 * mp + static_cast<int>(crank)) * kBM + static_cast<int>(tq);
        const size_t slot = (static_cast<size_t>(pair) * nspq + (nspq == 2 ? set : 0u)) * p.nq_pad + qi;
        p.cand_count[slot] = (qi < p.nq) ? sel->count[set][mp][tq] : 0u;
      }
    }
  }

  // teardown: nobody leaves while the peer can still signal its barriers or read its shared memory
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace vfi
