// dense_small.cuh — K1s: the fused GEMM + top-k' for SMALL query batches (9..64 queries), operands swapped.
//
// K1 (dense_fused.cuh) puts queries on the M side of the MMA (128 per CTA) and streams corpus tiles as the N operand; for a
// batch of a few dozen queries that wastes the tensor pipe (harmless) but also re-fetches the 256 KB query block from L2
// for EVERY corpus tile: the L2->SM fill is 1.5 x the corpus bytes and bounds the kernel at 0.68-0.81 of the HBM peak
// (measured, 16/64/128 queries over 1M x 1024).  Here the roles are swapped:
//   A = corpus tile   [128 rows x 64] bf16, K-major, 128B-swizzled, streamed from HBM by TMA through a 5-deep ring
//   B = query block   [NQ <= 64 rows x K'] RESIDENT in shared memory for the whole kernel (loaded once)
//   D = scores        TMEM lane = corpus row, column = query (tcgen05.mma cta_group::1, M = 128, N = NQ padded to 16)
// so the only traffic is the corpus stream: the kernel is HBM-bound like K1b, but on the tensor cores and for up to 64 queries.
//
// Selection: a TMEM lane holds ONE corpus row's scores against all queries, so the per-query state (threshold, key count) is
// shared by the CTA in shared memory; a passing (score, row) key is appended to that query's buffer in global memory through a
// shared-memory counter.  After every tile the four epilogue warps compact the buffers that could overflow during the next
// tile (warp_select_compact: exact k' best, threshold rises).  Buffers and counts have the layout K1c reads
// ([cta][nq_pad][cap]), the admission hint comes from the same row-sample pass as K1's.
#pragma once
#include "ptx.cuh"
#include "topk_common.cuh"

namespace vfi {

constexpr int kSmTileRows = 128;                 // corpus rows per MMA (M)
constexpr int kSmMaxQ = 64;                      // queries (N)
constexpr int kSmStages = 5;
constexpr int kSmABytes = kSmTileRows * 64 * 2;  // 16 KB per stage
constexpr int kSmAccCols = 64;                   // TMEM columns per accumulator stage
constexpr int kSmThreads = 256;                  // warps 0-3: TMA / MMA / TMEM alloc / spare, 4-7: epilogue
constexpr int kSmQBudget = 128 * 1024;           // resident query block: nq_pad * K' * 2 bytes
struct SmallSel {                                // per-query selection state shared by the CTA
  unsigned long long tau_key[kSmMaxQ];
  float tau_f[kSmMaxQ];
  uint32_t count[kSmMaxQ];
};
__host__ __device__ inline size_t dense_small_smem(int nq_pad, int n_kblocks) {
  return 1024 + static_cast<size_t>(nq_pad) * 128 * n_kblocks + kSmStages * kSmABytes + 256 + 4 * 256 * 4 + sizeof(SmallSel);
}

struct SmallParams {
  int nq;                 // queries
  int nq_pad;             // multiple of 16, <= 64
  int n_rows;
  int n_kblocks;          // K' / 64
  int n_tiles;            // ceil(n_rows / 128)
  int keep;               // k'
  int cap;                // keys per (cta, query) buffer, >= keep + 160
  uint64_t* cand;         // [gridDim.x][nq_pad][cap]
  uint32_t* cand_count;   // [gridDim.x][nq_pad]
  const float* tau_init;  // [nq] admission hints or nullptr
};

__global__ void __launch_bounds__(kSmThreads, 1)
dense_small_kernel(const __grid_constant__ CUtensorMap tmap_q /* box = nq_pad rows */,
                   const __grid_constant__ CUtensorMap tmap_d /* box = 128 rows */, const SmallParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int q_kb_bytes = p.nq_pad * 128;                              // one k-block of the query block
  uint8_t* sQ = smem;
  uint8_t* sA = smem + ((static_cast<size_t>(q_kb_bytes) * p.n_kblocks + 1023) & ~static_cast<size_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + kSmStages * kSmABytes);
  uint64_t* full_bar = bars;                       // [kSmStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kSmStages;          // [kSmStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kSmStages;      // [2]          MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kSmStages + 2; // [2]          epilogue -> MMA
  uint64_t* qfull_bar = bars + 2 * kSmStages + 4;  // query block resident
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kSmStages + 5);
  uint32_t* hist = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [4 warps][256]
  SmallSel* sel = reinterpret_cast<SmallSel*>(reinterpret_cast<uint8_t*>(hist) + 4 * 256 * 4);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_d);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kSmStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 4);
    }
    ptx::mbar_init(qfull_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 2 * kSmAccCols);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x < kSmMaxQ) {
    const int j = threadIdx.x;
    float tf = -INFINITY;
    unsigned long long tk = kKeyNone;
    if (j < p.nq && p.tau_init != nullptr) {
      tf = p.tau_init[j];
      tk = make_key(tf, 0u);                       // admits score > hint only
    }
    sel->tau_f[j] = tf;
    sel->tau_key[j] = tk;
    sel->count[j] = 0u;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nkb = p.n_kblocks;
  const int cta = static_cast<int>(blockIdx.x);
  const int n_ctas = static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer: the query block once, then the corpus stream =====================
    ptx::mbar_expect_tx(qfull_bar, static_cast<uint32_t>(q_kb_bytes) * nkb);
    for (int kb = 0; kb < nkb; ++kb)
      ptx::tma_load_2d(sQ + static_cast<size_t>(kb) * q_kb_bytes, &tmap_q, qfull_bar, kb * 64, 0, ptx::kEvictLast);
    uint32_t stage = 0, phase = 0;
    for (int t = cta; t < p.n_tiles; t += n_ctas) {
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        ptx::mbar_expect_tx(&full_bar[stage], kSmABytes);
        ptx::tma_load_2d(sA + stage * kSmABytes, &tmap_d, &full_bar[stage], kb * 64, t * kSmTileRows, ptx::kEvictFirst);
        if (++stage == kSmStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ptx::umma_idesc_bf16(kSmTileRows, static_cast<uint32_t>(p.nq_pad));
    ptx::mbar_wait(qfull_bar, 0);
    ptx::tc_fence_after();
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int t = cta; t < p.n_tiles; t += n_ctas) {
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kSmAccCols;
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(sA + stage * kSmABytes);
        const uint32_t b_addr = ptx::smem_u32(sQ + static_cast<size_t>(kb) * q_kb_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_bf16_ss(d_tmem, ptx::umma_desc_k128(a_addr + k * 32), ptx::umma_desc_k128(b_addr + k * 32), idesc,
                            (kb | k) != 0 ? 1u : 0u);
        ptx::tc_commit(&empty_bar[stage]);
        if (kb == nkb - 1) ptx::tc_commit(&tfull_bar[acc]);
        if (++stage == kSmStages) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread = corpus row of the tile =====================
    const uint32_t wq = warp - 4;                    // TMEM lane quadrant (= warp % 4)
    uint32_t* my_hist = hist + wq * 256;
    uint32_t acc = 0, acc_phase = 0;
    const uint32_t cap = static_cast<uint32_t>(p.cap);
    uint64_t* my_cand = p.cand + static_cast<size_t>(cta) * p.nq_pad * p.cap;
    for (int t = cta; t < p.n_tiles; t += n_ctas) {
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t row = static_cast<uint32_t>(t) * kSmTileRows + wq * 32 + lane;
      const bool valid_row = row < static_cast<uint32_t>(p.n_rows);
      const uint32_t taddr = tmem_base + ((wq * 32u) << 16) + acc * kSmAccCols;
#pragma unroll 1
      for (int c = 0; c * 32 < p.nq_pad; ++c) {
        float v[32];
        ptx::tmem_ld_32x32(taddr + c * 32, v);
        ptx::tmem_ld_wait();
        if (valid_row) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = c * 32 + j;
            if (q < p.nq && v[j] >= sel->tau_f[q]) {
              const uint64_t key = make_key(v[j], row);
              if (key > sel->tau_key[q]) {
                const uint32_t pos = atomicAdd(&sel->count[q], 1u);
                if (pos < cap) my_cand[static_cast<size_t>(q) * p.cap + pos] = key;   // (cannot overflow: see below)
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      // A tile adds at most 128 keys to a buffer: whatever could not take another 128 is compacted now to its exact k'
      // best keys (and the threshold rises).  The appends above must be visible to the compacting warp: block-scope fence
      // + barrier among the 128 epilogue threads.
      __threadfence_block();
      ptx::named_bar_sync(1, 128);
      for (int q = static_cast<int>(wq); q < p.nq; q += 4) {
        const uint32_t cnt = sel->count[q];
        if (cnt + kSmTileRows > cap) {
          const uint64_t kth = warp_select_compact(my_cand + static_cast<size_t>(q) * p.cap, min(cnt, cap), p.keep, my_hist, lane);
          if (lane == 0) {
            sel->count[q] = p.keep;
            sel->tau_key[q] = kth;
            sel->tau_f[q] = key_score(kth);
          }
        }
      }
      __threadfence_block();
      ptx::named_bar_sync(1, 128);
    }
    for (int q = static_cast<int>(threadIdx.x) - 128; q < p.nq_pad; q += 128)
      p.cand_count[static_cast<size_t>(cta) * p.nq_pad + q] = (q < p.nq) ? min(sel->count[q], cap) : 0u;
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * kSmAccCols);
  }
}

}  // namespace vfi
