// radix_sort.cuh — stable LSD radix sort of (u32 key, u32 value) pairs on the device, 8 bits per pass.
// Serves vfi_bm25_rank_all: bm25s retrieve(k = N) as the reference calls it (/root/reference/src/utils/
// ensembleRetriever.py:189) wants EVERY doc ranked.  Keys are ~orderable(score) (ascending key = descending score),
// values the doc ids in ascending order on entry: a stable sort then leaves ties in id order, i.e. the total order
// (score desc, id asc), with the id outside the sort key (4 passes instead of 8).
//
// Per pass: rs_hist_kernel (digit histogram of every 4096-key tile) -> rs_scan_kernel (exclusive scan of the
// counters in digit-major order: where each tile's share of a digit starts) -> rs_scatter_kernel (stable rank of
// every key inside its tile by warp match + per-warp counters, then the scatter).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vfi {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;                       // keys per thread
constexpr int kRsTile = kRsThreads * kRsItems;     // 4096 keys per CTA

__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                             uint32_t* __restrict__ tile_hist /* [256][n_tiles] */, int n_tiles) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kRsTile;
#pragma unroll
  for (int i = 0; i < kRsItems; ++i) {
    const int64_t j = base + i * kRsThreads + threadIdx.x;
    if (j < n) atomicAdd(&sh[(keys[j] >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  tile_hist[static_cast<size_t>(threadIdx.x) * n_tiles + blockIdx.x] = sh[threadIdx.x];
}

// exclusive scan of m counters by one CTA (m = 256 * n_tiles; a few hundred thousand at most)
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t* __restrict__ v, int64_t m) {
  __shared__ uint32_t part[1024];
  const int64_t per = (m + 1023) / 1024;
  const int64_t lo = per * threadIdx.x, hi = min(m, lo + per);
  uint32_t sum = 0;
  for (int64_t i = lo; i < hi; ++i) sum += v[i];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {           // Hillis-Steele inclusive scan of the per-thread sums
    const uint32_t t = (static_cast<int>(threadIdx.x) >= o) ? part[threadIdx.x - o] : 0u;
    __syncthreads();
    part[threadIdx.x] += t;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - sum;
  for (int64_t i = lo; i < hi; ++i) {
    const uint32_t c = v[i];
    v[i] = run;
    run += c;
  }
}

__global__ void __launch_bounds__(kRsThreads) rs_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                                int64_t n, int shift, const uint32_t* __restrict__ tile_base,
                                                                int n_tiles, uint32_t* __restrict__ out_keys,
                                                                uint32_t* __restrict__ out_vals) {
  constexpr int kWarps = kRsThreads / 32;
  __shared__ uint32_t cnt[kWarps][256];          // per-warp digit counters, then exclusive warp bases
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kWarps * 256; i += kRsThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  // warp w owns the contiguous slice [w*512, (w+1)*512) of the tile, walked 32 keys at a time: tile order = (warp, step, lane)
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kRsTile + warp * (32 * kRsItems);
  uint32_t key[kRsItems], rank[kRsItems];
#pragma unroll
  for (int i = 0; i < kRsItems; ++i) {
    const int64_t j = base + i * 32 + lane;
    const bool valid = j < n;
    key[i] = valid ? keys[j] : 0xFFFFFFFFu;
    const uint32_t d = (key[i] >> shift) & 0xFF;
    // lanes with the same digit (invalid lanes form their own group through bit 8)
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, valid ? d : 256u);
    const uint32_t before = __popc(peers & ((1u << lane) - 1u));
    rank[i] = valid ? cnt[warp][d] + before : 0u;
    __syncwarp();
    if (valid && before == 0) cnt[warp][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {   // thread d: exclusive scan of digit d over the warps
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kRsItems; ++i) {
    const int64_t j = base + i * 32 + lane;
    if (j < n) {
      const uint32_t d = (key[i] >> shift) & 0xFF;
      const uint32_t pos = tile_base[static_cast<size_t>(d) * n_tiles + blockIdx.x] + cnt[warp][d] + rank[i];
      out_keys[pos] = key[i];
      out_vals[pos] = vals[j];
    }
  }
}

}  // namespace vfi
