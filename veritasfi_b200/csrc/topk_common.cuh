// topk_common.cuh — the ordering contract and the candidate-buffer machinery shared by the
// dense, BM25, merge and fusion kernels.
//
// Total order: score descending, then id ascending.
// A (score, id) pair is packed into one 64-bit key whose unsigned order IS that total order:
//   key = orderable(score) << 32 | (0xFFFFFFFF - id)          (larger key = better)
// ids are shard-local row numbers < 2^32 - 1 (the id offset is added on output).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vfi {

constexpr uint64_t kKeyNone = 0ull;  // sorts after every real key (a real id is < 2^32-1)

__host__ __device__ __forceinline__ uint32_t float_orderable(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f + 0.0f);  // -0 -> +0
#else
  union { float f; uint32_t u; } c;
  c.f = f + 0.0f;
  uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float orderable_float(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c;
  c.u = u;
  return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t id) {
  return (static_cast<uint64_t>(float_orderable(score)) << 32) | (0xFFFFFFFFu - id);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) {
  return orderable_float(static_cast<uint32_t>(k >> 32));
}
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t k) {
  return 0xFFFFFFFFu - static_cast<uint32_t>(k);
}

#ifdef __CUDACC__

// -------------------------------------------------------------------------------------------
// Warp-cooperative exact selection on a small key buffer in global memory (L2 resident).
// Finds the `keep`-th largest key of buf[0..n) by an 8-bit MSD radix walk over the 64-bit keys
// (keys are unique because ids are), then compacts buf in place to exactly the `keep` largest.
// Returns that `keep`-th key (the new admission threshold).  hist: 256 u32 of shared memory
// private to the calling warp.  All 32 lanes must call.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t warp_select_compact(uint64_t* buf, uint32_t n, uint32_t keep,
                                                        uint32_t* hist, uint32_t lane) {
  uint64_t prefix = 0;       // the bits fixed so far (high bits of the answer)
  uint32_t remaining = keep; // rank of the answer among keys matching the prefix
  uint64_t result = 0;
  bool done = false;
#pragma unroll 1
  for (int shift = 56; shift >= 0 && !done; shift -= 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) hist[lane * 8 + i] = 0;
    __syncwarp();
    const uint64_t himask = (shift == 56) ? 0ull : (~0ull << (shift + 8));
    for (uint32_t i = lane; i < n; i += 32) {
      uint64_t k = buf[i];
      if ((k & himask) == prefix) atomicAdd(&hist[(k >> shift) & 0xFF], 1u);
    }
    __syncwarp();
    // suffix counts: lane l owns bins [8l, 8l+8); walk from bin 255 down
    uint32_t c[8];
    uint32_t mine = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i] = hist[lane * 8 + i]; mine += c[i]; }
    // above = number of keys in bins owned by higher lanes
    uint32_t above = 0;
    {
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_down_sync(0xFFFFFFFFu, incl, o);
        if (lane + o < 32) incl += t;
      }
      above = incl - mine;
    }
    // the answer's bin is in the unique lane with above < remaining <= above + mine
    bool owner = (above < remaining) && (remaining <= above + mine);
    uint32_t digit = 0, before = 0, cnt = 0;
    if (owner) {
      uint32_t run = above;
#pragma unroll
      for (int i = 7; i >= 0; --i) {
        if (run < remaining && remaining <= run + c[i]) { digit = lane * 8 + i; before = run; cnt = c[i]; }
        run += c[i];
      }
    }
    uint32_t src = __ffs(__ballot_sync(0xFFFFFFFFu, owner)) - 1;
    digit = __shfl_sync(0xFFFFFFFFu, digit, src);
    before = __shfl_sync(0xFFFFFFFFu, before, src);
    cnt = __shfl_sync(0xFFFFFFFFu, cnt, src);
    prefix |= static_cast<uint64_t>(digit) << shift;
    remaining -= before;
    if (cnt == remaining) {
      // every key in this bin is kept: the threshold is the smallest key with this prefix
      result = prefix;  // low bits zero => "key >= result" keeps exactly the bin and above
      done = true;
    } else if (shift == 0) {
      result = prefix;
      done = true;
    }
    __syncwarp();
  }
  // in-place stable compaction of keys >= result (exactly `keep` of them)
  uint32_t out = 0;
  uint64_t kept_min = ~0ull;
  for (uint32_t base = 0; base < n; base += 32) {
    uint32_t i = base + lane;
    uint64_t k = (i < n) ? buf[i] : 0ull;
    bool keepit = (i < n) && (k >= result);
    uint32_t m = __ballot_sync(0xFFFFFFFFu, keepit);
    if (keepit) {
      buf[out + __popc(m & ((1u << lane) - 1u))] = k;
      kept_min = k < kept_min ? k : kept_min;
    }
    out += __popc(m);
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t t = __shfl_xor_sync(0xFFFFFFFFu, kept_min, o);
    kept_min = t < kept_min ? t : kept_min;
  }
  return kept_min;  // exact keep-th largest key
}

// Block-wide bitonic sort of n (power of two) keys in shared memory, descending.
__device__ __forceinline__ void block_bitonic_desc(uint64_t* keys, uint32_t n) {
  for (uint32_t size = 2; size <= n; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (uint32_t t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        uint32_t lo = 2 * t - (t & (stride - 1));
        uint32_t hi = lo + stride;
        bool desc = ((lo & size) == 0);
        uint64_t a = keys[lo], b = keys[hi];
        if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// The last CTA of a grid to get here copies *counter to *host_out (pinned host memory mapped into the device address
// space): the host reads the batch's certificate count after the event behind the kernel without a copy operation in the
// stream (a 4-byte cudaMemcpyAsync costs the stream 8-15 us on B200, the kernel epilogue nothing measurable).  Thread 0
// of every CTA must be the thread that updated *counter.  done_ctas starts at zero (prep_queries_kernel).
__device__ __forceinline__ void publish_flag_count(int* counter, int* done_ctas, int* host_out) {
  if (done_ctas == nullptr || threadIdx.x != 0) return;
  __threadfence();
  if (atomicAdd(done_ctas, 1) == static_cast<int>(gridDim.x * gridDim.y) - 1) {
    __threadfence();
    *reinterpret_cast<volatile int*>(host_out) = *reinterpret_cast<volatile int*>(counter);
    __threadfence_system();
  }
}

#endif  // __CUDACC__
}  // namespace vfi
