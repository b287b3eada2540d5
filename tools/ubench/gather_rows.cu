// Micro-benchmark: how fast can a B200 gather random 2 KB rows from a multi-GB array, as a function of the size of the
// contiguous piece each request asks for?  (The rescoring kernel reads k' candidate rows per query this way.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/gather_rows tools/ubench/gather_rows.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <random>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// (a) one warp per row, whole row with coalesced 16-byte loads, ROWS_IN_FLIGHT rows per warp at a time
template <int RIF>
__global__ void gather_warp_rows(const uint8_t* __restrict__ base, const uint32_t* __restrict__ ids, int n_rows, int row_bytes,
                                 uint32_t* sink) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  uint32_t acc = 0;
  for (int r0 = warp * RIF; r0 < n_rows; r0 += n_warps * RIF) {
    uint4 v[RIF][4];
#pragma unroll
    for (int r = 0; r < RIF; ++r) {
      const uint8_t* p = base + static_cast<size_t>(ids[min(r0 + r, n_rows - 1)]) * row_bytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) v[r][i] = *reinterpret_cast<const uint4*>(p + (i * 32 + lane) * 16);
    }
#pragma unroll
    for (int r = 0; r < RIF; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc ^= v[r][i].x ^ v[r][i].y ^ v[r][i].z ^ v[r][i].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// (b) one lane per row, PIECE bytes per request via 1D bulk copies into shared memory, STAGES-deep ring per warp
template <int PIECE, int STAGES>
__global__ void gather_lane_rows(const uint8_t* __restrict__ base, const uint32_t* __restrict__ ids, int n_rows, int row_bytes,
                                 uint32_t* sink) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int kPitch = PIECE + 16;
  const int warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp_in_cta * 4;
  uint8_t* ring = smem + 8 * 4 * 8 + static_cast<size_t>(warp_in_cta) * STAGES * 32 * kPitch;
  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int n_pieces = row_bytes / PIECE;
  uint32_t acc = 0;
  uint32_t phase_base = 0;
  for (int r0 = gw * 32; r0 < n_rows; r0 += n_warps * 32) {
    const uint8_t* my = base + static_cast<size_t>(ids[min(r0 + lane, n_rows - 1)]) * row_bytes;
    auto issue = [&](int c) {
      if (c < n_pieces) {
        const int s = c % STAGES;
        if (lane == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(PIECE * 32) : "memory");
        __syncwarp();
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(ring + (s * 32 + lane) * kPitch)),
                     "l"(my + static_cast<size_t>(c) * PIECE), "r"(PIECE), "r"(smem_u32(&bars[s]))
                     : "memory");
      }
    };
    for (int s = 0; s < STAGES - 1; ++s) issue(s);
    for (int c = 0; c < n_pieces; ++c) {
      issue(c + STAGES - 1);
      const uint32_t parity = ((phase_base + c) / STAGES) & 1u;
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(&bars[c % STAGES])), "r"(parity) : "memory");
      }
      acc ^= *reinterpret_cast<const uint32_t*>(ring + ((c % STAGES) * 32 + lane) * kPitch);
      __syncwarp();
    }
    phase_base += n_pieces;   // n_pieces is a multiple of STAGES in this benchmark
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <class F>
static float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

template <int PIECE, int STAGES>
static void run_lane(const uint8_t* base, const uint32_t* ids, int n, int row_bytes, uint32_t* sink, int sms, int warps_per_cta) {
  const size_t smem = 8 * 4 * 8 + static_cast<size_t>(warps_per_cta) * STAGES * 32 * (PIECE + 16);
  cudaFuncSetAttribute(gather_lane_rows<PIECE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gather_lane_rows<PIECE, STAGES>, warps_per_cta * 32, smem);
  if (occ < 1) { printf("lane-per-row piece %4d stages %d: does not fit\n", PIECE, STAGES); return; }
  const int grid = sms * occ;
  const float ms = time_ms([&] { gather_lane_rows<PIECE, STAGES><<<grid, warps_per_cta * 32, smem>>>(base, ids, n, row_bytes, sink); });
  printf("lane-per-row piece %4d stages %d (%2d warps/SM, %3zu KB in flight/SM): %7.3f ms  %7.1f GB/s\n", PIECE, STAGES,
         occ * warps_per_cta, static_cast<size_t>(occ) * warps_per_cta * 32 * PIECE * (STAGES - 1) / 1024, ms,
         double(n) * row_bytes / ms * 1e-6);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const int row_bytes = 2048;
  const size_t n_corpus = 1250000;            // one 1/8 shard of C3
  const int n = 1024 * 128;                   // candidate rows of one batch
  uint8_t* base; uint32_t *ids, *sink;
  cudaMalloc(&base, n_corpus * row_bytes); cudaMemset(base, 1, n_corpus * row_bytes);
  cudaMalloc(&ids, n * 4); cudaMalloc(&sink, 4);
  std::vector<uint32_t> h(n);
  std::mt19937 rng(7);
  for (auto& v : h) v = rng() % n_corpus;
  cudaMemcpy(ids, h.data(), n * 4, cudaMemcpyHostToDevice);
  printf("%s: gather of %d random rows of %d B out of %.2f GB (%.1f MB per pass)\n", p.name, n, row_bytes,
         n_corpus * row_bytes * 1e-9, double(n) * row_bytes * 1e-6);
  for (int ctas : {2, 4, 8}) {
    float ms = time_ms([&] { gather_warp_rows<1><<<sms * ctas, 256>>>(base, ids, n, row_bytes, sink); });
    printf("warp-per-row, 1 row in flight, %2d warps/SM: %7.3f ms %7.1f GB/s\n", ctas * 8, ms, double(n) * row_bytes / ms * 1e-6);
    ms = time_ms([&] { gather_warp_rows<4><<<sms * ctas, 256>>>(base, ids, n, row_bytes, sink); });
    printf("warp-per-row, 4 rows in flight, %2d warps/SM: %7.3f ms %7.1f GB/s\n", ctas * 8, ms, double(n) * row_bytes / ms * 1e-6);
  }
  run_lane<64, 2>(base, ids, n, row_bytes, sink, sms, 4);
  run_lane<128, 2>(base, ids, n, row_bytes, sink, sms, 4);
  run_lane<128, 4>(base, ids, n, row_bytes, sink, sms, 4);
  run_lane<256, 2>(base, ids, n, row_bytes, sink, sms, 4);
  run_lane<256, 4>(base, ids, n, row_bytes, sink, sms, 4);
  run_lane<512, 2>(base, ids, n, row_bytes, sink, sms, 4);
  run_lane<1024, 2>(base, ids, n, row_bytes, sink, sms, 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
