// Micro-benchmark: issue rates of the instructions the canonical rescoring loop is made of (B200, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/fp64_rates tools/ubench/fp64_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

template <int CHAINS>
__global__ void dfma_kernel(double* out, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x + c;
  for (int i = 0; i < kIters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// F2F.F64.F32 stream: CHAINS independent conversions per iteration, folded with integer xor (cheap) so the
// conversion itself is what the loop is made of.
template <int CHAINS>
__global__ void f2f_kernel(uint32_t* out, uint32_t seed) {
  uint32_t v[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) v[c] = seed + threadIdx.x * 977u + c * 131u;
  for (int i = 0; i < kIters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      const double d = static_cast<double>(__uint_as_float((v[c] & 0x3FFFFFFFu) | 0x20000000u));
      v[c] ^= static_cast<uint32_t>(__double2hiint(d)) + static_cast<uint32_t>(__double2loint(d));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s ^= v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// integer widening of a bf16 (upper half of a word) to the high word of the equal double (normal numbers)
template <int CHAINS>
__global__ void widen_kernel(uint32_t* out, uint32_t seed) {
  uint32_t v[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) v[c] = seed + threadIdx.x * 977u + c * 131u;
  for (int i = 0; i < kIters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      const uint32_t w = v[c];
      const uint32_t hi = ((((w >> 3) & 0x0FFFE000u) + 0x38000000u) | (w & 0x80000000u));
      v[c] ^= hi;
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s ^= v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the candidate inner loop: one sequential DFMA chain per thread, x from registers (bf16 pairs), q as doubles from
// shared memory; MODE 0 = F2F conversion, 1 = integer widening (normal numbers only)
template <int MODE>
__global__ void chain_kernel(double* out, uint32_t seed) {
  __shared__ double q[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) q[i] = 1.0 / (i + 1);
  __syncthreads();
  uint32_t w = seed + threadIdx.x * 2654435761u;
  double acc = 0.0;
  for (int rep = 0; rep < kIters / 1024; ++rep) {
#pragma unroll 8
    for (int j = 0; j < 1024; j += 2) {
      w = w * 1664525u + 1013904223u;
      const uint32_t ww = (w & 0x3FFF3FFFu) | 0x30003000u;
      double x0, x1;
      if (MODE == 0) {
        x0 = static_cast<double>(__uint_as_float(ww << 16));
        x1 = static_cast<double>(__uint_as_float(ww & 0xFFFF0000u));
      } else {
        const uint32_t h0 = ((((ww << 13) & 0x0FFFE000u) + 0x38000000u) | ((ww << 16) & 0x80000000u));
        const uint32_t h1 = ((((ww >> 3) & 0x0FFFE000u) + 0x38000000u) | (ww & 0x80000000u));
        x0 = __hiloint2double(h0, 0);
        x1 = __hiloint2double(h1, 0);
      }
      const double2 qq = *reinterpret_cast<const double2*>(&q[j]);
      acc = fma(qq.x, x0, acc);
      acc = fma(qq.y, x1, acc);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
static float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int sms = p.multiProcessorCount;
  void* buf;
  cudaMalloc(&buf, size_t(sms) * 8 * 1024 * 8);
  printf("device %s, %d SMs, max clock %d kHz\n", p.name, sms, clk);
  for (int threads : {128, 256, 512, 1024}) {
    for (int ctas : {1, 2}) {
      if (threads * ctas > 2048) continue;
      const int grid = sms * ctas;
      const double lanes = double(grid) * threads * kIters;
      float ms;
      ms = time_ms([&] { dfma_kernel<1><<<grid, threads>>>((double*)buf, 1.0000001, 1e-9); });
      printf("threads/SM %4d  DFMA 1 chain : %8.3f ms  %7.2f lane-ops/ns  (%.1f per clk per SM at 1.9 GHz)\n", threads * ctas, ms,
             lanes / ms * 1e-6, lanes / ms * 1e-6 / sms / 1.9);
      ms = time_ms([&] { dfma_kernel<4><<<grid, threads>>>((double*)buf, 1.0000001, 1e-9); });
      printf("threads/SM %4d  DFMA 4 chains: %8.3f ms  %7.2f lane-ops/ns  (%.1f per clk per SM)\n", threads * ctas, ms,
             4 * lanes / ms * 1e-6, 4 * lanes / ms * 1e-6 / sms / 1.9);
      ms = time_ms([&] { f2f_kernel<4><<<grid, threads>>>((uint32_t*)buf, 17); });
      printf("threads/SM %4d  F2F 4 chains : %8.3f ms  %7.2f lane-ops/ns  (%.1f per clk per SM)\n", threads * ctas, ms,
             4 * lanes / ms * 1e-6, 4 * lanes / ms * 1e-6 / sms / 1.9);
      ms = time_ms([&] { widen_kernel<4><<<grid, threads>>>((uint32_t*)buf, 17); });
      printf("threads/SM %4d  widen 4 ch   : %8.3f ms  %7.2f lane-ops/ns  (%.1f per clk per SM)\n", threads * ctas, ms,
             4 * lanes / ms * 1e-6, 4 * lanes / ms * 1e-6 / sms / 1.9);
      ms = time_ms([&] { chain_kernel<0><<<grid, threads>>>((double*)buf, 17); });
      printf("threads/SM %4d  chain F2F    : %8.3f ms  %7.2f elem/ns\n", threads * ctas, ms, lanes / ms * 1e-6);
      ms = time_ms([&] { chain_kernel<1><<<grid, threads>>>((double*)buf, 17); });
      printf("threads/SM %4d  chain widen  : %8.3f ms  %7.2f elem/ns\n", threads * ctas, ms, lanes / ms * 1e-6);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
