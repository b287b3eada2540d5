// How many thread-block clusters of 1/2/4/8 CTAs with K1's footprint (384 threads, ~217 KB dynamic shared memory, one CTA per SM)
// does the device hold at once?  Decides whether a 4-CTA cluster variant of K1 (two CTA pairs sharing corpus tiles by TMA
// multicast) can keep all 148 SMs busy.   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occupancy cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(384, 1) probe_kernel(int* out) {
  extern __shared__ unsigned char smem[];
  if (out != nullptr && threadIdx.x == 0) out[blockIdx.x] = smem[0];
}

int main() {
  const int smem = 217 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  for (int c : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(prop.multiProcessorCount / c * c);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    printf("cluster size %2d: max active clusters %3d (%3d SMs busy)  %s\n", c, n, n * c, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
