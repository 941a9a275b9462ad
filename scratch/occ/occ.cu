// How many 8- and 16-CTA clusters of the decode kernel's footprint (256 threads, 128 registers, ~110 KB smem) fit on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) k(float* p) {
  extern __shared__ float s[];
  float r[96];
#pragma unroll
  for (int i = 0; i < 96; ++i) r[i] = p[threadIdx.x + i * 256];
  float a = 0;
#pragma unroll
  for (int i = 0; i < 96; ++i) a += r[i] * r[(i + 7) % 96];
  s[threadIdx.x] = a;
  __syncthreads();
  p[threadIdx.x] = s[(threadIdx.x + 1) & 255];
}
int main() {
  for (int smem_kb : {110, 150, 200}) {
    int smem = smem_kb * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {8, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %d KB, cluster %2d: max active clusters %d (%s) -> %d rows of 8 per CTA-group\n", smem_kb, cs, n, cudaGetErrorString(e), n * cs);
    }
  }
  return 0;
}
