#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  int sizes[] = {1, 2, 4, 8, 16};
  int smems[] = {207 * 1024, 112 * 1024, 100 * 1024, 72 * 1024};
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int sm : smems)
    for (int cs : sizes) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = sm;
      cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension;
      a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
      cfg.attrs = a; cfg.numAttrs = 1;
      int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %3d KB cluster %2d -> max active clusters %3d (CTAs %3d) %s\n", sm / 1024, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
