// How many thread-block clusters of each size can be resident at once on this GPU for a kernel that takes a whole SM
// (200 KB dynamic shared memory, 192 threads)?  Decides which cluster shapes the persistent sweeps may use.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/cluster_query tools/cluster_query.cu
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void __launch_bounds__(192, 1) k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  printf("%s: %d SMs\n", pr.name, pr.multiProcessorCount);
  for (int cs : {1, 2, 4, 6, 8, 10, 12, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 64);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %3d  (= %3d CTAs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    (void)cudaGetLastError();
  }
  return 0;
}
