// How many thread-block clusters of G CTAs (one CTA per SM: 200 KB of dynamic shared memory each) can be resident on
// this device, for G = 1 .. 16?  G x that = the SMs a cluster kernel of that shape can use.
// nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_probe cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dummy(int* p) { extern __shared__ char sm[]; if (p) p[0] = sm[0]; }
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("SMs %d\n", sms);
  for (int G = 1; G <= 16; ++G) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G * 64, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_dummy, &cfg);
    printf("G=%2d  resident clusters %3d  SMs used %3d  (%s)\n", G, n, n * G, cudaGetErrorString(e));
    cudaGetLastError();
  }
  return 0;
}
