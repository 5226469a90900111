// dfma_chain.cu -- FP64 DFMA/DADD dependent-issue latency and throughput vs. independent chains per warp
// and warps per SM sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 dfma_chain.cu -o dfma_chain
#include <cstdio>
#include <cuda_runtime.h>
template <int CH, bool ADD>
__global__ void k(double* out, int iters, long long* cyc) {
  double f[CH];
  for (int i = 0; i < CH; ++i) f[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (ADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[i]) : "d"(1e-9));
        else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(1.0000001), "d"(1e-9));
      }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH, bool ADD>
void run(int wps, double* out, long long* cyc) {
  const int iters = 2000;
  for (int r = 0; r < 2; ++r) { k<CH, ADD><<<148, wps * 128>>>(out, iters, cyc); cudaDeviceSynchronize(); }
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / ((double)iters * 4 * CH);
  printf("%s chains=%2d warps/SMSP=%d : %.2f clk per op per warp -> %.2f clk per op per SMSP (peak 2.0)\n", ADD ? "DADD" : "DFMA", CH, wps, per, per / wps);
}
int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  long long* cyc; cudaMalloc(&cyc, 8);
  for (int w : {1, 2, 4}) {
    run<1, false>(w, out, cyc); run<2, false>(w, out, cyc); run<4, false>(w, out, cyc); run<8, false>(w, out, cyc); run<16, false>(w, out, cyc);
    run<1, true>(w, out, cyc); run<4, true>(w, out, cyc); run<8, true>(w, out, cyc);
  }
  return 0;
}
