// dmma_chain.cu -- DMMA m8n8k4 issue rate vs. number of independent accumulator chains and warps per SM
// sub-partition, plus the same with interleaved DFMA/DADD.  Answers: how much ILP does a 2-warps-per-SMSP
// kernel need to keep the FP64 pipe full?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 dmma_chain.cu -o dmma_chain
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH, int NF>   // CH independent DMMA chains; NF dependent-free DFMAs per DMMA (on NF separate chains)
__global__ void k(double* out, int iters, long long* cyc) {
  double c[CH][2], f[8];
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
  for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-3 + i;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      dmma(c[i][0], c[i][1], a, b);
#pragma unroll
      for (int j = 0; j < NF; ++j) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[(i * NF + j) & 7]) : "d"(1.0000001), "d"(1e-9));
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  for (int i = 0; i < 8; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CH, int NF>
void run(int warps_per_sm, double* out, long long* cyc) {
  const int iters = 4000;
  k<CH, NF><<<148, warps_per_sm * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  k<CH, NF><<<148, warps_per_sm * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per_dmma = (double)h / ((double)iters * CH);          // cycles per DMMA of one warp
  double warps_per_smsp = warps_per_sm / 4.0;
  double pipe = (16.0 + 2.0 * NF) * warps_per_smsp / per_dmma;  // fraction of the shared FP64 issue rate
  printf("chains=%d dfma/dmma=%d warps/SMSP=%.0f : %.1f clk per DMMA(+%d DFMA) per warp -> pipe %.0f%%\n", CH, NF, warps_per_smsp, per_dmma, NF, 100 * pipe);
}

int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  long long* cyc; cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 16}) {
    run<1, 0>(w, out, cyc); run<2, 0>(w, out, cyc); run<4, 0>(w, out, cyc); run<8, 0>(w, out, cyc);
    run<1, 3>(w, out, cyc); run<2, 3>(w, out, cyc); run<4, 3>(w, out, cyc); run<8, 3>(w, out, cyc);
    run<4, 1>(w, out, cyc);
  }
  return 0;
}
