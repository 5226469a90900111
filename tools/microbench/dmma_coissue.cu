// dmma_coissue.cu -- does a DMMA in flight block the other warps of its SM sub-partition from issuing?
// Warps 0-3 (one per sub-partition) run a DMMA stream (4 independent accumulators); warps 4-7 run a stream of
// another kind (integer IMAD chain / LDS loop / DADD chain / DFMA independent) for a fixed count and time themselves.
// Compare the second group's time alone and next to the DMMA stream.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int KIND>
__global__ void k(double* out, int iters, int with_dmma, long long* cyc, int nother) {
  __shared__ double sm[4096];
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp < 4) {
    if (!with_dmma) return;
    double c[4][2] = {};
    double a = 1.0 + lane, b = 0.5 + lane;
    long long n = 0;
    while (!stop) {
#pragma unroll
      for (int u = 0; u < 8; ++u) dmma(c[u & 3][0], c[u & 3][1], a, b);
      n += 8;
    }
    out[blockIdx.x * 1024 + threadIdx.x] = c[0][0] + c[1][0] + c[2][0] + c[3][0];
    if (lane == 0 && blockIdx.x == 0) cyc[1 + warp] = n;
  } else if (warp < 4 + nother) {
    long long t0 = clock64();
    double acc = 0.0;
    if (KIND == 0) {          // dependent integer chain
      unsigned x = lane;
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = x * 1664525u + 1013904223u;
      }
      acc = x;
    } else if (KIND == 1) {   // independent integer ops (4 chains)
      unsigned x0 = lane, x1 = lane + 1, x2 = lane + 2, x3 = lane + 3;
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { x0 = x0 * 1664525u + 1013904223u; x1 = x1 * 22695477u + 1u; x2 = x2 * 1103515245u + 12345u; x3 = x3 * 134775813u + 1u; }
      }
      acc = x0 + x1 + x2 + x3;
    } else if (KIND == 2) {   // shared loads (8 in flight)
      const double* p = sm + lane;
      for (int i = 0; i < iters; ++i) {
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[u]) : "r"((unsigned)__cvta_generic_to_shared(p + ((u * 32 + i * 8) & 4064))));
        unsigned long long bits = 0;
#pragma unroll
        for (int u = 0; u < 16; ++u) bits ^= __double_as_longlong(v[u]);
        acc += (double)(bits & 1);
      }
    } else if (KIND == 3) {   // independent DFMA (4 chains)
      double x0 = lane, x1 = 1, x2 = 2, x3 = 3;
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { x0 = fma(x0, 1.0000001, 1e-9); x1 = fma(x1, 1.0000001, 1e-9); x2 = fma(x2, 1.0000001, 1e-9); x3 = fma(x3, 1.0000001, 1e-9); }
      }
      acc = x0 + x1 + x2 + x3;
    }
    long long t1 = clock64();
    out[blockIdx.x * 1024 + threadIdx.x] = acc;
    if (lane == 0 && warp == 4 && blockIdx.x == 0) cyc[0] = t1 - t0;
    __syncwarp();
    if (lane == 0) atomicAdd((int*)&stop, 1);
  }
  // let the DMMA warps see `stop` only when every "other" warp is done
  if (warp >= 4 && warp < 4 + nother && lane == 0) { while (stop < nother) {} stop = 1000; }
}
template <int KIND>
void run(const char* name, double* out, long long* cyc, int nother) {
  const int iters = 20000;
  for (int w = 0; w < 2; ++w) {
    cudaMemset(cyc, 0, 64);
    for (int r = 0; r < 2; ++r) { k<KIND><<<148, 256>>>(out, iters, w, cyc, nother); cudaDeviceSynchronize(); }
    long long h[5]; cudaMemcpy(h, cyc, 40, cudaMemcpyDeviceToHost);
    printf("%-34s x%d warps/SMSP %s : %.2f clk per instruction", name, nother / 4, w ? "next to a DMMA stream" : "alone                ", (double)h[0] / ((double)iters * 16));
    if (w) printf("   (DMMA warp: %.1f clk per DMMA)", (double)h[0] / (double)h[1]);
    printf("\n");
  }
}
int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  long long* cyc; cudaMalloc(&cyc, 64);
  run<0>("dependent IMAD chain", out, cyc, 4);
  run<1>("4 independent IMAD chains", out, cyc, 4);
  run<2>("16 LDS.64 in flight", out, cyc, 4);
  run<3>("4 independent DFMA chains", out, cyc, 4);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  return 0;
}
