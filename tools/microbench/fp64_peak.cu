// fp64_peak.cu -- measures the FP64 issue rates that bound the kernels of this repo on the box:
// DFMA (CUDA-core pipe), DMMA m8n8k4 (tensor pipe), and fast_exp throughput.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../include -I../../uni_b200/csrc fp64_peak.cu -o fp64_peak
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"

__global__ void k_dfma(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mix(double* out, int iters) {  // 1 DMMA : 4 DFMA per lane, independent chains
  double c[4][2], f[8];
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-3 + i;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      f[2 * i] = fma(f[2 * i], 1.0000001, 1e-9);
      f[2 * i + 1] = fma(f[2 * i + 1], 1.0000001, 1e-9);
    }
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  for (int i = 0; i < 8; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, int iters, int fast) {
  double x[4];
  for (int i = 0; i < 4; ++i) x[i] = -1.0 + 1e-3 * threadIdx.x + i;
  double s = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double e = fast ? um::fast_exp(x[i]) : exp(x[i]);
      s += e;
      x[i] = x[i] * 0.999 + 1e-4;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, (size_t)sms * 8 * 1024 * 8);
  const int blocks = sms * 8, threads = 256, iters = 20000;
  double lanes = (double)blocks * threads;
  float ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
  double fma_per_s = lanes * iters * 8 / (ms * 1e-3);
  printf("DFMA  : %.3f ms  %.2f TFLOP/s  (%.1f FMA lanes/clk/SM at %d MHz nominal)\n", ms, 2 * fma_per_s / 1e12, fma_per_s / sms / (clk * 1e3), clk / 1000);
  ms = timeit([&] { k_dmma<<<blocks, threads>>>(out, iters); });
  double mac_per_s = (double)blocks * (threads / 32) * iters * 8 * 256 / (ms * 1e-3);
  printf("DMMA  : %.3f ms  %.2f TFLOP/s  (%.1f MAC/clk/SM)\n", ms, 2 * mac_per_s / 1e12, mac_per_s / sms / (clk * 1e3));
  ms = timeit([&] { k_mix<<<blocks, threads>>>(out, iters); });
  double mix_mac = (double)blocks * (threads / 32) * iters * 4 * 256 / (ms * 1e-3);
  double mix_fma = lanes * iters * 8 / (ms * 1e-3);
  printf("MIX   : %.3f ms  DMMA %.2f TFLOP/s + DFMA %.2f TFLOP/s concurrently\n", ms, 2 * mix_mac / 1e12, 2 * mix_fma / 1e12);
  for (int fast = 0; fast < 2; ++fast) {
    ms = timeit([&] { k_exp<<<blocks, threads>>>(out, iters / 10, fast); });
    double eps = lanes * (iters / 10) * 4 / (ms * 1e-3);
    printf("%s: %.3f ms  %.1f G exp/s  (%.2f exp/clk/SM)\n", fast ? "fast_exp" : "exp     ", ms, eps / 1e9, eps / sms / (clk * 1e3));
  }
  return 0;
}
