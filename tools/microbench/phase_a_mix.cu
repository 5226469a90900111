// phase_a_mix.cu -- the exact FP64 instruction mix of phase A of the 3PRF sweep, from registers and from
// shared memory: per 32 elements 1 DMMA + DADD (d = a - k) + DADD (s1 += d) + DFMA (s2 += d*d).
// Ideal: 16 + 3*2 = 22 cycles per step per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <bool SMEM, bool STATS>
__global__ void k(double* out, int iters, long long* cyc) {
  __shared__ double sm[2][2304];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * 2304; i += blockDim.x) (&sm[0][0])[i] = i * 1e-4;
  __syncthreads();
  double q[2][2][2] = {}, s1[2][2] = {}, s2[2][2] = {}, zf[8];
  for (int i = 0; i < 8; ++i) zf[i] = 1.0 + i + lane * 1e-3;
  const double ksh[2] = {0.5, 0.25};
  double areg[4][2];
  for (int i = 0; i < 4; ++i) for (int m = 0; m < 2; ++m) areg[i][m] = lane + i + m * 0.5;
  const double* base = &sm[warp & 1][lane];
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kb = 0; kb < 8; kb += 4) {
      double a[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (SMEM) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[i][mt]) : "r"((unsigned)__cvta_generic_to_shared(base + ((kb + i) * 2 + mt) * 32 + (it & 3) * 512)));
          else a[i][mt] = areg[i][mt] + it;
        }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          dmma(q[mt][i & 1][0], q[mt][i & 1][1], a[i][mt], zf[kb + i]);
          if (STATS) {
            double d = a[i][mt] - ksh[mt];
            s1[mt][i & 1] += d;
            s2[mt][i & 1] = fma(d, d, s2[mt][i & 1]);
          }
        }
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) s += q[a][b][0] + q[a][b][1] + s1[a][b] + s2[a][b];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <bool SMEM, bool STATS>
void run(int wps, double* out, long long* cyc, const char* name) {
  const int iters = 4000;
  for (int r = 0; r < 2; ++r) { k<SMEM, STATS><<<148, wps * 128>>>(out, iters, cyc); cudaDeviceSynchronize(); }
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / ((double)iters * 16) / wps;
  printf("%-28s warps/SMSP=%d : %.1f clk per step per SMSP (ideal %d)\n", name, wps, per, STATS ? 22 : 16);
}
int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  long long* cyc; cudaMalloc(&cyc, 8);
  for (int w : {1, 2}) {
    run<false, false>(w, out, cyc, "DMMA only, registers");
    run<false, true>(w, out, cyc, "DMMA + stats, registers");
    run<true, false>(w, out, cyc, "DMMA only, shared loads");
    run<true, true>(w, out, cyc, "DMMA + stats, shared loads");
  }
  return 0;
}
