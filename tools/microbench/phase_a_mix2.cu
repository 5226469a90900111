// phase_a_mix2.cu -- scheduling variants of the phase-A mix (1 warp per SM sub-partition, operands from shared
// memory): V0 batch of 4 k-steps (as the kernel), V1 double-buffered batches (loads of b+1 before math of b),
// V2 batch of 8, V3 V1 + stats deferred by one batch.  Ideal 22 clk per (DMMA + 3 D-ops).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds(unsigned a) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
template <int V>
__global__ void k(double* out, int iters, long long* cyc) {
  __shared__ double sm[2][2560];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * 2560; i += blockDim.x) (&sm[0][0])[i] = i * 1e-4;
  __syncthreads();
  double q[2][2][2] = {}, s1[2][2] = {}, s2[2][2] = {}, zf[32];
  for (int i = 0; i < 32; ++i) zf[i] = 1.0 + i + lane * 1e-3;
  const double ksh[2] = {0.5, 0.25};
  const unsigned base0 = (unsigned)__cvta_generic_to_shared(&sm[warp & 1][lane]);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const unsigned base = base0 + (it & 1) * 8;   // defeat hoisting across iterations
    auto ld = [&](int ks, int mt) { return lds(base + ((ks * 2 + mt) * 32) * 8); };
    auto math = [&](int ks, int mt, double a) {
      dmma(q[mt][ks & 1][0], q[mt][ks & 1][1], a, zf[ks]);
      double d = a - ksh[mt];
      s1[mt][ks & 1] += d;
      s2[mt][ks & 1] = fma(d, d, s2[mt][ks & 1]);
    };
    constexpr int KS = 32;
    if (V == 0) {
#pragma unroll
      for (int kb = 0; kb < KS; kb += 4) {
        double a[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) for (int mt = 0; mt < 2; ++mt) a[i][mt] = ld(kb + i, mt);
#pragma unroll
        for (int i = 0; i < 4; ++i) for (int mt = 0; mt < 2; ++mt) math(kb + i, mt, a[i][mt]);
      }
    } else if (V == 1) {
      double a[2][4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) for (int mt = 0; mt < 2; ++mt) a[0][i][mt] = ld(i, mt);
#pragma unroll
      for (int kb = 0; kb < KS; kb += 4) {
        const int cur = (kb / 4) & 1;
        if (kb + 4 < KS) {
#pragma unroll
          for (int i = 0; i < 4; ++i) for (int mt = 0; mt < 2; ++mt) a[cur ^ 1][i][mt] = ld(kb + 4 + i, mt);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) for (int mt = 0; mt < 2; ++mt) math(kb + i, mt, a[cur][i][mt]);
      }
    } else if (V == 2) {
#pragma unroll
      for (int kb = 0; kb < KS; kb += 8) {
        double a[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) for (int mt = 0; mt < 2; ++mt) a[i][mt] = ld(kb + i, mt);
#pragma unroll
        for (int i = 0; i < 8; ++i) for (int mt = 0; mt < 2; ++mt) math(kb + i, mt, a[i][mt]);
      }
    } else {   // V3: all 64 loads up front (register-rich), then the math
      double a[32][2];
#pragma unroll
      for (int i = 0; i < 32; ++i) for (int mt = 0; mt < 2; ++mt) a[i][mt] = ld(i, mt);
#pragma unroll
      for (int i = 0; i < 32; ++i) for (int mt = 0; mt < 2; ++mt) math(i, mt, a[i][mt]);
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) s += q[a][b][0] + q[a][b][1] + s1[a][b] + s2[a][b];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int V>
void run(int wps, double* out, long long* cyc) {
  const int iters = 2000;
  for (int r = 0; r < 2; ++r) { k<V><<<148, wps * 128>>>(out, iters, cyc); cudaDeviceSynchronize(); }
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("variant %d warps/SMSP=%d : %.1f clk per step per SMSP (ideal 22)  %s\n", V, wps, (double)h / ((double)iters * 64) / wps, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  long long* cyc; cudaMalloc(&cyc, 8);
  for (int w : {1, 2}) { run<0>(w, out, cyc); run<1>(w, out, cyc); run<2>(w, out, cyc); run<3>(w, out, cyc); }
  return 0;
}
