// phase_a_mix3.cu -- phase A of the 3PRF sweep in isolation, in the code shapes tried for tprf_fused.cuh: one warp per
// SM sub-partition reads its 128 rows x 16 columns of a 64 KiB swizzled slot and does, per 32 elements, one DMMA + d = x - k,
// s1 += d, s2 += d*d.  Ideal: 64 steps x 22 = 1408 clocks per piece.  Prints clocks per piece for each shape.
#include <cstdio>
#include <type_traits>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double ldp(unsigned addr) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }
__device__ __forceinline__ double2 ldp2(unsigned addr) { double2 v; asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)); return v; }
__device__ __forceinline__ unsigned tok() { unsigned t; asm volatile("mov.u32 %0, 0;" : "=r"(t) :: "memory"); return t; }
__device__ __forceinline__ int col_of_m(int m) { return ((m & 2) << 2) | (m & 1) | ((m & 4) >> 1); }
constexpr int KS = 32;
// SHAPE 0: x into DMMA, d after (round 1)   1: d into DMMA, loads one group ahead, fence per group
//       2: as 1 without fence                3: LDS.128 (columns 2j, 2j+1 -> m-tiles even / odd), d into DMMA, fence
//       4: as 1, loads two groups ahead      5: as 3, loads two groups ahead
template <int SHAPE>
__global__ void __launch_bounds__(128, 1) k(double* out, int pieces, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sm = reinterpret_cast<double*>(smem);
  for (int i = threadIdx.x; i < 3 * 8192; i += blockDim.x) sm[i] = (i % 977) * 1e-3;
  __syncthreads();
  const int fm = lane >> 2, fk = lane & 3;
  const int rowbase = warp * 128;
  unsigned offa[2][2]; int cola[2];
  for (int mt = 0; mt < 2; ++mt) {
    cola[mt] = col_of_m(fm) + 4 * mt;
    for (int pb = 0; pb < 2; ++pb)
      offa[mt][pb] = (unsigned)(((((cola[mt] >> 1) ^ (4 * pb + fk)) & 7) << 4) | ((cola[mt] & 1) << 3)) + (unsigned)(rowbase + 4 * pb + fk) * 128u;
  }
  // LDS.128 shape: lane (fm, fk) reads row 4pb + fk, 16-byte chunk c(fm): quarter-warps (fm pairs) must hit 8 distinct chunks
  const int chunk = ((fm & 1) << 2) | (fm >> 1);   // fm 0,1 -> chunks 0,4 ; 2,3 -> 1,5 ; ...
  unsigned offv[2];
  for (int pb = 0; pb < 2; ++pb) offv[pb] = (unsigned)(((chunk ^ (4 * pb + fk)) & 7) << 4) + (unsigned)(rowbase + 4 * pb + fk) * 128u;
  double zf[KS];
  for (int i = 0; i < KS; ++i) zf[i] = 1.0 + i + lane * 1e-3;
  double tot = 0.0;
  const unsigned base = (unsigned)__cvta_generic_to_shared(smem);
  long long t0 = clock64();
  for (int it = 0; it < pieces; ++it) {
    const unsigned slot = base + (unsigned)(it % 3) * 65536u + tok();
    double q[2][2][2] = {}, s1[2][2] = {}, s2[2][2] = {};
    const double ksh[2] = {ldp(slot + cola[0] * 8), ldp(slot + cola[1] * 8)};
    if (SHAPE == 0) {
#pragma unroll
      for (int kb = 0; kb < KS; kb += 4) {
        double a[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) a[i][mt] = ldp(slot + offa[mt][i & 1] + (unsigned)(kb + (i & ~1)) * 512u);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            dmma(q[mt][i & 1][0], q[mt][i & 1][1], a[i][mt], zf[kb + i]);
            double d = a[i][mt] - ksh[mt];
            s1[mt][i & 1] += d;
            s2[mt][i & 1] = fma(d, d, s2[mt][i & 1]);
          }
      }
    } else if (SHAPE == 1 || SHAPE == 2 || SHAPE == 4) {
      constexpr int AHEAD = SHAPE == 4 ? 2 : 1;
      double a[AHEAD + 1][2][2];
#pragma unroll
      for (int h = 0; h < AHEAD; ++h)
#pragma unroll
        for (int pb = 0; pb < 2; ++pb)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) a[h][pb][mt] = ldp(slot + offa[mt][pb] + (unsigned)h * 1024u);
#pragma unroll
      for (int gk = 0; gk < KS / 2; ++gk) {
        if (gk + AHEAD < KS / 2) {
#pragma unroll
          for (int pb = 0; pb < 2; ++pb)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) a[AHEAD][pb][mt] = ldp(slot + offa[mt][pb] + (unsigned)(gk + AHEAD) * 1024u);
        }
        double d[2][2];
#pragma unroll
        for (int pb = 0; pb < 2; ++pb)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) d[pb][mt] = a[0][pb][mt] - ksh[mt];
#pragma unroll
        for (int pb = 0; pb < 2; ++pb)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            dmma(q[mt][pb][0], q[mt][pb][1], d[pb][mt], zf[2 * gk + pb]);
            s1[mt][pb] += d[pb][mt];
            s2[mt][pb] = fma(d[pb][mt], d[pb][mt], s2[mt][pb]);
          }
        if (SHAPE != 2) __syncwarp();
#pragma unroll
        for (int h = 0; h < AHEAD; ++h)
#pragma unroll
          for (int pb = 0; pb < 2; ++pb)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) a[h][pb][mt] = a[h + 1][pb][mt];
      }
    } else {
      constexpr int AHEAD = SHAPE == 5 ? 2 : 1;
      double2 a[AHEAD + 1][2];
#pragma unroll
      for (int h = 0; h < AHEAD; ++h)
#pragma unroll
        for (int pb = 0; pb < 2; ++pb) a[h][pb] = ldp2(slot + offv[pb] + (unsigned)h * 1024u);
#pragma unroll
      for (int gk = 0; gk < KS / 2; ++gk) {
        if (gk + AHEAD < KS / 2) {
#pragma unroll
          for (int pb = 0; pb < 2; ++pb) a[AHEAD][pb] = ldp2(slot + offv[pb] + (unsigned)(gk + AHEAD) * 1024u);
        }
        double d[2][2];
#pragma unroll
        for (int pb = 0; pb < 2; ++pb) { d[pb][0] = a[0][pb].x - ksh[0]; d[pb][1] = a[0][pb].y - ksh[1]; }
#pragma unroll
        for (int pb = 0; pb < 2; ++pb)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            dmma(q[mt][pb][0], q[mt][pb][1], d[pb][mt], zf[2 * gk + pb]);
            s1[mt][pb] += d[pb][mt];
            s2[mt][pb] = fma(d[pb][mt], d[pb][mt], s2[mt][pb]);
          }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < AHEAD; ++h)
#pragma unroll
          for (int pb = 0; pb < 2; ++pb) a[h][pb] = a[h + 1][pb];
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int pb = 0; pb < 2; ++pb) tot += q[mt][pb][0] + q[mt][pb][1] + s1[mt][pb] + s2[mt][pb];
  }
  long long t1 = clock64();
  out[blockIdx.x * 128 + threadIdx.x] = tot;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int SHAPE>
void run(double* out, long long* cyc, const char* name) {
  const int pieces = 600;
  cudaFuncSetAttribute(k<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536);
  for (int r = 0; r < 2; ++r) { k<SHAPE><<<148, 128, 3 * 65536>>>(out, pieces, cyc); cudaDeviceSynchronize(); }
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("shape %d %-58s: %.0f clk per piece (ideal 1408), %.1f per step  %s\n", SHAPE, name, (double)h / pieces, (double)h / pieces / 64, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double* out; cudaMalloc(&out, 148 * 128 * 8);
  long long* cyc; cudaMalloc(&cyc, 8);
  run<0>(out, cyc, "x into DMMA, d after (round 1)");
  run<1>(out, cyc, "d into DMMA, loads 1 group ahead, fence");
  run<2>(out, cyc, "d into DMMA, loads 1 group ahead, no fence");
  run<4>(out, cyc, "d into DMMA, loads 2 groups ahead, fence");
  run<3>(out, cyc, "LDS.128, d into DMMA, loads 1 group ahead, fence");
  run<5>(out, cyc, "LDS.128, d into DMMA, loads 2 groups ahead, fence");
  return 0;
}
