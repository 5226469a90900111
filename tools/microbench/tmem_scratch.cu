// tmem_scratch.cu -- can TMEM (256 KB per SM) be used as a plain FP64 stash between two warps of the same SM
// sub-partition?  Warp w (0..3) writes a pattern with tcgen05.st, warp 4+w reads it back with tcgen05.ld
// after an mbarrier hand-off; checks values and times the st / ld throughput.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tmem_scratch.cu -o tmem_scratch
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mb_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mb_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const double (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
               "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3]))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, double (&v)[4]) {
  unsigned r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __hiloint2double(r[2 * i + 1], r[2 * i]);
}

__global__ void __launch_bounds__(256, 1) k(int iters, int* errors, long long* cyc) {
  __shared__ unsigned tmem_base;
  __shared__ __align__(8) unsigned long long bars[8];  // full[4], free[4]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mb_init(sa(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned base = tmem_base;
  const int q = warp & 3;
  const unsigned lane_base = base + ((unsigned)(32 * q) << 16);
  int bad = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (warp < 4) {
      if (it > 0) mb_wait(sa(&bars[4 + q]), (unsigned)((it - 1) & 1));
      // 16 "m-tiles" x 8 columns = 128 columns of this quarter's buffer (it & 3)
#pragma unroll
      for (int mt = 0; mt < 16; ++mt) {
        double v[4];
        for (int i = 0; i < 4; ++i) v[i] = (double)it * 1000.0 + warp * 100.0 + mt * 4 + i + lane * 1e-3;
        tmem_st8(lane_base + (unsigned)((it & 3) * 128 + mt * 8), v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mb_arrive(sa(&bars[q]));
    } else {
      mb_wait(sa(&bars[q]), (unsigned)(it & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int mt = 0; mt < 16; ++mt) {
        double v[4];
        tmem_ld8(lane_base + (unsigned)((it & 3) * 128 + mt * 8), v);
        for (int i = 0; i < 4; ++i) {
          double want = (double)it * 1000.0 + (warp - 4) * 100.0 + mt * 4 + i + lane * 1e-3;
          if (v[i] != want) ++bad;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mb_arrive(sa(&bars[4 + q]));
    }
  }
  long long t1 = clock64();
  if (bad) atomicAdd(errors, bad);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

int main() {
  int* err; long long* cyc;
  cudaMalloc(&err, 4); cudaMalloc(&cyc, 8); cudaMemset(err, 0, 4);
  const int iters = 2000;
  k<<<148, 256>>>(iters, err, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  int h; long long c;
  cudaMemcpy(&h, err, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("status %s, mismatches %d, %.1f cycles per hand-off of 16 KB per warp pair (%d iterations)\n", cudaGetErrorString(e), h, (double)c / iters, iters);
  return 0;
}
