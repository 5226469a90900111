// tma_stream.cu -- how fast can 148 persistent CTAs stream a row-major T x N FP64 panel when every CTA
// pulls (R rows x CW columns) pieces with TMA, i.e. row segments of only CW*8 bytes?  This is the access
// pattern of the single-read fused 3PRF sweep (a piece must cover whole columns across a CTA group).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tma_stream.cu -o tma_stream
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef long long i64;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned s_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_addr(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(void* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_addr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_addr(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* b, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(s_addr(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, void* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   s_addr(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(s_addr(bar))
               : "memory");
}

constexpr int SLOT_BYTES = 65536;

// CTA c: group g = c / G handles column tiles g, g+ngroups, ...; part p = c % G handles rows [p*R, (p+1)*R)
template <int NSLOT>
__global__ void __launch_bounds__(512, 1)
k_stream(const __grid_constant__ CUtensorMap map, int G, int R, int CW, int box_rows, i64 ntiles, int consume, double* out, int BW) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[NSLOT], empty[NSLOT];
  const int tid = threadIdx.x;
  const int ngroups = gridDim.x / G;
  const int g = blockIdx.x / G, p = blockIdx.x % G;
  if (g >= ngroups) return;
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 512 / 32 - 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const i64 my_tiles = (ntiles - g + ngroups - 1) / ngroups;
  const int nbox = R / box_rows;
  const unsigned piece_bytes = (unsigned)R * CW * 8;
  double acc = 0.0;
  if (tid == 0) {
    for (i64 i = 0; i < my_tiles; ++i) {
      const int s = (int)(i % NSLOT);
      const unsigned ph = (unsigned)((i / NSLOT) & 1);
      if (i >= NSLOT) mbar_wait(&empty[s], ph ^ 1);
      mbar_expect(&full[s], piece_bytes);
      const i64 ct = g + i * ngroups;
      unsigned char* dst = smem + (size_t)s * SLOT_BYTES;
      for (int b = 0; b < nbox; ++b)
        for (int h = 0; h < CW / BW; ++h) {
          tma_2d(dst, &map, (int)(ct * CW + h * BW), p * R + b * box_rows, &full[s]);
          dst += (size_t)box_rows * BW * 8;
        }
    }
  }
  // consumers (all threads, including thread 0 after it has issued everything it can... keep it simple:
  // thread 0's warp consumes too, but only after the producer loop; with NSLOT slots this would deadlock,
  // so warp 0 is producer-only and arrives on `empty` without reading)
  if (tid >= 32) {
    for (i64 i = 0; i < my_tiles; ++i) {
      const int s = (int)(i % NSLOT);
      const unsigned ph = (unsigned)((i / NSLOT) & 1);
      {
        mbar_wait(&full[s], ph);
        if (consume) {
          const double* d = reinterpret_cast<const double*>(smem + (size_t)s * SLOT_BYTES);
          const int n = R * CW;
          for (int e = tid - 32; e < n; e += 480) acc += d[e];
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }
  }
  if (consume) out[blockIdx.x * 512 + tid] = acc;
}

int main(int argc, char** argv) {
  const i64 T = 8192, N = argc > 1 ? atoll(argv[1]) : 65536;
  double* X;
  cudaMalloc(&X, (size_t)T * N * 8);
  cudaMemset(X, 0, (size_t)T * N * 8);
  double* out; cudaMalloc(&out, 148 * 512 * 8);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  cudaFuncSetAttribute(k_stream<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * SLOT_BYTES);
  // per-SM vs chip-wide limit: the same 128-byte-row stream on fewer and fewer CTAs
  {
    const int G = 16, R = 512, CW = 16, box_rows = 256;
    int grids[] = {144, 128, 112, 96, 64, 32, 16};
    for (int grid : grids) {
      CUtensorMap map;
      cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)T};
      cuuint64_t strides[1] = {(cuuint64_t)N * 8};
      cuuint32_t box[2] = {(cuuint32_t)CW, (cuuint32_t)box_rows};
      cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, X, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
      const i64 ntiles = N / CW;
      cudaEvent_t a, b;
      cudaEventCreate(&a); cudaEventCreate(&b);
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        k_stream<3><<<grid, 512, 3 * SLOT_BYTES>>>(map, G, R, CW, box_rows, ntiles, 0, out, CW);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
      }
      printf("G=16 R=512 CW=16 grid=%3d : %.3f ms  %.0f GB/s  %.1f GB/s per SM\n", grid, best, (double)T * N * 8 / best / 1e6, (double)T * N * 8 / best / 1e6 / grid);
    }
  }
  return 0;
}
