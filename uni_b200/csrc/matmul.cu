// matmul.cu -- `a *@ b` for Mat[Double] (FP64 tensor pipe: DMMA m8n8k4) and Mat[Float].
//
// FP64: tcgen05 has no f64 kind, so the FP64 tensor path on sm_100a is the warp-level
// `mma.sync.aligned.m8n8k4.f64` (SASS DMMA.8x8x4).  CTA tile 128x128x16, 8 warps as 2x4
// (warp tile 64x32 = 8x4 DMMA tiles, 64 accumulator doubles per lane), 3-stage cp.async
// pipeline into padded shared memory whose pitches (20 / 132 doubles) make every fragment
// load bank-conflict free for both operand orientations.  Operands are read THROUGH their
// stride descriptor: a transposed view (rs == 1) is staged with its contiguous dimension
// along the shared-memory row, so `m.T *@ m` costs no copy (Mat.scala:2815-2881; the
// reference copies non-BLAS-safe operands first, Mat.scala:1169).
//
// FP32: register-tiled FFMA kernel (float accumulate, as multiplyFloat Mat.scala:3236-3268).
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace um {

// Two CTA tiles: 128 x 128 (one CTA per SM) and 64 x 64 (warp tile 32 x 16, two CTAs per SM).  DMMA is slow enough
// (16 FMA per clock per sub-partition) that the small tile loses nothing in the inner loop; it exists because a
// 1024^2 product is only 64 large tiles on 148 SMs and a 2048^2 one is 1.73 waves of them (launch_dgemm picks).
constexpr int MM_BK = 16, MM_STAGES = 3, MM_THREADS = 256;
constexpr int MM_PK = MM_BK + 4;     // pitch when the k dimension is contiguous (20 doubles)
template <int TILE> struct MMShape {
  static constexpr int PMN = TILE + 4;                 // pitch when the m/n dimension is contiguous (132 / 68 doubles)
  static constexpr int TILE_ELEMS = TILE * MM_PK;      // 2560 >= 16 * 132, 1280 >= 16 * 68: one operand stage
  static constexpr int SMEM_BYTES = MM_STAGES * 2 * TILE_ELEMS * 8;  // 122880 / 61440
  static constexpr int MI = TILE / 16, NJ = TILE / 32; // DMMA tiles per warp (2 x 4 warps)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Stage one operand tile.  The operand is a (outer x inner) box where `inner` is the
// memory-contiguous dimension: KMAJOR -> outer = m (or n), inner = k (box 128 x 16, pitch 20);
// otherwise outer = k, inner = m (or n) (box 16 x 128, pitch 132).  g(o, i) = p[o*ld + i].
template <int TILE, bool KMAJOR, bool VEC16>
__device__ __forceinline__ void stage_tile(double* sm, const double* __restrict__ p, i64 ld, i64 o0, i64 i0,
                                           i64 o_lim, i64 i_lim, int tid) {
  constexpr int OUTER = KMAJOR ? TILE : MM_BK;
  constexpr int INNER = KMAJOR ? MM_BK : TILE;
  constexpr int PITCH = KMAJOR ? MM_PK : MMShape<TILE>::PMN;
  constexpr int CH = VEC16 ? 2 : 1;               // doubles per cp.async
  constexpr int CPR = INNER / CH;                 // chunks per box row
  constexpr int TOTAL = OUTER * CPR;
#pragma unroll
  for (int c = tid; c < TOTAL; c += MM_THREADS) {
    int o = c / CPR, i = (c % CPR) * CH;
    i64 go = o0 + o, gi = i0 + i;
    i64 rem = (go < o_lim) ? (i_lim - gi) : 0;    // valid elements from gi on
    int nvalid = rem <= 0 ? 0 : (rem >= CH ? CH : (int)rem);
    const double* src = nvalid ? (p + go * ld + gi) : p;
    double* dst = sm + o * PITCH + i;
    if (VEC16) cp_async16(dst, src, nvalid * 8);
    else cp_async8(dst, src, nvalid * 8);
  }
}

// One pass of stage_tile's loop (iteration `it` of TOTAL / MM_THREADS): the main loop spreads the copies of the next
// k-tile between the DMMA batches of the current one, so the two warps of a sub-partition never spend the start of a
// k-tile on address arithmetic together with the tensor pipe idle.
template <int TILE, bool KMAJOR, bool VEC16>
__device__ __forceinline__ void stage_tile_part(double* sm, const double* __restrict__ p, i64 ld, i64 o0, i64 i0,
                                                i64 o_lim, i64 i_lim, int tid, int it) {
  constexpr int INNER = KMAJOR ? MM_BK : TILE;
  constexpr int PITCH = KMAJOR ? MM_PK : MMShape<TILE>::PMN;
  constexpr int CH = VEC16 ? 2 : 1;
  constexpr int CPR = INNER / CH;
  const int c = tid + it * MM_THREADS;
  int o = c / CPR, i = (c % CPR) * CH;
  i64 go = o0 + o, gi = i0 + i;
  i64 rem = (go < o_lim) ? (i_lim - gi) : 0;
  int nvalid = rem <= 0 ? 0 : (rem >= CH ? CH : (int)rem);
  const double* src = nvalid ? (p + go * ld + gi) : p;
  double* dst = sm + o * PITCH + i;
  if (VEC16) cp_async16(dst, src, nvalid * 8);
  else cp_async8(dst, src, nvalid * 8);
}

// C[m][n] = sum_k A(m,k) B(k,n).  A(m,k) = pa[m*a_rs + k*a_cs] with one of the strides == 1,
// same for B.  AK: A is k-contiguous (a_cs == 1); BK_: B is k-contiguous (b_rs == 1).
template <int TILE, bool AK, bool BKM, bool VEC16, bool SPLIT>
__global__ void __launch_bounds__(MM_THREADS, TILE == 128 ? 1 : 2)
k_dgemm(const double* __restrict__ pa, i64 lda, const double* __restrict__ pb, i64 ldb, double* __restrict__ pc, i64 ldc,
        i64 M, i64 N, i64 K, int kt_per, i64 split_stride) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps
  using S = MMShape<TILE>;
  constexpr int MM_PMN = S::PMN, MM_TILE_ELEMS = S::TILE_ELEMS, MI = S::MI, NJ = S::NJ;
  const i64 m0 = (i64)blockIdx.y * TILE, n0 = (i64)blockIdx.x * TILE;
  // split-K (few output tiles, long K): blockIdx.z takes k-tiles [kt0, kt1) and writes its own partial C
  const int KT_all = (int)((K + MM_BK - 1) / MM_BK);
  // (a separate instantiation: the extra index arithmetic cost the unsplit kernel 3 % at 4096^2)
  const int kt0 = SPLIT ? (int)blockIdx.z * kt_per : 0;
  const int kt1 = SPLIT ? (kt0 + kt_per < KT_all ? kt0 + kt_per : KT_all) : KT_all;
  if (SPLIT) pc += (i64)blockIdx.z * split_stride;

  auto stage = [&](int s, int kt) {
    double* sa = smem + (size_t)s * 2 * MM_TILE_ELEMS;
    double* sb = sa + MM_TILE_ELEMS;
    const i64 k0 = (i64)kt * MM_BK;
    if (AK) stage_tile<TILE, true, VEC16>(sa, pa, lda, m0, k0, M, K, tid);     // outer m, inner k
    else stage_tile<TILE, false, VEC16>(sa, pa, lda, k0, m0, K, M, tid);       // outer k, inner m
    if (BKM) stage_tile<TILE, true, VEC16>(sb, pb, ldb, n0, k0, N, K, tid);    // outer n, inner k
    else stage_tile<TILE, false, VEC16>(sb, pb, ldb, k0, n0, K, N, tid);       // outer k, inner n
  };

  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < MM_STAGES - 1; ++s) {
    if (kt0 + s < kt1) stage(s, kt0 + s);
    cp_async_commit();
  }

  // copies of one operand tile per thread, and how many of the 2 * ITERS (A then B) go after each of the 4 k-steps
  constexpr int ITERS = TILE * MM_BK / (VEC16 ? 2 : 1) / MM_THREADS;
  constexpr int PER = 2 * ITERS / (MM_BK / 4);
  static_assert(2 * ITERS % (MM_BK / 4) == 0, "staging units must split evenly over the k-steps");
  auto stage_part = [&](int s, int kt, int u) {
    double* sa = smem + (size_t)s * 2 * MM_TILE_ELEMS;
    double* sb = sa + MM_TILE_ELEMS;
    const i64 k0 = (i64)kt * MM_BK;
    if (u < ITERS) {
      if (AK) stage_tile_part<TILE, true, VEC16>(sa, pa, lda, m0, k0, M, K, tid, u);
      else stage_tile_part<TILE, false, VEC16>(sa, pa, lda, k0, m0, K, M, tid, u);
    } else {
      if (BKM) stage_tile_part<TILE, true, VEC16>(sb, pb, ldb, n0, k0, N, K, tid, u - ITERS);
      else stage_tile_part<TILE, false, VEC16>(sb, pb, ldb, k0, n0, K, N, tid, u - ITERS);
    }
  };

  const int fr = lane >> 2, fk = lane & 3;  // fragment row (m or n) and k index
  for (int kt = kt0; kt < kt1; ++kt) {
    cp_async_wait<MM_STAGES - 2>();
    __syncthreads();
    const int nk = kt + MM_STAGES - 1;
    const bool more = nk < kt1;
    const int ns = (nk - kt0) % MM_STAGES;   // the stage read in iteration kt - 1: free since the barrier above
    const double* sa = smem + (size_t)((kt - kt0) % MM_STAGES) * 2 * MM_TILE_ELEMS;
    const double* sb = sa + MM_TILE_ELEMS;
#pragma unroll
    for (int kk = 0; kk < MM_BK; kk += 4) {
      double af[MI], bf[NJ];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        int m = wm * (TILE / 2) + i * 8 + fr;
        af[i] = AK ? sa[m * MM_PK + kk + fk] : sa[(kk + fk) * MM_PMN + m];
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int n = wn * (TILE / 4) + j * 8 + fr;
        bf[j] = BKM ? sb[n * MM_PK + kk + fk] : sb[(kk + fk) * MM_PMN + n];
      }
#pragma unroll
      for (int i = 0; i < MI; ++i) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        if (i == MI / 2 - 1 && more) {
#pragma unroll
          for (int u = 0; u < PER; ++u) stage_part(ns, nk, (kk / 4) * PER + u);
        }
      }
    }
    cp_async_commit();
  }
  cp_async_wait<0>();

  // epilogue: lane holds C[row = fr][col = 2*fk, 2*fk+1] of each 8x8 tile
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    i64 m = m0 + wm * (TILE / 2) + i * 8 + fr;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      i64 n = n0 + wn * (TILE / 4) + j * 8 + 2 * fk;
      double* dst = pc + m * ldc + n;
      if (n + 1 < N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
      } else {
        if (n < N) dst[0] = acc[i][j][0];
        if (n + 1 < N) dst[1] = acc[i][j][1];
      }
    }
  }
}

// generic-stride fallback (neither stride of an operand is 1): one thread per C element
template <class T>
__global__ void __launch_bounds__(256) k_gemm_naive(In<T> a, In<T> b, T* __restrict__ pc, i64 ldc, i64 M, i64 N, i64 K) {
  i64 n = (i64)blockIdx.x * 16 + (threadIdx.x & 15);
  i64 m = (i64)blockIdx.y * 16 + (threadIdx.x >> 4);
  if (m >= M || n >= N) return;
  T s = 0;
  for (i64 k = 0; k < K; ++k) s += a.p[m * a.rs + k * a.cs] * b.p[k * b.rs + n * b.cs];
  pc[m * ldc + n] = s;
}

// ---- FP32: 128x128x8 tiles, 8x8 register micro-tile, float accumulate -----------------------------
constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8;
__global__ void __launch_bounds__(256, 2)
k_sgemm(In<float> a, In<float> b, float* __restrict__ pc, i64 ldc, i64 M, i64 N, i64 K) {
  __shared__ float As[2][SG_BK][SG_BM + 4];
  __shared__ float Bs[2][SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8x8 outputs (strided by 16)
  const i64 m0 = (i64)blockIdx.y * SG_BM, n0 = (i64)blockIdx.x * SG_BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  // loader mapping: 1024 elements per operand tile, 4 per thread.  Pick the thread->element
  // map so that the memory-contiguous dimension runs along consecutive threads.
  const bool a_kc = a.cs == 1;  // A contiguous along k
  const bool b_nc = b.cs == 1;  // B contiguous along n
  float ra[4], rb[4];
  auto gload = [&](i64 k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int e = tid + q * 256;
      int am, ak;
      if (a_kc) { ak = e & 7; am = e >> 3; } else { am = e & 127; ak = e >> 7; }
      i64 gm = m0 + am, gk = k0 + ak;
      ra[q] = (gm < M && gk < K) ? a.p[gm * a.rs + gk * a.cs] : 0.f;
      int bn, bk;
      if (b_nc) { bn = e & 127; bk = e >> 7; } else { bk = e & 7; bn = e >> 3; }
      i64 gn = n0 + bn;
      gk = k0 + bk;
      rb[q] = (gn < N && gk < K) ? b.p[gk * b.rs + gn * b.cs] : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int e = tid + q * 256;
      int am, ak;
      if (a_kc) { ak = e & 7; am = e >> 3; } else { am = e & 127; ak = e >> 7; }
      As[buf][ak][am] = ra[q];
      int bn, bk;
      if (b_nc) { bn = e & 127; bk = e >> 7; } else { bk = e & 7; bn = e >> 3; }
      Bs[buf][bk][bn] = rb[q];
    }
  };
  const int KT = (int)((K + SG_BK - 1) / SG_BK);
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) gload((i64)(kt + 1) * SG_BK);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float av[8], bv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = As[buf][k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) bv[j] = Bs[buf][k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < KT) sstore(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    i64 m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      i64 n = n0 + tx + 16 * j;
      if (n < N) pc[m * ldc + n] = acc[i][j];
    }
  }
}

static std::atomic<int> g_f64_tile{0};   // um_matmul_set_f64_tile: 0 = automatic, 64 / 128 forced (tests, profiling)

template <int TILE, bool AK, bool BKM, bool VEC16>
static int launch_dgemm_tile(const double* pa, i64 lda, const double* pb, i64 ldb, double* pc, i64 ldc, i64 M, i64 N, i64 K,
                             int splits, int kt_per, i64 split_stride) {
  dim3 grid((unsigned)((N + TILE - 1) / TILE), (unsigned)((M + TILE - 1) / TILE), (unsigned)splits);
  static std::once_flag once;   // one device per process (um_init)
  cudaError_t attr = cudaSuccess;
  std::call_once(once, [&] {
    attr = cudaFuncSetAttribute(k_dgemm<TILE, AK, BKM, VEC16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MMShape<TILE>::SMEM_BYTES);
    if (attr == cudaSuccess)
      attr = cudaFuncSetAttribute(k_dgemm<TILE, AK, BKM, VEC16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MMShape<TILE>::SMEM_BYTES);
  });
  UM_CUDA(attr);
  if (splits > 1)
    k_dgemm<TILE, AK, BKM, VEC16, true><<<grid, MM_THREADS, MMShape<TILE>::SMEM_BYTES, ctx()->stream>>>(pa, lda, pb, ldb, pc, ldc, M, N, K, kt_per,
                                                                                                        split_stride);
  else
    k_dgemm<TILE, AK, BKM, VEC16, false><<<grid, MM_THREADS, MMShape<TILE>::SMEM_BYTES, ctx()->stream>>>(pa, lda, pb, ldb, pc, ldc, M, N, K, kt_per,
                                                                                                         split_stride);
  UM_LAUNCHED();
  return UM_OK;
}

// Tile choice: time ~ waves x work per tile.  128 x 128 tiles run one per SM; 64 x 64 tiles run two per SM sharing
// its FP64 pipe, i.e. a quarter of the work each at (about) the same per-SM rate, with a few per cent more L2 traffic
// and prologue per flop.  The small tile wins whenever the large one leaves SMs idle or its last wave mostly empty.
static int pick_dgemm_tile(i64 M, i64 N) {
  const int forced = g_f64_tile.load();
  if (forced == 64 || forced == 128) return forced;
  const i64 sms = ctx()->sm_count > 0 ? ctx()->sm_count : 148;
  const i64 t128 = ((M + 127) / 128) * ((N + 127) / 128), t64 = ((M + 63) / 64) * ((N + 63) / 64);
  const double c128 = 4.0 * (double)((t128 + sms - 1) / sms);
  const double c64 = 1.06 * (double)((t64 + sms - 1) / sms);
  return c64 < c128 ? 64 : 128;
}

template <bool AK, bool BKM>
static int launch_dgemm_split(int tile, const double* pa, i64 lda, const double* pb, i64 ldb, double* pc, i64 ldc, i64 M, i64 N, i64 K,
                              bool vec16, int splits, int kt_per, i64 split_stride) {
  if (tile == 64) {
    if (vec16) return launch_dgemm_tile<64, AK, BKM, true>(pa, lda, pb, ldb, pc, ldc, M, N, K, splits, kt_per, split_stride);
    return launch_dgemm_tile<64, AK, BKM, false>(pa, lda, pb, ldb, pc, ldc, M, N, K, splits, kt_per, split_stride);
  }
  if (vec16) return launch_dgemm_tile<128, AK, BKM, true>(pa, lda, pb, ldb, pc, ldc, M, N, K, splits, kt_per, split_stride);
  return launch_dgemm_tile<128, AK, BKM, false>(pa, lda, pb, ldb, pc, ldc, M, N, K, splits, kt_per, split_stride);
}

// Few output tiles and a long contraction (A' A of a tall matrix, the normal equations of an OLS: the commonest
// product in the reference's statistics code) would leave the whole of K to a handful of CTAs -- ONE for an L x L
// Gram, 2.2 us per 16 k.  Such products are split along K over blockIdx.z into partial C's, summed by the library's
// fixed-tree axis reduction (run-to-run reproducible; the split depends only on the shape and the SM count).
template <bool AK, bool BKM>
static int launch_dgemm(const double* pa, i64 lda, const double* pb, i64 ldb, double* pc, i64 ldc, i64 M, i64 N, i64 K,
                        bool vec16) {
  const int tile = pick_dgemm_tile(M, N);
  const i64 sms = ctx()->sm_count > 0 ? ctx()->sm_count : 148;
  const i64 tiles = ((M + tile - 1) / tile) * ((N + tile - 1) / tile);
  const i64 target = (tile == 64 ? 2 : 1) * sms;       // resident CTAs
  const int KT = (int)((K + MM_BK - 1) / MM_BK);
  i64 splits = 1;
  if (ldc == N && K >= 2048 && 2 * tiles <= target) {
    splits = target / tiles;
    if (splits > K / 512) splits = K / 512;            // at least 32 k-tiles per split
    if (splits > 256) splits = 256;
    while (splits > 1 && (double)splits * (double)M * (double)N * 8.0 > 268435456.0) --splits;
  }
  if (splits <= 1) return launch_dgemm_split<AK, BKM>(tile, pa, lda, pb, ldb, pc, ldc, M, N, K, vec16, 1, KT, 0);
  const int kt_per = (int)((KT + splits - 1) / splits);
  splits = (KT + kt_per - 1) / kt_per;
  void* part = nullptr;
  int s = um_malloc((size_t)splits * (size_t)M * (size_t)N * 8, &part);
  if (s != UM_OK) return launch_dgemm_split<AK, BKM>(tile, pa, lda, pb, ldb, pc, ldc, M, N, K, vec16, 1, KT, 0);   // no room: unsplit
  s = launch_dgemm_split<AK, BKM>(tile, pa, lda, pb, ldb, static_cast<double*>(part), N, M, N, K, vec16, (int)splits, kt_per, M * N);
  if (s == UM_OK) {
    um_view pv, ov;
    um_view_init(&pv, part, splits, M * N, UM_F64);
    um_view_init(&ov, pc, 1, M * N, UM_F64);
    s = um_reduce_axis(UM_SUM, &pv, 0, &ov);
  }
  um_free(part);   // stream-ordered
  return s;
}

}  // namespace um

#include "matmul_tf32.cuh"

using namespace um;

static int g_f32_path = 0;   // um_matmul_set_f32_path

extern "C" int32_t um_matmul_set_f32_path(int32_t path) {
  if (path < 0 || path > 3) return um::fail(UM_ERR_ARG, "um_matmul_set_f32_path: 0 auto, 1 CUDA cores, 2 tensor cores (operands split in shared memory), 3 tensor cores (operands split once in global memory)");
  g_f32_path = path;
  return UM_OK;
}

extern "C" int32_t um_matmul_set_f64_tile(int32_t tile) {
  if (tile != 0 && tile != 64 && tile != 128) return um::fail(UM_ERR_ARG, "um_matmul_set_f64_tile: 0 (automatic), 64 or 128");
  um::g_f64_tile.store(tile);
  return UM_OK;
}

extern "C" int32_t um_matmul(const um_view* a, const um_view* b, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(b); UM_VIEW(out);
  // require(m.cols == other.rows) -> IllegalArgumentException (Mat.scala:2816-2817)
  UM_REQUIRE(a->cols == b->rows, "matmul shape mismatch: %lldx%lld *@ %lldx%lld", (i64)a->rows, (i64)a->cols,
             (i64)b->rows, (i64)b->cols);
  UM_REQUIRE(out->rows == a->rows && out->cols == b->cols, "matmul output must be %lldx%lld", (i64)a->rows, (i64)b->cols);
  UM_REQUIRE(a->dtype == b->dtype && a->dtype == out->dtype, "matmul operands must share a dtype");
  UM_REQUIRE(out->cs == 1 || out->cols == 1, "matmul output must be row-contiguous");
  const i64 M = a->rows, N = b->cols, K = a->cols;
  if (M == 0 || N == 0) return UM_OK;
  if (K == 0) return um_fill(out, 0.0);
  const i64 ldc = out->rs;
  if (a->dtype == UM_F64) {
    In<double> A = in_of<double>(a), B = in_of<double>(b);
    double* pc = static_cast<double*>(out->data) + out->offset;
    // degenerate dimensions make either stride usable as "the contiguous one"
    bool a_k = A.cs == 1 || K == 1, a_m = A.rs == 1 || M == 1;
    bool b_n = B.cs == 1 || N == 1, b_k = B.rs == 1 || K == 1;
    if ((a_k || a_m) && (b_n || b_k)) {
      bool AK = a_k;
      bool BKM = !b_n;
      i64 lda = AK ? (M == 1 ? K : A.rs) : (K == 1 ? M : A.cs);
      i64 ldb = BKM ? (N == 1 ? K : B.cs) : (K == 1 ? N : B.rs);
      bool vec16 = aligned16(A.p) && aligned16(B.p) && (lda % 2 == 0) && (ldb % 2 == 0);
      if (AK && !BKM) return launch_dgemm<true, false>(A.p, lda, B.p, ldb, pc, ldc, M, N, K, vec16);
      if (AK && BKM) return launch_dgemm<true, true>(A.p, lda, B.p, ldb, pc, ldc, M, N, K, vec16);
      if (!AK && !BKM) return launch_dgemm<false, false>(A.p, lda, B.p, ldb, pc, ldc, M, N, K, vec16);
      return launch_dgemm<false, true>(A.p, lda, B.p, ldb, pc, ldc, M, N, K, vec16);
    }
    dim3 grid((unsigned)((N + 15) / 16), (unsigned)((M + 15) / 16));
    k_gemm_naive<double><<<grid, 256, 0, ctx()->stream>>>(A, B, pc, ldc, M, N, K);
    UM_LAUNCHED();
    return UM_OK;
  }
  if (a->dtype == UM_F32) {
    // Few output tiles and a long contraction (A' A of a tall MatF): both FP32 kernels would leave all of K to a handful
    // of CTAs (8 ms for 33 x 262144 x 33).  Such a product is promoted to FP64 and takes the split-K DMMA path above --
    // every FP32 product is exact in FP64 and the sum is rounded to FP32 once, which is at least as accurate as the
    // reference's float accumulation (Mat.scala:3236-3268).  Automatic choice only; the operands' copies are capped at 1 GiB.
    {
      const i64 sms = ctx()->sm_count > 0 ? ctx()->sm_count : 148;
      const i64 t128 = ((M + 127) / 128) * ((N + 127) / 128);
      if (g_f32_path == 0 && K >= 8192 && 2 * t128 <= sms && (double)(M + N) * (double)K * 8.0 <= 1073741824.0) {
        void* blk = nullptr;
        int s = um_malloc((size_t)(M * K + K * N + M * N) * 8, &blk);
        if (s == UM_OK) {
        double* p64 = static_cast<double*>(blk);
        um_view a64, b64, c64;
        um_view_init(&a64, p64, M, K, UM_F64);
        um_view_init(&b64, p64 + M * K, K, N, UM_F64);
        um_view_init(&c64, p64 + M * K + K * N, M, N, UM_F64);
        s = um_copy(a, &a64);
        if (s == UM_OK) s = um_copy(b, &b64);
        if (s == UM_OK) s = um_matmul(&a64, &b64, &c64);
        if (s == UM_OK) s = um_copy(&c64, out);
        um_free(blk);   // stream-ordered
        return s;
        }   // no room for the copies: the FP32 kernels below
      }
    }
    In<float> A = in_of<float>(a), B = in_of<float>(b);
    float* pc = static_cast<float*>(out->data) + out->offset;
    // row-major, 16-byte aligned operands: 3xTF32 on the tcgen05 tensor cores (matmul_tf32.cuh)
    static const bool tensor_ok = [] { const char* e = getenv("UM_MATF_TENSOR"); return !(e && e[0] == '0'); }();
    const int path = g_f32_path;
    if (path >= 2 || (path == 0 && tensor_ok)) {
      int r = (A.cs == 1 && B.cs == 1) ? tf32::launch(A.p, A.rs, B.p, B.rs, pc, ldc, M, N, K, path) : -1;
      if (r != -1) return r;
      if (path >= 2) return fail(UM_ERR_ARG, "matmul: operands do not fit the tensor-core FP32 kernel (row-major, 16-byte aligned, pitches and N multiples of 4, M*N*K >= 2^18)");
    }
    dim3 grid((unsigned)((N + SG_BN - 1) / SG_BN), (unsigned)((M + SG_BM - 1) / SG_BM));
    k_sgemm<<<grid, 256, 0, ctx()->stream>>>(A, B, pc, ldc, M, N, K);
    UM_LAUNCHED();
    return UM_OK;
  }
  return fail(UM_ERR_ARG, "matmul needs f64 or f32 operands");
}
