// matmul_tf32.cuh -- Mat[Float] `a *@ b` on the 5th-generation tensor cores (included by matmul.cu).
//
// tcgen05 has no FP32 kind; kind::tf32 keeps 11 significant bits of each operand.  To stay at FP32-level
// accuracy (the reference's multiplyFloat accumulates in float, Mat.scala:3236-3268) every operand is split
// on the fly into hi = x rounded to TF32 and lo = x - hi (rounded to TF32 as well), and the product is
// formed as  hi*hi + hi*lo + lo*hi  (3xTF32; the dropped lo*lo term is ~2^-22 relative) with FP32
// accumulation in TENSOR MEMORY.
//
// One CTA (192 threads) per 128 x 128 tile of C, K in blocks of 32 floats (= one 128-byte swizzle row):
//   warp 0      TMA producer: A tile 128 x 32 (K-major) and B tile 32 x 128 (four 32 x 32 boxes, N-major as it
//               lies in memory), SWIZZLE_128B, into a 3-entry landing ring, `full` mbarriers
//   warps 2-9   splitter (two per SM sub-partition): landing entry -> 2-stage operand ring; A as `hi` with `lo`
//               beside it (element-wise, so the swizzle does not matter); B is TRANSPOSED on the way into K-major swizzled hi / lo tiles (an N-major TF32
//               operand made tcgen05.mma return zeros on this part: measured, so both operands are fed
//               K-major); fence.proxy.async, `ready` mbarrier; afterwards the epilogue: tcgen05.ld of the
//               128 x 128 FP32 accumulator -> global memory
//   warp 1      MMA issuer (one lane): 4 k-steps x 3 tcgen05.mma.kind::tf32 (M = 128, N = 128, K = 8) per stage,
//               tcgen05.commit frees the stage / hands the accumulator to the epilogue
// Operands must be row-contiguous, 16-byte aligned, with M, N multiples of 128 and K a multiple of 32;
// everything else stays on the register-tiled FFMA kernel.
//
// From 2^30 multiply-adds up the split does not happen in shared memory at all (k_sgemm_tf32x3<true>): k_split_a /
// k_split_bt write A_hi, A_lo and the TRANSPOSED B_hi', B_lo' to global memory once (12 bytes of traffic per operand
// element, 3 % of the product's time at 8192^3), and the product kernel is TMA -> three 64 KB operand stages -> MMA,
// with warps 2-9 left to promote and store.  Shared-memory traffic per k-block falls from 224 KB (TMA 32 + split
// 32 + 64 + MMA reads 96) to 160 KB, the splitter's barriers disappear, and the bound moves to the L2: 64 KB per
// k-block and CTA is 1 500 clocks at the ~6 300 B / clk the L2 delivers to 148 SMs.  Results are bit-identical
// to the shared-memory split (same rounding, same MMA order; tools/matf_paths.py).  Tiles are rasterised in bands
// of 8 tile rows so that the resident CTAs share their A and B panels in L2 (16384^3: 133 -> 143 TFLOP/s for the
// shared-memory split).
#pragma once
#include <cuda.h>

namespace um {
namespace tf32 {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 2;
constexpr int TILE_BYTES = BM * BK * 4;                    // 16384: A tile; B tile has the same size
constexpr int STAGE_BYTES = 4 * TILE_BYTES;                // operand ring entry: A_hi | A_lo | B_hi | B_lo
constexpr int NRAW = 3;                                    // landing ring (TMA destinations): A | B as they lie in memory
constexpr int RAW_BYTES = 2 * TILE_BYTES;
constexpr int OFF_RAW = STAGES * STAGE_BYTES;
constexpr int OFF_BAR = OFF_RAW + NRAW * RAW_BYTES;        // raw_full[3] raw_empty[3] ready[2] empty[2] acc_full[2] acc_free[2]
constexpr int CHUNK = 4;                                   // k-blocks accumulated in tensor memory before promotion to registers
constexpr int OFF_TMEM = OFF_BAR + 16 * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;
constexpr int RASTER = 8;                                  // tile rows per rasterisation band
constexpr int NSPLIT = 256;                                // splitter / epilogue threads: 8 warps, two per SM sub-partition
constexpr int NTHREADS = 64 + NSPLIT;

__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mb_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46),
// version = 1 [46,48), layout type [61,64): 2 = SWIZZLE_128B)
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  return (unsigned long long)((addr & 0x3FFFF) >> 4) | ((unsigned long long)(lbo_bytes >> 4) << 16) |
         ((unsigned long long)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), A and B K-major (bits 15, 16 = 0),
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((unsigned)(BN >> 3) << 17) |
                           ((unsigned)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// PRE = the operands arrive already split and K-major (k_split_a / k_split_bt below wrote A_hi, A_lo, B_hi', B_lo' to
// global memory): the TMA fills the operand ring directly (three 64 KB stages, map_a / map_a2 = A_hi / A_lo, map_b /
// map_b2 = B_hi' / B_lo'), nothing is split in shared memory and warps 2-9 only promote and store.
template <bool PRE>
__global__ void __launch_bounds__(NTHREADS, 1)
k_sgemm_tf32x3(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2, float* __restrict__ C, i64 ldc,
               int kblocks, int M, int N, int tiles_m, int tiles_n) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const unsigned s0 = sa(smem);
  const unsigned bar0 = s0 + OFF_BAR;
  auto bar_full = [&](int r) { return bar0 + 8u * r; };                  // landing entry r holds block kb (TMA bytes)
  auto bar_raw_empty = [&](int r) { return bar0 + 8u * (NRAW + r); };    // every splitter thread has read entry r
  auto bar_ready = [&](int s) { return bar0 + 8u * (2 * NRAW + s); };    // operand stage s is split and visible to the tensor core
  auto bar_empty = [&](int s) { return bar0 + 8u * (2 * NRAW + 2 + s); };  // the MMAs that read operand stage s have retired
  auto bar_acc_full = [&](int b) { return bar0 + 8u * (2 * NRAW + 4 + b); };
  auto bar_acc_free = [&](int b) { return bar0 + 8u * (2 * NRAW + 6 + b); };
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + OFF_TMEM);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tiles in bands of RASTER tile rows, column by column inside a band: the CTAs resident at one time share RASTER
  // row panels of A and ~148 / RASTER column panels of B in L2 (row-major order re-read all of B per tile row)
  int tm, tn;
  {
    const int bid = (int)blockIdx.x, band = bid / (RASTER * tiles_n), rem = bid % (RASTER * tiles_n);
    const int rows = tiles_m - band * RASTER < RASTER ? tiles_m - band * RASTER : RASTER;
    tm = band * RASTER + rem % rows;
    tn = rem / rows;
  }
  const int m0 = tm * BM, n0 = tn * BN;
  constexpr int PSTAGES = 3;   // PRE: operand ring entries

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int r = 0; r < NRAW; ++r) {
      mb_init(bar_full(r), 1);
      mb_init(bar_raw_empty(r), PRE ? 1 : NSPLIT);   // PRE: operand stage r, released by the MMA warp's commit
    }
    for (int s = 0; s < STAGES; ++s) {
      mb_init(bar_ready(s), NSPLIT);
      mb_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mb_init(bar_acc_full(b), 1);
      mb_init(bar_acc_free(b), NSPLIT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_d = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // The landing ring is released by the splitters, not by the MMAs: the loads run NRAW blocks ahead of the
      // split instead of waiting for the tensor core to retire a stage (with one ring for both, a k-block cost
      // (TMA latency + split + MMA) / 2)
      if (PRE) {
        for (int kb = 0; kb < kblocks; ++kb) {
          const int s = kb % PSTAGES;
          if (kb >= PSTAGES) mb_wait(bar_raw_empty(s), (unsigned)((kb / PSTAGES - 1) & 1));
          mb_expect_tx(bar_full(s), 4 * TILE_BYTES);
          const unsigned st = s0 + (unsigned)s * STAGE_BYTES;
          tma_2d(st, &map_a, kb * BK, m0, bar_full(s));
          tma_2d(st + TILE_BYTES, &map_a2, kb * BK, m0, bar_full(s));
          tma_2d(st + 2 * TILE_BYTES, &map_b, kb * BK, n0, bar_full(s));
          tma_2d(st + 3 * TILE_BYTES, &map_b2, kb * BK, n0, bar_full(s));
        }
      } else
      for (int kb = 0; kb < kblocks; ++kb) {
        const int r = kb % NRAW;
        if (kb >= NRAW) mb_wait(bar_raw_empty(r), (unsigned)((kb / NRAW - 1) & 1));
        mb_expect_tx(bar_full(r), 2 * TILE_BYTES);
        const unsigned st = s0 + OFF_RAW + (unsigned)r * RAW_BYTES;
        tma_2d(st, &map_a, kb * BK, m0, bar_full(r));                                // A: [128 rows m][32 k]
#pragma unroll
        for (int j = 0; j < 4; ++j)                                                 // B: 4 boxes of [32 rows k][32 n]
          tma_2d(st + TILE_BYTES + j * 4096, &map_b, n0 + 32 * j, kb * BK, bar_full(r));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // The tensor core adds into its FP32 accumulator with truncation: summed over thousands of k the bias is
      // visible (2e-4 relative at k = 4096).  So tensor memory only accumulates CHUNK k-blocks (128 k); the
      // epilogue warps promote each chunk into FP32 registers with round-to-nearest adds.  Two accumulator
      // buffers alternate so the next chunk's MMAs never wait for the promotion.  The cross terms hi*lo + lo*hi
      // are 2^-11 of the main term: they get an accumulator of their own (columns 256..383), where the same
      // truncation is 2^-11 smaller and may run over the whole of K, so the main accumulator only sees one
      // truncating add per 8 k instead of three.
      const unsigned tmem_corr = tmem_d + 2u * BN;
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % (PRE ? PSTAGES : STAGES);
        const int chunk = kb / CHUNK, buf = chunk & 1;
        if (kb % CHUNK == 0 && chunk >= 2) {
          mb_wait(bar_acc_free(buf), (unsigned)((chunk / 2 - 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const unsigned tmem_acc = tmem_d + (unsigned)(buf * BN);
        if (PRE) mb_wait(bar_full(s), (unsigned)((kb / PSTAGES) & 1));
        else mb_wait(bar_ready(s), (unsigned)((kb / STAGES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned st = s0 + (unsigned)s * STAGE_BYTES;
        const unsigned a_hi = st, a_lo = st + TILE_BYTES, b_hi = st + 2 * TILE_BYTES, b_lo = st + 3 * TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {   // K = 8 per instruction: 32 bytes along the K-major rows
          const unsigned long long dah = smem_desc(a_hi + kk * 32, 16, 1024), dal = smem_desc(a_lo + kk * 32, 16, 1024);
          const unsigned long long dbh = smem_desc(b_hi + kk * 32, 16, 1024), dbl = smem_desc(b_lo + kk * 32, 16, 1024);
          umma_tf32(tmem_acc, dah, dbh, ((kb % CHUNK) | kk) ? 1u : 0u);
          umma_tf32(tmem_corr, dah, dbl, (kb | kk) ? 1u : 0u);
          umma_tf32(tmem_corr, dal, dbh, 1u);
        }
        umma_commit(PRE ? bar_raw_empty(s) : bar_empty(s));   // the stage may be refilled once these MMAs have read it
        if (kb % CHUNK == CHUNK - 1 || kb == kblocks - 1) umma_commit(bar_acc_full(buf));
      }
    }
  } else {
    // ---- splitter: 256 threads, 2 x 4096 floats per stage ----
    const int t = tid - 64;
    const int q = warp & 3;            // the TMEM lane quarter this warp may access = 32 rows of the tile
    const int half = (warp - 2) >> 2;  // two warps share a lane quarter: columns [64 half, 64 half + 64) of the tile each
    constexpr int HN = BN / 2;
    float acc[HN];
#pragma unroll
    for (int j = 0; j < HN; ++j) acc[j] = 0.0f;
    auto add_tmem = [&](int col) {     // acc += this warp's 64 of the 128 tensor-memory columns from `col` (this thread: one row)
#pragma unroll
      for (int c0 = 0; c0 < HN; c0 += 32) {
        unsigned r[32];
        const unsigned taddr = tmem_d + ((unsigned)(32 * q) << 16) + (unsigned)(col + HN * half + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
            "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
      }
    };
    auto promote = [&](int chunk) {    // acc += the main accumulator of `chunk`, then hand its buffer back
      const int buf = chunk & 1;
      mb_wait(bar_acc_full(buf), (unsigned)((chunk / 2) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      add_tmem(buf * BN);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mb_arrive(bar_acc_free(buf));
    };
    const int nchunks = (kblocks + CHUNK - 1) / CHUNK;
    int promoted = 0;
    if (!PRE)
    for (int kb = 0; kb < kblocks; ++kb) {
      const int s = kb % STAGES;
      // chunk c is complete once the stage ring has moved STAGES blocks past its last block (the slot of that
      // block was refilled, so its MMAs have retired): promote it here, without ever waiting
      if (promoted < nchunks - 1 && kb >= (promoted + 1) * CHUNK + STAGES) promote(promoted++);
      const int r = kb % NRAW;
      mb_wait(bar_full(r), (unsigned)((kb / NRAW) & 1));
      if (kb >= STAGES) {   // the MMAs of block kb - STAGES have read this operand stage
        mb_wait(bar_empty(s), (unsigned)((kb / STAGES - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      unsigned char* stg = smem + (size_t)s * STAGE_BYTES;
      const unsigned char* raw = smem + OFF_RAW + (size_t)r * RAW_BYTES;
      const float4* araw = reinterpret_cast<const float4*>(raw);
      float4* hi = reinterpret_cast<float4*>(stg);
      float4* lo = reinterpret_cast<float4*>(stg + TILE_BYTES);
      // hi = x rounded to nearest TF32 (11 significant bits), lo = (x - hi) rounded to nearest TF32: with
      // truncation instead, every dropped tail has the sign of x and the error grows linearly with K
      auto rn = [](float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); };
      auto split4 = [&](float4 x, float4& h, float4& l) {
        h.x = rn(x.x); h.y = rn(x.y); h.z = rn(x.z); h.w = rn(x.w);
        l.x = rn(x.x - h.x); l.y = rn(x.y - h.y); l.z = rn(x.z - h.z); l.w = rn(x.w - h.w);
      };
#pragma unroll 4
      for (int i = t; i < TILE_BYTES / 16; i += NSPLIT) {   // A: element-wise (the swizzle does not matter)
        float4 h, l;
        split4(araw[i], h, l);
        hi[i] = h;
        lo[i] = l;
      }
      // B: landed as 4 boxes [32 k][32 n] (128-byte rows, 16-byte chunk c of row k at chunk c ^ (k & 7)); written
      // K-major [128 n][32 k] with the same swizzle rule on the row index n
      const unsigned char* braw = raw + TILE_BYTES;
      unsigned char* bhi = stg + 2 * TILE_BYTES;
      unsigned char* blo = stg + 3 * TILE_BYTES;
#pragma unroll 4
      for (int idx = t; idx < BN * (BK / 4); idx += NSPLIT) {
        const int n = idx & (BN - 1), kc = idx >> 7;   // consecutive threads: consecutive n, same 4 k's
        const unsigned char* src = braw + (n >> 5) * 4096 + (n & 3) * 4;
        const int nc = (n & 31) >> 2;
        float4 x;
        x.x = *reinterpret_cast<const float*>(src + (4 * kc + 0) * 128 + ((nc ^ ((4 * kc + 0) & 7)) << 4));
        x.y = *reinterpret_cast<const float*>(src + (4 * kc + 1) * 128 + ((nc ^ ((4 * kc + 1) & 7)) << 4));
        x.z = *reinterpret_cast<const float*>(src + (4 * kc + 2) * 128 + ((nc ^ ((4 * kc + 2) & 7)) << 4));
        x.w = *reinterpret_cast<const float*>(src + (4 * kc + 3) * 128 + ((nc ^ ((4 * kc + 3) & 7)) << 4));
        float4 h, l;
        split4(x, h, l);
        const int off = n * 128 + ((kc ^ (n & 7)) << 4);
        *reinterpret_cast<float4*>(bhi + off) = h;
        *reinterpret_cast<float4*>(blo + off) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      mb_arrive(bar_ready(s));
      mb_arrive(bar_raw_empty(r));
    }
    // ---- epilogue: the remaining chunks, then this thread's row of C ----
    while (promoted < nchunks) promote(promoted++);
    add_tmem(2 * BN);                  // the cross terms (complete: the last chunk's commit covers every MMA issued)
    // edge tiles: TMA zero-filled what lies outside A and B, only the stores need the bounds (N % 4 == 0)
    const int row = m0 + 32 * q + lane;
    if (row < M) {
      const int nh = n0 + HN * half;
      float* crow = C + (i64)row * ldc + nh;
#pragma unroll
      for (int j = 0; j < HN; j += 4)
        if (nh + j < N) *reinterpret_cast<float4*>(crow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
}

// ---- pre-pass of the PRE kernel: operands split once, in global memory ----
__device__ __forceinline__ float rn_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

// A (M x K, pitch lda, 16-byte aligned rows) -> hi, lo (M x Kp, Kp = K rounded up to 4; the pad is zero)
__global__ void __launch_bounds__(256) k_split_a(const float* __restrict__ A, i64 lda, i64 M, i64 K, i64 Kp, float* __restrict__ hi,
                                                  float* __restrict__ lo) {
  const i64 cpr = Kp >> 2, total = M * cpr;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
    const i64 m = i / cpr, k = (i % cpr) << 2;
    float4 x;
    if (k + 4 <= K) x = *reinterpret_cast<const float4*>(A + m * lda + k);
    else {
      const float* r = A + m * lda;
      x.x = k < K ? r[k] : 0.0f; x.y = k + 1 < K ? r[k + 1] : 0.0f; x.z = k + 2 < K ? r[k + 2] : 0.0f; x.w = 0.0f;
    }
    float4 h, l;
    h.x = rn_tf32(x.x); h.y = rn_tf32(x.y); h.z = rn_tf32(x.z); h.w = rn_tf32(x.w);
    l.x = rn_tf32(x.x - h.x); l.y = rn_tf32(x.y - h.y); l.z = rn_tf32(x.z - h.z); l.w = rn_tf32(x.w - h.w);
    *reinterpret_cast<float4*>(hi + m * Kp + k) = h;
    *reinterpret_cast<float4*>(lo + m * Kp + k) = l;
  }
}

// B (K x N, pitch ldb) -> hi', lo' (N x Kp): transposed through a 32 x 33 shared-memory tile, both sides coalesced
__global__ void __launch_bounds__(256) k_split_bt(const float* __restrict__ B, i64 ldb, i64 K, i64 N, i64 Kp, float* __restrict__ hi,
                                                   float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const i64 tiles_n = (N + 31) >> 5, tiles_k = (Kp + 31) >> 5, total = tiles_n * tiles_k;
  for (i64 t = blockIdx.x; t < total; t += gridDim.x) {
    const i64 n0 = (t % tiles_n) << 5, k0 = (t / tiles_n) << 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const i64 k = k0 + ty + 8 * j, n = n0 + tx;
      tile[ty + 8 * j][tx] = (k < K && n < N) ? B[k * ldb + n] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const i64 n = n0 + ty + 8 * j, k = k0 + tx;
      if (n < N && k < Kp) {
        const float x = tile[tx][ty + 8 * j], h = rn_tf32(x);
        hi[n * Kp + k] = h;
        lo[n * Kp + k] = rn_tf32(x - h);
      }
    }
    __syncthreads();
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  std::lock_guard<std::mutex> plan_lock(plan_mutex());
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// returns UM_OK if launched, -1 if the shape / layout is not eligible (caller falls back to the FFMA kernel)
// mode: 0 = automatic (operands split once in global memory from PRE_MIN_WORK multiply-adds up, in shared memory below),
// 2 = split in shared memory, 3 = split in global memory
constexpr double PRE_MIN_WORK = 1073741824.0;   // 2^30 = 1024^3: measured faster from there up, pre-pass included
inline int launch(const float* A, i64 lda, const float* B, i64 ldb, float* C, i64 ldc, i64 M, i64 N, i64 K, int mode) {
  // any M and K (partial tiles are zero-filled by TMA), N a multiple of 4 (16-byte stores); products under 2^18 multiply-adds stay on the FFMA kernel
  if (M <= 0 || N <= 0 || K <= 0 || (N % 4) || (double)M * (double)N * (double)K < 262144.0) return -1;
  if (!aligned16(A) || !aligned16(B) || !aligned16(C) || (lda % 4) || (ldb % 4) || (ldc % 4)) return -1;
  if ((K + BK - 1) / BK > 0x7fffffff || M >= (i64(1) << 31) || N >= (i64(1) << 31)) return -1;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return -1;
  {
    std::lock_guard<std::mutex> plan_lock(plan_mutex());
    static bool attr = false;
    if (!attr) {
      if (cudaFuncSetAttribute(k_sgemm_tf32x3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC) != cudaSuccess ||
          cudaFuncSetAttribute(k_sgemm_tf32x3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC) != cudaSuccess) {
        cudaGetLastError();
        return -1;
      }
      attr = true;
    }
  }
  const i64 tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  if (tiles_m * tiles_n > 0x7fffffff) return -1;
  const int kblocks = (int)((K + BK - 1) / BK);
  CUtensorMap ma, mb;
  cuuint32_t es[2] = {1, 1};
  if (mode == 3 || (mode == 0 && (double)M * (double)N * (double)K >= PRE_MIN_WORK)) {
    // ---- operands split (and B transposed) once, in global memory; the product kernel is then TMA -> MMA only ----
    const i64 Kp = (K + 3) & ~i64(3);
    void* tmp = nullptr;
    const size_t a_elems = (size_t)M * Kp, b_elems = (size_t)N * Kp;
    const bool have = um_malloc((2 * a_elems + 2 * b_elems) * sizeof(float), &tmp) == UM_OK;
    if (!have && mode == 3) return UM_ERR_OOM;   // automatic: no room for the temporary -> the shared-memory split below
    if (have) {
    float* a_hi = static_cast<float*>(tmp);
    float* a_lo = a_hi + a_elems;
    float* b_hi = a_lo + a_elems;
    float* b_lo = b_hi + b_elems;
    const int sms = ctx()->sm_count;
    k_split_a<<<sms * 8, 256, 0, ctx()->stream>>>(A, lda, M, K, Kp, a_hi, a_lo);
    k_split_bt<<<sms * 8, 256, 0, ctx()->stream>>>(B, ldb, K, N, Kp, b_hi, b_lo);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    CUtensorMap m4[4];
    float* ptr[4] = {a_hi, a_lo, b_hi, b_lo};
    bool ok = true;
    for (int i = 0; i < 4 && ok; ++i) {
      cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)(i < 2 ? M : N)};
      cuuint64_t strides[1] = {(cuuint64_t)Kp * 4};
      cuuint32_t box[2] = {BK, (cuuint32_t)(i < 2 ? BM : BN)};
      ok = enc(&m4[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr[i], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    int rc = UM_OK;
    if (ok) {
      k_sgemm_tf32x3<true><<<(unsigned)(tiles_m * tiles_n), NTHREADS, SMEM_ALLOC, ctx()->stream>>>(m4[0], m4[2], m4[1], m4[3], C, ldc, kblocks, (int)M,
                                                                                                 (int)N, (int)tiles_m, (int)tiles_n);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      if (cudaGetLastError() != cudaSuccess) rc = fail(UM_ERR_CUDA, "k_sgemm_tf32x3<pre-split> launch failed");
    }
    um_free(tmp);   // stream-ordered: after the product
    if (!ok) return mode == 3 ? fail(UM_ERR_ARG, "matmul: tensor map for the pre-split operands rejected") : -1;
    return rc;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)lda * 4};
    cuuint32_t box[2] = {BK, BM};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -1;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)K};
    cuuint64_t strides[1] = {(cuuint64_t)ldb * 4};
    cuuint32_t box[2] = {32, BK};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(B), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -1;
  }
  k_sgemm_tf32x3<false><<<(unsigned)(tiles_m * tiles_n), NTHREADS, SMEM_ALLOC, ctx()->stream>>>(ma, mb, ma, mb, C, ldc, kblocks, (int)M,
                                                                                            (int)N, (int)tiles_m, (int)tiles_n);
  UM_LAUNCHED();
  return UM_OK;
}

}  // namespace tf32
}  // namespace um
