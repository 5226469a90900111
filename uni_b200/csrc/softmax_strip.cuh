// softmax_strip.cuh -- softmax(axis = 0) of a row-major FP64 matrix with ONE read and ONE write of every element
// (included by reduce.cu; reference semantics: MatMathOps.scala:156-226).
//
// A lane of softmax(axis 0) is a COLUMN: 32768 elements 256 KB apart at BASELINE config 2, so the streaming form reads
// the matrix twice (online max / sum, then normalise: 24E bytes).  Here a thread-block CLUSTER of G CTAs keeps a whole
// strip of 16 columns (128-byte row segments, the narrowest that still streams well from HBM) on chip: CTA `rank`
// owns rows [rank * R, rank * R + R), R <= 2048, and its 2048 x 16 doubles = 256 KB fill the SM's TENSOR MEMORY
// exactly (tcgen05.st / tcgen05.ld, a double as two 32-bit columns; shared memory is only the TMA landing ring).
//
//   phase A  the strip streams in through a ring of 8 TMA boxes (128 rows x 16 columns); the 16 compute warps form
//            e = exp(x - k) (k = row 0 of the column, the same shift on every CTA: the sums add up without a
//            re-base), keep sum e per column and the range of x - k, and park e in tensor memory;
//   exchange CTA partials -> every CTA of the cluster through distributed shared memory, ONE cluster barrier per strip
//            (double-buffered slots), rank-order sums: S = sum e, and whether anyone saw an out-of-range element;
//   phase B  out = e * (1 / S) straight from tensor memory to HBM (16-byte stores, full 128-byte row segments), while
//            the TMA producer already fills the ring with the next strip.
//
// exp(x - k) / sum exp(x - k) is the reference's exp(x - max) / sum exp(x - max) with the common factor exp(max - k)
// cancelled.  That is exact arithmetic as long as nothing leaves the normal range, so a strip is only finished this
// way when every element has -690 <= x - k <= 300 (every e a normal number, the sum far from overflow).  Anything else
// -- a wider spread, NaN, +-Inf, a non-finite k -- sends its strip down the reference's own arithmetic: the cluster
// re-reads the strip from global memory three times (max, sum exp(x - max), exp(x - max) / sum).  The range test
// runs on the integer pipe, on the high word of x - k.
#pragma once
#include <cuda.h>

namespace um {
namespace sms {

constexpr int CW = 16;                    // columns per strip: one 128-byte row segment
constexpr int PR = 128;                   // rows per TMA box
constexpr int NSLOT = 8;                  // landing ring: 8 x 16 KB in flight per SM
constexpr int PIECE_BYTES = PR * CW * 8;  // 16384
constexpr int RMAX = 2048;                // rows per CTA: 2048 x 16 doubles = all 512 tensor-memory columns
constexpr int GMAX = 16;
constexpr int NCW = 16;                   // compute warps: four per sub-partition (8 measured the same; 16 keeps 95 registers per thread)
constexpr int NCOMP = NCW * 32;
constexpr int NTHREADS = NCOMP + 32;      // + TMA producer warp
constexpr int RG = NCOMP / 8;             // row groups of a box: thread = (row group, column pair)
constexpr int RPT = PR / RG;              // rows per thread per box (2)
constexpr int TW = 4 * RPT;               // 32-bit tensor-memory columns per thread per box
constexpr int TCOLS_PER_WARP = 512 / (NCW / 4);
static_assert(RPT == 2 && TW == 8 && TCOLS_PER_WARP * (NCW / 4) == 512 && (RMAX / PR) * TW == TCOLS_PER_WARP, "strip kernel geometry");

struct Smem {
  alignas(128) unsigned char slot[NSLOT][PIECE_BYTES];
  alignas(16) double red[NCW][2][CW];     // per-warp column partials: sum e, max x
  alignas(16) double exch[2][GMAX][2][CW];   // CTA partials of every rank, double-buffered by strip parity
  double exch2[GMAX][CW];                 // slow path: sum exp(x - max) per rank
  double exptab[32];                      // 2^(j/32)
  unsigned long long full[NSLOT], empty[NSLOT];
  unsigned tmem_base;
};
constexpr int SMEM_ALLOC = (int)sizeof(Smem) + 128;

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
__device__ __forceinline__ double2 lds_v2(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster(const double* local_addr, unsigned rank, double v) {
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sa(local_addr)), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(NCOMP) : "memory"); }
// 8 doubles <-> 16 consecutive 32-bit tensor-memory columns of this thread's lane
__device__ __forceinline__ void tmem_st16(unsigned taddr, const double (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
      "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3])),
      "r"(__double2loint(v[4])), "r"(__double2hiint(v[4])), "r"(__double2loint(v[5])), "r"(__double2hiint(v[5])),
      "r"(__double2loint(v[6])), "r"(__double2hiint(v[6])), "r"(__double2loint(v[7])), "r"(__double2hiint(v[7]))
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, double (&v)[8]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const double (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
               "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3]))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8_raw(unsigned taddr, unsigned (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
#define SMS_R8(n) "+r"(n[0]), "+r"(n[1]), "+r"(n[2]), "+r"(n[3]), "+r"(n[4]), "+r"(n[5]), "+r"(n[6]), "+r"(n[7])
// NB boxes' worth of this thread's stash (NB x 8 columns) with one wait; every register passes through the wait or
// through an empty volatile asm right behind it, so no use of them can be scheduled ahead of the wait
template <int NB>
__device__ __forceinline__ void tmem_ld8xN(unsigned taddr, double (&v)[NB][4]) {
  unsigned r[NB][8];
#pragma unroll
  for (int b = 0; b < NB; ++b) tmem_ld8_raw(taddr + (unsigned)b * 8u, r[b]);
  asm volatile("tcgen05.wait::ld.sync.aligned;" : SMS_R8(r[0]) : : "memory");
#pragma unroll
  for (int b = 1; b < NB; ++b) asm volatile("" : SMS_R8(r[b]) : : "memory");
#pragma unroll
  for (int b = 0; b < NB; ++b)
#pragma unroll
    for (int i = 0; i < 4; ++i) v[b][i] = __hiloint2double((int)r[b][2 * i + 1], (int)r[b][2 * i]);
}
// two loads in flight, one wait: the registers of both are tied to the wait so nothing reads them early
__device__ __forceinline__ void tmem_ld16_raw(unsigned taddr, unsigned (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
#define SMS_R16(n) "+r"(n[0]), "+r"(n[1]), "+r"(n[2]), "+r"(n[3]), "+r"(n[4]), "+r"(n[5]), "+r"(n[6]), "+r"(n[7]), "+r"(n[8]), "+r"(n[9]), \
                   "+r"(n[10]), "+r"(n[11]), "+r"(n[12]), "+r"(n[13]), "+r"(n[14]), "+r"(n[15])
// two 16-column loads in flight, one wait; every register of both loads passes through the wait (or through an empty
// volatile asm right behind it), so no use of them can be scheduled ahead of the wait
__device__ __forceinline__ void tmem_ld16x2(unsigned ta, unsigned tb, double (&a)[8], double (&b)[8]) {
  unsigned ra[16], rb[16];
  tmem_ld16_raw(ta, ra);
  tmem_ld16_raw(tb, rb);
  asm volatile("tcgen05.wait::ld.sync.aligned;" : SMS_R16(ra) : : "memory");
  asm volatile("" : SMS_R16(rb) : : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = __hiloint2double((int)ra[2 * i + 1], (int)ra[2 * i]);
    b[i] = __hiloint2double((int)rb[2 * i + 1], (int)rb[2 * i]);
  }
}
__device__ __forceinline__ void st_global_v2(double* p, double a, double b) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}


// fast_exp's arithmetic (common.cuh: 12 FP64 instructions, 32-entry table) with the table in shared memory and
// WITHOUT its range branch.  Branch-free on purpose: the eight elements a thread handles per box are eight
// independent chains the scheduler can interleave (with fast_exp's range test every element sat in its own
// reconvergence region and, at two warps per sub-partition, the phase ran at the latency of one Horner chain).
// Outside [-700, 700] (and for NaN / Inf) the value is garbage: such an element always sends its strip down the slow
// path (the range tracking below), which recomputes everything from global memory.
__device__ __forceinline__ double exp_shifted(double x, unsigned tab) {
  int k;
  double p = exp32_core(x, &k);
  double tj;
  asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(tab + (((unsigned)k << 3) & 0xf8u)));
  return exp32_scale(p * tj, k);
}
// Range tracking on the high word of d = x - k, on the integer pipe (the FP64 pipe and the issue slots are what phase A
// runs out of): as a SIGNED int the high word orders the positive doubles, +Inf and the positive NaNs above everything
// else; as an UNSIGNED int it orders the negative doubles by magnitude, -Inf and the negative NaNs on top.  Two
// integer max per element replace the FP64 max, the NaN test and the range select.
constexpr int HI_POS_LIMIT = 0x4072C000;        // high word of  300.0: d above it (or +Inf, +NaN) -> slow path
constexpr unsigned HI_NEG_LIMIT = 0xC0859000u;  // high word of -690.0: d below it (or -Inf, -NaN) -> slow path

struct Args {
  const double* x; i64 ldx;     // input, row pitch in elements (plain loads: shift row, slow path)
  double* out; i64 ldo;
  i64 T, N;
  int G, R;                     // cluster size, rows per CTA (multiple of PR)
  i64 nstrip;
};

__global__ void __launch_bounds__(NTHREADS, 1) k_softmax_strip(const __grid_constant__ CUtensorMap mx_map, const Args g) {
  extern __shared__ unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_rank();
  const int G = g.G;
  const i64 cluster_id = blockIdx.x / G, nclusters = gridDim.x / G;
  const i64 row_lo = (i64)rank * g.R;
  i64 rows_here = g.T - row_lo;
  if (rows_here > g.R) rows_here = g.R;
  if (rows_here < 0) rows_here = 0;
  const int npiece = (int)((rows_here + PR - 1) / PR);

  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mb_init(sa(&sm.full[s]), 1);
      mb_init(sa(&sm.empty[s]), NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) sm.exptab[tid] = k_exp2_32[tid];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa(&sm.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_sync_all();   // every CTA's barriers and exchange buffers exist before anyone pushes into them

  if (warp == NCW) {
    // ---- TMA producer: one lane issues, the whole warp takes part in the CTA / cluster barriers ----
    unsigned it = 0;   // pieces issued so far (ring position and parity)
    unsigned strip_it = 0;
    for (i64 strip = cluster_id; strip < g.nstrip; strip += nclusters, ++strip_it) {
      if (lane == 0) {
        for (int p = 0; p < npiece; ++p, ++it) {
          const unsigned s = it % NSLOT, use = it / NSLOT;
          if (use > 0) mb_wait(sa(&sm.empty[s]), (use - 1) & 1);
          mb_expect_tx(sa(&sm.full[s]), PIECE_BYTES);
          tma_2d(sa(sm.slot[s]), &mx_map, (int)(strip * CW), (int)(row_lo + (i64)p * PR), sa(&sm.full[s]));
        }
      }
      __syncwarp();
      cluster_sync_all();
      // (the cluster barrier also orders this CTA: the flags every thread reads below were pushed before it)
      bool slow = false;
      {
        const int par = (int)(strip_it & 1);
        for (int r = 0; r < G; ++r) slow |= sm.exch[par][r][1][0] != 0.0;
      }
      if (slow) {   // the strip goes down the slow path: its three extra cluster barriers
        cluster_sync_all();
        cluster_sync_all();
        cluster_sync_all();
      }
      // on to the next strip: the ring fills while the compute warps write this one out
    }
  } else {
    // ---- compute warps ----
    const int c2 = tid & 7;            // column pair of the strip
    const int rg = tid >> 3;           // row of the box, rows rg + RG i
    const unsigned tbase = sm.tmem_base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(warp >> 2) * (unsigned)TCOLS_PER_WARP;
    const unsigned tab = sa(sm.exptab);
    unsigned it = 0;
    unsigned strip_it = 0;
    for (i64 strip = cluster_id; strip < g.nstrip; strip += nclusters, ++strip_it) {
      const i64 col = strip * CW + 2 * c2;
      const bool col_ok = col < g.N;   // N is even: a pair never straddles the edge
      double k0 = 0.0, k1 = 0.0;
      if (col_ok) {
        const double2 kk = *reinterpret_cast<const double2*>(g.x + col);
        k0 = kk.x; k1 = kk.y;
      }
      double s0 = 0.0, s1 = 0.0;
      int hpos = (int)0x80000000;   // signed max of the high words of d
      unsigned hneg = 0u;           // unsigned max of the high words of d
      // ---- phase A ----
      for (int p = 0; p < npiece; ++p, ++it) {
        const unsigned s = it % NSLOT, use = it / NSLOT;
        mb_wait(sa(&sm.full[s]), use & 1);
        const unsigned base = sa(sm.slot[s]) + (unsigned)rg * (CW * 8) + (unsigned)c2 * 16;
        double2 v[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) v[i] = lds_v2(base + (unsigned)i * RG * CW * 8);
        __syncwarp();
        if (lane == 0) mb_arrive(sa(&sm.empty[s]));   // the values are in registers: the slot goes back to the producer
        double e[2 * RPT];
        if ((i64)(p + 1) * PR <= rows_here) {
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const double d0 = v[i].x - k0, d1 = v[i].y - k1;
            e[2 * i] = exp_shifted(d0, tab);
            e[2 * i + 1] = exp_shifted(d1, tab);
            s0 += e[2 * i]; s1 += e[2 * i + 1];
            hpos = max(hpos, max(__double2hiint(d0), __double2hiint(d1)));
            hneg = max(hneg, max((unsigned)__double2hiint(d0), (unsigned)__double2hiint(d1)));
          }
        } else {   // the box hangs over the last row: TMA filled the rest with zeros, which are not part of the lane
          const i64 r0 = (i64)p * PR + rg;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const bool ok = r0 + RG * i < rows_here;
            const double d0 = ok ? v[i].x - k0 : 0.0, d1 = ok ? v[i].y - k1 : 0.0;
            e[2 * i] = ok ? exp_shifted(d0, tab) : 0.0;
            e[2 * i + 1] = ok ? exp_shifted(d1, tab) : 0.0;
            s0 += e[2 * i]; s1 += e[2 * i + 1];
            hpos = max(hpos, max(__double2hiint(d0), __double2hiint(d1)));
            hneg = max(hneg, max((unsigned)__double2hiint(d0), (unsigned)__double2hiint(d1)));
          }
        }
        tmem_st8(tbase + (unsigned)p * (unsigned)TW, e);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      // ---- CTA partial: the rows of one column pair sit in the lanes with equal (lane & 7) ----
      s0 += __shfl_xor_sync(0xffffffffu, s0, 8); s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
      s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
      // k of a column past the edge is 0 and its elements are TMA's zero fill: never out of range
      const bool odd = __any_sync(0xffffffffu, hpos > HI_POS_LIMIT || hneg > HI_NEG_LIMIT);
      if (lane < 8) { sm.red[warp][0][2 * lane] = s0; sm.red[warp][0][2 * lane + 1] = s1; }
      if (lane == 8) sm.red[warp][1][0] = odd ? 1.0 : 0.0;
      bar_compute();
      const int par = (int)(strip_it & 1);
      if (tid < 2 * CW) {   // thread = (value, column): warp totals in warp order, then one store per rank of the cluster
        const int val = tid >> 4, cc = tid & 15;
        if (val == 0 || cc == 0) {
          double a = sm.red[0][val][cc];
#pragma unroll
          for (int w = 1; w < NCW; ++w) a += sm.red[w][val][cc];   // val 1: the number of warps that saw an odd element
#pragma unroll
          for (int r = 0; r < GMAX; ++r)
            if (r < G) st_cluster(&sm.exch[par][rank][val][cc], (unsigned)r, a);
        }
      }
      cluster_sync_all();
      // every thread adds the ranks' partials of its own two columns, in rank order: bit-identical everywhere
      double S0 = 0.0, S1 = 0.0;
      bool slow = false;
#pragma unroll
      for (int r = 0; r < GMAX; ++r)   // unrolled: the loads go out together, the adds keep their order
        if (r < G) {
          const double2 ps = *reinterpret_cast<const double2*>(&sm.exch[par][r][0][2 * c2]);
          S0 += ps.x; S1 += ps.y;
          slow |= sm.exch[par][r][1][0] != 0.0;
        }
      if (!slow) {
        // ---- phase B ----
        const double c0 = 1.0 / S0, c1 = 1.0 / S1;
        double* orow = g.out + (row_lo + rg) * g.ldo + col;
        int p = 0;
        for (; p + 4 <= npiece; p += 4) {
          double e[4][4];
          tmem_ld8xN<4>(tbase + (unsigned)p * (unsigned)TW, e);
          if (col_ok) {
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
              for (int i = 0; i < RPT; ++i) {
                const i64 r = (i64)(p + b) * PR + rg + RG * i;
                if (r < rows_here) st_global_v2(orow + ((i64)(p + b) * PR + RG * i) * g.ldo, e[b][2 * i] * c0, e[b][2 * i + 1] * c1);
              }
          }
        }
        for (; p < npiece; ++p) {
          double e[1][4];
          tmem_ld8xN<1>(tbase + (unsigned)p * (unsigned)TW, e);
          if (col_ok) {
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              const i64 r = (i64)p * PR + rg + RG * i;
              if (r < rows_here) st_global_v2(orow + ((i64)p * PR + RG * i) * g.ldo, e[0][2 * i] * c0, e[0][2 * i + 1] * c1);
            }
          }
        }
      } else {
        // ---- the reference's arithmetic, from global memory: max (NaN never wins, MatMathOps.scala:163-170),
        //      sum exp(x - max), exp(x - max) / sum ----
        double M0 = -INFINITY, M1 = -INFINITY;
        if (col_ok)
          for (i64 r = rg; r < rows_here; r += RG) {
            const double2 xv = *reinterpret_cast<const double2*>(g.x + (row_lo + r) * g.ldx + col);
            M0 = fmax(M0, xv.x); M1 = fmax(M1, xv.y);
          }
        M0 = fmax(M0, __shfl_xor_sync(0xffffffffu, M0, 8)); M1 = fmax(M1, __shfl_xor_sync(0xffffffffu, M1, 8));
        M0 = fmax(M0, __shfl_xor_sync(0xffffffffu, M0, 16)); M1 = fmax(M1, __shfl_xor_sync(0xffffffffu, M1, 16));
        if (lane < 8) { sm.red[warp][0][2 * lane] = M0; sm.red[warp][0][2 * lane + 1] = M1; }
        bar_compute();
        if (tid < CW) {
          double pm = -INFINITY;
          for (int w = 0; w < NCW; ++w) pm = fmax(pm, sm.red[w][0][tid]);
          for (int r = 0; r < G; ++r) st_cluster(&sm.exch2[rank][tid], (unsigned)r, pm);
        }
        cluster_sync_all();
        M0 = -INFINITY; M1 = -INFINITY;
        for (int r = 0; r < G; ++r) { M0 = fmax(M0, sm.exch2[r][2 * c2]); M1 = fmax(M1, sm.exch2[r][2 * c2 + 1]); }
        cluster_sync_all();   // every CTA has read the maxima (and its own `red`) before the sums overwrite them
        double t0 = 0.0, t1 = 0.0;
        if (col_ok)
          for (i64 r = rg; r < rows_here; r += RG) {
            const double2 xv = *reinterpret_cast<const double2*>(g.x + (row_lo + r) * g.ldx + col);
            t0 += exp(xv.x - M0);
            t1 += exp(xv.y - M1);
          }
        t0 += __shfl_xor_sync(0xffffffffu, t0, 8); t1 += __shfl_xor_sync(0xffffffffu, t1, 8);
        t0 += __shfl_xor_sync(0xffffffffu, t0, 16); t1 += __shfl_xor_sync(0xffffffffu, t1, 16);
        if (lane < 8) { sm.red[warp][0][2 * lane] = t0; sm.red[warp][0][2 * lane + 1] = t1; }
        bar_compute();
        if (tid < CW) {
          double ps = 0.0;
          for (int w = 0; w < NCW; ++w) ps += sm.red[w][0][tid];
          for (int r = 0; r < G; ++r) st_cluster(&sm.exch2[rank][tid], (unsigned)r, ps);
        }
        cluster_sync_all();
        double d0 = 0.0, d1 = 0.0;
        for (int r = 0; r < G; ++r) { d0 += sm.exch2[r][2 * c2]; d1 += sm.exch2[r][2 * c2 + 1]; }
        if (col_ok)
          for (i64 r = rg; r < rows_here; r += RG) {
            const double2 xv = *reinterpret_cast<const double2*>(g.x + (row_lo + r) * g.ldx + col);
            st_global_v2(g.out + (row_lo + r) * g.ldo + col, exp(xv.x - M0) / d0, exp(xv.y - M1) / d1);
          }
      }
    }
  }
  // nobody leaves while a peer may still push into its shared memory; then give the tensor memory back
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem_base), "r"(512) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Cluster shape for T rows: the smallest power of two G with ceil(T / G) <= 2048 rows per CTA.
static bool pick_shape(i64 T, int* G, int* R) {
  for (int g = 1; g <= GMAX; g *= 2) {
    const i64 r = ((T + g - 1) / g + PR - 1) / PR * PR;
    if (r <= RMAX) { *G = g; *R = (int)r; return true; }
  }
  return false;
}

static int resident_clusters(int G) {
  std::lock_guard<std::mutex> plan_lock(plan_mutex());
  static int cache[GMAX + 1];
  static bool have[GMAX + 1];
  if (have[G]) return cache[G];
  cudaFuncSetAttribute(k_softmax_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC);
  if (G > 8) cudaFuncSetAttribute(k_softmax_strip, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(G * 64), 1, 1);
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM_ALLOC;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k_softmax_strip, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  cache[G] = n; have[G] = true;
  return n;
}

// UM_OK when the strip kernel ran, -1 when the operands do not fit it (the caller takes the two-read path)
static int launch(const double* x, i64 ldx, i64 T, i64 N, double* out, i64 ldo) {
  int G = 0, R = 0;
  if (!pick_shape(T, &G, &R)) return -1;
  if ((N & 1) || (ldx & 1) || (ldo & 1) || !aligned16(x) || !aligned16(out) || N >= (i64(1) << 31) || T >= (i64(1) << 31)) return -1;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return -1;
  const int ncl = resident_clusters(G);
  if (ncl <= 0) return -1;
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)T};
  cuuint64_t strides[1] = {(cuuint64_t)ldx * 8};
  cuuint32_t box[2] = {(cuuint32_t)CW, (cuuint32_t)PR};
  cuuint32_t es[2] = {1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return -1;
  Args a;
  a.x = x; a.ldx = ldx; a.out = out; a.ldo = ldo; a.T = T; a.N = N; a.G = G; a.R = R;
  a.nstrip = (N + CW - 1) / CW;
  i64 nclusters = a.nstrip < ncl ? a.nstrip : ncl;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nclusters * G), 1, 1);
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM_ALLOC;
  cfg.stream = ctx()->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_softmax_strip, map, a);
  if (e != cudaSuccess) return fail(UM_ERR_CUDA, "softmax strip kernel launch failed: %s", cudaGetErrorString(e));
  UM_LAUNCHED();
  return UM_OK;
}

}  // namespace sms
}  // namespace um
