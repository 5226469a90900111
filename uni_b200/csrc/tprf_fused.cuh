// tprf_fused.cuh -- the single-read 3PRF sweep (included by tprf.cu).
//
// The two-sweep path reads X twice because sweep 2 (PX = Xn W0) needs W0_j, which needs the whole of
// column j.  Here a thread-block CLUSTER keeps a whole column tile (T rows x 16 columns) on chip: CTA
// `rank` of the cluster owns rows [rank*R, rank*R + R) and pulls its (R x 16) piece with TMA
// (SWIZZLE_128B boxes of 128 rows, one mbarrier per slot, three 64 KiB slots).  Warp roles per CTA (round 2:
// profiles/README.md has the measurements behind each choice):
//
//   A-warps 0-3   phase A of piece i: Zc' (x - k) (DMMA m8n8k4, M = 8 columns, K = rows; k = the column's first row)
//                 + shifted sum / sum of squares straight out of shared memory; per-lane partials go to `scratch`.
//                 The same warp copies its 128 rows, in the phase-B fragment layout, into TENSOR MEMORY (four
//                 64 KiB pieces deep) for the B-warp of its sub-partition, so the slot goes back to the producer
//                 as soon as phase A is done and the B-warps never touch a slot.
//   B-warps 4-7   phase B of piece i, as soon as every column's record is there (normally one or two pieces behind
//                 the A-warps): PX += x V, h0 += x / sd out of tensor memory (DMMA, M = 8 rows, K = columns); the
//                 accumulators stay in registers for the whole kernel.  A-warp a and B-warp 4+a share an SM
//                 sub-partition (= a tensor-memory lane quarter), so its FP64 pipe has a warp in each phase to draw on.
//   warp 8        TMA producer (one lane).
//   warps 9, 11   finalisers: owner of column c (rank c % G) adds the G CTA partials in a fixed tree, forms
//                 mu / sd / W0 / V, pushes V_c and 1/sd_c to every CTA of the cluster, writes W0 etc. for
//                 the per-column outputs and accumulates sw / SWW / smw / smu.  The two warps take the pieces in
//                 turn (large clusters) or the rounds of one piece (clusters of 1 or 2 CTAs).
//   warp 10       reducer: per-lane partials in `scratch` -> CTA partial -> owner CTA (few, larger DSMEM
//                 messages: 8-byte remote stores are expensive, 160 per piece instead of 1024).
// The A-warps run with 208 registers, the special warps with 128 (setmaxnreg): the Zc fragments of 128 rows alone
// are 64 registers, and with less ptxas sinks the shared-memory loads to just before their use.
//
// All cross-CTA traffic is DISTRIBUTED SHARED MEMORY written with `st.async`: the store itself completes
// transaction bytes on the destination CTA's mbarrier, so there are no fences, no cluster barriers and no
// polling in the steady state.  Every hand-off inside a CTA is an mbarrier as well; warps never wait for
// each other except through the data they need.  X is read from HBM exactly once; every sum has a fixed
// order, so results are reproducible run to run.
#pragma once
#include <cuda.h>
#include <type_traits>

namespace um {
namespace fused {

constexpr int CW = 16;                       // columns per tile = one 128-byte swizzle row
constexpr int NSLOT = 3;                     // shared-memory TMA landing slots
constexpr int NT = 4;                        // TMEM stash buffers per B-warp
// The A-warps are throttled by the three landing slots and by the NT stash buffers: phase A of piece a needs the
// B-warp done with piece a - NT, i.e. every finaliser's record of it.  No CTA is more than NT pieces ahead of any
// finaliser (the TMA producer NT + NSLOT); the exchange rings must cover that.
constexpr int NXR = 8;                       // ring entries of the partial / record buffers and their barriers
constexpr int NK = 16;                       // ring of shift rows (written by TMA, read by the finaliser)
constexpr int RMAX = 512;                    // rows per piece (per CTA of the cluster)
constexpr int SLOT_BYTES = RMAX * CW * 8;    // 65536
constexpr int BOX_ROWS = 128;
constexpr int NAW = 4;                       // A-warps (= B-warps): one of each per SM sub-partition
constexpr int NTHREADS = (2 * NAW + 4) * 32; // + TMA producer, finaliser, reducer, second finaliser (small clusters)
constexpr int NSTAT = 10;                    // per column: s1, s2, q[8]
constexpr int NREC = 10;                     // per column record: V[8], 1/sd, pad
constexpr int GMAX = 16;                     // largest cluster (non-portable size)
constexpr int SCR_W = 256;                   // scratch doubles per A-warp: q[16][8] | s1[16][4] | s2[16][4]
constexpr int REG_A = 208, REG_S = 128;       // setmaxnreg targets: A-warps up, special warps down (B-warps keep 168)
// the A-warps can only take what the special warps gave back to the CTA's pool (the kernel starts at 168 per thread,
// 65536 / 384 rounded down to a multiple of 8): asking for more spins forever in setmaxnreg.inc
static_assert(REG_A - 168 <= 168 - REG_S && REG_A % 8 == 0 && REG_S % 8 == 0, "setmaxnreg budget");
constexpr int STAT_SLOT = CW * NSTAT;        // doubles per slot of the partial buffer: [owned col][src rank][NSTAT]

constexpr int OFF_SLOTS = 0;
constexpr int OFF_SCRATCH = NSLOT * SLOT_BYTES;                 // double[NAW][SCR_W]
constexpr int OFF_STAT = OFF_SCRATCH + NAW * SCR_W * 8;         // double[NXR][STAT_SLOT]
constexpr int OFF_VREC = OFF_STAT + NXR * STAT_SLOT * 8;        // double[NXR][CW][NREC]
constexpr int OFF_KSH = OFF_VREC + NXR * CW * NREC * 8;         // double[NK][CW]   (row 0 of the panel: variance shift)
constexpr int OFF_BAR = OFF_KSH + NK * CW * 8;                  // full[3] empty[3] scr_full scr_free statready[NXR] vready[NXR]
constexpr int NBAR = 8 + 2 * NXR + 2 * NAW * NT;              // + stash_full[NAW][NT] stash_free[NAW][NT]
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;                    // u32: TMEM base address
constexpr int OFF_FIN2 = OFF_TMEM + 16;                         // double[11][8]: second finaliser's small sums -> first finaliser
constexpr int SMEM_BYTES = OFF_FIN2 + 11 * 8 * 8;
static_assert(SMEM_BYTES + 1024 <= 232448, "shared memory of the cluster sweep exceeds 227 KB");
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;                   // slack to align the base to 1024 (swizzle atom)

// ---- PTX wrappers ------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void mb_arrive_remote(unsigned cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mb_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mb_wait_cluster(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(
          bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ unsigned map_rank(unsigned local_addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_remote(unsigned cluster_addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(cluster_addr), "d"(v) : "memory");
}
// remote store that completes `8` transaction bytes on the destination CTA's mbarrier: the data is visible
// to whoever observes that barrier phase complete -- no fence, no separate arrive
__device__ __forceinline__ void st_async(unsigned cluster_addr, double v, unsigned cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(cluster_addr),
               "l"(__double_as_longlong(v)), "r"(cluster_bar)
               : "memory");
}
// mbarrier wait that also yields a zero the compiler cannot see through: adding it to a shared-memory
// address makes the (pure, freely schedulable) loads below data-dependent on the wait
__device__ __forceinline__ unsigned mb_wait_token(unsigned bar, unsigned parity) {
  unsigned tok;
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\nmov.u32 %0, 0;\n}"
      : "=r"(tok)
      : "r"(bar), "r"(parity)
      : "memory");
  return tok;
}
// pure loads of the piece: schedulable between the DMMAs; ordered after the wait through the token and
// before the slot is released through `consume`
__device__ __forceinline__ double ldp_f64(unsigned addr) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double2 ldp_v2f64(unsigned addr) {
  double2 v;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void consume(double v) { asm volatile("" ::"d"(v) : "memory"); }
// ptxas reorders freely inside a basic block but not across a warp-level barrier: a (free, the warp is converged)
// WARPSYNC pins the instruction order chosen in the source where the order matters (independent DMMA chains)
__device__ __forceinline__ void sched_fence() { __syncwarp(); }
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double2 lds_v2f64(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v2f64(unsigned addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void tma_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// ---- tensor memory as an FP64 stash (tools/microbench/tmem_scratch.cu): 32x32b shape = lane i of the warp
// owns TMEM lane 32*(warp%4)+i; a double is two consecutive 32-bit columns ----
__device__ __forceinline__ void tmem_st8(unsigned taddr, double2 a, double2 b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__double2loint(a.x)), "r"(__double2hiint(a.x)), "r"(__double2loint(a.y)), "r"(__double2hiint(a.y)),
               "r"(__double2loint(b.x)), "r"(__double2hiint(b.x)), "r"(__double2loint(b.y)), "r"(__double2hiint(b.y))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, unsigned (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

struct Args {
  i64 T, N, npad, tpad;
  int L, G, R;              // cluster size, rows per piece
  int batch, cpp;           // panels, chunks per panel (a chunk = the tiles c, c + cpp, ... of one panel)
  i64 ntile;                // column tiles per panel (this launch sweeps tiles [tile_lo, ntile))
  i64 tile_lo;
  int cpp_out, cidx_off;    // chunk slots per panel in the output partials, and this launch's first slot
  const double* zc;         // [batch][T][8]
  double* w0;               // [batch][npad][8]
  double* invsd;            // [batch][npad]
  double* mun;              // [batch][npad]
  double* blockpart;        // [batch][cpp*G][81]
  double* pxp;              // [batch][cpp][tpad][8]
  double* h0p;              // [batch][cpp][tpad]
  unsigned* started;        // optional: += 1 per cluster once it is resident (split sweep)
  long long* dbg;           // optional: [rank][warp][iter][8] clock64 stamps of cluster 0 (profiling aid)
};

// logical column of the phase-A fragment row m (0..7) of m-tile 0; m-tile 1 is +4.  Chosen so the
// 16 lanes of a half-warp touch 16 distinct 8-byte bank pairs under the 128-byte swizzle.
__device__ __forceinline__ int col_of_m(int m) { return ((m & 2) << 2) | (m & 1) | ((m & 4) >> 1); }  // 0,1,8,9,2,3,10,11

// pure (non-volatile) DMMA so the compiler may schedule it between the shared-memory loads
__device__ __forceinline__ void f_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// J = R / 128 (rows per piece); each A-warp and each B-warp owns R / 4 = 32*J rows of the piece
template <int J>
__global__ void __launch_bounds__(NTHREADS, 1)
k_tprf_fused(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_row0, const Args g) {
  constexpr int R = 128 * J;
  constexpr int RW = R / NAW;      // rows per A-/B-warp: 32, 64, 96 or 128
  constexpr int KS = RW / 4;       // phase-A k-steps (4 rows each)
  constexpr int MT = RW / 8;       // phase-B m-tiles (8 rows each)
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  double* scratch = reinterpret_cast<double*>(smem + OFF_SCRATCH);
  double* statbuf = reinterpret_cast<double*>(smem + OFF_STAT);
  double* vrec = reinterpret_cast<double*>(smem + OFF_VREC);
  double* kshs = reinterpret_cast<double*>(smem + OFF_KSH);
  const unsigned bar0 = sa(smem + OFF_BAR);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };          // TMA landed the piece in slot s
  auto bar_empty = [&](int s) { return bar0 + 8u * (3 + s); };   // the A-warps are done with slot s
  auto bar_st_full = [&](int q, int t) { return bar0 + 8u * (8 + 2 * NXR + q * NT + t); };             // A-warp q stashed a piece in buffer t
  auto bar_st_free = [&](int q, int t) { return bar0 + 8u * (8 + 2 * NXR + NAW * NT + q * NT + t); };  // B-warp 4+q is done with it
  const unsigned bar_scr_full = bar0 + 8u * 6, bar_scr_free = bar0 + 8u * 7;
  auto bar_stat = [&](int x) { return bar0 + 8u * (8 + x); };        // every CTA's partials of my columns arrived
  auto bar_vrdy = [&](int x) { return bar0 + 8u * (8 + NXR + x); };  // every column's record arrived
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + OFF_TMEM);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = g.G;
  const int rank = (int)cluster_rank();
  const int cl = blockIdx.x / G, ncl = gridDim.x / G;
  const int owned = CW / G;
  const i64 T = g.T;
  const int nchunks = g.batch * g.cpp;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mb_init(bar_full(s), 1);
      mb_init(bar_empty(s), NAW);       // the A-warps (phase A + copy to tensor memory)
    }
    for (int x = 0; x < NXR; ++x) {
      mb_init(bar_stat(x), 1);   // armed per piece: expect_tx(STAT_SLOT doubles) = owned columns x G ranks x NSTAT
      mb_init(bar_vrdy(x), 1);   // armed per piece: expect_tx(CW * 9 doubles)     = 16 columns x (V[8], 1/sd)
    }
    mb_init(bar_scr_full, NAW);
    mb_init(bar_scr_free, 1);
    for (int q = 0; q < NAW; ++q)
      for (int t = 0; t < NT; ++t) {
        mb_init(bar_st_full(q, t), 1);
        mb_init(bar_st_free(q, t), 1);
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const unsigned tmem_base = *tmem_slot;
  cluster_sync_all();
  // this cluster is resident: tell the host-side stream wait that holds back the grid sweep which is to
  // fill the SMs the clusters cannot use (it must be placed AFTER them)
  if (g.started != nullptr && rank == 0 && tid == 0) atomicAdd(g.started, 1u);

  const bool dbg_on = g.dbg != nullptr && cl == 0 && lane == 0;
  auto stamp = [&](int iter, int slot_) {
    if (dbg_on && iter < 64) g.dbg[(((i64)rank * 16 + warp) * 64 + iter) * 8 + slot_] = clock64();
  };
  const unsigned slots0 = sa(smem + OFF_SLOTS);
  const int fm = lane >> 2, fk = lane & 3;
  const int rm = (fm >> 1) + 4 * (fm & 1);   // row of an 8-row m-tile held by this lane in the phase-B fragment

  // Register re-allocation between the three warpgroups (warps 0-3 A, 4-7 B, 8-11 producer / finaliser / reducer):
  // the A-warps hold the Zc fragments of their 128 rows (64 registers) on top of the accumulators, and ptxas, capped
  // at 168 registers, sank every shared-memory load to just in front of its use -- a lookahead of ~50 clocks, less
  // than the load's latency under the TMA writes and the B-warps' reads.  The special warps give registers back.
  if (warp < NAW) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_A));
  } else if (warp >= 2 * NAW) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_S));
  }
  if (warp < NAW) {
    // ================= A-warps: column statistics of piece `it` =================
    const int rowbase = warp * RW;   // first local row of this warp (multiple of 32)
    // one 16-byte load feeds both m-tiles: lane (m = fm, k = fk) reads x[row 4ks + k][columns 2c, 2c + 1] with
    // c = chunk(fm); m-tile 0 = the even columns, m-tile 1 = the odd ones.  chunk() makes the 8 lanes of a
    // quarter-warp (two fm values x four rows) touch 8 distinct 16-byte bank groups under the 128-byte swizzle.
    const int chunk = ((fm & 1) << 2) | (fm >> 1);
    const int cola[2] = {2 * chunk, 2 * chunk + 1};
    unsigned offv[2];
#pragma unroll
    for (int pb = 0; pb < 2; ++pb) offv[pb] = (unsigned)(((chunk ^ (4 * pb + fk)) & 7) << 4) + (unsigned)(rowbase + 4 * pb + fk) * 128u;
    // The A-warp also copies its rows, in the phase-B fragment layout, into TENSOR MEMORY for B-warp 4 + warp (same
    // sub-partition = same tensor-memory lane quarter): the slot is released by the A-warps alone, the B-warps never
    // touch shared-memory slots and run phase B up to NT pieces later.  lane (rm, fk) holds x[row 8m + rm][cols 2fk,
    // 2fk+1 and 8+2fk, 8+2fk+1].
    unsigned offb[2];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) offb[s2] = (unsigned)((((4 * s2 + fk) ^ rm) & 7) << 4) + (unsigned)(rowbase + rm) * 128u;
    const unsigned tm_warp = tmem_base + ((unsigned)(32 * warp) << 16);
    int it = 0;
    for (int ch = cl; ch < nchunks; ch += ncl) {
      const int b = ch / g.cpp, cidx = ch % g.cpp;
      // Zc fragments of this warp's rows (B operand): lane holds zc[row 4ks + k][l = fm]
      double zf[KS];
      const i64 grow0 = (i64)rank * R + rowbase;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const i64 t = grow0 + 4 * ks + fk;
        zf[ks] = t < T ? g.zc[((i64)b * T + t) * 8 + fm] : 0.0;
      }
      const i64 tv = T - grow0 - fk;                 // row 4ks + k is inside the panel iff 4ks < tv
      const int tvalid = tv > RW ? RW : (tv < 0 ? 0 : (int)tv);
      const bool need_mask = grow0 + RW > T;
      for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) {
        const int s = it % NSLOT, tb = it % NT;
        stamp(it, 0);
        const unsigned tokf = mb_wait_token(bar_full(s), (unsigned)((it / NSLOT) & 1));
        stamp(it, 1);
        const unsigned slot = slots0 + (unsigned)s * SLOT_BYTES + tokf;
        if (it >= NT) {   // the B-warp has finished with this stash buffer
          mb_wait(bar_st_free(warp, tb), (unsigned)((it / NT - 1) & 1));
          tmem_fence_after();
        }
        const unsigned tm = tm_warp + (unsigned)(tb * (MT * 8));
        double q[2][2][2], s1[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, s2[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int pb = 0; pb < 2; ++pb) q[mt][pb][0] = q[mt][pb][1] = 0.0;
        const double2 kshv = ldp_v2f64(sa(kshs + (it % NK) * CW + cola[0]) + tokf);
        const double ksh[2] = {kshv.x, kshv.y};
        // One GROUP = two k-steps (rows 8g + fk and 8g + 4 + fk) x two m-tiles = four DMMAs on four INDEPENDENT
        // accumulators, then a scheduling fence.  Without the fence ptxas sorts the fully unrolled body chain by
        // chain (16 dependent DMMAs in a row; a dependent DMMA issues every ~26 clocks instead of every 16,
        // profiles/r1_dmma_chain.log).  The loads of group g + 1 are issued before the math of group g.
        auto phase_a = [&](auto masked) {
          constexpr bool MASK = decltype(masked)::value;
          // The DMMA takes the SHIFTED value d = x - k, not x: Zc is centred, so sum_t (x_t - k) zc_t = sum_t x_t zc_t,
          // and d -- which the two sums need anyway -- then has to be ready BEFORE the DMMA issues, whose 16 clocks
          // cover the latency of d for the sums behind it (with x as the operand ptxas emitted DMMA, d = x - k,
          // 8 idle clocks, s1 += d).  Loads run AHEAD groups (of two k-steps x two m-tiles) ahead: under the TMA
          // writes and the B-warps' reads a shared-memory load takes far longer than in isolation (ncu: 20 % of the
          // A-warp's time was short-scoreboard stall with one group of lookahead).
          constexpr int NG = KS / 2;
          constexpr int AHEAD = NG >= 3 ? 3 : NG;
          double2 a[AHEAD + 1][2], xb[AHEAD + 1][2];
#pragma unroll
          for (int h = 0; h < AHEAD; ++h) {
#pragma unroll
            for (int pb = 0; pb < 2; ++pb) a[h][pb] = ldp_v2f64(slot + offv[pb] + (unsigned)h * 1024u);
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) xb[h][s2] = ldp_v2f64(slot + offb[s2] + (unsigned)h * 1024u);
          }
#pragma unroll
          for (int gk = 0; gk < NG; ++gk) {
            if (gk + AHEAD < NG) {
#pragma unroll
              for (int pb = 0; pb < 2; ++pb) a[AHEAD][pb] = ldp_v2f64(slot + offv[pb] + (unsigned)(gk + AHEAD) * 1024u);
#pragma unroll
              for (int s2 = 0; s2 < 2; ++s2) xb[AHEAD][s2] = ldp_v2f64(slot + offb[s2] + (unsigned)(gk + AHEAD) * 1024u);
            }
            double d[2][2];
#pragma unroll
            for (int pb = 0; pb < 2; ++pb) {
              d[pb][0] = a[0][pb].x - ksh[0];
              d[pb][1] = a[0][pb].y - ksh[1];
              if (MASK && 4 * (2 * gk + pb) >= tvalid) d[pb][0] = d[pb][1] = 0.0;
            }
#pragma unroll
            for (int pb = 0; pb < 2; ++pb)
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) {
#ifndef ABL_NODMMA_A
                f_dmma(q[mt][pb][0], q[mt][pb][1], d[pb][mt], zf[2 * gk + pb]);
#else
                q[mt][pb][0] += d[pb][mt];
#endif
#ifndef ABL_NOSTATS
                s1[mt][pb] += d[pb][mt];
                s2[mt][pb] = fma(d[pb][mt], d[pb][mt], s2[mt][pb]);
#endif
              }
            tmem_st8(tm + (unsigned)(gk * 8), xb[0][0], xb[0][1]);   // group gk = rows 8 gk .. 8 gk + 7 = m-tile gk of the stash
            sched_fence();   // keeps the four accumulator chains interleaved (ptxas would sort them chain by chain)
#pragma unroll
            for (int h = 0; h < AHEAD; ++h) {
#pragma unroll
              for (int pb = 0; pb < 2; ++pb) a[h][pb] = a[h + 1][pb];
#pragma unroll
              for (int s2 = 0; s2 < 2; ++s2) xb[h][s2] = xb[h + 1][s2];
            }
          }
        };
        if (need_mask) phase_a(std::true_type{});
        else phase_a(std::false_type{});
        // done with the slot; the stash buffer goes to the B-warp
        tmem_wait_st();
        tmem_fence_before();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) consume(q[mt][0][0] + q[mt][1][0]);   // every load of the slot has landed in a register
        __syncwarp();
        if (lane == 0) {
          mb_arrive(bar_st_full(warp, tb));
          mb_arrive(bar_empty(s));
        }
        // per-lane partials -> scratch (the reducer adds the 4 row phases and the 4 A-warps)
        mb_wait(bar_scr_free, (unsigned)((it & 1) ^ 1));  // the reducer is done with piece it-1 (passes at once for it = 0)
        {
          const unsigned sp = sa(scratch + warp * SCR_W);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            sts_v2f64(sp + (cola[mt] * 8 + 2 * fk) * 8, q[mt][0][0] + q[mt][1][0], q[mt][0][1] + q[mt][1][1]);
            sts_f64(sp + (128 + cola[mt] * 4 + fk) * 8, s1[mt][0] + s1[mt][1]);
            sts_f64(sp + (192 + cola[mt] * 4 + fk) * 8, s2[mt][0] + s2[mt][1]);
          }
        }
        __syncwarp();
        if (lane == 0) mb_arrive(bar_scr_full);
        stamp(it, 2);
      }
    }
  } else if (warp < 2 * NAW) {
    // ================= B-warps: PX / h0 of piece n out of tensor memory, once every column's record is there =================
    const int bw = warp - NAW;
    const int rowbase = bw * RW;
    const unsigned tm_warp = tmem_base + ((unsigned)(32 * bw) << 16);
    double px[MT][2], hh[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) px[i][0] = px[i][1] = hh[i] = 0.0;

    auto phase_b = [&](int n) {
      const int tb = n % NT, xr = n % NXR;
      stamp(n, 3);
      const unsigned tokv = mb_wait_token(bar_vrdy(xr), (unsigned)((n / NXR) & 1));
      stamp(n, 4);
      const unsigned vr = sa(vrec + xr * CW * NREC) + tokv;
      double vf[2][2], isd[2][2];
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = 8 * s2 + 2 * fk + e;
          vf[s2][e] = ldp_f64(vr + (c * NREC + fm) * 8);
          isd[s2][e] = ldp_f64(vr + (c * NREC + 8) * 8);
        }
      mb_wait(bar_st_full(bw, tb), (unsigned)((n / NT) & 1));   // A-warp bw has stashed piece n (long ago, normally)
      tmem_fence_after();
      const unsigned tm = tm_warp + (unsigned)(tb * (MT * 8));
      // the tensor-memory loads of the next pair of m-tiles are in flight while this pair is multiplied (waiting for
      // each pair in turn left the warp idle for the load latency eight times per piece: 0.4 % of the fit)
      unsigned rr[2][2][8];
#pragma unroll
      for (int i = 0; i < 2; ++i) tmem_ld8(tm + (unsigned)(i * 8), rr[0][i]);
#pragma unroll
      for (int m0 = 0; m0 < MT; m0 += 2) {
        tmem_wait_ld();
        unsigned (&r)[2][8] = rr[(m0 >> 1) & 1];
        if (m0 + 2 < MT) {
#pragma unroll
          for (int i = 0; i < 2; ++i) tmem_ld8(tm + (unsigned)((m0 + 2 + i) * 8), rr[((m0 >> 1) & 1) ^ 1][i]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const double x0 = __hiloint2double(r[i][4 * s2 + 1], r[i][4 * s2]);
            const double x1 = __hiloint2double(r[i][4 * s2 + 3], r[i][4 * s2 + 2]);
#ifndef ABL_NODMMA_B
            f_dmma(px[m0 + i][0], px[m0 + i][1], x0, vf[s2][0]);
            f_dmma(px[m0 + i][0], px[m0 + i][1], x1, vf[s2][1]);
#else
            px[m0 + i][0] += x0; px[m0 + i][1] += x1;
#endif
#ifndef ABL_NOHH
            hh[m0 + i] = fma(x0, isd[s2][0], hh[m0 + i]);
            hh[m0 + i] = fma(x1, isd[s2][1], hh[m0 + i]);
#endif
          }
      }
      tmem_fence_before();
      __syncwarp();
      if (lane == 0) mb_arrive(bar_st_free(bw, tb));
      stamp(n, 5);
    };

    int it = 0;
    for (int ch = cl; ch < nchunks; ch += ncl) {
      const int b = ch / g.cpp, cidx = ch % g.cpp;
      const i64 grow0 = (i64)rank * R + rowbase;
      for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) phase_b(it);
      // flush this chunk's PX / h0 partial
      {
        double* pxo = g.pxp + ((i64)(b * g.cpp_out + g.cidx_off + cidx) * g.tpad) * 8;
        double* h0o = g.h0p + (i64)(b * g.cpp_out + g.cidx_off + cidx) * g.tpad;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          double h = hh[mt];
          h += __shfl_xor_sync(0xffffffffu, h, 1);
          h += __shfl_xor_sync(0xffffffffu, h, 2);
          const i64 t = grow0 + mt * 8 + rm;
          if (t < T) {
            *reinterpret_cast<double2*>(pxo + t * 8 + 2 * fk) = make_double2(px[mt][0], px[mt][1]);
            if (fk == 0) h0o[t] = h;
          }
          px[mt][0] = px[mt][1] = hh[mt] = 0.0;
        }
      }
    }
  } else if (warp == 2 * NAW) {
    // ================= TMA producer =================
    if (lane == 0) {
      constexpr unsigned piece_bytes = (unsigned)R * CW * 8 + CW * 8;
      int it = 0;
      for (int ch = cl; ch < nchunks; ch += ncl) {
        const int b = ch / g.cpp, cidx = ch % g.cpp;
        for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) {
          const int s = it % NSLOT;
          if (it >= NSLOT) mb_wait(bar_empty(s), (unsigned)((it / NSLOT - 1) & 1));
          mb_expect_tx(bar_full(s), piece_bytes);
          const unsigned dst = slots0 + (unsigned)s * SLOT_BYTES;
#pragma unroll
          for (int bx = 0; bx < J; ++bx)
            tma_3d(dst + bx * BOX_ROWS * CW * 8, &map_x, (int)(tile * CW), rank * R + bx * BOX_ROWS, b, bar_full(s));
          tma_3d(sa(kshs + (it % NK) * CW), &map_row0, (int)(tile * CW), 0, b, bar_full(s));
        }
      }
    }
  } else if (warp == 2 * NAW + 1 || warp == 2 * NAW + 3) {
    // ================= finaliser: owner of columns {oc*G + rank} =================
    // Clusters of 1 or 2 CTAs own 16 / 8 columns per CTA = 4 / 2 serial rounds per piece, and this latency-bound
    // warp then paces the whole sweep (profiles/r1_fused_timeline_g1.log): a second finaliser warp takes every
    // other round.  Larger clusters (one round per piece) run the first one alone; the second stays idle.
    // Larger clusters (one round per piece): the two finaliser warps take the PIECES in turn -- the per-piece chain
    // (sum the G partials, 1/sd, V, push, outputs, small sums: ~2500 clocks on a sub-partition whose FP64 pipe is busy
    // with DMMAs) is otherwise the serial stage that paces the sweep once phase A is fast.
    const int fw = warp == 2 * NAW + 3 ? 1 : 0;
    const bool piece_split = owned <= 4;
    const int nfw = piece_split ? 1 : 2;
    const int fwr = piece_split ? 0 : fw;      // round offset inside a piece (small clusters only)
    double* fin2 = reinterpret_cast<double*>(smem + OFF_FIN2);
    // 8 lanes (l = value index) per lane group; `sub` lane groups share one column and split its G
    // source ranks (<= 4 each); up to 4 columns per round
    const int lg = lane >> 3, l = lane & 7;
    const int cpr = owned < 4 ? owned : 4;
    const int sub = 4 / cpr;
    const int cs = lg / sub, rs = lg % sub;
    const int nr = G / sub;               // ranks this lane sums / pushes to: min(G, 4)
    unsigned rvrec[4], rvbar[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const unsigned dstrank = (unsigned)(rs * nr + (r < nr ? r : 0));
      rvrec[r] = map_rank(sa(vrec), dstrank);
      rvbar[r] = map_rank(bar_vrdy(0), dstrank);
    }
    double a_sw = 0.0, a_smw = 0.0, a_smu = 0.0, a_sww[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a_sww[c] = 0.0;
    const double invT = 1.0 / (double)T, invTm1 = T > 1 ? 1.0 / (double)(T - 1) : 0.0;
    int it = 0;
    for (int ch = cl; ch < nchunks; ch += ncl) {
      const int b = ch / g.cpp, cidx = ch % g.cpp;
      for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) {
        const int s = it % NXR;  // ring entry of the exchange buffers and barriers
        if (piece_split && (it & 1) != fw) continue;
        if (lane == 0 && (piece_split || fw == 0)) {
          mb_expect_tx(bar_vrdy(s), CW * 9 * 8);
          mb_expect_tx(bar_stat(s), STAT_SLOT * 8);
        }
        stamp(it, 0);
        mb_wait(bar_stat(s), (unsigned)((it / NXR) & 1));
        stamp(it, 1);
        const unsigned sb = sa(statbuf + s * STAT_SLOT);
        // Small groups own many columns (16 at G = 1): the long chain S, SS -> 1/sd is then done ONCE, one
        // column per lane, before the rounds, instead of 8 times redundantly in each of 4 serial rounds.
        const bool pre = owned > 4;
        double inv_pre = 1.0, mun_pre = 0.0;
        if (pre && lane < owned) {
          double S = 0.0, SS = 0.0;
          for (int r = 0; r < G; ++r) {   // G <= 2: rank order
            S += lds_f64(sb + (unsigned)((lane * G + r) * NSTAT) * 8);
            SS += lds_f64(sb + (unsigned)((lane * G + r) * NSTAT + 1) * 8);
          }
          const double k0 = lds_f64(sa(kshs + (it % NK) * CW + lane * G + rank));
          const double dm = S * invT;
          if (T > 1) {
            double var = fma(-S, dm, SS) * invTm1;
            if (var < 0.0) var = 0.0;
            if (var != 0.0) {
              double y;
              asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(var));
              const double hv = 0.5 * var;
              y = y * fma(-hv * y, y, 1.5);
              y = y * fma(-hv * y, y, 1.5);
              inv_pre = (var < 1e-290 || var > 1e290) ? 1.0 / sqrt(var) : y;
            }
          }
          mun_pre = (k0 + dm) * inv_pre;
        }
        for (int oc0 = fwr * cpr; oc0 < owned; oc0 += cpr * nfw) {
          const int oc = oc0 + cs;              // < owned always (owned is a multiple of cpr)
          const int col = oc * G + rank;
          double S, SS, q;
          {
            double ps[4], pss[4], pq[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const bool on = r < nr;
              const unsigned pa = sb + (unsigned)((oc * G + rs * nr + (on ? r : 0)) * NSTAT) * 8;
              ps[r] = on ? lds_f64(pa) : 0.0;
              pss[r] = on ? lds_f64(pa + 8) : 0.0;
              pq[r] = on ? lds_f64(pa + (2 + l) * 8) : 0.0;
            }
            // fixed tree: ranks inside the lane, then lane groups (xor 8, 16)
            S = (ps[0] + ps[1]) + (ps[2] + ps[3]);
            SS = (pss[0] + pss[1]) + (pss[2] + pss[3]);
            q = (pq[0] + pq[1]) + (pq[2] + pq[3]);
            if (sub >= 2) {
              S += __shfl_xor_sync(0xffffffffu, S, 8);
              SS += __shfl_xor_sync(0xffffffffu, SS, 8);
              q += __shfl_xor_sync(0xffffffffu, q, 8);
            }
            if (sub == 4) {
              S += __shfl_xor_sync(0xffffffffu, S, 16);
              SS += __shfl_xor_sync(0xffffffffu, SS, 16);
              q += __shfl_xor_sync(0xffffffffu, q, 16);
            }
          }
          const i64 gcol = tile * CW + col;
          const bool valid = gcol < g.N;
          const double k0 = lds_f64(sa(kshs + (it % NK) * CW + col));
          // short dependent chain (this warp shares the FP64 pipe with the DMMA stream, every dependent
          // FP64 instruction costs a pipe round trip): reciprocals instead of divisions, 1/sd from an
          // rsqrt seed + two Newton steps (relative error ~1e-16; the parity bar is 1e-9)
          double inv = 1.0, m_n;
          if (pre) {
            inv = __shfl_sync(0xffffffffu, inv_pre, oc);
            m_n = __shfl_sync(0xffffffffu, mun_pre, oc);
          } else {
            const double dm = S * invT;
            const double mu = k0 + dm;
            if (T > 1) {
              double var = fma(-S, dm, SS) * invTm1;
              if (var < 0.0) var = 0.0;
              if (var != 0.0) {                      // sd == 0 -> 1 (Tprf3.scala:493-501); NaN propagates
                double y;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(var));
                const double hv = 0.5 * var;
                y = y * fma(-hv * y, y, 1.5);   // seed is good to ~2^-23: 2^-45 after one step,
                y = y * fma(-hv * y, y, 1.5);   // full double after two
                inv = (var < 1e-290 || var > 1e290) ? 1.0 / sqrt(var) : y;
              }
            }
            m_n = mu * inv;
          }
          const double w = q * inv;
          const double v = w * inv;
          // record of this column -> every CTA of the cluster; this lane serves ranks [rs*nr, rs*nr + nr)
          {
            const unsigned o_v = (unsigned)((s * CW + col) * NREC + l) * 8;
            const unsigned o_i = (unsigned)((s * CW + col) * NREC + 8) * 8;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              if (r < nr) {
                st_async(rvrec[r] + o_v, v, rvbar[r] + 8u * s);
                if (l == 0) st_async(rvrec[r] + o_i, inv, rvbar[r] + 8u * s);
              }
            }
          }
          stamp(it, 2);
          const bool mine = valid && rs == 0;   // one lane per (column, l) writes / accumulates
          if (mine) {
            const i64 cix = (i64)b * g.npad + gcol;
            g.w0[cix * 8 + l] = w;
            if (l == 0) {
              g.invsd[cix] = inv;
              g.mun[cix] = m_n;
            }
          }
          if (!pre) {   // (small groups: k_tprf_smallsums forms these sums from W0 / mun after the sweep)
            const double wv = mine ? w : 0.0, mv = mine ? m_n : 0.0;
            a_sw += wv;
            a_smw += mv * wv;
            if (l == 0) a_smu += mv;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const double wc = __shfl_sync(0xffffffffu, wv, (lane & 24) | c);
              a_sww[c] += wv * wc;
            }
          }
        }
      }
      // chunk done: flush the small sums of this (panel, chunk, rank); fixed order over the 4 lane groups
      {
        double vals[11];
        vals[0] = a_sw; vals[9] = a_smw; vals[10] = a_smu;
#pragma unroll
        for (int c = 0; c < 8; ++c) vals[1 + c] = a_sww[c];
#pragma unroll
        for (int i = 0; i < 11; ++i) {
          double x0 = __shfl_sync(0xffffffffu, vals[i], l);
          double x1 = __shfl_sync(0xffffffffu, vals[i], 8 + l);
          double x2 = __shfl_sync(0xffffffffu, vals[i], 16 + l);
          double x3 = __shfl_sync(0xffffffffu, vals[i], 24 + l);
          vals[i] = ((x0 + x1) + x2) + x3;
        }
        if (piece_split) {   // second finaliser -> first, through shared memory; the sum order is fixed (first + second)
          if (fw == 1 && lg == 0) {
#pragma unroll
            for (int i = 0; i < 11; ++i) fin2[i * 8 + l] = vals[i];
          }
          asm volatile("bar.sync 1, 64;" ::: "memory");
          if (fw == 0) {
#pragma unroll
            for (int i = 0; i < 11; ++i) vals[i] += fin2[i * 8 + l];
          }
          asm volatile("bar.sync 1, 64;" ::: "memory");   // the buffer may be rewritten at the end of the next chunk
        }
        if (lg == 0 && fw == 0) {
          double* bp = g.blockpart + ((i64)(b * g.cpp_out + g.cidx_off + cidx) * G + rank) * NSMALL;
          bp[l] = vals[0];
#pragma unroll
          for (int c = 0; c < 8; ++c) bp[8 + l * 8 + c] = vals[1 + c];
          bp[72 + l] = vals[9];
          if (l == 0) bp[80] = vals[10];
        }
        a_sw = a_smw = a_smu = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) a_sww[c] = 0.0;
      }
    }
  } else if (warp == 2 * NAW + 2) {
    // ================= reducer: lane partials in `scratch` -> CTA partial -> owner CTA =================
    // lane handles 5 of the 160 (column, statistic) values: value 0 is s1 / s2 of column lane >> 1 (4 row
    // phases x 4 A-warps), values 1..4 are q[column (e >> 3)][l = e & 7], e = lane + 32 (j - 1) (4 A-warps)
    unsigned rdst[5], rbar[5];
    {
      const int c = lane >> 1, st = lane & 1;
      const int owner = c % G, oc = c / G;
      rdst[0] = map_rank(sa(statbuf + (oc * G + rank) * NSTAT + st), owner);
      rbar[0] = map_rank(bar_stat(0), owner);
    }
#pragma unroll
    for (int j = 1; j < 5; ++j) {
      const int e = lane + 32 * (j - 1);
      const int c = e >> 3, l = e & 7;
      const int owner = c % G, oc = c / G;
      rdst[j] = map_rank(sa(statbuf + (oc * G + rank) * NSTAT + 2 + l), owner);
      rbar[j] = map_rank(bar_stat(0), owner);
    }
    const unsigned sbase = sa(scratch + 128 + (lane & 1) * 64 + (lane >> 1) * 4);
    const unsigned qbase = sa(scratch + lane);
    int it = 0;
    for (int ch = cl; ch < nchunks; ch += ncl) {
      const int cidx = ch % g.cpp;
      for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) {
        const int s = it % NXR;
        stamp(it, 0);
        mb_wait(bar_scr_full, (unsigned)(it & 1));
        stamp(it, 1);
        double p[NAW][4], pq[4][NAW];
#pragma unroll
        for (int w = 0; w < NAW; ++w)
#pragma unroll
          for (int k = 0; k < 4; ++k) p[w][k] = lds_f64(sbase + (w * SCR_W + k) * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int w = 0; w < NAW; ++w) pq[j][w] = lds_f64(qbase + (w * SCR_W + 32 * j) * 8);
        double acc[5];
        {
          double pw[NAW];
#pragma unroll
          for (int w = 0; w < NAW; ++w) pw[w] = (p[w][0] + p[w][1]) + (p[w][2] + p[w][3]);
          acc[0] = (pw[0] + pw[1]) + (pw[2] + pw[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[1 + j] = (pq[j][0] + pq[j][1]) + (pq[j][2] + pq[j][3]);
        consume(acc[0]); consume(acc[1]); consume(acc[2]); consume(acc[3]); consume(acc[4]);
        __syncwarp();
        if (lane == 0) mb_arrive(bar_scr_free);
#pragma unroll
        for (int j = 0; j < 5; ++j) st_async(rdst[j] + (unsigned)s * (STAT_SLOT * 8), acc[j], rbar[j] + 8u * s);
        stamp(it, 2);
      }
    }
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  cluster_sync_all();  // nobody leaves while a peer may still write into its shared memory
}

typedef void (*FusedKernel)(const CUtensorMap, const CUtensorMap, const Args);
inline FusedKernel kernel_for(int R) {
  switch (R / 128) {
    case 1: return k_tprf_fused<1>;
    case 2: return k_tprf_fused<2>;
    case 3: return k_tprf_fused<3>;
    default: return k_tprf_fused<4>;
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
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// cluster size / rows per piece for a T-row panel; G == 0: the fused sweep does not apply
inline void pick_shape(i64 T, int* G, int* R) {
  *G = 0; *R = 0;
  for (int gsz = 1; gsz <= GMAX; gsz *= 2) {
    i64 rows = (T + gsz - 1) / gsz;
    if (rows <= RMAX) {
      *G = gsz;
      *R = (int)((rows + BOX_ROWS - 1) / BOX_ROWS) * BOX_ROWS;
      return;
    }
  }
}

}  // namespace fused
}  // namespace um
