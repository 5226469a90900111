// tprf_sym.cuh -- the single-read 3PRF cluster sweep with SYMMETRIC compute warps (included by tprf.cu after
// tprf_fused.cuh, whose PTX wrappers, fragment maps and Args it reuses).
//
// Geometry is that of tprf_fused.cuh: a cluster of G CTAs keeps a column tile (T rows x 16 columns) on chip, CTA
// `rank` owns rows [rank*R, rank*R + R) and pulls its (R x 16) piece with TMA into one of three 64 KiB slots.
// What changed is who does the arithmetic.  tprf_fused.cuh had one A-warp (column statistics) and one B-warp
// (projection) per SM sub-partition; ncu (profiles/r2_fused_asym_stalls.txt) showed the A-warp on the critical path
// with its own fixed-latency stalls (1800 clocks per piece) and the B-warp's FP64-pipe time (1150 clocks) simply
// ADDING UP, while the B-warp idled a third of the time waiting for work: 2560 pipe clocks took 3900.
//
// Here all 16 compute warps (four per sub-partition) are alike.  Warp w owns rows [w*RW, w*RW + RW) of every piece
// (RW = R/16 = 8, 16, 24 or 32) and does, for piece n:
//     phase A  column statistics of its rows: Zc' x by DMMA (M = 8 columns, K = rows) + shifted sum / sum of squares,
//              and, from the same slot, its rows once more in the phase-B fragment layout -> TENSOR MEMORY
//     phase B  of piece n - LAG out of tensor memory: PX += x V (DMMA, M = 8 rows, K = columns), h0 += x / sd
// so every warp always has work, the four warps of a sub-partition cover one another's latencies, and the slot goes
// back to the TMA producer as soon as phase A has read it.  The exchange is unchanged: per-warp partials -> `scratch`
// -> reducer warp -> owner CTA (st.async into distributed shared memory, completing bytes on the owner's mbarrier)
// -> finaliser warp (mu, sd, W0, V) -> record to every CTA of the cluster.  Every sum has a fixed order.
#pragma once

namespace um {
namespace sym {

using fused::Args;
using fused::CW;
using fused::NSLOT;
using fused::NT;
using fused::NK;
using fused::RMAX;
using fused::SLOT_BYTES;
using fused::BOX_ROWS;
using fused::NSTAT;
using fused::NREC;
using fused::GMAX;
using fused::STAT_SLOT;
using fused::sa;
using fused::mb_init;
using fused::mb_expect_tx;
using fused::mb_arrive;
using fused::mb_wait;
using fused::mb_wait_token;
using fused::map_rank;
using fused::st_async;
using fused::ldp_f64;
using fused::ldp_v2f64;
using fused::lds_f64;
using fused::sts_f64;
using fused::sts_v2f64;
using fused::consume;
using fused::tma_3d;
using fused::cluster_sync_all;
using fused::cluster_rank;
using fused::tmem_st8;
using fused::tmem_ld8;
using fused::tmem_wait_ld;
using fused::tmem_wait_st;
using fused::tmem_fence_before;
using fused::tmem_fence_after;
using fused::col_of_m;
using fused::f_dmma;

constexpr int NAW = 8;                       // A-warps (= B-warps): TWO of each per SM sub-partition
constexpr int NCW = 2 * NAW;                 // compute warps
constexpr int W_TMA = NCW, W_FIN = NCW + 1, W_RED = NCW + 2, W_FIN2 = NCW + 3;
constexpr int NTHREADS = (NCW + 4) * 32;     // 640
constexpr int LAG = 2;                       // phase B runs this many pieces behind phase A (NT = 4 stash buffers)
// Ring depth of the partial / record buffers and their barriers.  The A-warps are throttled only by the three landing
// slots, whose release needs the B-warps' copy of piece a-3, which follows phase B of piece a-4-LAG in program order:
// the finaliser may be up to 6 pieces behind the A-warps.  The exchange rings must cover that.
constexpr int NXR = 8;
constexpr int SCR_W = 160;                   // scratch doubles per compute warp: q[16 cols][8] | s1[16] | s2[16]
constexpr int DBG_WARPS = 32;                // warps per rank in the debug stamp buffer

constexpr int OFF_SLOTS = 0;
constexpr int OFF_SCRATCH = NSLOT * SLOT_BYTES;                 // double[NAW][SCR_W]
constexpr int OFF_STAT = OFF_SCRATCH + NAW * SCR_W * 8;         // double[NXR][STAT_SLOT]
constexpr int OFF_VREC = OFF_STAT + NXR * STAT_SLOT * 8;        // double[NXR][CW][NREC]
constexpr int OFF_KSH = OFF_VREC + NXR * CW * NREC * 8;         // double[NK][CW]   (row 0 of the panel: variance shift)
constexpr int OFF_BAR = OFF_KSH + NK * CW * 8;                  // full[3] empty[3] scr_full scr_free statready[NXR] vready[NXR]
constexpr int NBAR = 8 + 2 * NXR;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;                    // u32: TMEM base address
constexpr int SMEM_BYTES = OFF_TMEM + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;                   // slack to align the base to 1024 (swizzle atom)
static_assert(SMEM_ALLOC <= 232448, "shared memory of the symmetric sweep exceeds 227 KB");
static_assert(NSLOT + 2 + LAG < NXR && LAG < NT, "exchange rings / stash buffers too shallow for LAG");

// J = R / 128 (rows per piece); every compute warp owns RW = 8 J rows of the piece
template <int J>
__global__ void __launch_bounds__(NTHREADS, 1)
k_tprf_sym(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_row0, const Args g) {
  constexpr int R = 128 * J;
  constexpr int RW = R / NAW;      // rows per A-warp and per B-warp: 16, 32, 48 or 64
  constexpr int KS = RW / 4;       // phase-A k-steps (4 rows each) = 4 J
  constexpr int MT = RW / 8;       // m-tiles (8 rows each) = 2 J
  constexpr int TCOLS = MT * 8;    // 32-bit TMEM columns of one stash buffer of one warp
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  double* scratch = reinterpret_cast<double*>(smem + OFF_SCRATCH);
  double* statbuf = reinterpret_cast<double*>(smem + OFF_STAT);
  double* vrec = reinterpret_cast<double*>(smem + OFF_VREC);
  double* kshs = reinterpret_cast<double*>(smem + OFF_KSH);
  const unsigned bar0 = sa(smem + OFF_BAR);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };          // TMA landed the piece in slot s
  auto bar_empty = [&](int s) { return bar0 + 8u * (3 + s); };   // every compute warp has read its rows of slot s
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
      mb_init(bar_empty(s), NCW);
    }
    for (int x = 0; x < NXR; ++x) {
      mb_init(bar_stat(x), 1);   // armed per piece: expect_tx(STAT_SLOT doubles) = owned columns x G ranks x NSTAT
      mb_init(bar_vrdy(x), 1);   // armed per piece: expect_tx(CW * 9 doubles)     = 16 columns x (V[8], 1/sd)
    }
    mb_init(bar_scr_full, NAW);
    mb_init(bar_scr_free, 1);
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
    if (dbg_on && iter < 64) g.dbg[(((i64)rank * DBG_WARPS + warp) * 64 + iter) * 8 + slot_] = clock64();
  };
  const unsigned slots0 = sa(smem + OFF_SLOTS);
  const int fm = lane >> 2, fk = lane & 3;
  const int rm = (fm >> 1) + 4 * (fm & 1);   // row of an 8-row m-tile held by this lane in the phase-B fragment

  if (warp < NAW) {
    // ================= A-warps: column statistics of piece `it` (rows [warp*RW, +RW)) =================
    const int rowbase = warp * RW;   // first local row of this warp (multiple of 16)
    // one 16-byte load feeds both m-tiles: lane (m = fm, k = fk) reads x[row 4ks + k][columns 2c, 2c + 1] with
    // c = chunk(fm); m-tile 0 = the even columns, m-tile 1 = the odd ones.  chunk() makes the 8 lanes of a
    // quarter-warp (two fm values x four rows) touch 8 distinct 16-byte bank groups under the 128-byte swizzle.
    const int chunk = ((fm & 1) << 2) | (fm >> 1);
    const int cola[2] = {2 * chunk, 2 * chunk + 1};
    unsigned offv[2];
#pragma unroll
    for (int pb = 0; pb < 2; ++pb) offv[pb] = (unsigned)(((chunk ^ (4 * pb + fk)) & 7) << 4) + (unsigned)(rowbase + 4 * pb + fk) * 128u;
    const unsigned scr = sa(scratch + warp * SCR_W);
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
        const int s = it % NSLOT;
        stamp(it, 0);
        const unsigned tokf = mb_wait_token(bar_full(s), (unsigned)((it / NSLOT) & 1));
        stamp(it, 1);
        const unsigned slot = slots0 + (unsigned)s * SLOT_BYTES + tokf;
        double q[2][2][2], s1[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, s2[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int pb = 0; pb < 2; ++pb) q[mt][pb][0] = q[mt][pb][1] = 0.0;
        const double2 kshv = ldp_v2f64(sa(kshs + (it % NK) * CW + cola[0]) + tokf);
        const double ksh[2] = {kshv.x, kshv.y};
        // The DMMA takes the SHIFTED value d = x - k, not x: Zc is centred, so sum_t (x_t - k) zc_t = sum_t x_t zc_t,
        // and d -- which the two sums need anyway -- is then ready before the DMMA issues, whose 16 clocks cover its
        // latency for the sums behind it.  One group = two k-steps x two m-tiles = four DMMAs on four independent
        // accumulators; the fence keeps ptxas from sorting them chain by chain; loads run one group ahead.
        auto phase_a = [&](auto masked) {
          constexpr bool MASK = decltype(masked)::value;
          double2 a_cur[2], a_nxt[2];
#pragma unroll
          for (int pb = 0; pb < 2; ++pb) a_cur[pb] = ldp_v2f64(slot + offv[pb]);
#pragma unroll
          for (int gk = 0; gk < KS / 2; ++gk) {
            if (gk + 1 < KS / 2) {
#pragma unroll
              for (int pb = 0; pb < 2; ++pb) a_nxt[pb] = ldp_v2f64(slot + offv[pb] + (unsigned)(gk + 1) * 1024u);
            }
            double d[2][2];
#pragma unroll
            for (int pb = 0; pb < 2; ++pb) {
              d[pb][0] = a_cur[pb].x - ksh[0];
              d[pb][1] = a_cur[pb].y - ksh[1];
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
            fused::sched_fence();
#pragma unroll
            for (int pb = 0; pb < 2; ++pb) a_cur[pb] = a_nxt[pb];
          }
        };
        if (need_mask) phase_a(std::true_type{});
        else phase_a(std::false_type{});
        // done with the slot (the B-warps release it too, once they have copied their rows to tensor memory)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) consume(q[mt][0][0] + q[mt][1][0]);   // every load of the slot has landed in a register
        __syncwarp();
        if (lane == 0) mb_arrive(bar_empty(s));
        // warp partials -> scratch: q per (column, l); s1 / s2 per column after the fixed xor tree over the 4 row phases
        double t1[2], t2[2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          t1[mt] = s1[mt][0] + s1[mt][1];
          t2[mt] = s2[mt][0] + s2[mt][1];
          t1[mt] += __shfl_xor_sync(0xffffffffu, t1[mt], 1);
          t2[mt] += __shfl_xor_sync(0xffffffffu, t2[mt], 1);
          t1[mt] += __shfl_xor_sync(0xffffffffu, t1[mt], 2);
          t2[mt] += __shfl_xor_sync(0xffffffffu, t2[mt], 2);
        }
        mb_wait(bar_scr_free, (unsigned)((it & 1) ^ 1));  // the reducer is done with piece it-1 (passes at once for it = 0)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          sts_v2f64(scr + (cola[mt] * 8 + 2 * fk) * 8, q[mt][0][0] + q[mt][1][0], q[mt][0][1] + q[mt][1][1]);
          if (fk == 0) {
            sts_f64(scr + (128 + cola[mt]) * 8, t1[mt]);
            sts_f64(scr + (144 + cola[mt]) * 8, t2[mt]);
          }
        }
        __syncwarp();
        if (lane == 0) mb_arrive(bar_scr_full);
        stamp(it, 2);
      }
    }
  } else if (warp < NCW) {
    // ================= B-warps: copy piece n into tensor memory, then PX / h0 of piece n - LAG =================
    const int bw = warp - NAW;
    const int rowbase = bw * RW;
    // tensor memory: a warp may touch lanes [32 (warp % 4), +32); the two B-warps of a quarter split its 512 columns
    const unsigned tm_warp = tmem_base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)((bw >> 2) * (NT * TCOLS));
    unsigned offb[2];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) offb[s2] = (unsigned)((((4 * s2 + fk) ^ rm) & 7) << 4) + (unsigned)(rowbase + rm) * 128u;
    double px[MT][2], hh[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) px[i][0] = px[i][1] = hh[i] = 0.0;

    auto stash = [&](int n) {
      const int s = n % NSLOT, tb = n % NT;
      const unsigned tokf = mb_wait_token(bar_full(s), (unsigned)((n / NSLOT) & 1));
      const unsigned slot = slots0 + (unsigned)s * SLOT_BYTES + tokf;
      const unsigned tm = tm_warp + (unsigned)(tb * TCOLS);
#pragma unroll
      for (int m0 = 0; m0 < MT; m0 += 4) {
        double2 xb[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
            if (m0 + i < MT) xb[i][s2] = ldp_v2f64(slot + offb[s2] + (unsigned)(m0 + i) * 1024u);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (m0 + i < MT) tmem_st8(tm + (unsigned)((m0 + i) * 8), xb[i][0], xb[i][1]);
      }
      __syncwarp();
      if (lane == 0) mb_arrive(bar_empty(s));   // tmem_st8 consumed the loaded registers: the slot is free
    };
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
      tmem_wait_st();   // this warp's own stash stores (issued >= LAG pieces ago) are complete
      const unsigned tm = tm_warp + (unsigned)(tb * TCOLS);
#pragma unroll
      for (int m0 = 0; m0 < MT; m0 += 2) {
        unsigned r[2][8];
#pragma unroll
        for (int i = 0; i < 2; ++i) tmem_ld8(tm + (unsigned)((m0 + i) * 8), r[i]);
        tmem_wait_ld();
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const double x = __hiloint2double(r[i][4 * s2 + 2 * e + 1], r[i][4 * s2 + 2 * e]);
#ifndef ABL_NODMMA_B
              f_dmma(px[m0 + i][0], px[m0 + i][1], x, vf[s2][e]);
#else
              px[m0 + i][0] += x;
#endif
#ifndef ABL_NOHH
              hh[m0 + i] = fma(x, isd[s2][e], hh[m0 + i]);
#endif
            }
      }
      stamp(n, 5);
    };

    int it = 0;
    for (int ch = cl; ch < nchunks; ch += ncl) {
      const int b = ch / g.cpp, cidx = ch % g.cpp;
      const i64 grow0 = (i64)rank * R + rowbase;
      const int it_first = it;
      for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) {
        stash(it);
        if (it - LAG >= it_first) phase_b(it - LAG);
      }
      for (int n = (it - LAG > it_first ? it - LAG : it_first); n < it; ++n) phase_b(n);   // drain
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
  } else if (warp == W_TMA) {
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
  } else if ((warp == W_FIN || warp == W_FIN2) && (warp == W_FIN || owned > 4)) {
    // ================= finaliser: owner of columns {oc*G + rank} =================
    // Clusters of 1 or 2 CTAs own 16 / 8 columns per CTA = 4 / 2 serial rounds per piece, and this latency-bound
    // warp then paces the whole sweep: a second finaliser warp takes every other round.  Larger clusters (one round
    // per piece) run the first one alone; the second stays idle.
    const int fw = warp == W_FIN2 ? 1 : 0;
    const int nfw = owned > 4 ? 2 : 1;
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
        if (lane == 0 && fw == 0) {
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
        for (int oc0 = fw * cpr; oc0 < owned; oc0 += cpr * nfw) {
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
  } else if (warp == W_RED) {
    // ================= reducer: warp partials in `scratch` -> CTA partial -> owner CTA =================
    // lane handles values v = lane + 32 j (j = 0..4) of the 160 per warp: v < 128 is q[column v >> 3][l = v & 7],
    // 128 + c is s1 of column c, 144 + c is s2; the 8 A-warps are added in a fixed tree
    unsigned rdst[5], rbar[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int v = lane + 32 * j;
      const int c = v < 128 ? (v >> 3) : ((v - 128) & 15);
      const int st = v < 128 ? 2 + (v & 7) : (v < 144 ? 0 : 1);
      const int owner = c % G, oc = c / G;
      rdst[j] = map_rank(sa(statbuf + (oc * G + rank) * NSTAT + st), owner);
      rbar[j] = map_rank(bar_stat(0), owner);
    }
    const unsigned vbase = sa(scratch + lane);
    int it = 0;
    for (int ch = cl; ch < nchunks; ch += ncl) {
      const int cidx = ch % g.cpp;
      for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) {
        const int s = it % NXR;
        stamp(it, 0);
        mb_wait(bar_scr_full, (unsigned)(it & 1));
        stamp(it, 1);
        double acc[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          double p[NAW];
#pragma unroll
          for (int w = 0; w < NAW; ++w) p[w] = lds_f64(vbase + (unsigned)(w * SCR_W + 32 * j) * 8);
          acc[j] = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
        }
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

typedef void (*SymKernel)(const CUtensorMap, const CUtensorMap, const Args);
inline SymKernel kernel_for(int R) {
  switch (R / 128) {
    case 1: return k_tprf_sym<1>;
    case 2: return k_tprf_sym<2>;
    case 3: return k_tprf_sym<3>;
    default: return k_tprf_sym<4>;
  }
}

}  // namespace sym
}  // namespace um
