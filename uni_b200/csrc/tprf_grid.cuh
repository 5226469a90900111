// tprf_grid.cuh -- the single-read 3PRF sweep for tall panels, on ALL SMs (included by tprf.cu after
// tprf_fused.cuh, whose PTX wrappers and fragment maps it reuses).
//
// tprf_fused.cuh keeps a column tile on chip across a hardware thread-block cluster, but only 7 clusters of
// 16 CTAs (112 of 148 SMs) are co-resident on a B200, and its three shared-memory slots leave the B-warps
// one piece of slack against a ~3000-cycle exchange.  This kernel keeps the geometry (CTA `rank` of a GROUP
// of G CTAs owns rows [rank*R, rank*R+R) of every 16-column tile the group sweeps) and changes two things:
//
//   * the group is VIRTUAL: any G CTAs of a cooperative launch (one CTA per SM, floor(SMs / G) groups: 144 SMs
//     at G = 16).  Partials and records travel through small rings in global memory (L2) with a
//     flag-in-data protocol: every double is written as one 16-byte store {lo, tag, hi, tag} and read until
//     both tags match the expected use count, so a message costs one store and one load -- no fence, no
//     separate flag, no release/acquire round trips (under a 6 TB/s stream each of those costs thousands
//     of cycles: a fence + counter version of this exchange took ~20k cycles per piece).
//   * the piece does not wait in shared memory for phase B.  While an A-warp runs phase A it also copies its
//     128 rows, already in the phase-B fragment layout, into TENSOR MEMORY (tcgen05.st; 256 KB per SM = four
//     64 KiB pieces, FP64 stored as two 32-bit columns).  The shared-memory slot is released as soon as phase
//     A is done, so the three slots are pure TMA landing buffers, and B-warp 4+a reads "its" rows back with
//     tcgen05.ld up to four pieces later -- enough slack to hide the two L2 round trips of the exchange.
//     A-warp a and B-warp 4+a sit on the same SM sub-partition, which is also the TMEM lane quarter a warp
//     may touch, so the stash is private to the pair and handed over with two mbarriers.
//
// Everything else (DMMA fragments, swizzled TMA boxes, fixed summation trees, the finaliser's algebra, the
// outputs) is that of tprf_fused.cuh; results agree with it and with the two-sweep path to ~1e-12.
#pragma once

namespace um {
namespace gridk {

using namespace fused;

constexpr int NX = 8;                        // entries of the global exchange rings
constexpr int PSTRIDE = 12;                  // doubles per (rank, column) partial: s1, s2, q[8], shift, pad
constexpr int PART_ENTRY = GMAX * CW * PSTRIDE;   // 16-byte cells per ring entry of the partial buffer (G <= GMAX ranks)
constexpr int REC_ENTRY = CW * NREC;              // 16-byte cells per ring entry of the record buffer
constexpr int G_NTHREADS = (2 * NAW + 3) * 32;

constexpr int GOFF_SLOTS = 0;
constexpr int GOFF_SCRATCH = NSLOT * SLOT_BYTES;               // double[NAW][SCR_W + 16]  (+16: the tile's shift row)
constexpr int GSCR_W = SCR_W + 16;
constexpr int GOFF_KSH = GOFF_SCRATCH + NAW * GSCR_W * 8;      // double[NSLOT][CW]
constexpr int GOFF_BAR = GOFF_KSH + NSLOT * CW * 8;            // full[3] empty[3] scr_full scr_free stash_full[4][NT] stash_free[4][NT]
constexpr int GOFF_TMEM = GOFF_BAR + (8 + 2 * NAW * NT) * 8;   // u32: TMEM base address
constexpr int GSMEM_BYTES = GOFF_TMEM + 16;
constexpr int GSMEM_ALLOC = GSMEM_BYTES + 1024;

struct GArgs {
  Args a;                       // same meaning as in tprf_fused.cuh (a.G = group size)
  uint4* part;                  // [groups][NX][GMAX][CW][PSTRIDE] cells {lo, tag, hi, tag}; zeroed before every launch
  uint4* rec;                   // [groups][NX][CW][NREC]
};

// flag-in-data cell: {lo, tag, hi, tag}.  The 16-byte store may be performed as two 8-byte halves; each
// half carries its own tag, so a reader that sees both tags has both halves (the LL protocol of NCCL).
__device__ __forceinline__ void ll_store(uint4* cell, double v, unsigned tag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"(__double2loint(v)), "r"(tag),
               "r"(__double2hiint(v)), "r"(tag)
               : "memory");
}
__device__ __forceinline__ bool ll_try(const uint4* cell, unsigned tag, double& v) {
  unsigned a, f1, b, f2;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(f1), "=r"(b), "=r"(f2) : "l"(cell) : "memory");
  v = __hiloint2double(b, a);
  return f1 == tag && f2 == tag;
}
// N cells at once: all loads are in flight together (one L2 round trip per attempt, not one per cell)
template <int N>
__device__ __forceinline__ void ll_load_n(const uint4* const (&cell)[N], unsigned tag, double (&v)[N]) {
  for (;;) {
    unsigned a[N], f1[N], b[N], f2[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
      asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a[i]), "=r"(f1[i]), "=r"(b[i]), "=r"(f2[i]) : "l"(cell[i]) : "memory");
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      ok = ok && f1[i] == tag && f2[i] == tag;
      v[i] = __hiloint2double(b[i], a[i]);
    }
    if (ok) break;
    __nanosleep(200);
  }
}
template <int J>
__global__ void __launch_bounds__(G_NTHREADS, 1)
k_tprf_grid(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_row0, const GArgs ga) {
  constexpr int R = 128 * J;
  constexpr int RW = R / NAW;      // rows per A-/B-warp: 32, 64, 96 or 128
  constexpr int KS = RW / 4;       // phase-A k-steps (4 rows each)
  constexpr int MT = RW / 8;       // m-tiles (8 rows each)
  constexpr int TCOLS = MT * 8;    // TMEM columns of one stash buffer of one warp pair (<= 128)
  const Args& g = ga.a;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  double* scratch = reinterpret_cast<double*>(smem + GOFF_SCRATCH);
  double* kshs = reinterpret_cast<double*>(smem + GOFF_KSH);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + GOFF_TMEM);
  const unsigned bar0 = sa(smem + GOFF_BAR);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };          // TMA landed the piece in slot s
  auto bar_empty = [&](int s) { return bar0 + 8u * (3 + s); };   // the A-warps are done with slot s
  const unsigned bar_scr_full = bar0 + 8u * 6, bar_scr_free = bar0 + 8u * 7;
  auto bar_st_full = [&](int q, int t) { return bar0 + 8u * (8 + q * NT + t); };             // A-warp q stashed a piece in buffer t
  auto bar_st_free = [&](int q, int t) { return bar0 + 8u * (8 + NAW * NT + q * NT + t); };  // B-warp 4+q is done with it

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = g.G;
  const int ngroups = gridDim.x / G;
  const int cl = blockIdx.x / G, rank = blockIdx.x % G;
  const bool active = cl < ngroups;       // CTAs past the last whole group have nothing to do
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
      mb_init(bar_empty(s), NAW);
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

  const bool dbg_on = g.dbg != nullptr && cl == 0 && lane == 0;
  auto stamp = [&](int iter, int slot_) {
    if (dbg_on && iter < 64) g.dbg[(((i64)rank * 16 + warp) * 64 + iter) * 8 + slot_] = clock64();
  };
  const unsigned slots0 = sa(smem + GOFF_SLOTS);
  const int fm = lane >> 2, fk = lane & 3;
  // phase-B fragment geometry, shared by the A-warp (stash) and the B-warp (use)
  const int rm = (fm >> 1) + 4 * (fm & 1);
  uint4* part_g = ga.part + (size_t)cl * NX * PART_ENTRY;
  uint4* rec_g = ga.rec + (size_t)cl * NX * REC_ENTRY;

  if (!active) {
    // fall through to the common exit
  } else if (warp < NAW) {
    // ================= A-warps: column statistics of piece `it` + stash for the B-warp =================
    const int rowbase = warp * RW;
    // one 16-byte load feeds both m-tiles of phase A (even / odd columns), as in tprf_fused.cuh
    const int chunk = ((fm & 1) << 2) | (fm >> 1);
    const int cola[2] = {2 * chunk, 2 * chunk + 1};
    unsigned offv[2];
#pragma unroll
    for (int pb = 0; pb < 2; ++pb) offv[pb] = (unsigned)(((chunk ^ (4 * pb + fk)) & 7) << 4) + (unsigned)(rowbase + 4 * pb + fk) * 128u;
    unsigned offb[2];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) offb[s2] = (unsigned)((((4 * s2 + fk) ^ rm) & 7) << 4) + (unsigned)(rowbase + rm) * 128u;
    const unsigned tm_warp = tmem_base + ((unsigned)(32 * warp) << 16);
    int it = 0;
    for (int ch = cl; ch < nchunks; ch += ngroups) {
      const int b = ch / g.cpp, cidx = ch % g.cpp;
      double zf[KS];
      const i64 grow0 = (i64)rank * R + rowbase;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const i64 t = grow0 + 4 * ks + fk;
        zf[ks] = t < T ? g.zc[((i64)b * T + t) * 8 + fm] : 0.0;
      }
      const i64 tv = T - grow0 - fk;
      const int tvalid = tv > RW ? RW : (tv < 0 ? 0 : (int)tv);
      const bool need_mask = grow0 + RW > T;
      for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) {
        const int s = it % NSLOT, tb = it % NT;
        stamp(it, 0);
        const unsigned tokf = mb_wait_token(bar_full(s), (unsigned)((it / NSLOT) & 1));
        stamp(it, 1);
        const unsigned slot = slots0 + (unsigned)s * SLOT_BYTES + tokf;
        double q[2][2][2], s1[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, s2[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int pb = 0; pb < 2; ++pb) q[mt][pb][0] = q[mt][pb][1] = 0.0;
        const double2 kshv = ldp_v2f64(sa(kshs + s * CW + cola[0]) + tokf);
        const double ksh[2] = {kshv.x, kshv.y};
        if (it >= NT) {   // the B-warp has finished with this stash buffer
          mb_wait(bar_st_free(warp, tb), (unsigned)((it / NT - 1) & 1));
          tmem_fence_after();
        }
        stamp(it, 2);
        const unsigned tm = tm_warp + (unsigned)(tb * TCOLS);
        // groups of two k-steps (8 rows): four DMMAs on four independent accumulators, fed with the shifted value
        // d = x - k (see tprf_fused.cuh), + the same 8 rows once more in the phase-B layout -> tensor memory (one
        // m-tile); loads run AHEAD groups ahead, the fence keeps the accumulator chains interleaved
        auto phase_a = [&](auto masked) {
          constexpr bool MASK = decltype(masked)::value;
          constexpr int NG = KS / 2;
          constexpr int AHEAD = NG >= 2 ? 2 : NG;
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
                f_dmma(q[mt][pb][0], q[mt][pb][1], d[pb][mt], zf[2 * gk + pb]);
                s1[mt][pb] += d[pb][mt];
                s2[mt][pb] = fma(d[pb][mt], d[pb][mt], s2[mt][pb]);
              }
            tmem_st8(tm + (unsigned)(gk * 8), xb[0][0], xb[0][1]);
            sched_fence();
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
        tmem_wait_st();
        tmem_fence_before();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) consume(q[mt][0][0]);   // every load of the slot has landed in a register
        __syncwarp();
        if (lane == 0) {
          mb_arrive(bar_st_full(warp, tb));
          mb_arrive(bar_empty(s));
        }
        // per-lane partials -> scratch (the reducer adds the 4 row phases and the 4 A-warps)
        mb_wait(bar_scr_free, (unsigned)((it & 1) ^ 1));
        {
          const unsigned sp = sa(scratch + warp * GSCR_W);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            sts_v2f64(sp + (cola[mt] * 8 + 2 * fk) * 8, q[mt][0][0] + q[mt][1][0], q[mt][0][1] + q[mt][1][1]);
            sts_f64(sp + (128 + cola[mt] * 4 + fk) * 8, s1[mt][0] + s1[mt][1]);
            sts_f64(sp + (192 + cola[mt] * 4 + fk) * 8, s2[mt][0] + s2[mt][1]);
            if (warp == 0 && fk == 0) sts_f64(sp + (SCR_W + cola[mt]) * 8, ksh[mt]);
          }
        }
        __syncwarp();
        if (lane == 0) mb_arrive(bar_scr_full);
        stamp(it, 3);
      }
    }
  } else if (warp < 2 * NAW) {
    // ================= B-warps: PX / h0 of piece `it` from the stash, once every column's record is out =================
    const int bw = warp - NAW;
    const int rowbase = bw * RW;
    const unsigned tm_warp = tmem_base + ((unsigned)(32 * bw) << 16);
    double px[MT][2], hh[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) px[i][0] = px[i][1] = hh[i] = 0.0;
    int it = 0;
    for (int ch = cl; ch < nchunks; ch += ngroups) {
      const int b = ch / g.cpp, cidx = ch % g.cpp;
      const i64 grow0 = (i64)rank * R + rowbase;
      for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) {
        const int tb = it % NT, e = it % NX;
        const unsigned tag = (unsigned)(it / NX + 1);
        stamp(it, 0);
        const uint4* rr = rec_g + (size_t)e * REC_ENTRY;
        double vf[2][2], isd[2][2];
        {
          const uint4* cells[8];
          double vals[8];
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int ee = 0; ee < 2; ++ee) {
              const int c = 8 * s2 + 2 * fk + ee;
              cells[(s2 * 2 + ee) * 2] = rr + c * NREC + fm;
              cells[(s2 * 2 + ee) * 2 + 1] = rr + c * NREC + 8;
            }
          ll_load_n<8>(cells, tag, vals);
          __syncwarp();
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int ee = 0; ee < 2; ++ee) {
              vf[s2][ee] = vals[(s2 * 2 + ee) * 2];
              isd[s2][ee] = vals[(s2 * 2 + ee) * 2 + 1];
            }
        }
        stamp(it, 1);
        mb_wait(bar_st_full(bw, tb), (unsigned)((it / NT) & 1));
        tmem_fence_after();
        stamp(it, 2);
        const unsigned tm = tm_warp + (unsigned)(tb * TCOLS);
        unsigned tr2[2][2][8];   // next pair's tensor-memory loads in flight during this pair's math (as tprf_fused.cuh)
#pragma unroll
        for (int i = 0; i < 2; ++i) tmem_ld8(tm + (unsigned)(i * 8), tr2[0][i]);
#pragma unroll
        for (int m0 = 0; m0 < MT; m0 += 2) {
          tmem_wait_ld();
          unsigned (&r)[2][8] = tr2[(m0 >> 1) & 1];
          if (m0 + 2 < MT) {
#pragma unroll
            for (int i = 0; i < 2; ++i) tmem_ld8(tm + (unsigned)((m0 + 2 + i) * 8), tr2[((m0 >> 1) & 1) ^ 1][i]);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
              const double x0 = __hiloint2double(r[i][4 * s2 + 1], r[i][4 * s2]);
              const double x1 = __hiloint2double(r[i][4 * s2 + 3], r[i][4 * s2 + 2]);
              f_dmma(px[m0 + i][0], px[m0 + i][1], x0, vf[s2][0]);
              f_dmma(px[m0 + i][0], px[m0 + i][1], x1, vf[s2][1]);
              hh[m0 + i] = fma(x0, isd[s2][0], hh[m0 + i]);
              hh[m0 + i] = fma(x1, isd[s2][1], hh[m0 + i]);
            }
        }
        tmem_fence_before();
        __syncwarp();
        if (lane == 0) mb_arrive(bar_st_free(bw, tb));
        stamp(it, 3);
      }
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
      for (int ch = cl; ch < nchunks; ch += ngroups) {
        const int b = ch / g.cpp, cidx = ch % g.cpp;
        for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) {
          const int s = it % NSLOT;
          if (it >= NSLOT) mb_wait(bar_empty(s), (unsigned)((it / NSLOT - 1) & 1));
          mb_expect_tx(bar_full(s), piece_bytes);
          const unsigned dst = slots0 + (unsigned)s * SLOT_BYTES;
#pragma unroll
          for (int bx = 0; bx < J; ++bx)
            tma_3d(dst + bx * BOX_ROWS * CW * 8, &map_x, (int)(tile * CW), rank * R + bx * BOX_ROWS, b, bar_full(s));
          tma_3d(sa(kshs + s * CW), &map_row0, (int)(tile * CW), 0, b, bar_full(s));
        }
      }
    }
  } else if (warp == 2 * NAW + 1) {
    // ================= finaliser: owner of columns {oc*G + rank} =================
    const int lg = lane >> 3, l = lane & 7;
    const int cpr = owned < 4 ? owned : 4;
    const int sub = 4 / cpr;
    const int cs = lg / sub, rs = lg % sub;
    const int nr = G / sub;               // ranks this lane sums: min(G, 4)
    double a_sw = 0.0, a_smw = 0.0, a_smu = 0.0, a_sww[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a_sww[c] = 0.0;
    const double invT = 1.0 / (double)T, invTm1 = T > 1 ? 1.0 / (double)(T - 1) : 0.0;
    int it = 0;
    for (int ch = cl; ch < nchunks; ch += ngroups) {
      const int b = ch / g.cpp, cidx = ch % g.cpp;
      for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) {
        const int e = it % NX;
        const unsigned tag = (unsigned)(it / NX + 1);
        stamp(it, 0);
        const uint4* pb = part_g + (size_t)e * PART_ENTRY;
        uint4* rr = rec_g + (size_t)e * REC_ENTRY;
        for (int oc0 = 0; oc0 < owned; oc0 += cpr) {
          const int oc = oc0 + cs;
          const int col = oc * G + rank;
          double S, SS, q, k0;
          {
            double ps[4], pss[4], pq[4];
            {
              const uint4* cells[13];
              double vals[13];
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const uint4* pa = pb + ((size_t)(rs * nr + (r < nr ? r : 0)) * CW + col) * PSTRIDE;
                cells[3 * r] = pa;
                cells[3 * r + 1] = pa + 1;
                cells[3 * r + 2] = pa + 2 + l;
              }
              cells[12] = pb + (size_t)col * PSTRIDE + 10;   // rank 0 forwards the tile's shift row
              ll_load_n<13>(cells, tag, vals);
              __syncwarp();
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const bool on = r < nr;
                ps[r] = on ? vals[3 * r] : 0.0;
                pss[r] = on ? vals[3 * r + 1] : 0.0;
                pq[r] = on ? vals[3 * r + 2] : 0.0;
              }
              k0 = vals[12];
            }
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
          if (oc0 == 0) stamp(it, 1);
          const double dm = S * invT;
          const double mu = k0 + dm;
          double inv = 1.0;
          if (T > 1) {
            double var = fma(-S, dm, SS) * invTm1;
            if (var < 0.0) var = 0.0;
            if (var != 0.0) {                      // sd == 0 -> 1 (Tprf3.scala:493-501); NaN propagates
              double y;
              asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(var));
              const double hv = 0.5 * var;
              y = y * fma(-hv * y, y, 1.5);
              y = y * fma(-hv * y, y, 1.5);
              inv = (var < 1e-290 || var > 1e290) ? 1.0 / sqrt(var) : y;
            }
          }
          const double m_n = mu * inv;
          const double w = q * inv;
          const double v = w * inv;
          if (rs == 0) {
            ll_store(rr + col * NREC + l, v, tag);
            if (l == 0) ll_store(rr + col * NREC + 8, inv, tag);
          }
          const bool mine = valid && rs == 0;
          if (mine) {
            const i64 cix = (i64)b * g.npad + gcol;
            g.w0[cix * 8 + l] = w;
            if (l == 0) {
              g.invsd[cix] = inv;
              g.mun[cix] = m_n;
            }
          }
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
        stamp(it, 2);
      }
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
        if (lg == 0) {
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
    // ================= reducer: lane partials in `scratch` -> CTA partial -> ring entry in global memory =================
    const unsigned sbase = sa(scratch + 128 + (lane & 1) * 64 + (lane >> 1) * 4);
    const unsigned qbase = sa(scratch + lane);
    int it = 0;
    for (int ch = cl; ch < nchunks; ch += ngroups) {
      const int cidx = ch % g.cpp;
      for (i64 tile = g.tile_lo + cidx; tile < g.ntile; tile += g.cpp, ++it) {
        const int e = it % NX;
        mb_wait(bar_scr_full, (unsigned)(it & 1));
        stamp(it, 1);
        double p[NAW][4], pq[4][NAW];
#pragma unroll
        for (int w = 0; w < NAW; ++w)
#pragma unroll
          for (int k = 0; k < 4; ++k) p[w][k] = lds_f64(sbase + (w * GSCR_W + k) * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int w = 0; w < NAW; ++w) pq[j][w] = lds_f64(qbase + (w * GSCR_W + 32 * j) * 8);
        const double k0 = lane < CW ? lds_f64(sa(scratch + SCR_W + lane)) : 0.0;
        double acc[5];
        {
          double pw[NAW];
#pragma unroll
          for (int w = 0; w < NAW; ++w) pw[w] = (p[w][0] + p[w][1]) + (p[w][2] + p[w][3]);
          acc[0] = (pw[0] + pw[1]) + (pw[2] + pw[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[1 + j] = (pq[j][0] + pq[j][1]) + (pq[j][2] + pq[j][3]);
        consume(acc[0]); consume(acc[1]); consume(acc[2]); consume(acc[3]); consume(acc[4]); consume(k0);
        __syncwarp();
        if (lane == 0) mb_arrive(bar_scr_free);
        const unsigned tag = (unsigned)(it / NX + 1);
        uint4* pe = part_g + (size_t)e * PART_ENTRY + (size_t)rank * CW * PSTRIDE;
        ll_store(pe + (lane >> 1) * PSTRIDE + (lane & 1), acc[0], tag);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ee = lane + 32 * j;
          ll_store(pe + (ee >> 3) * PSTRIDE + 2 + (ee & 7), acc[1 + j], tag);
        }
        if (lane < CW && rank == 0) ll_store(pe + lane * PSTRIDE + 10, k0, tag);
        stamp(it, 2);
      }
    }
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

typedef void (*GridKernel)(const CUtensorMap, const CUtensorMap, const GArgs);
inline GridKernel grid_kernel_for(int R) {
  switch (R / 128) {
    case 1: return k_tprf_grid<1>;
    case 2: return k_tprf_grid<2>;
    case 3: return k_tprf_grid<3>;
    default: return k_tprf_grid<4>;
  }
}

}  // namespace gridk
}  // namespace um
