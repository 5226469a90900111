// rng.cu -- NumPy-compatible PCG64 generator on the device: MatD.setSeed / rand /
// randn / normal / uniform / randint (reference: src/main/scala/uni/data/NumPyRNG.scala:22-160 generator,
// :584-770 SeedSequence + PCG64 seeding; src/main/scala/uni/data/Mat.scala:1461-1530 the Mat factories).
//
// The reference draws one element after the other from a single 128-bit LCG.  Here every thread JUMPS to its own
// place in that stream (x -> x*A^k + C_k, the power-of-two maps are precomputed on the host and travel as kernel
// parameters) and then walks it with the block-stride map, so consecutive threads hold consecutive draws and the
// stores coalesce.  rand / uniform / randint consume a fixed number of draws per element and are one kernel.
//
// randn is NumPy's 256-level ziggurat: an ATTEMPT at stream position p consumes c(p) draws (1 for the 98.8 % that land
// inside a strip, 2 for a wedge test, 1 + 2k for the tail loop) and yields 0 or 1 values, so the element -> position
// map is a chain p -> p + c(p).  It is resolved exactly, in parallel:
//   k_rng_block_states  generator state at the first draw of every segment (S = 4096 draws);
//   k_zig_scan          per segment: classify every draw; the ~50 non-simple ones are listed and each is evaluated by
//                       its own thread (jump to p + 1 for the extra draws); each then decides whether the chain from
//                       the segment's first draw starts an attempt AT it (walk back to the nearest position no earlier
//                       attempt can reach over, replay the few attempts in between) and marks the draws it consumes;
//                       -> yield mask of the segment (stored), and for every entry offset e < 16 (the overshoot a
//                       previous segment can leave behind; its chain merges with the one from 0 within a few draws)
//                       the table e -> (exit offset, number of yields);
//   k_zig_pf_reduce / _scan / _apply   compose the tables in segment order (a three-level scan of 16-entry maps)
//                       -> true entry offset and first output index of every segment;
//   k_zig_emit_fast     segments entered at offset 0 (98.8 %): regenerate the draws, store value at base + rank
//                       (rank = popcounts of the stored mask); tail values are recomputed in place;
//   k_zig_emit          the other segments: full resolution again from the true entry.
// The stream is therefore consumed exactly as the sequential generator consumes it, and the generator's state after
// the call is the state NumPy (and NumPyRNG.scala) would have.  exp / log1p in the wedge and tail branches are CUDA's
// (<= 1 ulp); they decide accept / reject and the tail magnitude (|z| > 3.654, 2.7e-4 of the draws), every other
// draw is integer arithmetic and one exact multiply.  An attempt that would consume more than 32 draws, or an
// overshoot of 16 or more (a tail loop of 8+ rounds at a segment end, p < 1e-12 per 2^31 draws) is reported as an
// error, never emitted wrong.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "ziggurat_tables.cuh"

namespace um {
namespace {

typedef unsigned __int128 u128;
typedef unsigned long long u64;

constexpr int RNG_THREADS = 256;
constexpr int RNG_PER_THREAD = 16;
constexpr int SEG = RNG_THREADS * RNG_PER_THREAD;  // positions per block / per ziggurat segment
constexpr int SEG_WORDS = SEG / 64;
constexpr int ZIG_ENTRIES = 16;                    // entry offsets tabulated per segment
constexpr int ZIG_MAX_NS = 512;                    // non-simple positions a segment can list (mean 50, sd 7)
constexpr int ZIG_MAX_CONS = 32;                   // draws one attempt may consume (tail loop of 15 rounds: p ~ 1e-17)

inline u128 mk128(u64 hi, u64 lo) { return ((u128)hi << 64) | lo; }
const u128 PCG_MULT = mk128(0x2360ED051FC65DA4ULL, 0x4385DF649FCCF645ULL);  // NumPyRNG.scala:36-38

// x -> x * m[k] + p[k] advances the generator by 2^k draws
struct Jump {
  u64 mh[64], ml[64], ph[64], pl[64];
};

Jump make_jump(u128 inc) {
  Jump j;
  u128 m = PCG_MULT, p = inc;
  for (int k = 0; k < 64; k++) {
    j.mh[k] = (u64)(m >> 64); j.ml[k] = (u64)m; j.ph[k] = (u64)(p >> 64); j.pl[k] = (u64)p;
    p = (m + 1) * p;
    m = m * m;
  }
  return j;
}

u128 advance(u128 state, u128 inc, u64 delta) {
  u128 m = PCG_MULT, p = inc;
  while (delta) {
    if (delta & 1) state = state * m + p;
    p = (m + 1) * p;
    m = m * m;
    delta >>= 1;
  }
  return state;
}

inline u64 host_output(u128 s) {  // NumPyRNG.scala:53-56
  u64 hi = (u64)(s >> 64), lo = (u64)s, x = hi ^ lo;
  unsigned r = (unsigned)(hi >> 58);
  return r ? (x >> r) | (x << (64 - r)) : x;
}

__device__ __forceinline__ void lcg(u64& hi, u64& lo, u64 mh, u64 ml, u64 ph, u64 pl) {
  const u64 nlo = lo * ml;
  const u64 nhi = __umul64hi(lo, ml) + lo * mh + hi * ml;
  asm("add.cc.u64 %0, %2, %4;\n\taddc.u64 %1, %3, %5;" : "=l"(lo), "=l"(hi) : "l"(nlo), "l"(nhi), "l"(pl), "l"(ph));  // 128-bit add
}
// The block-stride map: A^256 mod 2^128 is a property of the multiplier alone, so it is a compile-time constant and the
// multiplies take immediates; only the additive part depends on the stream (checked against make_jump on the host).
constexpr u64 STRIDE256_MH = 0x61ecb5c24d95b058ULL, STRIDE256_ML = 0xf04c80a23697bc01ULL;
__device__ __forceinline__ void lcg_stride256(u64& hi, u64& lo, u64 ph, u64 pl) { lcg(hi, lo, STRIDE256_MH, STRIDE256_ML, ph, pl); }
// (A hand-laid 32-bit limb product -- six wide and four narrow multiply-adds -- was tried: the zero-extended addends cost
//  a register move each and the instruction count came out the same as the compiler's 64-bit expansion.)
__device__ __forceinline__ u64 pcg_out(u64 hi, u64 lo) {
  // rotr64(hi ^ lo, hi >> 58) with two funnel shifts: a rotation by >= 32 swaps the halves first
  const u64 x = hi ^ lo;
  const unsigned r = (unsigned)(hi >> 58);
  const unsigned xl = (unsigned)x, xh = (unsigned)(x >> 32);
  const bool sw = (r & 32u) != 0u;
  const unsigned a = sw ? xh : xl, b = sw ? xl : xh;
  const unsigned rl = __funnelshift_r(a, b, r), rh = __funnelshift_r(b, a, r);  // shift amount taken mod 32
  return ((u64)rh << 32) | rl;
}
__device__ __forceinline__ void jump_to(u64& hi, u64& lo, u64 delta, const Jump& J) {
  for (int k = 0; delta != 0; k++, delta >>= 1)
    if (delta & 1) lcg(hi, lo, J.mh[k], J.ml[k], J.ph[k], J.pl[k]);
}
__device__ __forceinline__ double u01(u64 raw) { return (double)(raw >> 11) * (1.0 / 9007199254740992.0); }  // NumPyRNG.scala:68-71

// State of the generator at the first position of every block (segment): one full jump per BS_CHUNK blocks, then
// block-stride steps.  The per-thread jump inside a block is then only log2(SEG) maps deep.
constexpr int BS_CHUNK = 16;
__global__ void __launch_bounds__(128) k_rng_block_states(const __grid_constant__ Jump J, u64 s_hi, u64 s_lo, i64 nblocks,
                                                         int log2_span, ulonglong2* __restrict__ bs) {
  i64 b0 = ((i64)blockIdx.x * 128 + threadIdx.x) * BS_CHUNK;
  if (b0 >= nblocks) return;
  u64 hi = s_hi, lo = s_lo;
  jump_to(hi, lo, (u64)b0 << log2_span, J);
  const u64 mh = J.mh[log2_span], ml = J.ml[log2_span], ph = J.ph[log2_span], pl = J.pl[log2_span];
  for (int i = 0; i < BS_CHUNK && b0 + i < nblocks; i++) {
    bs[b0 + i] = make_ulonglong2(hi, lo);
    lcg(hi, lo, mh, ml, ph, pl);  // 2^log2_span draws
  }
}
static_assert(SEG == 4096, "ziggurat segments are 2^12 draws (16-bit positions, block-state stride)");
// For the ziggurat kernels: the state at the first draw of every WARP of every segment (draw seg * 4096 + 32 w, w < 8),
// so a thread's own jump is at most 32 draws (6 maps instead of 9).
constexpr int ZIG_WARPS = RNG_THREADS / 32;
__global__ void __launch_bounds__(128) k_zig_warp_states(const __grid_constant__ Jump J, u64 s_hi, u64 s_lo, i64 nseg,
                                                        ulonglong2* __restrict__ ws) {
  i64 b0 = ((i64)blockIdx.x * 128 + threadIdx.x) * BS_CHUNK;
  if (b0 >= nseg) return;
  u64 hi = s_hi, lo = s_lo;
  jump_to(hi, lo, (u64)b0 << 12, J);
  for (int i = 0; i < BS_CHUNK && b0 + i < nseg; i++) {
    u64 h = hi, l = lo;
#pragma unroll
    for (int w = 0; w < ZIG_WARPS; w++) {
      ws[(b0 + i) * ZIG_WARPS + w] = make_ulonglong2(h, l);
      lcg(h, l, J.mh[5], J.ml[5], J.ph[5], J.pl[5]);
    }
    lcg(hi, lo, J.mh[12], J.ml[12], J.ph[12], J.pl[12]);
  }
}
// the fixed-consumption fills amortise the per-thread jump over more draws
constexpr int DRAW_PER_THREAD = 128;
constexpr int DRAW_LOG2_SPAN = 15;
constexpr int DRAW_SPAN = RNG_THREADS * DRAW_PER_THREAD;
static_assert(DRAW_SPAN == 1 << DRAW_LOG2_SPAN, "block-state stride");

// ---- fixed-consumption draws: rand, uniform, randint ----------------------------------------------------------
enum { DRAW_U01 = 0, DRAW_UNIFORM = 1, DRAW_INT = 2 };

template <int KIND>
__device__ __forceinline__ void draw_store(u64 raw, i64 p, double* __restrict__ outd, int* __restrict__ outi, i64 n_out, double low,
                                           double span, int ilow, unsigned ibound) {
  if (KIND == DRAW_U01) {
    outd[p] = u01(raw);
  } else if (KIND == DRAW_UNIFORM) {
    outd[p] = __dadd_rn(low, __dmul_rn(u01(raw), span));  // low + nextDouble() * (high - low), NumPyRNG.scala:73-74
  } else {
    // nextBoundedInt (NumPyRNG.scala:82-110): low half first, then the high half, (bits32 * bound) >>> 32
    const i64 o = 2 * p;
    outi[o] = ilow + (int)(((raw & 0xffffffffULL) * (u64)ibound) >> 32);
    if (o + 1 < n_out) outi[o + 1] = ilow + (int)(((raw >> 32) * (u64)ibound) >> 32);
  }
}

template <int KIND>
__global__ void __launch_bounds__(RNG_THREADS) k_rng_draw(const __grid_constant__ Jump J, const ulonglong2* __restrict__ bs, i64 n_raw,
                                                         double* __restrict__ outd, int* __restrict__ outi, i64 n_out,
                                                         double low, double span, int ilow, unsigned ibound, int pre_flag,
                                                         int pre_val) {
  if (KIND == DRAW_INT && pre_flag && blockIdx.x == 0 && threadIdx.x == 0) outi[-1] = pre_val;
  const i64 block0 = (i64)blockIdx.x * DRAW_SPAN;
  const i64 base = block0 + threadIdx.x;
  if (base >= n_raw) return;
  const ulonglong2 b = bs[blockIdx.x];
  u64 hi = b.x, lo = b.y;
  jump_to(hi, lo, (u64)threadIdx.x + 1, J);
  const u64 ph = J.ph[8], pl = J.pl[8];  // stride 256 = RNG_THREADS
  if (block0 + DRAW_SPAN <= n_raw) {
    // the whole block is inside the fill (all but the last block): no per-draw bounds arithmetic
#pragma unroll 8
    for (int j = 0; j < DRAW_PER_THREAD; j++) {
      draw_store<KIND>(pcg_out(hi, lo), base + (i64)j * RNG_THREADS, outd, outi, n_out, low, span, ilow, ibound);
      lcg_stride256(hi, lo, ph, pl);
    }
    return;
  }
#pragma unroll 4
  for (int j = 0; j < DRAW_PER_THREAD; j++) {
    const i64 p = base + (i64)j * RNG_THREADS;
    if (p >= n_raw) break;
    draw_store<KIND>(pcg_out(hi, lo), p, outd, outi, n_out, low, span, ilow, ibound);
    lcg_stride256(hi, lo, ph, pl);
  }
}

// ---- ziggurat ------------------------------------------------------------------------------------------------
__device__ u64 d_zig_ki[256];
__device__ double d_zig_wi[256];
__device__ double d_zig_fi[256];

struct ZigShared {
  u64 ki[256];
  double wi[256];
  u64 ns[SEG_WORDS];        // non-simple positions (attempt needs more than its own draw)
  u64 yes[SEG_WORDS];       // non-simple positions whose attempt yields a value
  u64 covered[SEG_WORDS];   // positions consumed by an earlier attempt (emit only)
  unsigned short cons[ZIG_MAX_NS];   // draws consumed, by list slot
  unsigned short lpos[ZIG_MAX_NS];   // position, by list slot
  double tailv[ZIG_MAX_NS];          // value of a tail attempt, by list slot (emit only)
  unsigned short slot_of[SEG];       // list slot of a non-simple position (written for those only)
  u64 ym[SEG_WORDS];        // positions whose attempt yields a value (chain from entry 0, then the true entry)
  int nlist;
  int overflow;
  int reach;                // max over the chain's non-simple attempts of (position + draws consumed)
  int half;
  int prefix[SEG_WORDS + 1];
};

// one draw -> (idx, sign, rabs), NumPyRNG.scala:129-135
__device__ __forceinline__ void zig_split(u64 raw, int& idx, int& sign, u64& rabs) {
  idx = (int)(raw & 0xff);
  u64 r = raw >> 8;
  sign = (int)(r & 1);
  rabs = (r >> 1) & 0x000fffffffffffffULL;
}
__device__ __forceinline__ double zig_x(u64 rabs, double w, int sign) {
  // rabs < 2^52: exact conversion through the 2^52 exponent trick, then ONE rounded multiply (rabs.toDouble * wi)
  double m = __longlong_as_double((long long)(rabs | 0x4330000000000000ULL)) - 4503599627370496.0;
  double x = __dmul_rn(m, w);  // >= 0: the sign is one bit to set (-0.0 for x == 0, like `sign * x`)
  return __longlong_as_double(__double_as_longlong(x) | ((long long)sign << 63));
}

// Phase 1 (both kernels): generate this thread's draws, flag the non-simple ones, list them.
// Returns the draws through raws[].
__device__ __forceinline__ void zig_generate(ZigShared& sh, const Jump& J, const ulonglong2* __restrict__ wsb,
                                             u64 (&raws)[RNG_PER_THREAD]) {
  const int t = threadIdx.x;
  const ulonglong2 b = wsb[t >> 5];
  u64 hi = b.x, lo = b.y;
  jump_to(hi, lo, (u64)(t & 31) + 1, J);
  const u64 ph = J.ph[8], pl = J.pl[8];
#pragma unroll
  for (int j = 0; j < RNG_PER_THREAD; j++) {
    u64 raw = pcg_out(hi, lo);
    raws[j] = raw;
    int idx, sign;
    u64 rabs;
    zig_split(raw, idx, sign, rabs);
    bool nonsimple = !(rabs < sh.ki[idx]);
    unsigned b = __ballot_sync(0xffffffffu, nonsimple);
    if ((t & 31) == 0) reinterpret_cast<unsigned*>(sh.ns)[(j * RNG_THREADS + t) >> 5] = b;
    if (nonsimple) {
      int slot = atomicAdd(&sh.nlist, 1);
      if (slot < ZIG_MAX_NS) {
        sh.lpos[slot] = (unsigned short)(j * RNG_THREADS + t);
        sh.slot_of[j * RNG_THREADS + t] = (unsigned short)slot;
      } else {
        sh.overflow = 1;
      }
    }
    lcg_stride256(hi, lo, ph, pl);
  }
}

// Phase 2: one thread per listed position evaluates its attempt as if it started there (NumPyRNG.scala:139-157).
__device__ __forceinline__ void zig_evaluate(ZigShared& sh, const Jump& J, ulonglong2 b, bool want_values) {
  int n = min(sh.nlist, ZIG_MAX_NS);
  for (int slot = threadIdx.x; slot < n; slot += RNG_THREADS) {
    int p = sh.lpos[slot];
    u64 hi = b.x, lo = b.y;
    jump_to(hi, lo, (u64)p + 1, J);  // state whose output is draw p
    int idx, sign;
    u64 rabs;
    zig_split(pcg_out(hi, lo), idx, sign, rabs);
    int consumed = 1;
    bool yields;
    double val = 0.0;
    if (idx == 0) {
      double xx, yy;
      do {
        lcg(hi, lo, J.mh[0], J.ml[0], J.ph[0], J.pl[0]);
        xx = -UM_ZIG_INV_R * log1p(-u01(pcg_out(hi, lo)));
        lcg(hi, lo, J.mh[0], J.ml[0], J.ph[0], J.pl[0]);
        yy = -log1p(-u01(pcg_out(hi, lo)));
        consumed += 2;
      } while (!(__dadd_rn(yy, yy) > __dmul_rn(xx, xx)) && consumed < ZIG_MAX_CONS);
      if (!(__dadd_rn(yy, yy) > __dmul_rn(xx, xx))) sh.overflow = 1;  // reported by the host, never silently wrong
      yields = true;
      val = ((rabs >> 8) & 1) ? -__dadd_rn(UM_ZIG_R, xx) : __dadd_rn(UM_ZIG_R, xx);
    } else {
      lcg(hi, lo, J.mh[0], J.ml[0], J.ph[0], J.pl[0]);
      consumed = 2;
      double x = zig_x(rabs, sh.wi[idx], sign);
      double f1 = d_zig_fi[idx], f0 = d_zig_fi[idx - 1];
      double lhs = __dadd_rn(__dmul_rn(__dadd_rn(f0, -f1), u01(pcg_out(hi, lo))), f1);
      yields = lhs < exp(__dmul_rn(__dmul_rn(-0.5, x), x));
    }
    sh.cons[slot] = (unsigned short)consumed;
    if (want_values) sh.tailv[slot] = val;
    if (yields) atomicOr(&sh.yes[p >> 6], 1ULL << (p & 63));
  }
}

// first non-simple position >= p (SEG if none)
__device__ __forceinline__ int zig_next_ns(const ZigShared& sh, int p) {
  if (p >= SEG) return SEG;
  int w = p >> 6;
  u64 m = sh.ns[w] & (~0ULL << (p & 63));
  while (m == 0) {
    if (++w >= SEG_WORDS) return SEG;
    m = sh.ns[w];
  }
  return (w << 6) + __ffsll((long long)m) - 1;
}
__device__ __forceinline__ int zig_cons(const ZigShared& sh, int q) { return sh.cons[sh.slot_of[q]]; }
__device__ __forceinline__ bool zig_bit(const u64* m, int p) { return (m[p >> 6] >> (p & 63)) & 1; }

// non-simple flags of the 32 positions before q: bit i <-> position q - 32 + i (positions < 0 read as 0)
__device__ __forceinline__ unsigned zig_window(const ZigShared& sh, int q) {
  const int w = q >> 6, b = q & 63;
  const u64 cur = sh.ns[w], prev = w > 0 ? sh.ns[w - 1] : 0ULL;
  const u64 before = b == 0 ? prev : (prev >> b) | (cur << (64 - b));  // bit i <-> position q - 64 + i
  return (unsigned)(before >> 32);
}
// HEAD: no earlier non-simple position could reach past q even if its attempt took place (draws consumed <= 32, so
// 32 positions back is far enough) -> q starts an attempt whatever happened before.
__device__ __forceinline__ bool zig_is_head(const ZigShared& sh, int q, int& prev_ns) {
  unsigned win = zig_window(sh, q);
  prev_ns = -1;
  while (win) {
    const int i = 31 - __clz(win);
    const int r = q - 32 + i;
    if (prev_ns < 0) prev_ns = r;
    if (r + zig_cons(sh, r) > q) return false;
    win &= ~(1u << i);
  }
  return true;
}
// Does the chain of attempts that starts at position 0 start an attempt at the non-simple position q?
// Walk back to the nearest head (a certain start), then replay the few attempts between it and q.
__device__ __forceinline__ bool zig_starts_at(const ZigShared& sh, int q) {
  int h = q, prev;
  for (int guard = 0; guard < SEG; guard++) {
    if (zig_is_head(sh, h, prev)) break;
    h = prev;  // not a head: some non-simple position within 32 before h exists
  }
  int p = h;
  while (p < q) p = zig_next_ns(sh, p + zig_cons(sh, p));
  return p == q;
}

// W1 + W2: resolve the chain from entry 0 in parallel (one thread per non-simple position), leaving covered[], ym[],
// reach.  Ends with a barrier.
__device__ __forceinline__ void zig_chain0(ZigShared& sh) {
  const int n = min(sh.nlist, ZIG_MAX_NS);
  if (!sh.overflow) {
    for (int slot = threadIdx.x; slot < n; slot += RNG_THREADS) {
      const int q = sh.lpos[slot];
      if (!zig_starts_at(sh, q)) continue;
      const int c = sh.cons[slot];
      for (int k = q + 1; k < q + c && k < SEG; k++) atomicOr(&sh.covered[k >> 6], 1ULL << (k & 63));
      atomicMax(&sh.reach, q + c);
    }
  }
  __syncthreads();
  if (threadIdx.x < SEG_WORDS) sh.ym[threadIdx.x] = ~sh.covered[threadIdx.x] & (~sh.ns[threadIdx.x] | sh.yes[threadIdx.x]);
  __syncthreads();
}

// exclusive prefix popcounts of ym[] (prefix[SEG_WORDS] = total).  Called by every thread; ends with a barrier.
__device__ __forceinline__ void zig_prefix(ZigShared& sh) {
  static_assert(SEG_WORDS == 64, "two warps scan the 64 mask words");
  const int t = threadIdx.x, lane = t & 31;
  int c = 0, incl = 0;
  if (t < SEG_WORDS) {
    c = __popcll(sh.ym[t]);
    incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (t == 31) sh.half = incl;
  }
  __syncthreads();
  if (t < SEG_WORDS) {
    const int excl = incl - c + (t >= 32 ? sh.half : 0);
    sh.prefix[t] = excl;
    if (t == SEG_WORDS - 1) sh.prefix[SEG_WORDS] = excl + c;
  }
  __syncthreads();
}

// The chain that ENTERS at position e (0 < e): walk it until it meets the chain from 0 (the first position it visits
// that the 0-chain also starts at), which happens within a few positions.  Returns yields and exit offset for entry e.
// FIX: rewrite ym[] for the positions before the meeting point (emit); prefix[] must be rebuilt afterwards.
template <bool FIX>
__device__ __forceinline__ int zig_enter(ZigShared& sh, int e, int& exit_off) {
  const int exit0 = max(sh.reach, SEG) - SEG;
  int p = e, cnt = 0;
  if (FIX)
    for (int k = 0; k < e; k++) sh.ym[k >> 6] &= ~(1ULL << (k & 63));
  while (p < SEG && zig_bit(sh.covered, p)) {  // consumed by the 0-chain, but an attempt of this one
    int step = 1, y = 1;
    if (zig_bit(sh.ns, p)) { step = zig_cons(sh, p); y = (int)zig_bit(sh.yes, p); }
    if (FIX) {
      for (int k = p; k < p + step && k < SEG; k++) sh.ym[k >> 6] &= ~(1ULL << (k & 63));
      if (y) sh.ym[p >> 6] |= 1ULL << (p & 63);
    }
    cnt += y;
    p += step;
  }
  if (p >= SEG) { exit_off = p - SEG; return cnt; }
  exit_off = exit0;
  const int before = sh.prefix[p >> 6] + __popcll(sh.ym[p >> 6] & ((1ULL << (p & 63)) - 1));  // 0-chain yields in [0, p)
  return cnt + sh.prefix[SEG_WORDS] - before;
}

__device__ __forceinline__ void zig_init_shared(ZigShared& sh) {
  for (int i = threadIdx.x; i < 256; i += RNG_THREADS) { sh.ki[i] = d_zig_ki[i]; sh.wi[i] = d_zig_wi[i]; }
  for (int i = threadIdx.x; i < SEG_WORDS; i += RNG_THREADS) { sh.yes[i] = 0; sh.covered[i] = 0; }
  if (threadIdx.x == 0) { sh.nlist = 0; sh.overflow = 0; sh.reach = 0; }
  __syncthreads();
}

// table entry: yields | exit << 16 ; exit 0xff = unsupported (overshoot >= ZIG_ENTRIES or list overflow)
__global__ void __launch_bounds__(RNG_THREADS) k_zig_scan(const __grid_constant__ Jump J, const ulonglong2* __restrict__ bs,
                                                         unsigned* __restrict__ table, u64* __restrict__ ym_all) {
  __shared__ ZigShared sh;
  zig_init_shared(sh);
  const ulonglong2* wsb = bs + (i64)blockIdx.x * ZIG_WARPS;
  const ulonglong2 b = wsb[0];
  u64 raws[RNG_PER_THREAD];
  zig_generate(sh, J, wsb, raws);
  __syncthreads();
  zig_evaluate(sh, J, b, false);
  __syncthreads();
  zig_chain0(sh);
  zig_prefix(sh);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + SEG_WORDS)  // the entry-0 yield mask, for k_zig_emit_fast
    ym_all[(i64)blockIdx.x * SEG_WORDS + threadIdx.x - 64] = sh.ym[threadIdx.x - 64];
  if (threadIdx.x < ZIG_ENTRIES) {
    int ex = 0xff, cnt = 0;
    if (!sh.overflow) {
      if (threadIdx.x == 0) { cnt = sh.prefix[SEG_WORDS]; ex = max(sh.reach, SEG) - SEG; }
      else cnt = zig_enter<false>(sh, threadIdx.x, ex);
      if (ex >= ZIG_ENTRIES) ex = 0xff;
    }
    table[(i64)blockIdx.x * ZIG_ENTRIES + threadIdx.x] = (unsigned)cnt | ((unsigned)ex << 16);
  }
}

// result block written by the prefix kernels / k_zig_emit*: [0] total yields of all segments, [1] exit offset of the
// last segment, [2] error flag, [3] stream position right after the draw that completed output n_out-1 (emit)
//
// Prefix over segments of the maps e -> (exit, yields): (1) k_zig_pf_reduce: a thread composes PF_RUN consecutive
// segments for all 16 entries, 16 threads then compose the block's 256 thread maps -> one map per block;
// (2) k_zig_pf_scan: one thread chains the block maps (staged through shared memory) -> every block's entry + base;
// (3) k_zig_pf_apply: the block rebuilds its thread maps, chains them from its entry, and every thread replays its run.
constexpr int PF_THREADS = 256;
constexpr int PF_RUN = 8;
constexpr int PF_SPAN = PF_THREADS * PF_RUN;  // segments per block
struct PfShared {
  unsigned cnt[PF_THREADS][ZIG_ENTRIES + 1];  // yields of a thread's run, per entry (<= 8 * 4096); padded against bank conflicts
  unsigned char ex[PF_THREADS][ZIG_ENTRIES];
  i64 tbase[PF_THREADS];
  unsigned char tentry[PF_THREADS];
};
struct BlockMap {
  i64 cnt[ZIG_ENTRIES];
  int ex[ZIG_ENTRIES];
};
// this thread's run of segments -> sh.cnt[t][*], sh.ex[t][*]; ends with a barrier
__device__ __forceinline__ void pf_thread_maps(PfShared& sh, const unsigned* __restrict__ table, i64 nseg) {
  const int t = threadIdx.x;
  const i64 s0 = min(((i64)blockIdx.x * PF_THREADS + t) * PF_RUN, nseg), s1 = min(s0 + PF_RUN, nseg);
#pragma unroll
  for (int e = 0; e < ZIG_ENTRIES; e++) {
    int cur = e;
    unsigned cnt = 0;
    for (i64 sg = s0; sg < s1 && cur != 0xff; sg++) {
      const unsigned v = __ldg(table + sg * ZIG_ENTRIES + cur);  // the 16 chains share the row's two sectors
      cnt += v & 0xffff;
      cur = (int)(v >> 16);
    }
    sh.cnt[t][e] = cnt;
    sh.ex[t][e] = (unsigned char)cur;
  }
  __syncthreads();
}
__global__ void __launch_bounds__(PF_THREADS) k_zig_pf_reduce(const unsigned* __restrict__ table, i64 nseg,
                                                             BlockMap* __restrict__ maps) {
  __shared__ PfShared sh;
  pf_thread_maps(sh, table, nseg);
  if (threadIdx.x < ZIG_ENTRIES) {
    int cur = threadIdx.x;
    i64 total = 0;
    for (int k = 0; k < PF_THREADS && cur != 0xff; k++) {
      total += sh.cnt[k][cur];
      cur = sh.ex[k][cur];
    }
    maps[blockIdx.x].cnt[threadIdx.x] = total;
    maps[blockIdx.x].ex[threadIdx.x] = cur;
  }
}
struct BlockStart {
  i64 base;
  int entry;
  int pad;
};
__global__ void __launch_bounds__(128) k_zig_pf_scan(const BlockMap* __restrict__ maps, int nblocks, BlockStart* __restrict__ starts,
                                                    i64* __restrict__ result) {
  constexpr int STAGE = 128;
  __shared__ BlockMap stage[STAGE];
  __shared__ int s_cur;
  __shared__ i64 s_total;
  if (threadIdx.x == 0) { s_cur = 0; s_total = 0; }
  for (int b0 = 0; b0 < nblocks; b0 += STAGE) {
    const int nb = min(STAGE, nblocks - b0);
    __syncthreads();
    const int words = nb * (int)(sizeof(BlockMap) / 8);
    for (int i = threadIdx.x; i < words; i += blockDim.x)
      reinterpret_cast<i64*>(stage)[i] = reinterpret_cast<const i64*>(maps + b0)[i];
    __syncthreads();
    if (threadIdx.x == 0) {
      int cur = s_cur;
      i64 total = s_total;
      for (int k = 0; k < nb; k++) {
        starts[b0 + k].base = total;
        starts[b0 + k].entry = cur;
        if (cur != 0xff) { total += stage[k].cnt[cur]; cur = stage[k].ex[cur]; }
      }
      s_cur = cur;
      s_total = total;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    result[0] = s_total;
    result[1] = s_cur;
    result[2] = s_cur == 0xff ? 1 : 0;
  }
}
__global__ void __launch_bounds__(PF_THREADS) k_zig_pf_apply(const unsigned* __restrict__ table, i64 nseg,
                                                            const BlockStart* __restrict__ starts,
                                                            unsigned char* __restrict__ seg_entry, i64* __restrict__ seg_base_out,
                                                            const i64* __restrict__ result) {
  __shared__ PfShared sh;
  if (result[2]) return;
  pf_thread_maps(sh, table, nseg);
  const int t = threadIdx.x;
  if (t == 0) {
    int cur = starts[blockIdx.x].entry;
    i64 base = starts[blockIdx.x].base;
    for (int k = 0; k < PF_THREADS; k++) {
      sh.tentry[k] = (unsigned char)cur;
      sh.tbase[k] = base;
      base += sh.cnt[k][cur];
      cur = sh.ex[k][cur];
    }
  }
  __syncthreads();
  const i64 s0 = min(((i64)blockIdx.x * PF_THREADS + t) * PF_RUN, nseg), s1 = min(s0 + PF_RUN, nseg);
  int cur = sh.tentry[t];
  i64 base = sh.tbase[t];
  for (i64 sg = s0; sg < s1; sg++) {
    seg_entry[sg] = (unsigned char)cur;
    seg_base_out[sg] = base;
    const unsigned v = __ldg(table + sg * ZIG_ENTRIES + cur);
    base += v & 0xffff;
    cur = (int)(v >> 16);
  }
}

// Emission, common case (the segment is entered at its first position): the yield mask k_zig_scan stored is all that
// is needed -- regenerate the draws and store each value at base + rank.  Tail draws (about one per segment) are
// recomputed in place.
__global__ void __launch_bounds__(RNG_THREADS) k_zig_emit_fast(const __grid_constant__ Jump J, const ulonglong2* __restrict__ bs,
                                                              const unsigned char* __restrict__ seg_entry,
                                                              const i64* __restrict__ seg_base_in, const u64* __restrict__ ym_all,
                                                              double* __restrict__ out, i64 n_out, double mean, double sd,
                                                              int affine, i64* __restrict__ result) {
  __shared__ u64 s_ki[256];
  __shared__ double s_wi[256];
  __shared__ u64 s_ym[SEG_WORDS];
  __shared__ int s_prefix[SEG_WORDS + 1];
  __shared__ int s_half;
  if (result[2]) return;
  if (seg_entry[blockIdx.x] != 0) return;  // k_zig_emit handles it
  const i64 out_base = seg_base_in[blockIdx.x];
  if (out_base >= n_out) return;
  const int t = threadIdx.x, lane = t & 31;
  s_ki[t] = d_zig_ki[t];
  s_wi[t] = d_zig_wi[t];
  int c = 0, incl = 0;
  if (t < SEG_WORDS) {
    const u64 m = ym_all[(i64)blockIdx.x * SEG_WORDS + t];
    s_ym[t] = m;
    c = __popcll(m);
    incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (t == 31) s_half = incl;
  }
  __syncthreads();
  if (t < SEG_WORDS) s_prefix[t] = incl - c + (t >= 32 ? s_half : 0);
  __syncthreads();
  const ulonglong2 b = bs[(i64)blockIdx.x * ZIG_WARPS + (t >> 5)];
  u64 hi = b.x, lo = b.y;
  jump_to(hi, lo, (u64)lane + 1, J);
  const u64 ph = J.ph[8], pl = J.pl[8];
  // INNER: every value of this segment lands before the last wanted element (all but the final segments), so the
  // output-range tests drop out of the loop
  auto body = [&](auto inner_tag) {
    constexpr bool INNER = decltype(inner_tag)::value;
#pragma unroll 4
    for (int j = 0; j < RNG_PER_THREAD; j++) {
      const int p = j * RNG_THREADS + t;
      const u64 raw = pcg_out(hi, lo);
      const u64 ym = s_ym[p >> 6];
      const i64 o = out_base + s_prefix[p >> 6] + __popcll(ym & ((1ULL << (p & 63)) - 1));
      if (((ym >> (p & 63)) & 1) && (INNER || o < n_out)) {
        int idx, sign;
        u64 rabs;
        zig_split(raw, idx, sign, rabs);
        double v = zig_x(rabs, s_wi[idx], sign);
        int consumed = 1;
        if (!(rabs < s_ki[idx])) {
          consumed = 2;  // an accepted wedge attempt
          if (idx == 0) {  // tail: replay its draws (NumPyRNG.scala:142-151)
            u64 th = hi, tl = lo;
            double xx, yy;
            consumed = 1;
            do {
              lcg(th, tl, J.mh[0], J.ml[0], J.ph[0], J.pl[0]);
              xx = -UM_ZIG_INV_R * log1p(-u01(pcg_out(th, tl)));
              lcg(th, tl, J.mh[0], J.ml[0], J.ph[0], J.pl[0]);
              yy = -log1p(-u01(pcg_out(th, tl)));
              consumed += 2;
            } while (!(__dadd_rn(yy, yy) > __dmul_rn(xx, xx)) && consumed < ZIG_MAX_CONS);
            v = ((rabs >> 8) & 1) ? -__dadd_rn(UM_ZIG_R, xx) : __dadd_rn(UM_ZIG_R, xx);
          }
        }
        if (affine) v = __dadd_rn(mean, __dmul_rn(sd, v));  // mean + std * rng.randn(), Mat.scala:1500-1507
        out[o] = v;
        if (!INNER && o == n_out - 1) result[3] = (i64)blockIdx.x * SEG + p + consumed;
      }
      lcg_stride256(hi, lo, ph, pl);
    }
  };
  if (out_base + SEG < n_out) body(std::true_type{});
  else body(std::false_type{});
}

__global__ void __launch_bounds__(RNG_THREADS, 3) k_zig_emit(const __grid_constant__ Jump J, const ulonglong2* __restrict__ bs,
                                                         const unsigned char* __restrict__ seg_entry,
                                                         const i64* __restrict__ seg_base_in, double* __restrict__ out,
                                                         i64 n_out, double mean, double sd, int affine,
                                                         i64* __restrict__ result) {
  __shared__ ZigShared sh;
  if (result[2]) return;  // the prefix met a segment it cannot chain: the host reports it
  if (seg_entry[blockIdx.x] == 0) return;  // k_zig_emit_fast handles it
  const i64 out_base = seg_base_in[blockIdx.x];
  if (out_base >= n_out) return;  // whole segment lies past the last wanted element
  zig_init_shared(sh);
  const i64 seg_base = (i64)blockIdx.x * SEG;
  const int entry = seg_entry[blockIdx.x];
  const int t = threadIdx.x;
  const ulonglong2* wsb = bs + (i64)blockIdx.x * ZIG_WARPS;
  const ulonglong2 b = wsb[0];
  u64 raws[RNG_PER_THREAD];
  zig_generate(sh, J, wsb, raws);
  __syncthreads();
  zig_evaluate(sh, J, b, true);
  __syncthreads();
  zig_chain0(sh);
  if (entry > 0) {  // block-uniform
    zig_prefix(sh);
    if (t == 0) {
      int ex;
      zig_enter<true>(sh, entry, ex);
    }
    __syncthreads();
  }
  zig_prefix(sh);
#pragma unroll
  for (int j = 0; j < RNG_PER_THREAD; j++) {
    const int p = j * RNG_THREADS + t;
    const u64 ym = sh.ym[p >> 6];
    if (!((ym >> (p & 63)) & 1)) continue;
    const i64 o = out_base + sh.prefix[p >> 6] + __popcll(ym & ((1ULL << (p & 63)) - 1));
    if (o >= n_out) continue;
    int idx, sign;
    u64 rabs;
    zig_split(raws[j], idx, sign, rabs);
    double v;
    int consumed = 1;
    if (rabs < sh.ki[idx]) {
      v = zig_x(rabs, sh.wi[idx], sign);
    } else {
      const int slot = sh.slot_of[p];
      consumed = sh.cons[slot];
      v = idx == 0 ? sh.tailv[slot] : zig_x(rabs, sh.wi[idx], sign);
    }
    if (affine) v = __dadd_rn(mean, __dmul_rn(sd, v));  // mean + std * rng.randn(), Mat.scala:1500-1507
    out[o] = v;
    if (o == n_out - 1) result[3] = seg_base + p + consumed;
  }
}

int upload_tables() {
  static std::atomic<unsigned long long> done{0};  // one bit per device
  Ctx* c = ctx();
  const unsigned long long bit = 1ULL << (c->device & 63);
  if (done.load() & bit) return UM_OK;
  UM_CUDA(cudaMemcpyToSymbol(d_zig_ki, h_zig_ki, sizeof(h_zig_ki)));
  UM_CUDA(cudaMemcpyToSymbol(d_zig_wi, h_zig_wi_bits, sizeof(h_zig_wi_bits)));
  UM_CUDA(cudaMemcpyToSymbol(d_zig_fi, h_zig_fi_bits, sizeof(h_zig_fi_bits)));
  UM_CUDA(cudaDeviceSynchronize());
  done.fetch_or(bit);
  return UM_OK;
}

int check_dst(const um_view* dst, int dtype, const char* what) {
  if (dst->dtype != dtype) return fail(UM_ERR_ARG, "%s: destination dtype %d, expected %d", what, dst->dtype, dtype);
  bool contiguous = dst->cs == 1 && (dst->rows <= 1 || dst->rs == dst->cols);
  if (!contiguous && dst->rows * dst->cols > 0)
    return fail(UM_ERR_ARG, "%s: destination must be a contiguous row-major matrix (the factories return fresh ones)", what);
  return UM_OK;
}

inline u128 st_of(const um_rng* g) { return mk128(g->state_hi, g->state_lo); }
inline u128 inc_of(const um_rng* g) { return mk128(g->inc_hi, g->inc_lo); }
inline void st_set(um_rng* g, u128 s) { g->state_hi = (u64)(s >> 64); g->state_lo = (u64)s; }

// ---- SeedSequence (NumPyRNG.scala:557-570 hashmix / mix, :635-666 mixEntropy, :685-718 generateState) ---------
const uint32_t SS_INIT_A = 0x43b0d7e5u, SS_MULT_A = 0x931e8875u, SS_INIT_B = 0x8b51f9ddu, SS_MULT_B = 0x58f38dedu,
               SS_MIX_L = 0xca01f9ddu, SS_MIX_R = 0x4973f715u;
inline uint32_t ss_hashmix(uint32_t v, uint32_t& hc) {
  v ^= hc; hc *= SS_MULT_A; v *= hc; v ^= v >> 16;
  return v;
}
inline uint32_t ss_mix(uint32_t x, uint32_t y) {
  uint32_t r = SS_MIX_L * x - SS_MIX_R * y;
  r ^= r >> 16;
  return r;
}

int launch_block_states(const Jump& J, const um_rng* g, i64 nblocks, int log2_span, ulonglong2* bs) {
  if (J.mh[8] != STRIDE256_MH || J.ml[8] != STRIDE256_ML) return fail(UM_ERR_CUDA, "rng: stride-256 multiplier constant is wrong");
  i64 threads = (nblocks + BS_CHUNK - 1) / BS_CHUNK;
  k_rng_block_states<<<(unsigned)((threads + 127) / 128), 128, 0, ctx()->stream>>>(J, g->state_hi, g->state_lo, nblocks, log2_span, bs);
  UM_LAUNCHED();
  return UM_OK;
}

template <int KIND>
int draw_impl(um_rng* g, const um_view* dst, double low, double span, int ilow, unsigned ibound, const char* what) {
  UM_CTX();
  UM_VIEW(dst);
  if (!g) return fail(UM_ERR_ARG, "%s: null generator", what);
  int s = check_dst(dst, KIND == DRAW_INT ? UM_I32 : UM_F64, what);
  if (s != UM_OK) return s;
  i64 n = dst->rows * dst->cols;
  if (n == 0) return UM_OK;
  double* outd = KIND == DRAW_INT ? nullptr : static_cast<double*>(dst->data) + dst->offset;
  int* outi = KIND == DRAW_INT ? static_cast<int*>(dst->data) + dst->offset : nullptr;
  i64 n_raw = n;
  int pre_flag = 0, pre_val = 0;
  if (KIND == DRAW_INT) {
    if (g->has_uint32) {  // the cached high half of the previous draw goes first (NumPyRNG.scala:86-101)
      pre_flag = 1;
      pre_val = ilow + (int)(((u64)g->uinteger * (u64)ibound) >> 32);
      g->has_uint32 = 0;
      outi += 1;
      n -= 1;
    }
    n_raw = (n + 1) / 2;
  }
  Jump J = make_jump(inc_of(g));
  i64 blocks = (n_raw + DRAW_SPAN - 1) / DRAW_SPAN;
  if (blocks == 0) blocks = 1;  // only the cached element
  if (blocks > 0x7fffffffLL) return fail(UM_ERR_ARG, "%s: %lld elements exceed one launch", what, (long long)n);
  ulonglong2* bs = static_cast<ulonglong2*>(scratch((size_t)blocks * sizeof(ulonglong2)));
  if (!bs) return UM_ERR_OOM;
  int s0 = launch_block_states(J, g, blocks, DRAW_LOG2_SPAN, bs);
  if (s0 != UM_OK) return s0;
  k_rng_draw<KIND><<<(unsigned)blocks, RNG_THREADS, 0, ctx()->stream>>>(J, bs, n_raw, outd, outi, n, low, span, ilow, ibound, pre_flag,
                                                                        pre_val);
  UM_LAUNCHED();
  u128 s2 = advance(st_of(g), inc_of(g), (u64)n_raw);
  if (KIND == DRAW_INT && (n & 1)) {
    g->has_uint32 = 1;
    g->uinteger = (uint32_t)(host_output(s2) >> 32);
  }
  st_set(g, s2);
  return UM_OK;
}

int randn_impl(um_rng* g, const um_view* dst, double mean, double sd, int affine, const char* what) {
  UM_CTX();
  UM_VIEW(dst);
  if (!g) return fail(UM_ERR_ARG, "%s: null generator", what);
  int s = check_dst(dst, UM_F64, what);
  if (s != UM_OK) return s;
  i64 n = dst->rows * dst->cols;
  if (n == 0) return UM_OK;
  s = upload_tables();
  if (s != UM_OK) return s;
  double* out = static_cast<double*>(dst->data) + dst->offset;
  cudaStream_t st = ctx()->stream;
  const Jump J = make_jump(inc_of(g));
  if (J.mh[8] != STRIDE256_MH || J.ml[8] != STRIDE256_ML) return fail(UM_ERR_CUDA, "rng: stride-256 multiplier constant is wrong");
  while (n > 0) {
    // 1.0173 draws per value on average; 3 % + two segments of margin makes a second round rare
    i64 nseg = (i64)((double)n * 1.03 / SEG) + 2;
    if (nseg > 0x7fffffffLL) return fail(UM_ERR_ARG, "%s: too many elements for one call", what);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const int pf_blocks = (int)((nseg + PF_SPAN - 1) / PF_SPAN);
    size_t off = 0;
    const size_t tab_off = off; off += up((size_t)nseg * ZIG_ENTRIES * sizeof(unsigned));
    const size_t base_off = off; off += up((size_t)nseg * sizeof(i64));
    const size_t entry_off = off; off += up((size_t)nseg);
    const size_t bs_off = off; off += up((size_t)nseg * ZIG_WARPS * sizeof(ulonglong2));
    const size_t ym_off = off; off += up((size_t)nseg * SEG_WORDS * sizeof(u64));
    const size_t map_off = off; off += up(std::max((size_t)pf_blocks * sizeof(BlockMap), (size_t)nseg * sizeof(int)));
    const size_t start_off = off; off += up((size_t)pf_blocks * sizeof(BlockStart));
    const size_t res_off = off; off += 64;
    char* ws = static_cast<char*>(scratch(off));
    if (!ws) return UM_ERR_OOM;
    unsigned* table = reinterpret_cast<unsigned*>(ws + tab_off);
    i64* seg_base = reinterpret_cast<i64*>(ws + base_off);
    unsigned char* seg_entry = reinterpret_cast<unsigned char*>(ws + entry_off);
    ulonglong2* bs = reinterpret_cast<ulonglong2*>(ws + bs_off);
    u64* ym_all = reinterpret_cast<u64*>(ws + ym_off);
    BlockMap* maps = reinterpret_cast<BlockMap*>(ws + map_off);
    BlockStart* starts = reinterpret_cast<BlockStart*>(ws + start_off);
    i64* result = reinterpret_cast<i64*>(ws + res_off);
    UM_CUDA(cudaMemsetAsync(result, 0, 64, st));
    {
      const i64 threads = (nseg + BS_CHUNK - 1) / BS_CHUNK;
      k_zig_warp_states<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(J, g->state_hi, g->state_lo, nseg, bs);
      UM_LAUNCHED();
    }
    k_zig_scan<<<(unsigned)nseg, RNG_THREADS, 0, st>>>(J, bs, table, ym_all);
    UM_LAUNCHED();
    k_zig_pf_reduce<<<pf_blocks, PF_THREADS, 0, st>>>(table, nseg, maps);
    UM_LAUNCHED();
    k_zig_pf_scan<<<1, 128, 0, st>>>(maps, pf_blocks, starts, result);
    UM_LAUNCHED();
    k_zig_pf_apply<<<pf_blocks, PF_THREADS, 0, st>>>(table, nseg, starts, seg_entry, seg_base, result);
    UM_LAUNCHED();
    k_zig_emit_fast<<<(unsigned)nseg, RNG_THREADS, 0, st>>>(J, bs, seg_entry, seg_base, ym_all, out, n, mean, sd, affine, result);
    UM_LAUNCHED();
    k_zig_emit<<<(unsigned)nseg, RNG_THREADS, 0, st>>>(J, bs, seg_entry, seg_base, out, n, mean, sd, affine, result);
    UM_LAUNCHED();
    i64 res[4];
    UM_CUDA(cudaMemcpyAsync(res, result, sizeof(res), cudaMemcpyDeviceToHost, st));
    UM_CUDA(cudaStreamSynchronize(st));
    if (res[2]) return fail(UM_ERR_ARG, "%s: a ziggurat tail loop ran past the %d-draw window this generator tabulates", what, ZIG_ENTRIES);
    if (res[0] >= n) {
      st_set(g, advance(st_of(g), inc_of(g), (u64)res[3]));
      n = 0;
    } else {  // the margin was too small: keep what was produced, continue behind the last segment's overshoot
      st_set(g, advance(st_of(g), inc_of(g), (u64)(nseg * SEG + res[1])));
      out += res[0];
      n -= res[0];
    }
  }
  return UM_OK;
}

}  // namespace
}  // namespace um

using namespace um;

extern "C" {

int32_t um_rng_seed(um_rng* g, int64_t seed) {
  if (!g) return fail(UM_ERR_ARG, "um_rng_seed: null generator");
  if (seed < 0) return fail(UM_ERR_ARG, "seed [%lld] must be non-negative", (long long)seed);  // NumPyRNG.scala:770-771
  // entropy -> little-endian 32-bit words, at least one (NumPyRNG.scala:484-501)
  uint32_t ent[2];
  int nent = 0;
  u64 v = (u64)seed;
  do { ent[nent++] = (uint32_t)v; v >>= 32; } while (v != 0);
  uint32_t pool[4], hc = SS_INIT_A;
  for (int i = 0; i < 4; i++) pool[i] = ss_hashmix(i < nent ? ent[i] : 0u, hc);
  for (int src = 0; src < 4; src++)
    for (int dst = 0; dst < 4; dst++)
      if (src != dst) pool[dst] = ss_mix(pool[dst], ss_hashmix(pool[src], hc));
  // generate_state(4, uint64) = 8 uint32 words paired little-endian
  uint32_t w32[8], hb = SS_INIT_B;
  for (int i = 0; i < 8; i++) {
    uint32_t d = pool[i & 3];
    d ^= hb; hb *= SS_MULT_B; d *= hb; d ^= d >> 16;
    w32[i] = d;
  }
  u64 w[4];
  for (int i = 0; i < 4; i++) w[i] = (u64)w32[2 * i] | ((u64)w32[2 * i + 1] << 32);
  // pcg64_set_seed / pcg_setseq_128_srandom_r (NumPyRNG.scala:761-767,780-786)
  u128 initstate = mk128(w[0], w[1]), initseq = mk128(w[2], w[3]);
  u128 inc = (initseq << 1) | 1;
  u128 s = 0 * PCG_MULT + inc;
  s += initstate;
  s = s * PCG_MULT + inc;
  g->state_hi = (u64)(s >> 64); g->state_lo = (u64)s;
  g->inc_hi = (u64)(inc >> 64); g->inc_lo = (u64)inc;
  g->has_uint32 = 0;
  g->uinteger = 0;
  return UM_OK;
}

int32_t um_rng_advance(um_rng* g, uint64_t draws) {
  if (!g) return fail(UM_ERR_ARG, "um_rng_advance: null generator");
  st_set(g, advance(st_of(g), inc_of(g), draws));
  return UM_OK;
}

int32_t um_rng_next_u64(um_rng* g, uint64_t* out) {
  if (!g || !out) return fail(UM_ERR_ARG, "um_rng_next_u64: null argument");
  u128 s = st_of(g) * PCG_MULT + inc_of(g);
  st_set(g, s);
  *out = host_output(s);
  return UM_OK;
}

int32_t um_rng_rand(um_rng* g, const um_view* dst) { return draw_impl<DRAW_U01>(g, dst, 0, 0, 0, 0, "um_rng_rand"); }

int32_t um_rng_uniform(um_rng* g, double low, double high, const um_view* dst) {
  return draw_impl<DRAW_UNIFORM>(g, dst, low, high - low, 0, 0, "um_rng_uniform");
}

int32_t um_rng_randint(um_rng* g, int32_t low, int32_t high, const um_view* dst_i32) {
  // nextBoundedInt: require(bound > 0) (NumPyRNG.scala:83); high - low wraps exactly like the JVM's Int
  int32_t bound = (int32_t)((uint32_t)high - (uint32_t)low);
  if (bound <= 0) return fail(UM_ERR_ARG, "bound must be positive");
  return draw_impl<DRAW_INT>(g, dst_i32, 0, 0, low, (unsigned)bound, "um_rng_randint");
}

int32_t um_rng_randn(um_rng* g, const um_view* dst) { return randn_impl(g, dst, 0.0, 1.0, 0, "um_rng_randn"); }

int32_t um_rng_normal(um_rng* g, double mean, double sd, const um_view* dst) {
  return randn_impl(g, dst, mean, sd, 1, "um_rng_normal");
}

}  // extern "C"
