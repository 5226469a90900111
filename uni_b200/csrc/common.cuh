// common.cuh -- shared host/device helpers for libunimat (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>

#include "unimat.h"

namespace um {

typedef long long i64;

// ---- per-host-thread context: device, stream, timing events, scratch ------------------
struct Ctx {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 0;
  void* scratch = nullptr;  // reduction partials etc. (stream-ordered reuse)
  size_t scratch_bytes = 0;
};

Ctx* ctx();  // lazily initialised on device 0 (or the device chosen by um_init)
int ensure_ctx();
unsigned long long op_epoch();  // content-changing library calls made by this host thread so far
void op_epoch_passive();
int fail(int code, const char* fmt, ...);
void* scratch(size_t bytes);  // nullptr + error set on failure
extern std::atomic<long long> g_launches;

// accumulators of a 3PRF fit with more than 8 proxies, filled group by group by the fused sweeps (tprf.cu) and finished
// with the library's operators (tprf_wide.cu): W0 N x L, PX T x L, h0 T x 1, isd = 1/sd and m = mu/sd 1 x N
struct TprfWideAcc { um_view W0, PX, h0, isd, m; };
int tprf_fit_wide(int mode, const um_view* y, const um_view* Z, i64 N, const TprfWideAcc& acc, double n_total, bool sharded,
                  um_tprf_out* out);

// per-kernel event timing (bench harness): UM_PROF_BEGIN(slot) ... launch ... UM_PROF_END(slot)
bool prof_on();
void prof_begin(int slot);
void prof_end(int slot);

#define UM_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return um::fail(_e == cudaErrorMemoryAllocation ? UM_ERR_OOM : UM_ERR_CUDA, "%s: %s", #expr, \
                      cudaGetErrorString(_e));                                             \
  } while (0)

#define UM_CTX()                      \
  do {                                \
    int _s = um::ensure_ctx();        \
    if (_s != UM_OK) return _s;       \
  } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define UM_LAUNCHED()                                                                 \
  do {                                                                                \
    um::g_launches.fetch_add(1, std::memory_order_relaxed);                           \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) return um::fail(UM_ERR_CUDA, "kernel launch failed (%s:%d): %s", __FILE__, __LINE__, \
                                           cudaGetErrorString(_e));                   \
  } while (0)

#define UM_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return um::fail(UM_ERR_ARG, __VA_ARGS__);      \
  } while (0)

inline size_t dtype_size(int dt) { return dt == UM_F64 ? 8 : dt == UM_F32 ? 4 : dt == UM_I32 ? 4 : 1; }

inline int check_view(const um_view* v, const char* name) {
  if (!v) return fail(UM_ERR_ARG, "%s: null view", name);
  if (v->rows < 0 || v->cols < 0) return fail(UM_ERR_ARG, "%s: negative shape %lldx%lld", name, (i64)v->rows, (i64)v->cols);
  if (v->dtype < 0 || v->dtype > UM_I32) return fail(UM_ERR_ARG, "%s: bad dtype %d", name, v->dtype);
  if (v->rows > 0 && v->cols > 0 && !v->data) return fail(UM_ERR_ARG, "%s: null data", name);
  return UM_OK;
}
#define UM_VIEW(v)                          \
  do {                                      \
    int _s = um::check_view((v), #v);       \
    if (_s != UM_OK) return _s;             \
  } while (0)

// typed operand: pointer already advanced to element (0,0)
template <class T>
struct In {
  const T* p;
  i64 rs, cs;
};
template <class T>
struct Out {
  T* p;
  i64 rs, cs;
};
template <class T>
inline In<T> in_of(const um_view* v) {
  return In<T>{static_cast<const T*>(v->data) + v->offset, v->rs, v->cs};
}
template <class T>
inline Out<T> out_of(const um_view* v) {
  return Out<T>{static_cast<T*>(v->data) + v->offset, v->rs, v->cs};
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// can rows of this operand be read/written with 16-byte vectors of VW elements?
template <class T>
inline bool vec_ok(const T* p, i64 rs, i64 cs, i64 rows, int vw) {
  if (cs == 0) return true;  // splat
  if (cs != 1) return false;
  if (!aligned16(p)) return false;
  if (rows > 1 && (rs % vw) != 0) return false;
  return true;
}

// ---- device helpers --------------------------------------------------------------------
__device__ __forceinline__ double shfl_down_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ i64 shfl_down_l(i64 v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

// java.lang.Double.compare order as a signed 64-bit key (Mat.scala:418-440,892-895):
// -Inf < ... < -0.0 < +0.0 < ... < +Inf < NaN, every NaN equal.
__device__ __forceinline__ i64 total_key(double x) {
  if (x != x) return 0x7ff8000000000000LL;
  i64 b = __double_as_longlong(x);
  return b < 0 ? (b ^ 0x7fffffffffffffffLL) : b;
}
__device__ __forceinline__ i64 total_key(float x) { return total_key((double)x); }

// exp(x) in 18 FP64 instructions (libdevice's exp: about 40): x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor / Horner
// (truncation 4e-18), 2^k applied to the exponent field.  Max error < 2 ulp on [-700, 700]; anything outside (and NaN)
// takes libdevice's exp.
// The coefficients live in CONSTANT memory on purpose: written as literals, ptxas materialises each 64-bit immediate
// with two UMOVs in front of its DFMA, and the sigmoid / softmax kernels -- 85 issued instructions per element, two
// thirds of them moves -- were bound by instruction issue (81 % issue-active, FP64 pipe 40 %; profiles/
// r1_sigmoid_softmax0_ncu_raw.csv), not by HBM.  From the constant bank each DFMA takes its coefficient as an operand.
__constant__ double k_expc[16] = {
    1.4426950408889634074,       // 0: 1/ln2
    -6.93147180369123816490e-01, // 1: -ln2 high part (32 significant bits: t * hi is exact)
    -1.90821492927058770002e-10, // 2: -ln2 low part
    1.6059043836821613e-10,      // 3: 1/13!
    2.08767569878681e-09,        // 4: 1/12!
    2.505210838544172e-08,       // 5: 1/11!
    2.755731922398589e-07,       // 6: 1/10!
    2.7557319223985893e-06,      // 7: 1/9!
    2.48015873015873e-05,        // 8: 1/8!
    1.984126984126984e-04,       // 9: 1/7!
    1.388888888888889e-03,       // 10: 1/6!
    8.333333333333333e-03,       // 11: 1/5!
    4.1666666666666664e-02,      // 12: 1/4!
    1.6666666666666666e-01,      // 13: 1/3!
    6755399441055744.0,          // 14: 1.5 * 2^52: low mantissa bits of (v + MAGIC) = rint(v)
    0.0};
__device__ __forceinline__ double fast_exp(double x) {
  if (!(fabs(x) <= 700.0)) return exp(x);
  const double* __restrict__ c = k_expc;
  double tt = fma(x, c[0], c[14]);
  int k = __double2loint(tt);
  double t = tt - c[14];
  double r = fma(t, c[1], x);
  r = fma(t, c[2], r);
  double p = c[3];
  p = fma(p, r, c[4]);
  p = fma(p, r, c[5]);
  p = fma(p, r, c[6]);
  p = fma(p, r, c[7]);
  p = fma(p, r, c[8]);
  p = fma(p, r, c[9]);
  p = fma(p, r, c[10]);
  p = fma(p, r, c[11]);
  p = fma(p, r, c[12]);
  p = fma(p, r, c[13]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// The table-based exp of the strip softmax kernel (softmax_strip.cuh): x = (32 q + j) ln2/32 + r, |r| <= ln2/64,
// 2^(j/32) from a 32-entry table the kernel copies into SHARED memory, degree-6 Taylor for exp(r): 12 FP64
// instructions, <= 2 ulp on [-700, 700].  WHERE THE TABLE LIVES decides everything: read per lane through the
// read-only global path (a drop-in replacement for fast_exp) it measured slower than the 18-instruction polynomial
// (sigmoid 3.19 -> 3.30 ms, softmax(1) 3.20 -> 3.56 ms at 32768^2); from shared memory it is faster (sigmoid 3.19 ->
// 2.79 ms = 0.94 of HBM).  So the kernels that can afford 256 bytes of shared memory use exp_tab_nonpos below.
__constant__ double k_exp2_32[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924,
    1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832, 1.4142135623730951, 1.4451808069770467,
    1.4768261459394993, 1.5091644275934228, 1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103,
    1.9152065613971474, 1.9571441241754002};
__constant__ double k_exp32c[10] = {
    46.16624130844683,        // 0: 32 / ln2
    -0.02166084938653512,     // 1: -(ln2 / 32) high part (32 significant bits: t * hi is exact for |t| < 2^21)
    -5.9631716539705866e-12,  // 2: -(ln2 / 32) low part
    6755399441055744.0,       // 3: 1.5 * 2^52: low mantissa bits of (v + MAGIC) = rint(v)
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,   // 4..7
    0.0, 0.0};
// the polynomial part: exp(x) / 2^(j/32) scaled into place, and j; valid for |x| <= 700
__device__ __forceinline__ double exp32_core(double x, int* k_out) {
  const double* __restrict__ c = k_exp32c;
  double tt = fma(x, c[0], c[3]);
  *k_out = __double2loint(tt);
  const double t = tt - c[3];
  double r = fma(t, c[1], x);
  r = fma(t, c[2], r);
  double p = c[4];
  p = fma(p, r, c[5]);
  p = fma(p, r, c[6]);
  p = fma(p, r, c[7]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return p;
}
__device__ __forceinline__ double exp32_scale(double p, int k) {   // p * 2^(k >> 5)
  return __hiloint2double(__double2hiint(p) + ((k << 15) & 0xfff00000), __double2loint(p));
}
// exp(d) for d <= 0 (or NaN), `tab` = the shared-memory address of a copy of k_exp2_32: the table form, then -- a
// branch normal data never takes -- libdevice's exp for d < -700 (denormal results), -Inf and NaN
__device__ __forceinline__ double exp_tab_nonpos(double d, unsigned tab) {
  int k;
  const double p = exp32_core(d, &k);
  double tj;
  asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(tab + (((unsigned)k << 3) & 0xf8u)));
  double z = exp32_scale(p * tj, k);
  if (!(d >= -700.0)) z = exp(d);
  return z;
}
// 1/d for normal, finite d (callers: d in [1, 2^40]): MUFU seed (2^-23) + two Newton steps.
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// One-time host-side setup shared by all host threads (kernel attributes, occupancy results, driver entry points):
// the functions that own such static state take this lock (one device per process, so the state is per process).
std::mutex& plan_mutex();

}  // namespace um
