// elementwise.cu -- broadcast binary ops, scalar ops, unary ops / activations, masks, copies,
// fills and synthetic-panel generators over strided views.
//
// One kernel shape for everything: a 2-D thread block (TX along columns, TY along rows,
// TX*TY = 256), each thread moving UNROLL 128-bit vectors per operand when every operand's
// rows are 16-byte addressable (cs == 1, or cs == 0 -> splat), otherwise one element at a
// time through the stride equation.  Dense same-shape operands collapse to one long row.
// HBM-bound: no shared memory, loads issued before any compute, grid >> 148 SMs.
#include <type_traits>

#include "common.cuh"

namespace um {

constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;

template <class T, int VW>
struct alignas(sizeof(T) * VW) Vec {
  T v[VW];
};

template <class T, int VW>
__device__ __forceinline__ Vec<T, VW> load_vec(const T* row, i64 c, i64 cs) {
  Vec<T, VW> r;
  if (VW > 1) {
    if (cs == 0) {
      T s = row[0];
#pragma unroll
      for (int i = 0; i < VW; ++i) r.v[i] = s;
    } else {
      r = *reinterpret_cast<const Vec<T, VW>*>(row + c);
    }
  } else {
    r.v[0] = row[c * cs];
  }
  return r;
}

// ---- functors ----------------------------------------------------------------------------
struct FAdd { template <class T> __device__ T operator()(T a, T b) const { return a + b; } };
struct FSub { template <class T> __device__ T operator()(T a, T b) const { return a - b; } };
struct FMul { template <class T> __device__ T operator()(T a, T b) const { return a * b; } };
struct FDiv { template <class T> __device__ T operator()(T a, T b) const { return a / b; } };
// maximum/minimum compare with Ordering[Double] = java.lang.Double.compare (Mat.scala:418-440)
struct FMaximum { template <class T> __device__ T operator()(T a, T b) const { return total_key(b) > total_key(a) ? b : a; } };
struct FMinimum { template <class T> __device__ T operator()(T a, T b) const { return total_key(b) < total_key(a) ? b : a; } };
struct FPow {
  __device__ double operator()(double a, double b) const { return pow(a, b); }
  __device__ float operator()(float a, float b) const { return (float)pow((double)a, (double)b); }
};

template <class F, class T, bool LEFT>
struct BindScalar {
  F f;
  T s;
  __device__ T operator()(T x) const { return LEFT ? f(s, x) : f(x, s); }
};

struct UNeg { template <class TI, class TO> __device__ TO apply(TI x) const { return (TO)(-x); } };
// abs = x < 0 ? -x : x, keeps -0.0 (Mat.scala:3557-3560)
struct UAbs { template <class TI, class TO> __device__ TO apply(TI x) const { return (TO)(x < (TI)0 ? -x : x); } };
struct UExp { template <class TI, class TO> __device__ TO apply(TI x) const { return (TO)fast_exp((double)x); } };
struct USqrt { template <class TI, class TO> __device__ TO apply(TI x) const { return (TO)sqrt((double)x); } };
struct ULog { template <class TI, class TO> __device__ TO apply(TI x) const { return (TO)log((double)x); } };
struct UTanh { template <class TI, class TO> __device__ TO apply(TI x) const { return (TO)tanh((double)x); } };
struct USquare { template <class TI, class TO> __device__ TO apply(TI x) const { return (TO)(x * x); } };
struct UCast { template <class TI, class TO> __device__ TO apply(TI x) const { return (TO)x; } };
// sigmoid: x >= 0 ? 1/(1+exp(-x)) : (e = exp(x); e/(1+e)); NaN takes the second branch
// (MatMathOps.scala:103-107); computed in double, as the reference widens to Double.
struct USigmoid {
  static constexpr bool exptab = true;   // k_unary gives the functor a copy of the 2^(j/32) table in shared memory
  unsigned tab;
  template <class TI, class TO>
  __device__ TO apply(TI xi) const {
    // same two formulas with z = exp(-|x|): x >= 0 -> 1/(1+z), else z/(1+z); the division is
    // num * rcp(1+z) with 1+z in [1,2] (<= 2 ulp from the reference's IEEE divide).
    // exp: the 12-instruction table form of common.cuh with the table in shared memory (per-lane index: the constant
    // bank would serialise it, the global read-only path measured slower than six more DFMA); arguments below -700
    // (denormal results) and NaN are redone by libdevice's exp afterwards -- a branch normal data never takes.
    double x = (double)xi;
    const double z = exp_tab_nonpos(-fabs(x), tab);
    double num = (x >= 0.0) ? 1.0 : z;
    return (TO)(num * fast_rcp(1.0 + z));
  }
};
// relu: Math.max(x, 0.0) -- NaN survives, -0.0 -> +0.0 (MatMathOps.scala:131); the generic
// branch `if gt(x, 0) x else 0` under the total order gives the same values (:133-146).
struct URelu {
  template <class TI, class TO>
  __device__ TO apply(TI x) const { return (TO)((x != x) ? x : (x > (TI)0 ? x : (TI)0)); }
};
template <class F, class T>
struct UFromBound {  // adapts a bound scalar functor T->T to the unary interface
  F f;
  template <class TI, class TO> __device__ TO apply(TI x) const { return (TO)f(x); }
};

// IEEE compares (Mat.scala:2576-2582): NaN makes everything false except `!=`
template <int OP>
struct FCmp {
  template <class T>
  __device__ uint8_t operator()(T a, T b) const {
    double x = (double)a, y = (double)b;
    return OP == UM_GT ? x > y : OP == UM_LT ? x < y : OP == UM_GTE ? x >= y : OP == UM_LTE ? x <= y : OP == UM_EQ ? x == y : x != y;
  }
};
template <int OP>
struct UCmpScalar {
  double s;
  template <class TI, class TO>
  __device__ TO apply(TI a) const {
    double x = (double)a;
    return (TO)(OP == UM_GT ? x > s : OP == UM_LT ? x < s : OP == UM_GTE ? x >= s : OP == UM_LTE ? x <= s : OP == UM_EQ ? x == s : x != s);
  }
};
template <int OP>
struct UClass {
  template <class TI, class TO>
  __device__ TO apply(TI a) const {
    double x = (double)a;
    return (TO)(OP == UM_ISNAN ? (x != x) : OP == UM_ISINF ? (isinf(x) != 0) : (isfinite(x) != 0));
  }
};
template <int OP>
struct FLogic {
  __device__ uint8_t operator()(uint8_t a, uint8_t b) const {
    bool x = a != 0, y = b != 0;
    return OP == UM_AND ? (x && y) : OP == UM_OR ? (x || y) : (x != y);
  }
};
struct UNot { template <class TI, class TO> __device__ TO apply(TI x) const { return (TO)(x == 0); } };

// ---- kernels -----------------------------------------------------------------------------
template <class F, class = void> struct wants_exptab : std::false_type {};
template <class F> struct wants_exptab<F, std::enable_if_t<F::exptab>> : std::true_type {};

template <int VW, class F, class TA, class TO>
__global__ void __launch_bounds__(EW_THREADS) k_unary(F f, In<TA> a, Out<TO> o, i64 rows, i64 cols) {
  if constexpr (wants_exptab<F>::value) {
    __shared__ double s_tab[32];
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    if (t < 32) s_tab[t] = k_exp2_32[t];
    __syncthreads();
    f.tab = (unsigned)__cvta_generic_to_shared(s_tab);
  }
  const int TX = blockDim.x, TY = blockDim.y;
  const i64 cbase = ((i64)blockIdx.x * EW_UNROLL * TX + threadIdx.x) * VW;
  for (i64 r = (i64)blockIdx.y * TY + threadIdx.y; r < rows; r += (i64)gridDim.y * TY) {
    const TA* ar = a.p + r * a.rs;
    TO* orow = o.p + r * o.rs;
    Vec<TA, VW> va[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      i64 c = cbase + (i64)u * TX * VW;
      if (c + VW <= cols) va[u] = load_vec<TA, VW>(ar, c, a.cs);
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      i64 c = cbase + (i64)u * TX * VW;
      if (c + VW <= cols) {
        Vec<TO, VW> vo;
#pragma unroll
        for (int i = 0; i < VW; ++i) vo.v[i] = f.template apply<TA, TO>(va[u].v[i]);
        if (VW > 1) *reinterpret_cast<Vec<TO, VW>*>(orow + c) = vo;
        else orow[c * o.cs] = vo.v[0];
      } else if (VW > 1 && c < cols) {
        for (i64 cc = c; cc < cols; ++cc) orow[cc] = f.template apply<TA, TO>(ar[cc * a.cs]);
      }
    }
  }
}

template <int VW, class F, class TA, class TB, class TO>
__global__ void __launch_bounds__(EW_THREADS) k_binary(F f, In<TA> a, In<TB> b, Out<TO> o, i64 rows, i64 cols) {
  const int TX = blockDim.x, TY = blockDim.y;
  const i64 cbase = ((i64)blockIdx.x * EW_UNROLL * TX + threadIdx.x) * VW;
  for (i64 r = (i64)blockIdx.y * TY + threadIdx.y; r < rows; r += (i64)gridDim.y * TY) {
    const TA* ar = a.p + r * a.rs;
    const TB* br = b.p + r * b.rs;
    TO* orow = o.p + r * o.rs;
    Vec<TA, VW> va[EW_UNROLL];
    Vec<TB, VW> vb[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      i64 c = cbase + (i64)u * TX * VW;
      if (c + VW <= cols) {
        va[u] = load_vec<TA, VW>(ar, c, a.cs);
        vb[u] = load_vec<TB, VW>(br, c, b.cs);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      i64 c = cbase + (i64)u * TX * VW;
      if (c + VW <= cols) {
        Vec<TO, VW> vo;
#pragma unroll
        for (int i = 0; i < VW; ++i) vo.v[i] = f(va[u].v[i], vb[u].v[i]);
        if (VW > 1) *reinterpret_cast<Vec<TO, VW>*>(orow + c) = vo;
        else orow[c * o.cs] = vo.v[0];
      } else if (VW > 1 && c < cols) {
        for (i64 cc = c; cc < cols; ++cc) orow[cc] = f(ar[cc * a.cs], br[cc * b.cs]);
      }
    }
  }
}

// ---- launch geometry -----------------------------------------------------------------------
struct Geom {
  dim3 grid, block;
};
static Geom geom(i64 rows, i64 cols, int vw) {
  i64 vcols = (cols + vw - 1) / vw;
  int tx = EW_THREADS;
  while (tx > 1 && tx / 2 >= vcols) tx /= 2;  // smallest power of two >= vcols, capped at 256
  int ty = EW_THREADS / tx;
  i64 gx = (vcols + (i64)tx * EW_UNROLL - 1) / ((i64)tx * EW_UNROLL);
  i64 gy = (rows + ty - 1) / ty;
  if (gy > 65535) gy = 65535;
  if (gx < 1) gx = 1;
  if (gy < 1) gy = 1;
  Geom g;
  g.grid = dim3((unsigned)gx, (unsigned)gy, 1);
  g.block = dim3(tx, ty, 1);
  return g;
}

static inline bool dense(i64 rs, i64 cs, i64 rows, i64 cols) { return cs == 1 && (rs == cols || rows <= 1); }

template <class F, class TA, class TO>
static int launch_unary(F f, const um_view* av, const um_view* ov) {
  i64 rows = ov->rows, cols = ov->cols;
  if (rows == 0 || cols == 0) return UM_OK;
  In<TA> a = in_of<TA>(av);
  Out<TO> o = out_of<TO>(ov);
  if (dense(a.rs, a.cs, rows, cols) && dense(o.rs, o.cs, rows, cols)) {
    cols = rows * cols;
    rows = 1;
    a.rs = o.rs = 0;
  }
  constexpr int VW = 16 / (sizeof(TA) > sizeof(TO) ? sizeof(TA) : sizeof(TO));
  bool v = VW > 1 && vec_ok(a.p, a.rs, a.cs, rows, VW) && o.cs == 1 && vec_ok(o.p, o.rs, o.cs, rows, VW);
  cudaStream_t st = ctx()->stream;
  if (v) {
    Geom g = geom(rows, cols, VW);
    k_unary<VW, F, TA, TO><<<g.grid, g.block, 0, st>>>(f, a, o, rows, cols);
  } else {
    Geom g = geom(rows, cols, 1);
    k_unary<1, F, TA, TO><<<g.grid, g.block, 0, st>>>(f, a, o, rows, cols);
  }
  UM_LAUNCHED();
  return UM_OK;
}

template <class F, class TA, class TB, class TO>
static int launch_binary(F f, const um_view* av, const um_view* bv, const um_view* ov) {
  i64 rows = ov->rows, cols = ov->cols;
  if (rows == 0 || cols == 0) return UM_OK;
  In<TA> a = in_of<TA>(av);
  In<TB> b = in_of<TB>(bv);
  Out<TO> o = out_of<TO>(ov);
  // broadcasting: an operand dimension of 1 against a larger target reads with stride 0
  if (av->rows == 1 && rows > 1) a.rs = 0;
  if (av->cols == 1 && cols > 1) a.cs = 0;
  if (bv->rows == 1 && rows > 1) b.rs = 0;
  if (bv->cols == 1 && cols > 1) b.cs = 0;
  if (dense(a.rs, a.cs, rows, cols) && dense(b.rs, b.cs, rows, cols) && dense(o.rs, o.cs, rows, cols)) {
    cols = rows * cols;
    rows = 1;
    a.rs = b.rs = o.rs = 0;
  }
  constexpr size_t big = sizeof(TA) > sizeof(TO) ? sizeof(TA) : sizeof(TO);
  constexpr int VW = 16 / big;
  bool v = VW > 1 && vec_ok(a.p, a.rs, a.cs, rows, VW) && vec_ok(b.p, b.rs, b.cs, rows, VW) && o.cs == 1 &&
           vec_ok(o.p, o.rs, o.cs, rows, VW);
  cudaStream_t st = ctx()->stream;
  if (v) {
    Geom g = geom(rows, cols, VW);
    k_binary<VW, F, TA, TB, TO><<<g.grid, g.block, 0, st>>>(f, a, b, o, rows, cols);
  } else {
    Geom g = geom(rows, cols, 1);
    k_binary<1, F, TA, TB, TO><<<g.grid, g.block, 0, st>>>(f, a, b, o, rows, cols);
  }
  UM_LAUNCHED();
  return UM_OK;
}

static int check_bcast(const um_view* a, const um_view* b, const um_view* o) {
  i64 tr = a->rows > b->rows ? a->rows : b->rows, tc = a->cols > b->cols ? a->cols : b->cols;
  bool ok = (a->rows == tr || a->rows == 1) && (a->cols == tc || a->cols == 1) && (b->rows == tr || b->rows == 1) &&
            (b->cols == tc || b->cols == 1);
  if (!ok)
    return fail(UM_ERR_ARG, "Cannot broadcast shapes (%lld, %lld) and (%lld, %lld)", (i64)a->rows, (i64)a->cols,
                (i64)b->rows, (i64)b->cols);
  if (o->rows != tr || o->cols != tc)
    return fail(UM_ERR_ARG, "output shape (%lld, %lld) != broadcast shape (%lld, %lld)", (i64)o->rows, (i64)o->cols, tr, tc);
  return UM_OK;
}

template <class F>
static int binary_typed(F f, const um_view* a, const um_view* b, const um_view* o) {
  if (a->dtype == UM_F64 && b->dtype == UM_F64 && o->dtype == UM_F64) return launch_binary<F, double, double, double>(f, a, b, o);
  if (a->dtype == UM_F32 && b->dtype == UM_F32 && o->dtype == UM_F32) return launch_binary<F, float, float, float>(f, a, b, o);
  return fail(UM_ERR_ARG, "binary op needs matching f64 or f32 operands (got %d, %d -> %d)", a->dtype, b->dtype, o->dtype);
}

template <class F>
static int scalar_typed(F f, const um_view* a, double s, int left, const um_view* o) {
  if (a->dtype == UM_F64 && o->dtype == UM_F64) {
    if (left) return launch_unary<UFromBound<BindScalar<F, double, true>, double>, double, double>({{f, s}}, a, o);
    return launch_unary<UFromBound<BindScalar<F, double, false>, double>, double, double>({{f, s}}, a, o);
  }
  if (a->dtype == UM_F32 && o->dtype == UM_F32) {
    float fs = (float)s;
    if (left) return launch_unary<UFromBound<BindScalar<F, float, true>, float>, float, float>({{f, fs}}, a, o);
    return launch_unary<UFromBound<BindScalar<F, float, false>, float>, float, float>({{f, fs}}, a, o);
  }
  return fail(UM_ERR_ARG, "scalar op needs f64->f64 or f32->f32 (got %d -> %d)", a->dtype, o->dtype);
}

template <class F>
static int unary_typed(F f, const um_view* a, const um_view* o) {
  if (a->dtype == UM_F64 && o->dtype == UM_F64) return launch_unary<F, double, double>(f, a, o);
  if (a->dtype == UM_F32 && o->dtype == UM_F32) return launch_unary<F, float, float>(f, a, o);
  if (a->dtype == UM_F32 && o->dtype == UM_F64) return launch_unary<F, float, double>(f, a, o);
  if (a->dtype == UM_F64 && o->dtype == UM_F32) return launch_unary<F, double, float>(f, a, o);
  return fail(UM_ERR_ARG, "unary op needs f64/f32 operands (got %d -> %d)", a->dtype, o->dtype);
}

template <class F>
static int tobool_typed(F f, const um_view* a, const um_view* o) {
  if (o->dtype != UM_BOOL) return fail(UM_ERR_ARG, "mask output must be bool");
  if (a->dtype == UM_F64) return launch_unary<F, double, uint8_t>(f, a, o);
  if (a->dtype == UM_F32) return launch_unary<F, float, uint8_t>(f, a, o);
  return fail(UM_ERR_ARG, "mask input must be f64 or f32");
}

template <class F>
static int cmp_typed(F f, const um_view* a, const um_view* b, const um_view* o) {
  if (o->dtype != UM_BOOL) return fail(UM_ERR_ARG, "mask output must be bool");
  if (a->dtype == UM_F64 && b->dtype == UM_F64) return launch_binary<F, double, double, uint8_t>(f, a, b, o);
  if (a->dtype == UM_F32 && b->dtype == UM_F32) return launch_binary<F, float, float, uint8_t>(f, a, b, o);
  return fail(UM_ERR_ARG, "compare needs matching f64 or f32 operands");
}

// ---- synthetic panels: counter-based generator (splitmix64 hash of seed+index) -------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

template <class T, bool NORMAL>
__global__ void __launch_bounds__(256) k_random(Out<T> o, i64 rows, i64 cols, uint64_t seed) {
  // one thread makes TWO values (Box-Muller pair) at logical indices 2p, 2p+1
  i64 n = rows * cols;
  i64 pairs = (n + 1) / 2;
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += (i64)gridDim.x * blockDim.x) {
    uint64_t h1 = mix64(seed * 0x2545f4914f6cdd1dULL + 2 * (uint64_t)p);
    uint64_t h2 = mix64(seed * 0x2545f4914f6cdd1dULL + 2 * (uint64_t)p + 1);
    double a, b;
    if (NORMAL) {
      double u1 = u01(h1) + (1.0 / 18014398509481984.0);  // (0,1]
      double u2 = u01(h2);
      double rad = sqrt(-2.0 * log(u1));
      double s, c;
      sincospi(2.0 * u2, &s, &c);
      a = rad * c;
      b = rad * s;
    } else {
      a = u01(h1);
      b = u01(h2);
    }
    i64 i0 = 2 * p, i1 = 2 * p + 1;
    o.p[(i0 / cols) * o.rs + (i0 % cols) * o.cs] = (T)a;
    if (i1 < n) o.p[(i1 / cols) * o.rs + (i1 % cols) * o.cs] = (T)b;
  }
}

template <class T>
__global__ void k_eye(Out<T> o, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o.p[i * o.rs + i * o.cs] = (T)1;
}

struct UConst {
  double c;
  template <class TI, class TO> __device__ TO apply(TI) const { return (TO)c; }
};

}  // namespace um

using namespace um;

extern "C" {

int32_t um_binary(int32_t op, const um_view* a, const um_view* b, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(b); UM_VIEW(out);
  int s = check_bcast(a, b, out);
  if (s != UM_OK) return s;
  switch (op) {
    case UM_ADD: return binary_typed(FAdd{}, a, b, out);
    case UM_SUB: return binary_typed(FSub{}, a, b, out);
    case UM_MUL: return binary_typed(FMul{}, a, b, out);
    case UM_DIV: return binary_typed(FDiv{}, a, b, out);
    case UM_MAXIMUM: return binary_typed(FMaximum{}, a, b, out);
    case UM_MINIMUM: return binary_typed(FMinimum{}, a, b, out);
    case UM_POW: return binary_typed(FPow{}, a, b, out);
  }
  return fail(UM_ERR_ARG, "unknown binary op %d", op);
}

int32_t um_binary_scalar(int32_t op, const um_view* a, double s, int32_t scalar_left, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(out);
  UM_REQUIRE(a->rows == out->rows && a->cols == out->cols, "scalar op: output shape mismatch");
  switch (op) {
    case UM_ADD: return scalar_typed(FAdd{}, a, s, scalar_left, out);
    case UM_SUB: return scalar_typed(FSub{}, a, s, scalar_left, out);
    case UM_MUL: return scalar_typed(FMul{}, a, s, scalar_left, out);
    case UM_DIV: return scalar_typed(FDiv{}, a, s, scalar_left, out);
    case UM_MAXIMUM: return scalar_typed(FMaximum{}, a, s, scalar_left, out);
    case UM_MINIMUM: return scalar_typed(FMinimum{}, a, s, scalar_left, out);
    case UM_POW: return scalar_typed(FPow{}, a, s, scalar_left, out);
  }
  return fail(UM_ERR_ARG, "unknown binary op %d", op);
}

int32_t um_unary(int32_t op, const um_view* a, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(out);
  UM_REQUIRE(a->rows == out->rows && a->cols == out->cols, "unary op: output shape mismatch");
  switch (op) {
    case UM_NEG: return unary_typed(UNeg{}, a, out);
    case UM_ABS: return unary_typed(UAbs{}, a, out);
    case UM_EXP: return unary_typed(UExp{}, a, out);
    case UM_SQRT: return unary_typed(USqrt{}, a, out);
    case UM_LOG: return unary_typed(ULog{}, a, out);
    case UM_SIGMOID: return unary_typed(USigmoid{}, a, out);
    case UM_RELU:
    case UM_RELU_GENERIC: return unary_typed(URelu{}, a, out);
    case UM_TANH: return unary_typed(UTanh{}, a, out);
    case UM_SQUARE: return unary_typed(USquare{}, a, out);
  }
  return fail(UM_ERR_ARG, "unknown unary op %d", op);
}

int32_t um_copy(const um_view* src, const um_view* dst) {
  UM_CTX();
  UM_VIEW(src); UM_VIEW(dst);
  UM_REQUIRE(src->rows == dst->rows && src->cols == dst->cols, "um_copy: shape mismatch (%lldx%lld -> %lldx%lld)",
             (i64)src->rows, (i64)src->cols, (i64)dst->rows, (i64)dst->cols);
  int a = src->dtype, o = dst->dtype;
  UCast f;
  if (a == UM_F64 && o == UM_F64) return launch_unary<UCast, double, double>(f, src, dst);
  if (a == UM_F32 && o == UM_F32) return launch_unary<UCast, float, float>(f, src, dst);
  if (a == UM_F32 && o == UM_F64) return launch_unary<UCast, float, double>(f, src, dst);
  if (a == UM_F64 && o == UM_F32) return launch_unary<UCast, double, float>(f, src, dst);
  if (a == UM_BOOL && o == UM_BOOL) return launch_unary<UCast, uint8_t, uint8_t>(f, src, dst);
  if (a == UM_I32 && o == UM_I32) return launch_unary<UCast, int32_t, int32_t>(f, src, dst);
  if (a == UM_BOOL && o == UM_F64) return launch_unary<UCast, uint8_t, double>(f, src, dst);
  if (a == UM_I32 && o == UM_F64) return launch_unary<UCast, int32_t, double>(f, src, dst);
  return fail(UM_ERR_ARG, "um_copy: unsupported conversion %d -> %d", a, o);
}

int32_t um_fill(const um_view* dst, double value) {
  UM_CTX();
  UM_VIEW(dst);
  UConst f{value};
  switch (dst->dtype) {
    case UM_F64: return launch_unary<UConst, double, double>(f, dst, dst);
    case UM_F32: return launch_unary<UConst, float, float>(f, dst, dst);
    case UM_BOOL: return launch_unary<UConst, uint8_t, uint8_t>(f, dst, dst);
    case UM_I32: return launch_unary<UConst, int32_t, int32_t>(f, dst, dst);
  }
  return fail(UM_ERR_ARG, "bad dtype");
}

int32_t um_eye(const um_view* dst) {
  UM_CTX();
  UM_VIEW(dst);
  int s = um_fill(dst, 0.0);
  if (s != UM_OK) return s;
  i64 n = dst->rows < dst->cols ? dst->rows : dst->cols;
  if (n == 0) return UM_OK;
  unsigned g = (unsigned)((n + 255) / 256);
  if (dst->dtype == UM_F64) k_eye<double><<<g, 256, 0, ctx()->stream>>>(out_of<double>(dst), n);
  else if (dst->dtype == UM_F32) k_eye<float><<<g, 256, 0, ctx()->stream>>>(out_of<float>(dst), n);
  else return fail(UM_ERR_ARG, "um_eye: f64/f32 only");
  UM_LAUNCHED();
  return UM_OK;
}

static int random_impl(const um_view* dst, uint64_t seed, bool normal) {
  UM_CTX();
  UM_VIEW(dst);
  i64 n = dst->rows * dst->cols;
  if (n == 0) return UM_OK;
  i64 pairs = (n + 1) / 2;
  i64 blocks = (pairs + 255) / 256;
  i64 cap = (i64)ctx()->sm_count * 32;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = ctx()->stream;
  if (dst->dtype == UM_F64) {
    if (normal) k_random<double, true><<<(unsigned)blocks, 256, 0, st>>>(out_of<double>(dst), dst->rows, dst->cols, seed);
    else k_random<double, false><<<(unsigned)blocks, 256, 0, st>>>(out_of<double>(dst), dst->rows, dst->cols, seed);
  } else if (dst->dtype == UM_F32) {
    if (normal) k_random<float, true><<<(unsigned)blocks, 256, 0, st>>>(out_of<float>(dst), dst->rows, dst->cols, seed);
    else k_random<float, false><<<(unsigned)blocks, 256, 0, st>>>(out_of<float>(dst), dst->rows, dst->cols, seed);
  } else {
    return fail(UM_ERR_ARG, "random fill: f64/f32 only");
  }
  UM_LAUNCHED();
  return UM_OK;
}
int32_t um_randn(const um_view* dst, uint64_t seed) { return random_impl(dst, seed, true); }
int32_t um_rand(const um_view* dst, uint64_t seed) { return random_impl(dst, seed, false); }

int32_t um_compare_scalar(int32_t op, const um_view* a, double s, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(out);
  UM_REQUIRE(a->rows == out->rows && a->cols == out->cols, "compare: output shape mismatch");
  // Float operand: the reference widens both sides to Double (Mat.scala:2576-2578); the
  // scalar arrives as a Float there, so round it through float first.
  if (a->dtype == UM_F32) s = (double)(float)s;
  switch (op) {
    case UM_GT: return tobool_typed(UCmpScalar<UM_GT>{s}, a, out);
    case UM_LT: return tobool_typed(UCmpScalar<UM_LT>{s}, a, out);
    case UM_GTE: return tobool_typed(UCmpScalar<UM_GTE>{s}, a, out);
    case UM_LTE: return tobool_typed(UCmpScalar<UM_LTE>{s}, a, out);
    case UM_EQ: return tobool_typed(UCmpScalar<UM_EQ>{s}, a, out);
    case UM_NE: return tobool_typed(UCmpScalar<UM_NE>{s}, a, out);
  }
  return fail(UM_ERR_ARG, "unknown compare op %d", op);
}

int32_t um_compare(int32_t op, const um_view* a, const um_view* b, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(b); UM_VIEW(out);
  int s = check_bcast(a, b, out);
  if (s != UM_OK) return s;
  switch (op) {
    case UM_GT: return cmp_typed(FCmp<UM_GT>{}, a, b, out);
    case UM_LT: return cmp_typed(FCmp<UM_LT>{}, a, b, out);
    case UM_GTE: return cmp_typed(FCmp<UM_GTE>{}, a, b, out);
    case UM_LTE: return cmp_typed(FCmp<UM_LTE>{}, a, b, out);
    case UM_EQ: return cmp_typed(FCmp<UM_EQ>{}, a, b, out);
    case UM_NE: return cmp_typed(FCmp<UM_NE>{}, a, b, out);
  }
  return fail(UM_ERR_ARG, "unknown compare op %d", op);
}

int32_t um_classify(int32_t op, const um_view* a, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(out);
  UM_REQUIRE(a->rows == out->rows && a->cols == out->cols, "classify: output shape mismatch");
  switch (op) {
    case UM_ISNAN: return tobool_typed(UClass<UM_ISNAN>{}, a, out);
    case UM_ISINF: return tobool_typed(UClass<UM_ISINF>{}, a, out);
    case UM_ISFINITE: return tobool_typed(UClass<UM_ISFINITE>{}, a, out);
  }
  return fail(UM_ERR_ARG, "unknown classify op %d", op);
}

int32_t um_mask_logic(int32_t op, const um_view* a, const um_view* b, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(out);
  UM_REQUIRE(a->dtype == UM_BOOL && out->dtype == UM_BOOL, "mask logic needs bool operands");
  if (op == UM_NOT) {
    UM_REQUIRE(a->rows == out->rows && a->cols == out->cols, "mask not: shape mismatch");
    return launch_unary<UNot, uint8_t, uint8_t>(UNot{}, a, out);
  }
  UM_VIEW(b);
  UM_REQUIRE(b->dtype == UM_BOOL, "mask logic needs bool operands");
  // && / || on Mat[Boolean] require equal shapes (Mat.scala:5157-5170)
  UM_REQUIRE(a->rows == b->rows && a->cols == b->cols && a->rows == out->rows && a->cols == out->cols,
             "mask logic: shape mismatch");
  switch (op) {
    case UM_AND: return launch_binary<FLogic<UM_AND>, uint8_t, uint8_t, uint8_t>({}, a, b, out);
    case UM_OR: return launch_binary<FLogic<UM_OR>, uint8_t, uint8_t, uint8_t>({}, a, b, out);
    case UM_XOR: return launch_binary<FLogic<UM_XOR>, uint8_t, uint8_t, uint8_t>({}, a, b, out);
  }
  return fail(UM_ERR_ARG, "unknown logic op %d", op);
}

}  // extern "C"
