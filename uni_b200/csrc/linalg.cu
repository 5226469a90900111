// linalg.cu -- `inverse` / `solve` (LU with partial pivoting) for one matrix or a batch.
// One CTA per matrix; the LU lives in shared memory up to n = 64 and in a global workspace
// beyond.  Same pivot rule (first strictly-larger |a| wins) and the same operation order per
// element as luDecomposeD (Mat.scala:970-1001), multiply/subtract rounded separately, so the
// inverse is bit-identical to the reference's for Mat[Double].
#include "common.cuh"

namespace um {

constexpr int LA_THREADS = 256;
constexpr int LA_SMEM_N = 64;

struct LaArgs {
  const double* a; i64 a_rs, a_cs, a_stride;   // input matrices
  const double* b; i64 b_rs, b_cs, b_stride;   // rhs (solve) or null (inverse)
  double* out; i64 o_rs, o_cs, o_stride;       // n x nrhs
  double* work;                                // batch * n * n doubles when n > LA_SMEM_N
  int* piv_work;                               // batch * n ints
  int* flag;                                   // set to 1 on an exact-zero pivot
  int n, nrhs;
};

__global__ void __launch_bounds__(LA_THREADS) k_lu_solve(LaArgs g) {
  extern __shared__ double sm_lu[];
  __shared__ int s_piv[LA_SMEM_N];
  __shared__ double s_best[LA_THREADS / 32];
  __shared__ int s_besti[LA_THREADS / 32];
  __shared__ int s_maxrow;
  __shared__ int s_singular;
  const int n = g.n, tid = threadIdx.x;
  const i64 bi = blockIdx.x;
  double* lu = n <= LA_SMEM_N ? sm_lu : g.work + bi * (i64)n * n;
  int* piv = n <= LA_SMEM_N ? s_piv : g.piv_work + bi * n;
  const double* A = g.a + bi * g.a_stride;
  for (int e = tid; e < n * n; e += LA_THREADS) lu[e] = A[(i64)(e / n) * g.a_rs + (i64)(e % n) * g.a_cs];
  for (int e = tid; e < n; e += LA_THREADS) piv[e] = e;
  if (tid == 0) s_singular = 0;
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    // pivot search: max |lu[k][i]|, k >= i, first occurrence on ties
    double best = -1.0;
    int besti = n;
    for (int k = i + tid; k < n; k += LA_THREADS) {
      double v = fabs(lu[k * n + i]);
      if (v != v) v = -0.5;  // `v > maxAbs` is false for NaN: a NaN never becomes the pivot row
      if (v > best) { best = v; besti = k; }
    }
    for (int d = 16; d > 0; d >>= 1) {
      double ob = __shfl_down_sync(0xffffffffu, best, d);
      int oi = __shfl_down_sync(0xffffffffu, besti, d);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_besti[tid >> 5] = besti; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < LA_THREADS / 32; ++w)
        if (s_best[w] > best || (s_best[w] == best && s_besti[w] < besti)) { best = s_best[w]; besti = s_besti[w]; }
      // the reference seeds with row i and replaces only on strictly greater
      double vi = fabs(lu[i * n + i]);
      int mr = (best > vi) ? besti : i;
      if (!(vi == vi)) mr = i;  // |a_ii| NaN: every `v > maxAbs` is false, the row stays
      s_maxrow = mr;
    }
    __syncthreads();
    const int mr = s_maxrow;
    if (mr != i) {
      for (int c = tid; c < n; c += LA_THREADS) { double t = lu[i * n + c]; lu[i * n + c] = lu[mr * n + c]; lu[mr * n + c] = t; }
      if (tid == 0) { int tp = piv[i]; piv[i] = piv[mr]; piv[mr] = tp; }
    }
    __syncthreads();
    const double pivot = lu[i * n + i];
    if (pivot == 0.0) {
      if (tid == 0) { s_singular = 1; atomicExch(g.flag, 1); }
      break;
    }
    for (int r = i + 1 + tid; r < n; r += LA_THREADS) lu[r * n + i] = __ddiv_rn(lu[r * n + i], pivot);
    __syncthreads();
    const int w = n - i - 1;
    for (int e = tid; e < w * w; e += LA_THREADS) {
      int r = i + 1 + e / w, c = i + 1 + e % w;
      lu[r * n + c] = __dsub_rn(lu[r * n + c], __dmul_rn(lu[r * n + i], lu[i * n + c]));
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_singular) return;
  // substitutions: one thread per right-hand-side column, x kept in the output column
  double* O = g.out + bi * g.o_stride;
  const double* B = g.b ? g.b + bi * g.b_stride : nullptr;
  for (int col = tid; col < g.nrhs; col += LA_THREADS) {
    for (int i = 0; i < n; ++i) {
      double xi = B ? B[(i64)piv[i] * g.b_rs + (i64)col * g.b_cs] : (piv[i] == col ? 1.0 : 0.0);
      for (int k = 0; k < i; ++k) xi = __dsub_rn(xi, __dmul_rn(lu[i * n + k], O[(i64)k * g.o_rs + (i64)col * g.o_cs]));
      O[(i64)i * g.o_rs + (i64)col * g.o_cs] = xi;
    }
    for (int i = n - 1; i >= 0; --i) {
      double xi = O[(i64)i * g.o_rs + (i64)col * g.o_cs];
      for (int k = i + 1; k < n; ++k) xi = __dsub_rn(xi, __dmul_rn(lu[i * n + k], O[(i64)k * g.o_rs + (i64)col * g.o_cs]));
      O[(i64)i * g.o_rs + (i64)col * g.o_cs] = __ddiv_rn(xi, lu[i * n + i]);
    }
  }
}

static int run_lu(const um_view* a, i64 a_stride, const um_view* b, i64 b_stride, const um_view* out, i64 o_stride,
                  i64 batch) {
  const i64 n = a->rows;
  UM_REQUIRE(a->dtype == UM_F64 && out->dtype == UM_F64 && (!b || b->dtype == UM_F64), "inverse/solve: f64 only on device");
  if (n == 0 || batch == 0) return UM_OK;
  UM_REQUIRE(n <= 4096, "inverse/solve: n = %lld too large for the single-CTA LU", n);
  LaArgs g;
  g.a = static_cast<const double*>(a->data) + a->offset; g.a_rs = a->rs; g.a_cs = a->cs; g.a_stride = a_stride;
  g.b = b ? static_cast<const double*>(b->data) + b->offset : nullptr;
  g.b_rs = b ? b->rs : 0; g.b_cs = b ? b->cs : 0; g.b_stride = b_stride;
  g.out = static_cast<double*>(out->data) + out->offset; g.o_rs = out->rs; g.o_cs = out->cs; g.o_stride = o_stride;
  g.n = (int)n;
  g.nrhs = (int)out->cols;
  size_t work_bytes = n > LA_SMEM_N ? (size_t)batch * n * n * 8 : 0;
  size_t piv_bytes = n > LA_SMEM_N ? (size_t)batch * n * 4 : 0;
  work_bytes = (work_bytes + 15) & ~size_t(15);
  piv_bytes = (piv_bytes + 15) & ~size_t(15);
  char* s = static_cast<char*>(scratch(work_bytes + piv_bytes + 16));
  if (!s) return UM_ERR_OOM;
  g.work = reinterpret_cast<double*>(s);
  g.piv_work = reinterpret_cast<int*>(s + work_bytes);
  g.flag = reinterpret_cast<int*>(s + work_bytes + piv_bytes);
  cudaStream_t st = ctx()->stream;
  UM_CUDA(cudaMemsetAsync(g.flag, 0, 4, st));
  size_t smem = n <= LA_SMEM_N ? (size_t)n * n * 8 : 0;
  k_lu_solve<<<(unsigned)batch, LA_THREADS, smem, st>>>(g);
  UM_LAUNCHED();
  int flag = 0;
  UM_CUDA(cudaMemcpyAsync(&flag, g.flag, 4, cudaMemcpyDeviceToHost, st));
  UM_CUDA(cudaStreamSynchronize(st));
  if (flag) return fail(UM_ERR_SINGULAR, "Matrix is singular or nearly singular");
  return UM_OK;
}

}  // namespace um

using namespace um;

extern "C" {

int32_t um_inverse(const um_view* a, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(out);
  UM_REQUIRE(a->rows == a->cols, "inverse requires square matrix, got (%lld, %lld)", (i64)a->rows, (i64)a->cols);
  UM_REQUIRE(out->rows == a->rows && out->cols == a->cols, "inverse: output shape mismatch");
  return run_lu(a, 0, nullptr, 0, out, 0, 1);
}

int32_t um_solve(const um_view* a, const um_view* b, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(b); UM_VIEW(out);
  UM_REQUIRE(a->rows == a->cols, "solve requires square matrix, got (%lld, %lld)", (i64)a->rows, (i64)a->cols);
  UM_REQUIRE(b->rows == a->rows, "bCol.rows %lld must match matrix rows %lld", (i64)b->rows, (i64)a->rows);
  UM_REQUIRE(out->rows == b->rows && out->cols == b->cols, "solve: output shape mismatch");
  return run_lu(a, 0, b, 0, out, 0, 1);
}

int32_t um_inverse_batched(const um_view* a0, int64_t stride_a, const um_view* out0, int64_t stride_out, int64_t batch) {
  UM_CTX();
  UM_VIEW(a0); UM_VIEW(out0);
  UM_REQUIRE(a0->rows == a0->cols, "inverse requires square matrix");
  UM_REQUIRE(out0->rows == a0->rows && out0->cols == a0->cols, "inverse: output shape mismatch");
  UM_REQUIRE(batch >= 0, "negative batch");
  return run_lu(a0, stride_a, nullptr, 0, out0, stride_out, batch);
}

}  // extern "C"
