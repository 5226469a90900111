// small_linalg.cuh -- sequential small dense solves, run by ONE thread (n <= 17).
// Mirrors `luDecomposeD` (Mat.scala:970-1001) + the per-column substitutions of `inverse`
// (Mat.scala:2979-3006) operation for operation, with multiply and subtract rounded
// separately (no FMA contraction), so the result is bit-identical to the JVM's.
#pragma once
#include "common.cuh"

namespace um {

constexpr int SL_MAX = 17;

// a: n x n row-major (overwritten with LU), inv: n x n out.  returns false on an exact-zero pivot.
__device__ inline bool lu_inverse_seq(double* a, int n, double* inv) {
  int piv[SL_MAX];
  for (int i = 0; i < n; ++i) piv[i] = i;
  for (int i = 0; i < n; ++i) {
    int max_row = i;
    double max_abs = fabs(a[i * n + i]);
    for (int k = i + 1; k < n; ++k) {
      double v = fabs(a[k * n + i]);
      if (v > max_abs) { max_abs = v; max_row = k; }
    }
    if (max_row != i) {
      for (int c = 0; c < n; ++c) { double t = a[i * n + c]; a[i * n + c] = a[max_row * n + c]; a[max_row * n + c] = t; }
      int tp = piv[i]; piv[i] = piv[max_row]; piv[max_row] = tp;
    }
    double pivot = a[i * n + i];
    if (pivot == 0.0) return false;
    for (int r = i + 1; r < n; ++r) {
      double f = __ddiv_rn(a[r * n + i], pivot);
      a[r * n + i] = f;
      for (int c = i + 1; c < n; ++c) a[r * n + c] = __dsub_rn(a[r * n + c], __dmul_rn(f, a[i * n + c]));
    }
  }
  double x[SL_MAX];
  for (int col = 0; col < n; ++col) {
    for (int i = 0; i < n; ++i) x[i] = piv[i] == col ? 1.0 : 0.0;
    for (int i = 1; i < n; ++i)
      for (int k = 0; k < i; ++k) x[i] = __dsub_rn(x[i], __dmul_rn(a[i * n + k], x[k]));
    for (int i = n - 1; i >= 0; --i) {
      for (int k = i + 1; k < n; ++k) x[i] = __dsub_rn(x[i], __dmul_rn(a[i * n + k], x[k]));
      x[i] = __ddiv_rn(x[i], a[i * n + i]);
    }
    for (int r = 0; r < n; ++r) inv[r * n + col] = x[r];
  }
  return true;
}

// The same inverse by one WARP on matrices in shared memory (`a` n x n is overwritten with LU, `x` is an
// n x n scratch, `piv` n ints): rows of an elimination step and columns of the substitutions are
// independent, so lanes take one each; every element sees exactly the operations, in the order, of
// lu_inverse_seq -- the result is bit-identical, ~8x sooner (the serial version is a chain of ~2000
// dependent local-memory accesses).  All 32 lanes must call; returns the same flag on every lane.
__device__ inline bool lu_inverse_warp(double* a, int n, double* inv, double* x, int* piv, int lane) {
  if (lane < n) piv[lane] = lane;
  __syncwarp();
  for (int i = 0; i < n; ++i) {
    int max_row = i;
    double max_abs = fabs(a[i * n + i]);
    for (int k = i + 1; k < n; ++k) {   // every lane scans the same column: first strictly-larger wins
      double v = fabs(a[k * n + i]);
      if (v > max_abs) { max_abs = v; max_row = k; }
    }
    __syncwarp();
    if (max_row != i) {
      if (lane < n) { double t = a[i * n + lane]; a[i * n + lane] = a[max_row * n + lane]; a[max_row * n + lane] = t; }
      if (lane == 0) { int tp = piv[i]; piv[i] = piv[max_row]; piv[max_row] = tp; }
      __syncwarp();
    }
    const double pivot = a[i * n + i];
    if (pivot == 0.0) return false;
    const int r = i + 1 + lane;
    if (r < n) {
      const double f = __ddiv_rn(a[r * n + i], pivot);
      a[r * n + i] = f;
      for (int c = i + 1; c < n; ++c) a[r * n + c] = __dsub_rn(a[r * n + c], __dmul_rn(f, a[i * n + c]));
    }
    __syncwarp();
  }
  if (lane < n) {
    const int col = lane;
    double* xc = x + col * n;
    for (int i = 0; i < n; ++i) xc[i] = piv[i] == col ? 1.0 : 0.0;
    for (int i = 1; i < n; ++i)
      for (int k = 0; k < i; ++k) xc[i] = __dsub_rn(xc[i], __dmul_rn(a[i * n + k], xc[k]));
    for (int i = n - 1; i >= 0; --i) {
      for (int k = i + 1; k < n; ++k) xc[i] = __dsub_rn(xc[i], __dmul_rn(a[i * n + k], xc[k]));
      xc[i] = __ddiv_rn(xc[i], a[i * n + i]);
    }
    for (int r2 = 0; r2 < n; ++r2) inv[r2 * n + col] = xc[r2];
  }
  __syncwarp();
  return true;
}

// y = M (n x n) * v (n)
__device__ inline void matvec_seq(const double* M, const double* v, int n, double* y) {
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += M[i * n + k] * v[k];
    y[i] = s;
  }
}

}  // namespace um
