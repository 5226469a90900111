// jvm_restated.cpp -- TEST / BENCH INFRASTRUCTURE, not product code.
//
// A C++ restatement of the reference's PURE-JVM Mat kernels (no BLAS), for the CPU baseline that bench.py prints
// beside every GPU figure ("pure-JVM path ... restated"): the image has no JVM, so the Scala
// cannot run; this file restates the same loops -- same blocking, same accumulator counts, same chunking of the
// ForkJoin work split -- in C++ (g++ -O2 -ffp-contract=off: HotSpot does not contract a*b+c into an FMA and does
// not auto-vectorise FP reductions).  Only bench.py's cpu legs and tests/test_jvm_restated.py load it.
//
//   jr_matmul_pure   MatmulPure.multiply          src/main/scala/uni/data/MatmulPure.scala:78-304 (pack4 + 4x4 microkernel,
//                                                 16-row blocks x column-panel ranges over the common pool)
//   jr_sum           Mat.sumD / sumRange          Mat.scala:1042-1085 (8 accumulators, <= 16 chunks, partials in order)
//   jr_col_sums      colSumsD via overLanes       Mat.scala:586-633, 836-845 (8 columns at a time, rows in order)
//   jr_row_sums      rowSumsD via overLanes       Mat.scala:635-647
//   jr_sub_rowvec    fastBinOp, 1 x N broadcast   Mat.scala:2013-2031, bcastRows :1031-1036
//   jr_add_same      fastBinOp, same shape        Mat.scala:2017-2021, fillD :1016-1023
//   jr_sigmoid       MatMathOps.sigmoid           MatMathOps.scala:97-122 (x >= 0 ? 1/(1+e^-x) : e^x/(1+e^x))
//   jr_softmax_rows  MatMathOps.softmax(axis=1)   MatMathOps.scala:156-191 (three scalar passes per lane, SINGLE-threaded)
//
// `threads` plays the role of ForkJoinPool.commonPool parallelism; work items are handed out round-robin.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace {

void run_items(int n, int threads, const std::function<void(int)>& item) {
  if (threads <= 1 || n <= 1) {
    for (int t = 0; t < n; ++t) item(t);
    return;
  }
  const int nt = std::min(threads, n);
  std::vector<std::thread> pool;
  pool.reserve(nt);
  for (int w = 0; w < nt; ++w)
    pool.emplace_back([=, &item] {
      for (int t = w; t < n; t += nt) item(t);
    });
  for (auto& th : pool) th.join();
}

constexpr int kBlockRows = 16;          // MatmulPure.scala:54
constexpr long kParOps = 262144;        // :61
constexpr long kMinItemOps = 131072;    // :62
constexpr int kMaxSumChunks = 16;       // Mat.scala:366
constexpr int kParallelThreshold = 4096;        // :356
constexpr long kLaneParallelThreshold = 1 << 19;  // :389
constexpr long kBcastParallelThreshold = 1 << 16; // :373

// MatmulPure.scala:120-126
int col_split(long ops, int row_blocks, int max_chunks, int pool) {
  if (ops < kParOps) return 1;
  const int want = (2 * pool + row_blocks - 1) / row_blocks;
  const long by_work = ops / ((long)row_blocks * kMinItemOps);
  return (int)std::max<long>(1, std::min<long>(max_chunks, std::min<long>(want, by_work)));
}

// 4x4 cells accumulated over the whole k range from 0.0, A rows read in place (micro44Direct, :200-216)
inline void micro44(const double* a0, long a_rs, const double* bp, double* out, long o0, long cb, int j, int w, long ca) {
  const double *a1 = a0 + a_rs, *a2 = a1 + a_rs, *a3 = a2 + a_rs;
  double c[4][4] = {};
  for (long k = 0; k < ca; ++k) {
    const double b0 = bp[k * 4], b1 = bp[k * 4 + 1], b2 = bp[k * 4 + 2], b3 = bp[k * 4 + 3];
    const double v0 = a0[k], v1 = a1[k], v2 = a2[k], v3 = a3[k];
    c[0][0] += v0 * b0; c[0][1] += v0 * b1; c[0][2] += v0 * b2; c[0][3] += v0 * b3;
    c[1][0] += v1 * b0; c[1][1] += v1 * b1; c[1][2] += v1 * b2; c[1][3] += v1 * b3;
    c[2][0] += v2 * b0; c[2][1] += v2 * b1; c[2][2] += v2 * b2; c[2][3] += v2 * b3;
    c[3][0] += v3 * b0; c[3][1] += v3 * b1; c[3][2] += v3 * b2; c[3][3] += v3 * b3;
  }
  for (int r = 0; r < 4; ++r)
    for (int x = 0; x < w; ++x) out[o0 + r * cb + j + x] = c[r][x];
}

// a row remainder (< 4 rows) against one panel (panelRows)
inline void panel_rows(const double* a, long a_rs, int rows, const double* bp, double* out, long i, long ca, long cb, int j, int w) {
  for (int r = 0; r < rows; ++r) {
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    const double* ar = a + (i + r) * a_rs;
    for (long k = 0; k < ca; ++k) {
      const double v = ar[k];
      c0 += v * bp[k * 4]; c1 += v * bp[k * 4 + 1]; c2 += v * bp[k * 4 + 2]; c3 += v * bp[k * 4 + 3];
    }
    const double cc[4] = {c0, c1, c2, c3};
    for (int x = 0; x < w; ++x) out[(i + r) * cb + j + x] = cc[x];
  }
}

double sum_range(const double* a, long from, long until) {   // Mat.scala:1042-1057
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
  long i = from;
  const long limit = until - 7;
  for (; i < limit; i += 8) {
    s0 += a[i]; s1 += a[i + 1]; s2 += a[i + 2]; s3 += a[i + 3];
    s4 += a[i + 4]; s5 += a[i + 5]; s6 += a[i + 6]; s7 += a[i + 7];
  }
  double s = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
  for (; i < until; ++i) s += a[i];
  return s;
}

}  // namespace

extern "C" {

// C (ra x cb) = A (ra x ca) * B (ca x cb), all contiguous row-major; the microkernel path of MatmulPure.multiply
// (ca >= 16 and cb >= 4; narrower products take its streaming path, which the benches do not exercise)
int jr_matmul_pure(const double* a, const double* b, double* out, long ra, long ca, long cb, int threads) {
  if (ra == 0 || cb == 0) return 0;
  const long ops = ra * ca * cb;
  const int row_blocks = (int)((ra + kBlockRows - 1) / kBlockRows);
  const int njb = (int)((cb + 3) / 4);
  std::vector<double> bp((size_t)njb * ca * 4, 0.0);   // pack4 (:140-155): bp[jb*ca*4 + k*4 + x] = b(k, jb*4 + x)
  for (int jb = 0; jb < njb; ++jb) {
    const long base = (long)jb * ca * 4, j0 = (long)jb * 4;
    const int w = (int)std::min<long>(4, cb - j0);
    for (long k = 0; k < ca; ++k)
      for (int x = 0; x < w; ++x) bp[base + k * 4 + x] = b[k * cb + j0 + x];
  }
  const int csplit = col_split(ops, row_blocks, njb, threads);
  const int panels_per = (njb + csplit - 1) / csplit;
  const int items = row_blocks * csplit;
  auto item = [&](int t) {
    const int rb = t / csplit, cc = t % csplit;
    const int jb_s = cc * panels_per, jb_e = std::min(jb_s + panels_per, njb);
    if (jb_s >= jb_e) return;
    const long i_s = (long)rb * kBlockRows, i_e = std::min<long>((long)(rb + 1) * kBlockRows, ra);
    long i = i_s;
    for (; i + 3 < i_e; i += 4)
      for (int jb = jb_s; jb < jb_e; ++jb) {
        const int j = jb * 4, w = (int)std::min<long>(4, cb - j);
        micro44(a + i * ca, ca, bp.data() + (size_t)jb * ca * 4, out, i * cb, cb, j, w, ca);
      }
    if (i < i_e)
      for (int jb = jb_s; jb < jb_e; ++jb) {
        const int j = jb * 4, w = (int)std::min<long>(4, cb - j);
        panel_rows(a, ca, (int)(i_e - i), bp.data() + (size_t)jb * ca * 4, out, i, ca, cb, j, w);
      }
  };
  run_items(items, ops < kParOps ? 1 : threads, item);
  return 0;
}

double jr_sum(const double* a, long n, int threads) {   // sumD, Mat.scala:1066-1085
  if (n < kParallelThreshold) return sum_range(a, 0, n);
  const int chunks = (int)std::min<long>(kMaxSumChunks, std::max<long>(n / kParallelThreshold, 1));
  const long step = (n + chunks - 1) / chunks;
  std::vector<double> partials(chunks, 0.0);
  run_items(chunks, threads, [&](int c) {
    const long from = c * step, until = std::min(from + step, n);
    if (from < until) partials[c] = sum_range(a, from, until);
  });
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += partials[c];
  return s;
}

// overLanes (Mat.scala:836-845): lanes split into <= 16 chunks of at least 8 lanes above 2^19 elements
static void over_lanes(long lanes, long total, int threads, const std::function<void(long, long)>& body) {
  if (total < kLaneParallelThreshold || lanes < 8) { body(0, lanes); return; }
  const int chunks = (int)std::min<long>(kMaxSumChunks, std::max<long>(lanes / 8, 1));
  const long step = (lanes + chunks - 1) / chunks;
  run_items(chunks, threads, [&](int c) {
    const long from = c * step, until = std::min(from + step, lanes);
    if (from < until) body(from, until);
  });
}

int jr_col_sums(const double* a, long rows, long cols, double* result, int threads) {   // colSumsD, cs == 1
  std::memset(result, 0, sizeof(double) * cols);
  over_lanes(cols, rows * cols, threads, [&](long c0, long c1) {
    long jb = c0;
    for (; jb + 8 <= c1; jb += 8) {
      double s0 = result[jb], s1 = result[jb + 1], s2 = result[jb + 2], s3 = result[jb + 3];
      double s4 = result[jb + 4], s5 = result[jb + 5], s6 = result[jb + 6], s7 = result[jb + 7];
      for (long i = 0; i < rows; ++i) {
        const double* p = a + i * cols + jb;
        s0 += p[0]; s1 += p[1]; s2 += p[2]; s3 += p[3]; s4 += p[4]; s5 += p[5]; s6 += p[6]; s7 += p[7];
      }
      result[jb] = s0; result[jb + 1] = s1; result[jb + 2] = s2; result[jb + 3] = s3;
      result[jb + 4] = s4; result[jb + 5] = s5; result[jb + 6] = s6; result[jb + 7] = s7;
    }
    for (; jb < c1; ++jb) {
      double s = result[jb];
      for (long i = 0; i < rows; ++i) s += a[i * cols + jb];
      result[jb] = s;
    }
  });
  return 0;
}

int jr_row_sums(const double* a, long rows, long cols, double* result, int threads) {   // rowSumsD, cs == 1
  over_lanes(rows, rows * cols, threads, [&](long r0, long r1) {
    for (long i = r0; i < r1; ++i) {
      const double* p = a + i * cols;
      double s = 0.0;
      for (long j = 0; j < cols; ++j) s += p[j];
      result[i] = s;
    }
  });
  return 0;
}

int jr_sub_rowvec(const double* a, const double* v, double* out, long rows, long cols, int threads) {
  const int par = rows * cols >= kBcastParallelThreshold ? threads : 1;
  run_items((int)rows, par, [&](int r) {
    const long base = (long)r * cols;
    for (long c = 0; c < cols; ++c) out[base + c] = a[base + c] - v[c];
  });
  return 0;
}

// fillD above the threshold is Arrays.parallelSetAll: the index range split evenly over the pool
static void fill_d(long n, int threads, const std::function<void(long, long)>& body) {
  if (n < kParallelThreshold || threads <= 1) { body(0, n); return; }
  const long step = (n + threads - 1) / threads;
  run_items(threads, threads, [&](int t) {
    const long from = t * step, until = std::min(from + step, n);
    if (from < until) body(from, until);
  });
}

int jr_add_same(const double* a, const double* b, double* out, long n, int threads) {
  fill_d(n, threads, [&](long f, long u) { for (long i = f; i < u; ++i) out[i] = a[i] + b[i]; });
  return 0;
}

int jr_sigmoid(const double* a, double* out, long n, int threads) {
  fill_d(n, threads, [&](long f, long u) {
    for (long i = f; i < u; ++i) {
      const double x = a[i];
      if (x >= 0) out[i] = 1.0 / (1.0 + std::exp(-x));
      else { const double e = std::exp(x); out[i] = e / (1.0 + e); }
    }
  });
  return 0;
}

// softmax(axis = 1): per row max, e = exp(x - max), left-fold sum, divide -- one thread, as the reference's loop is
int jr_softmax_rows(const double* a, double* out, long rows, long cols) {
  for (long r = 0; r < rows; ++r) {
    const double* p = a + r * cols;
    double* o = out + r * cols;
    double mx = p[0];
    for (long j = 1; j < cols; ++j) if (p[j] > mx) mx = p[j];
    double s = 0.0;
    for (long j = 0; j < cols; ++j) { const double e = std::exp(p[j] - mx); o[j] = e; s += e; }
    for (long j = 0; j < cols; ++j) o[j] = o[j] / s;
  }
  return 0;
}

}  // extern "C"
