/* unimat.h -- C ABI of libunimat.so: a B200 (sm_100a) device backend for the data-parallel
 * matrix hot path of philwalk/uni (`MatD`/`MatF` operators and the 3PRF fits).
 *
 * The reference has NO native seam for this path (its `Mat[T]` operators are Scala
 * extension methods over a JVM array, src/main/scala/uni/data/Mat.scala:18,125-133); the
 * only native calls it makes today are JNI BLAS (`netlib.dgemm` Mat.scala:3317,
 * `cblas_dgemm` Mat.scala:3336).  This header is therefore the interface a Scala 3 Panama
 * (java.lang.foreign) binding module downcalls into; each entry point names the reference
 * operator (file:line under /root/reference/src/main/scala/uni/) it stands in for.
 * INTEGRATION.md shows the Scala-side binding.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no C++/torch types.
 *   - every function returns an `um_status`; `um_last_error()` gives the thread-local text.
 *     Status -> reference exception: UM_ERR_ARG   -> IllegalArgumentException (`require`)
 *                                    UM_ERR_SINGULAR -> ArithmeticException (Mat.scala:990)
 *                                    UM_ERR_EMPTY -> UnsupportedOperationException (Mat.scala:2216)
 *   - device memory is owned by the CALLER (the host `Mat` object): `um_malloc`/`um_free`.
 *     Operators never allocate results: the caller passes an output view of the result
 *     shape (fresh contiguous buffer, or the receiver itself for the in-place family).
 *   - a matrix is a `um_view`: device pointer + the reference's stride descriptor
 *     `(offset, rows, cols, rs, cs, transposed)` (Mat.scala:125-133) with 64-bit sizes, so
 *     transposes / slices / stride-0 broadcasts stay zero-copy.
 *   - all work is issued on the calling host thread's stream (one stream per host thread
 *     per device); functions returning host scalars synchronise that stream.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *     with UM_ERR_CUDA.
 */
#ifndef UNIMAT_H
#define UNIMAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UM_VERSION_MAJOR 0
#define UM_VERSION_MINOR 1

typedef enum um_status {
  UM_OK = 0,
  UM_ERR_ARG = 1,
  UM_ERR_SINGULAR = 2,
  UM_ERR_EMPTY = 3,
  UM_ERR_CUDA = 4,
  UM_ERR_NCCL = 5,
  UM_ERR_OOM = 6
} um_status;

typedef enum um_dtype { UM_F64 = 0, UM_F32 = 1, UM_BOOL = 2, UM_I32 = 3 } um_dtype;

/* MatData(_tdata, _rows, _cols, _transposed, _offset, _rs, _cs): Mat.scala:125-133 */
typedef struct um_view {
  void* data;        /* device pointer to element 0 of the backing buffer            */
  int64_t offset;    /* element index of (0,0)                                        */
  int64_t rows, cols;
  int64_t rs, cs;    /* element strides; 0 = broadcast                                */
  int32_t dtype;     /* um_dtype                                                      */
  int32_t transposed; /* the reference's flag; only the layout guard reads it          */
} um_view;

/* ---- lifecycle / memory ------------------------------------------------------------ */
int32_t um_version(void);
const char* um_last_error(void);
int32_t um_device_count(int32_t* out);
/* Binds the calling host thread to `device` (own stream, own scratch).  One device per process: the first um_init
 * fixes it, a later um_init with another device returns UM_ERR_ARG, and host threads that never call um_init (a JVM
 * Cleaner or ForkJoin worker) inherit it on their first library call. */
int32_t um_init(int32_t device);
int32_t um_shutdown(void);
int32_t um_sync(void);                      /* synchronise the calling thread's stream   */
int32_t um_device_props(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* total_mem);
int32_t um_malloc(size_t bytes, void** out);      /* stream-ordered pool (cudaMallocAsync)   */
/* May be called from ANY host thread (JVM Cleaner, Python GC): the block remembers the device and stream it was
 * allocated on and the free is queued there, after everything the allocating thread launched on it.  Never opens a
 * context; a pointer that is not a live um_malloc block returns UM_ERR_ARG. */
int32_t um_free(void* ptr);
int32_t um_malloc_host(size_t bytes, void** out); /* pinned host memory for staging          */
int32_t um_free_host(void* ptr);
int32_t um_h2d(void* dst, const void* src, size_t bytes);
int32_t um_d2h(void* dst, const void* src, size_t bytes);  /* synchronises                 */
int32_t um_d2d(void* dst, const void* src, size_t bytes);
int32_t um_memset(void* dst, int32_t byte, size_t bytes);
/* strided 2-D host<->device copy (column shards of a row-major host panel) */
int32_t um_h2d_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t height);
int32_t um_launch_count(int64_t* out);      /* kernels launched by this library so far     */
/* event timing on the calling thread's stream (bench harness) */
int32_t um_timer_start(void);
int32_t um_timer_stop(float* ms_out);       /* synchronises                                */

/* per-kernel device timing of the 3PRF pipeline (CUDA events around each launch while enabled):
 * slot 0 colstats, 1 colfinal, 2 project, 3 gather, 4 finish, 5 outputs, 6 allreduce, 7 prep */
int32_t um_prof_enable(int32_t on);
int32_t um_prof_reset(void);
int32_t um_prof_read(int32_t slot, double* total_ms, int64_t* count);   /* synchronises */

/* ---- views: Mat.scala:170-199 (create / transposeView), :1865-1880 (layout guard),
 *      :1905-1927 (slice), :1933-1953 (broadcastTo).  Host-side descriptor arithmetic. */
int32_t um_view_init(um_view* v, void* data, int64_t rows, int64_t cols, int32_t dtype);
int32_t um_view_is_weird(const um_view* v, int32_t* out);       /* isWeirdLayout            */
int32_t um_view_is_contiguous(const um_view* v, int32_t* out);  /* rs == cols && cs == 1     */
int32_t um_view_transpose(const um_view* v, um_view* out);
/* `needs_copy` = 1 when Internal.create would materialise the slice (fragmented layout): the
 * caller must then allocate rows*cols elements and `um_copy` the returned view into it. */
int32_t um_view_slice(const um_view* v, int64_t r0, int64_t r1, int64_t c0, int64_t c1, um_view* out,
                      int32_t* needs_copy);
int32_t um_view_broadcast(const um_view* v, int64_t rows, int64_t cols, um_view* out);
/* result shape of `a op b` under the reference's broadcasting rule (Mat.scala:1989-2005) */
int32_t um_broadcast_shape(const um_view* a, const um_view* b, int64_t* rows, int64_t* cols);

/* ---- copies / fills ---------------------------------------------------------------- */
/* matCopy (Mat.scala:1886-1899) + dtype conversion (f32<->f64, bool/i32 -> f64): dst(r,c) = src(r,c) */
int32_t um_copy(const um_view* src, const um_view* dst);
int32_t um_fill(const um_view* dst, double value);              /* MatD.full / zeros / ones  */
int32_t um_eye(const um_view* dst);                             /* MatD.eye                  */
/* synthetic N(0,1) / U[0,1) panels generated in HBM (bench inputs; counter-based, seed+index) */
int32_t um_randn(const um_view* dst, uint64_t seed);
int32_t um_rand(const um_view* dst, uint64_t seed);
int32_t um_get(const um_view* v, int64_t r, int64_t c, double* out);   /* m(r, c)            */
int32_t um_set(const um_view* v, int64_t r, int64_t c, double value);  /* m(r, c) = value    */

/* ---- NumPy-compatible PCG64 generator on the device -------------------------------
 * The reference's global generator (data/NumPyRNG.scala:22-160: PCG64 XSL-RR 128, SeedSequence seeding :584-770,
 * nextDouble :68-71, nextBoundedInt :82-110, ziggurat randn :124-160) behind MatD.setSeed / rand / randn / normal /
 * uniform / randint (data/Mat.scala:1461-1530).  The state lives in a caller-owned POD (the Scala side keeps it where
 * it keeps `globalRNG`); every fill consumes the stream exactly as the sequential generator would -- elements in
 * row-major order -- and leaves the state where NumPy's would be, so fills can be mixed and continued.  `dst` must be
 * a contiguous row-major matrix (the factories return fresh ones).  Values are bit-identical to
 * numpy.random.default_rng(seed) except ziggurat tail draws (|z| > 3.654, 2.7e-4 of them), which go through CUDA's
 * log1p and agree to 1 ulp.  um_rng_randn / um_rng_normal synchronise (the number of draws consumed is data-dependent). */
typedef struct um_rng {
  uint64_t state_hi, state_lo;   /* 128-bit LCG state                                         */
  uint64_t inc_hi, inc_lo;       /* 128-bit increment (stream)                                */
  int32_t has_uint32;            /* nextBoundedInt's cached high half (useUpper / cachedRaw)   */
  uint32_t uinteger;
} um_rng;
int32_t um_rng_seed(um_rng* g, int64_t seed);         /* new NumPyRNG(seed); seed < 0 -> UM_ERR_ARG (NumPyRNG.scala:770) */
int32_t um_rng_advance(um_rng* g, uint64_t draws);    /* skip `draws` 64-bit outputs (host arithmetic)                    */
int32_t um_rng_next_u64(um_rng* g, uint64_t* out);    /* nextLong (host arithmetic; NumPyRNG.scala:58-61)                  */
int32_t um_rng_rand(um_rng* g, const um_view* dst);                             /* MatD.rand     Mat.scala:1473 */
int32_t um_rng_uniform(um_rng* g, double low, double high, const um_view* dst);  /* MatD.uniform  Mat.scala:1510 */
int32_t um_rng_randn(um_rng* g, const um_view* dst);                            /* MatD.randn    Mat.scala:1480,1488 */
int32_t um_rng_normal(um_rng* g, double mean, double sd, const um_view* dst);    /* MatD.normal   Mat.scala:1500 */
int32_t um_rng_randint(um_rng* g, int32_t low, int32_t high, const um_view* dst_i32); /* MatD.randint Mat.scala:1525 */

/* ---- on-disk format -----------------------------------------------------------------
 * `path.loadMatD` / `loadSmartD` / `loadMatF` (ext/PathExts.scala:583-590 -> io/FileOps.scala:139-188): CSV text ->
 * matrix.  The text is parsed on the host (all cores) into a staging table, then uploaded with one copy; two steps
 * because the caller owns the result buffer.  Cell = `big(s).toDouble` (data/Big.scala:59-64: trim, drop ',' and '$',
 * trailing '%' divides by 100, java.math.BigDecimal grammar, anything else NaN); rows without content dropped and
 * rows padded to the widest (io/FastCsv.scala:21,35-41); header = row 0 all non-numeric and row 1 has a numeric cell
 * (FileOps.scala:154-166), blank header names become colN (:117-121).  `delimiter` 0 = pick the most frequent of
 * , TAB ; | on the first content line (the reference's statistical detector io/Delimiter.scala is not reproduced). */
int32_t um_csv_open(const char* path, int32_t delimiter, uint64_t* handle, int64_t* rows, int64_t* cols, int32_t* has_header);
int32_t um_csv_header(uint64_t handle, int64_t col, char* buf, int64_t buflen);
int32_t um_csv_read_host(uint64_t handle, double* dst, int64_t count);       /* the parsed table, row-major (host) */
int32_t um_csv_upload(uint64_t handle, const um_view* dst);                  /* f64, or f32 (loadMatF: toDouble.toFloat) */
int32_t um_csv_close(uint64_t handle);
/* `m.saveCSV(path, sep, nanAs)` (data/Mat.scala:90-116): NaN -> nanAs, +-Infinity -> Inf / -Inf, other cells
 * String.valueOf: java.lang.Double/Float.toString layout (shortest round-trip digits, JDK 19+), Int decimal, Boolean
 * true/false; rows end with '\n'.  `m` must be contiguous (materialise views first). */
int32_t um_csv_save(const um_view* m, const char* path, const char* sep, const char* nan_as);
/* host-only helpers: one cell of saveCSV / loadMatD */
int32_t um_csv_format_f64(double v, const char* nan_as, char* buf, int64_t buflen);
int32_t um_csv_format_f32(float v, const char* nan_as, char* buf, int64_t buflen);
int32_t um_csv_parse_cell(const char* text, double* out);

/* ---- elementwise (Mat.scala:1989-2102 Mat(+)Mat, :2154-2209,:2514-2531 Mat(+)scalar,
 *      :1416-1419 scalar(+)Mat, :4513-4588 in-place family = `out` aliasing `a`) ------ */
typedef enum um_binop { UM_ADD = 0, UM_SUB = 1, UM_MUL = 2, UM_DIV = 3, UM_MAXIMUM = 4, UM_MINIMUM = 5,
                        UM_POW = 6 } um_binop;
int32_t um_binary(int32_t op, const um_view* a, const um_view* b, const um_view* out);
int32_t um_binary_scalar(int32_t op, const um_view* a, double s, int32_t scalar_left, const um_view* out);

/* unary ops: unary_- (Mat.scala:2086), abs (:3557), exp/sqrt/log (MatMathOps.scala:23-41),
 * sigmoid (MatMathOps.scala:97-122), relu (:125-146), tanh; `out` may be f64 for an f32 input */
typedef enum um_unop { UM_NEG = 0, UM_ABS = 1, UM_EXP = 2, UM_SQRT = 3, UM_LOG = 4, UM_SIGMOID = 5,
                       UM_RELU = 6, UM_TANH = 7, UM_SQUARE = 8, UM_RELU_GENERIC = 9 } um_unop;
int32_t um_unary(int32_t op, const um_view* a, const um_view* out);

/* softmax(axis) (MatMathOps.scala:156-226); out is f64 */
int32_t um_softmax(const um_view* a, int32_t axis, const um_view* out);
/* softmax over strided lanes (axis 0 of a row-major matrix): 0 = automatic (large FP64 row-major matrices of up to
 * 32768 rows take the single-read strip kernel: a column strip held in tensor memory across a thread-block cluster,
 * softmax_strip.cuh; everything else the two-read kernels), 1 = two-read kernels only, 2 = strip kernel or
 * UM_ERR_ARG.  Process-wide; for tests and profiling. */
int32_t um_softmax_set_axis0_path(int32_t path);

/* ---- masks: IEEE compare -> Mat[Boolean] (Mat.scala:2564-2582,3948-4003), isnan/isinf/
 *      isfinite (:4214-4225), && || ! (:5157-5170) ------------------------------------ */
typedef enum um_cmpop { UM_GT = 0, UM_LT = 1, UM_GTE = 2, UM_LTE = 3, UM_EQ = 4, UM_NE = 5 } um_cmpop;
typedef enum um_classop { UM_ISNAN = 0, UM_ISINF = 1, UM_ISFINITE = 2 } um_classop;
typedef enum um_logicop { UM_AND = 0, UM_OR = 1, UM_NOT = 2, UM_XOR = 3 } um_logicop;
int32_t um_compare_scalar(int32_t op, const um_view* a, double s, const um_view* out_bool);
int32_t um_compare(int32_t op, const um_view* a, const um_view* b, const um_view* out_bool);
int32_t um_classify(int32_t op, const um_view* a, const um_view* out_bool);
int32_t um_mask_logic(int32_t op, const um_view* a, const um_view* b /* NULL for NOT */, const um_view* out_bool);
/* all / any: axis = -1 whole mask -> *host_out; axis 0/1 -> out_bool lanes (Mat.scala:4907-5001) */
int32_t um_mask_reduce(int32_t want_all, const um_view* a, int32_t axis, const um_view* out_bool,
                       int32_t* host_out);

/* ---- running folds and mask-driven selection --------------
 * cumsum / cummax / cummin (Mat.scala:3635-3745, cumExtremumD :929-964): axis 0 / 1 -> `out` of the
 * operand's shape; axis -1 = the flat cumsum over row-major order -> `out` 1 x (rows*cols)
 * (Mat.scala:3676-3690).  Every lane is folded sequentially in increasing index, exactly like the
 * reference, so results are bit-identical (fixture rows cum0fnv cum1fnv cmax0fnv cmin1fnv cummaxfnv
 * cumlast); cummax / cummin use java.lang.Double.compare's order and replace on strictly better. */
typedef enum um_cumop { UM_CUMSUM = 0, UM_CUMMAX = 1, UM_CUMMIN = 2 } um_cumop;
int32_t um_cumulative(int32_t kind, const um_view* a, int32_t axis, const um_view* out);
/* Mat.where(condition, x, y) and Mat.where(condition, sx, sy) (Mat.scala:1305-1336): out(i,j) =
 * condition(i,j) ? x(i,j) : y(i,j); shapes must match exactly (no broadcasting) */
int32_t um_where(const um_view* cond, const um_view* x, const um_view* y, const um_view* out);
int32_t um_where_scalar(const um_view* cond, double x, double y, const um_view* out);
/* m(mask) (Mat.scala:4008-4021): the elements where the mask is true, in row-major order, as a
 * 1 x k row.  Two calls because the caller owns the result: um_mask_count returns k (synchronises),
 * um_mask_select fills a 1 x k `out`. */
int32_t um_mask_count(const um_view* mask, int64_t* host_count);
int32_t um_mask_select(const um_view* a, const um_view* mask, const um_view* out);

/* ---- reductions (Mat.scala:2251-2279 sum, :3469-3485 mean, :4784-4805 variance, :4699-4717
 *      std, :3414-3467 sum(axis), :3487-3492 mean(axis), :4720-4739 std(axis), :2215-2249
 *      min/max, :3503-3556 min/max(axis)).  Population statistics (divide by n).  Sums use a
 *      fixed two-level tree: results are run-to-run reproducible. ---------------------- */
typedef enum um_redop { UM_SUM = 0, UM_MEAN = 1, UM_VAR = 2, UM_STD = 3, UM_MIN = 4, UM_MAX = 5 } um_redop;
int32_t um_reduce_all(int32_t kind, const um_view* a, double* host_out);
int32_t um_reduce_axis(int32_t kind, const um_view* a, int32_t axis, const um_view* out);
/* argmin/argmax -> (row, col), first occurrence under java.lang.Double.compare (Mat.scala:2281-2334) */
int32_t um_arg_all(int32_t want_max, const um_view* a, int64_t* row, int64_t* col);
/* idxmin/idxmax(axis) -> Mat[Int] (MatPandasOps.scala:20-64) */
int32_t um_idx_axis(int32_t want_max, const um_view* a, int32_t axis, const um_view* out_i32);

/* ---- dense linear algebra ---------------------------------------------------------- */
/* `a *@ b` (Mat.scala:2815-2881,4811-4817): C fresh contiguous; operands any stride /
 * transposed view, f64 (DMMA tensor pipe) or f32.  Replaces netlib.dgemm (Mat.scala:3317),
 * cblas_dgemm (:3336), MatmulPure.multiply (MatmulPure.scala:78) and multiplyFloat (:3236). */
int32_t um_matmul(const um_view* a, const um_view* b, const um_view* out);
/* FP32 `*@` kernel choice (the reference's uni.mat.blas switch, Mat.scala:299-335, picks between a BLAS
 * sgemm and matmulPure in the same way): 0 = automatic (row-major, 16-byte aligned operands whose
 * pitches and N are multiples of 4, M*N*K >= 2^18, run 3xTF32 on the tcgen05 tensor cores with FP32
 * promotion every 128 k, partial tiles zero-filled by TMA; everything else on the FFMA kernel; from 2^30 multiply-adds up the
 * hi / lo operands are split once in global memory -- a temporary of 8 (M + N) K bytes from the pool -- instead of in shared
 * memory, same bits), 1 = FFMA kernel only, 2 = tensor-core kernel with the split in shared memory or UM_ERR_ARG,
 * 3 = tensor-core kernel with the split in global memory or UM_ERR_ARG.  Process-wide; for tests and profiling.
 * UM_MATF_TENSOR=0 in the environment makes automatic mean 1.  Automatic also promotes a product with few output
 * tiles and K >= 8192 (A' A of a tall MatF) to the split-K FP64 kernel and rounds the result to FP32 once. */
int32_t um_matmul_set_f32_path(int32_t path);
/* FP64 `*@` CTA tile: 0 = automatic (64 x 64 tiles, two CTAs per SM, whenever 128 x 128 tiles would leave SMs idle
 * or a mostly empty last wave -- 1024^2 and 2048^2 of BASELINE config 3; 128 x 128 otherwise), 64 / 128 = forced.
 * Process-wide; for tests and profiling. */
int32_t um_matmul_set_f64_tile(int32_t tile);
/* inverse (Mat.scala:2979-3006): LU with partial pivoting (luDecomposeD :970-1001); exact-zero
 * pivot -> UM_ERR_SINGULAR.  Batched form: `batch` matrices `stride_*` elements apart. */
int32_t um_inverse(const um_view* a, const um_view* out);
int32_t um_solve(const um_view* a, const um_view* b, const um_view* out);   /* Mat.scala:3598-3634 */
int32_t um_inverse_batched(const um_view* a0, int64_t stride_a, const um_view* out0, int64_t stride_out,
                           int64_t batch);

/* ---- 3PRF (stats/Tprf3.scala) -------------------------------------------------------
 * One fused sweep over the T x N panel produces every cross product both fits need
 * (DESIGN.md section 4); the fits differ only in the O(T*L^2) finish. */
typedef struct um_tprf_out {
  /* all device pointers into caller-owned f64 buffers; NULL = not wanted */
  double* forecasts;   /* T                                                            */
  double* residuals;   /* T                                                            */
  double* alpha;       /* N_local      closed form / IS Full `alpha` (Tprf3.scala:207,1240-1250) */
  /* the four pass-1/2/3 fields are filled for EVERY mode, UM_TPRF_CLOSED included: tprfClosedForm
   * returns them too (Tprf3.scala:211-231,243-252) and they equal t3prf's */
  double* phi;         /* N_local x L  pass-1 slopes (Tprf3.scala:263-264)                 */
  double* phi_hat;     /* (L+1) x N_local  B1 incl. intercept row (Tprf3.scala:263)        */
  double* sigma;       /* T x L        pass-2 factors (Tprf3.scala:266-267)                */
  double* beta3;       /* L+1          pass-3 coefficients (Tprf3.scala:268-269)           */
  double* xstd;        /* N_local      column sample std (nanStdCols, Tprf3.scala:469-502) */
  /* host scalars, written after a stream sync */
  double r_squared;    /* closed/t3prf: ssy==0 -> 0 (Tprf3.scala:239-241); IS Full: unguarded (:1214-1225) */
  int32_t singular;    /* 1 if a small solve hit an exact-zero pivot: the call then returns UM_ERR_SINGULAR
                        * (ArithmeticException in the reference, Mat.scala:990) and the outputs that needed
                        * that inverse hold NaN                                                          */
} um_tprf_out;

typedef enum um_tprf_mode {
  UM_TPRF_CLOSED = 0,   /* Tprf3.tprfClosedForm  (Tprf3.scala:199-252)                     */
  UM_TPRF_ITER = 1,     /* Tprf3.t3prf           (Tprf3.scala:260-289)                     */
  UM_TPRF_IS_FULL = 2,  /* Tprf3.estimate3prf(.., "IS Full") (Tprf3.scala:1056-1110,1206-1250): T < 10 -> NaN
                         * fit; pass 3 is nanOls (rows with a NaN in y or Sigma left out, :929-931); R^2 over
                         * the rows that have a residual, no zero guard                                  */
  UM_TPRF_WINDOW = 3    /* one out-of-sample window of estimate3prf: runT3prf -> t3prfFast -> t3prfTail
                         * (Tprf3.scala:883-943) = t3prf's passes 1 and 2 with the nanOls pass 3             */
} um_tprf_mode;

/* n_total <= 0: a fit of the panel held in X (communicator or not -- nothing is exchanged).  n_total > 0 with an
 * active communicator of more than one rank: X is this rank's column shard (N_local columns, y and Z replicated,
 * n_total = N over all ranks) and the packed partial sums are all-reduced once.  X: T x N_local f64
 * view with cs == 1; y: T x 1; Z: T x L, any L >= 1 (Tprf3.scala:199-289 take any proxy count): L <= 8 runs the
 * single-read fused sweep, L > 8 one such sweep per group of 8 proxies and a finish on the library's own operators
 * (tprf_wide.cu; column-sharded fits included: the sums over predictors are completed across the ranks). */
int32_t um_tprf_fit(int32_t mode, const um_view* y, const um_view* X, const um_view* Z, int64_t n_total,
                    um_tprf_out* out);
/* Same fit with y/X/Z in HOST memory (row-major, ldx elements between rows of X): the panel is
 * streamed to the device in column blocks overlapped with the sweep; outputs to host pointers. */
int32_t um_tprf_fit_host(int32_t mode, const double* y, const double* X, int64_t ldx, const double* Z,
                         int64_t T, int64_t N_local, int64_t L, int64_t n_total, double* forecasts_host,
                         double* alpha_host, double* r_squared);
/* Batched independent panels (BASELINE config 5): panel p at X + p*stride_x etc.; outputs
 * forecasts[p*T..], r2[p] (device), alpha[p*N..] (may be NULL). */
int32_t um_tprf_fit_batched(int32_t mode, const double* y, const double* X, const double* Z, int64_t T,
                            int64_t N, int64_t L, int64_t batch, double* forecasts, double* alpha,
                            double* r2);
/* Which sweep the fits use: 0 = auto (a single-read sweep whenever one applies: T <= 8192, X 16-byte
 * aligned with an even row pitch -- thread-block clusters, plus a concurrent grid sweep on the SMs a
 * 16- or 8-CTA cluster shape strands; otherwise two sweeps), 1 = two sweeps only, 2 = the cluster
 * sweep or UM_ERR_ARG, 3 = the grid sweep (virtual CTA groups over all SMs, tensor-memory stash,
 * exchange through L2) or UM_ERR_ARG.  Process-wide; for tests and profiling. */
int32_t um_tprf_set_path(int32_t path);
/* Shape of the single-read sweep for a T-row panel on the current device: CTAs per cluster (0 = does
 * not apply), rows of the panel each CTA keeps on chip, clusters resident at once. */
/* Profiling aid: when non-NULL, cluster 0 of the single-read sweep records clock64() stamps per
 * (rank, warp, piece) into this device buffer of 16*16*64*8 int64 (tools/fused_timeline.py). */
int32_t um_tprf_debug_stamps(long long* device_buf);
int32_t um_tprf_fused_info(int64_t T, int32_t* cluster_size, int32_t* rows_per_piece, int32_t* resident_clusters);
/* ALL out-of-sample windows of estimate3prf's Cross Val / Recursive / Rolling procedures in one launch (Tprf3.scala:1112-1199:
 * the reference's `.par` loop over runT3prf -> t3prfFast -> t3prfTail, :883-943).  Xn: the globally standardised panel (T x N,
 * cs == 1), or -- standardize != 0 -- the RAW panel, which the call first divides by nanStdCols (Tprf3.scala:469-502,1078-1081);
 * Z: the proxy matrix (T x L), or NULL for `n_auto` automatic proxies (each window rebuilds them from its own
 * residuals, newest column in front, :1134-1139).  Window k keeps rows [lo[k], hi[k]) (kind 0: Recursive, Rolling) or every row
 * BUT [lo[k], hi[k]) (kind 1: Cross Val) and forecasts row held[k]; it is re-standardised by its own column std (invStdCols).
 * Fewer than min_obs kept rows -> NaN.  Outputs (host arrays of nwin): the forecasts `yhatt` and the mean of the window's y
 * (`rollfore`, over its non-NaN entries: colMean(yt, hasNan), which on NaN-free y is the plain mean).  NaN semantics are the reference's: a NaN in X or Z
 * inside a window spreads to its forecast, a NaN in y only leaves pass 3 (nanOls).  One CTA per window, one synchronisation
 * at the end; UM_ERR_SINGULAR if any window hit an exact-zero pivot (ArithmeticException out of the parallel loop). */
int32_t um_tprf_oos_windows(const um_view* y, const um_view* Xn, const um_view* Z, int32_t n_auto, int32_t kind,
                            const int32_t* lo, const int32_t* hi, const int32_t* held, int64_t nwin, int32_t min_obs,
                            int32_t standardize, double* fore_host, double* roll_host);
/* Out-of-sample point forecast of one window (t3prfTail's `yhatt`, Tprf3.scala:913-943), from the outputs of
 * um_tprf_fit(UM_TPRF_ITER) on the window's kept rows: phi (N x L), xstd (N), beta3 (L+1), and the held-out row
 * xt (1 x N, any column stride).  estimate3prf's OOS procedures (Tprf3.scala:1112-1199) call it once per window.
 * Synchronises; UM_ERR_SINGULAR (and NaN) when [1|Phi]'[1|Phi] has an exact-zero pivot. */
int32_t um_tprf_oos_forecast(const um_view* phi, const um_view* xstd, const um_view* beta3, const um_view* xt, double* host_out);
/* accounting for bench.py: algorithmic flops/bytes of one fit */
int32_t um_tprf_work(int64_t T, int64_t N, int64_t L, double* flops, double* bytes);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink -------------------------------- */
int32_t um_comm_unique_id(uint8_t out[128]);
int32_t um_comm_init(int32_t rank, int32_t world, const uint8_t id[128]);
int32_t um_comm_destroy(void);
int32_t um_comm_info(int32_t* rank, int32_t* world);
/* *p2p = 1 when the exchange step runs over mapped peer memory (CUDA IPC, NVLink / NVSwitch), 0 when it runs over
 * NCCL (peer mapping unavailable, or UM_COMM_P2P=0 in the environment at um_comm_init). */
int32_t um_comm_transport(int32_t* p2p);
/* deterministic sum: all-gather of each rank's block + fixed rank-order add */
int32_t um_comm_allreduce_sum_f64(double* buf, int64_t count);
int32_t um_comm_barrier(void);

#ifdef __cplusplus
}
#endif
#endif /* UNIMAT_H */
