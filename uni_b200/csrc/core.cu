// core.cu -- lifecycle, memory, view descriptor arithmetic (host side), get/set, timers.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace um {

static thread_local char t_err[512] = "";
static thread_local Ctx t_ctx;
std::atomic<long long> g_launches{0};
static std::once_flag g_pool_once[64];
// The device the process chose with its first um_init: host threads that never called um_init themselves (a JVM
// Cleaner / ForkJoin worker, Python's GC running off the owning thread) inherit it instead of opening device 0.
static std::atomic<int> g_default_device{-1};
// Every um_malloc block remembers the device and stream it was allocated on, so um_free can be called from ANY host
// thread: the free is queued on the owner's stream (ordered after everything that thread submitted) on the owner's device.
struct Owner { int device; cudaStream_t stream; };
static std::mutex g_alloc_mu;
static std::unordered_map<void*, Owner> g_allocs;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
  return code;
}

Ctx* ctx() { return &t_ctx; }

std::mutex& plan_mutex() {
  static std::mutex m;
  return m;
}

static int init_ctx(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(UM_ERR_CUDA, "no CUDA device available (%s): libunimat has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return fail(UM_ERR_ARG, "device %d out of range [0,%d)", device, n);
  UM_CUDA(cudaSetDevice(device));
  Ctx& c = t_ctx;
  if (c.stream && c.device != device) {
    cudaStreamSynchronize(c.stream);
    if (c.scratch) cudaFree(c.scratch);
    c.scratch = nullptr;
    c.scratch_bytes = 0;
    cudaStreamDestroy(c.stream);
    cudaEventDestroy(c.ev0);
    cudaEventDestroy(c.ev1);
    c.stream = nullptr;
  }
  if (!c.stream) {
    UM_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    UM_CUDA(cudaEventCreate(&c.ev0));
    UM_CUDA(cudaEventCreate(&c.ev1));
    UM_CUDA(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, device));
  }
  c.device = device;
  int none = -1;
  g_default_device.compare_exchange_strong(none, device);
  if (device < 64) {
    std::call_once(g_pool_once[device], [device]() {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ULL;  // keep freed blocks in the pool
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
      }
    });
  }
  return UM_OK;
}

// Counts the library calls of the calling host thread that touch the device (every such entry point passes through
// ensure_ctx): lets um_mask_select recognise that nothing ran on this thread since the um_mask_count of the same mask.
static thread_local unsigned long long t_epoch = 0;
unsigned long long op_epoch() { return t_epoch; }
void op_epoch_passive() { --t_epoch; }  // the current call cannot change the contents of any existing buffer (malloc / free)

int ensure_ctx() {
  ++t_epoch;
  if (t_ctx.stream) {
    cudaSetDevice(t_ctx.device);
    return UM_OK;
  }
  const int d = g_default_device.load();
  return init_ctx(d >= 0 ? d : 0);
}

void* scratch(size_t bytes) {
  Ctx& c = t_ctx;
  if (bytes <= c.scratch_bytes) return c.scratch;
  size_t want = std::max<size_t>(bytes, 1u << 20);
  want = (want + (1u << 20) - 1) & ~size_t((1u << 20) - 1);
  if (c.scratch) {
    cudaStreamSynchronize(c.stream);
    cudaFree(c.scratch);
    c.scratch = nullptr;
    c.scratch_bytes = 0;
  }
  cudaError_t e = cudaMalloc(&c.scratch, want);
  if (e != cudaSuccess) {
    fail(UM_ERR_OOM, "scratch allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    return nullptr;
  }
  c.scratch_bytes = want;
  return c.scratch;
}

// ---- per-kernel profiling -------------------------------------------------------------------
constexpr int PROF_SLOTS = 8;
struct ProfPair { cudaEvent_t a, b; int slot; };
static thread_local bool t_prof = false;
static thread_local std::vector<ProfPair>* t_pending = nullptr;
static thread_local std::vector<cudaEvent_t>* t_free_ev = nullptr;
static thread_local double t_prof_ms[PROF_SLOTS];
static thread_local long long t_prof_n[PROF_SLOTS];
static thread_local cudaEvent_t t_cur = nullptr;

bool prof_on() { return t_prof; }
static cudaEvent_t prof_event() {
  if (!t_free_ev) t_free_ev = new std::vector<cudaEvent_t>();
  if (!t_free_ev->empty()) { cudaEvent_t e = t_free_ev->back(); t_free_ev->pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(int) {
  if (!t_prof) return;
  t_cur = prof_event();
  cudaEventRecord(t_cur, t_ctx.stream);
}
void prof_end(int slot) {
  if (!t_prof || !t_cur) return;
  cudaEvent_t b = prof_event();
  cudaEventRecord(b, t_ctx.stream);
  if (!t_pending) t_pending = new std::vector<ProfPair>();
  t_pending->push_back(ProfPair{t_cur, b, slot});
  t_cur = nullptr;
}
static void prof_drain() {
  if (!t_pending) return;
  for (auto& p : *t_pending) {
    cudaEventSynchronize(p.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess && p.slot >= 0 && p.slot < PROF_SLOTS) {
      t_prof_ms[p.slot] += ms;
      t_prof_n[p.slot] += 1;
    }
    t_free_ev->push_back(p.a);
    t_free_ev->push_back(p.b);
  }
  t_pending->clear();
}

}  // namespace um

using namespace um;

extern "C" int32_t um_prof_enable(int32_t on) { um::t_prof = on != 0; return UM_OK; }
extern "C" int32_t um_prof_reset(void) {
  um::prof_drain();
  for (int i = 0; i < um::PROF_SLOTS; ++i) { um::t_prof_ms[i] = 0.0; um::t_prof_n[i] = 0; }
  return UM_OK;
}
extern "C" int32_t um_prof_read(int32_t slot, double* total_ms, int64_t* count) {
  if (slot < 0 || slot >= um::PROF_SLOTS) return fail(UM_ERR_ARG, "bad profiling slot %d", slot);
  um::prof_drain();
  if (total_ms) *total_ms = um::t_prof_ms[slot];
  if (count) *count = um::t_prof_n[slot];
  return UM_OK;
}

extern "C" {

int32_t um_version(void) { return UM_VERSION_MAJOR * 1000 + UM_VERSION_MINOR; }
const char* um_last_error(void) { return um::t_err; }

int32_t um_device_count(int32_t* out) {
  if (!out) return fail(UM_ERR_ARG, "null out");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *out = n;
  return UM_OK;
}

int32_t um_init(int32_t device) {
  // One device per process (one process per GPU): kernel attributes, occupancy results and the peer-memory exchange
  // are cached process-wide, so a second device in the same process is refused instead of silently reusing them.
  const int d = g_default_device.load();
  if (d >= 0 && device != d)
    return fail(UM_ERR_ARG, "um_init(%d): this process is bound to device %d (libunimat is one process per GPU)", device, d);
  return init_ctx(device);
}

int32_t um_shutdown(void) {
  Ctx& c = *ctx();
  if (c.stream) {
    cudaStreamSynchronize(c.stream);
    if (c.scratch) cudaFree(c.scratch);
    c.scratch = nullptr;
    c.scratch_bytes = 0;
    cudaEventDestroy(c.ev0);
    cudaEventDestroy(c.ev1);
    cudaStreamDestroy(c.stream);
    c.stream = nullptr;
    c.device = -1;
  }
  return UM_OK;
}

int32_t um_sync(void) {
  UM_CTX();
  UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  return UM_OK;
}

int32_t um_device_props(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* total_mem) {
  UM_CTX();
  cudaDeviceProp p;
  UM_CUDA(cudaGetDeviceProperties(&p, ctx()->device));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (total_mem) *total_mem = (int64_t)p.totalGlobalMem;
  return UM_OK;
}

int32_t um_malloc(size_t bytes, void** out) {
  UM_CTX();
  op_epoch_passive();
  if (!out) return fail(UM_ERR_ARG, "null out");
  *out = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMallocAsync(out, bytes, ctx()->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(UM_ERR_OOM, "um_malloc(%zu): %s", bytes, cudaGetErrorString(e));
  }
  {
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    g_allocs[*out] = Owner{ctx()->device, ctx()->stream};
  }
  return UM_OK;
}

int32_t um_free(void* ptr) {
  if (!ptr) return UM_OK;
  // no UM_CTX(): a free must never open a context (a Cleaner thread would open device 0 on every rank)
  Owner o{-1, nullptr};
  {
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    auto it = g_allocs.find(ptr);
    if (it == g_allocs.end()) return fail(UM_ERR_ARG, "um_free(%p): not a live um_malloc block", ptr);
    o = it->second;
    g_allocs.erase(it);
  }
  int cur = -1;
  cudaGetDevice(&cur);
  if (cur != o.device) UM_CUDA(cudaSetDevice(o.device));
  // queued on the owner's stream: ordered after every kernel the owning thread launched on the block
  cudaError_t e = cudaFreeAsync(ptr, o.stream);
  if (e != cudaSuccess) {  // the owner re-initialised onto another device and its stream is gone
    cudaGetLastError();
    e = cudaFree(ptr);
  }
  if (cur >= 0 && cur != o.device) cudaSetDevice(cur);
  if (e != cudaSuccess) return fail(UM_ERR_CUDA, "um_free(%p): %s", ptr, cudaGetErrorString(e));
  return UM_OK;
}

int32_t um_malloc_host(size_t bytes, void** out) {
  UM_CTX();
  if (!out) return fail(UM_ERR_ARG, "null out");
  UM_CUDA(cudaMallocHost(out, bytes ? bytes : 16));
  return UM_OK;
}

int32_t um_free_host(void* ptr) {
  if (!ptr) return UM_OK;
  UM_CUDA(cudaFreeHost(ptr));
  return UM_OK;
}

int32_t um_h2d(void* dst, const void* src, size_t bytes) {
  UM_CTX();
  if (bytes == 0) return UM_OK;
  UM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx()->stream));
  // pageable sources may be reused by the caller right after return
  UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  return UM_OK;
}

int32_t um_d2h(void* dst, const void* src, size_t bytes) {
  UM_CTX();
  if (bytes == 0) return UM_OK;
  UM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx()->stream));
  UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  return UM_OK;
}

int32_t um_d2d(void* dst, const void* src, size_t bytes) {
  UM_CTX();
  if (bytes == 0) return UM_OK;
  UM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx()->stream));
  return UM_OK;
}

int32_t um_memset(void* dst, int32_t byte, size_t bytes) {
  UM_CTX();
  if (bytes == 0) return UM_OK;
  UM_CUDA(cudaMemsetAsync(dst, byte, bytes, ctx()->stream));
  return UM_OK;
}

int32_t um_h2d_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t height) {
  UM_CTX();
  if (width_bytes == 0 || height == 0) return UM_OK;
  UM_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height, cudaMemcpyHostToDevice, ctx()->stream));
  return UM_OK;
}

int32_t um_launch_count(int64_t* out) {
  if (!out) return fail(UM_ERR_ARG, "null out");
  *out = g_launches.load();
  return UM_OK;
}

int32_t um_timer_start(void) {
  UM_CTX();
  UM_CUDA(cudaEventRecord(ctx()->ev0, ctx()->stream));
  return UM_OK;
}

int32_t um_timer_stop(float* ms_out) {
  UM_CTX();
  UM_CUDA(cudaEventRecord(ctx()->ev1, ctx()->stream));
  UM_CUDA(cudaEventSynchronize(ctx()->ev1));
  float ms = 0.f;
  UM_CUDA(cudaEventElapsedTime(&ms, ctx()->ev0, ctx()->ev1));
  if (ms_out) *ms_out = ms;
  return UM_OK;
}

// ---- views (pure host arithmetic; no CUDA needed) ---------------------------------------
int32_t um_view_init(um_view* v, void* data, int64_t rows, int64_t cols, int32_t dtype) {
  if (!v) return fail(UM_ERR_ARG, "null view");
  if (rows < 0 || cols < 0) return fail(UM_ERR_ARG, "negative shape");
  v->data = data;
  v->offset = 0;
  v->rows = rows;
  v->cols = cols;
  v->rs = cols;
  v->cs = 1;
  v->dtype = dtype;
  v->transposed = 0;
  return UM_OK;
}

static bool is_standard_contiguous(const um_view* m) {  // Mat.scala:1878-1880
  return (m->rs == m->cols && m->cs == 1 && !m->transposed) || (m->rs == 1 && m->cs == m->rows && m->transposed);
}

int32_t um_view_is_weird(const um_view* m, int32_t* out) {  // Mat.scala:1865-1876
  if (!m || !out) return fail(UM_ERR_ARG, "null argument");
  if (is_standard_contiguous(m) || m->rs == 0 || m->cs == 0) {
    *out = 0;
    return UM_OK;
  }
  int64_t leading = m->transposed ? m->cols : m->rows;
  int64_t major = m->transposed ? m->cs : m->rs;
  *out = major > std::max<int64_t>(leading, 1) ? 1 : 0;
  return UM_OK;
}

int32_t um_view_is_contiguous(const um_view* m, int32_t* out) {  // Mat.scala:49-50
  if (!m || !out) return fail(UM_ERR_ARG, "null argument");
  *out = (m->rs == m->cols && m->cs == 1) ? 1 : 0;
  return UM_OK;
}

int32_t um_view_transpose(const um_view* v, um_view* out) {  // Mat.scala:198-199
  if (!v || !out) return fail(UM_ERR_ARG, "null argument");
  um_view t = *v;
  t.rows = v->cols;
  t.cols = v->rows;
  t.rs = v->cs;
  t.cs = v->rs;
  t.transposed = !v->transposed;
  *out = t;
  return UM_OK;
}

int32_t um_view_slice(const um_view* v, int64_t r0, int64_t r1, int64_t c0, int64_t c1, um_view* out,
                      int32_t* needs_copy) {  // Mat.scala:1905-1927
  if (!v || !out) return fail(UM_ERR_ARG, "null argument");
  if (!(0 <= r0 && r0 <= r1 && r1 <= v->rows && 0 <= c0 && c0 <= c1 && c1 <= v->cols))
    return fail(UM_ERR_ARG, "Slice [%lld,%lld) x [%lld,%lld) out of bounds for %lldx%lld", (i64)r0, (i64)r1, (i64)c0,
                (i64)c1, (i64)v->rows, (i64)v->cols);
  um_view s = *v;
  s.offset = v->offset + r0 * v->rs + c0 * v->cs;
  s.rows = r1 - r0;
  s.cols = c1 - c0;
  *out = s;
  int32_t w = 0;
  um_view_is_weird(&s, &w);
  if (needs_copy) *needs_copy = w;
  return UM_OK;
}

int32_t um_view_broadcast(const um_view* v, int64_t rows, int64_t cols, um_view* out) {  // Mat.scala:1933-1953
  if (!v || !out) return fail(UM_ERR_ARG, "null argument");
  if (v->rows == rows && v->cols == cols) {
    *out = *v;
    return UM_OK;
  }
  bool okr = v->rows == rows || v->rows == 1;
  bool okc = v->cols == cols || v->cols == 1;
  if (!okr || !okc)
    return fail(UM_ERR_ARG, "Cannot broadcast shape (%lld, %lld) to (%lld, %lld)", (i64)v->rows, (i64)v->cols, (i64)rows,
                (i64)cols);
  um_view b = *v;
  if (v->rows == 1 && rows > 1) b.rs = 0;
  if (v->cols == 1 && cols > 1) b.cs = 0;
  b.rows = rows;
  b.cols = cols;
  *out = b;
  return UM_OK;
}

int32_t um_broadcast_shape(const um_view* a, const um_view* b, int64_t* rows, int64_t* cols) {  // Mat.scala:1989-1993
  if (!a || !b || !rows || !cols) return fail(UM_ERR_ARG, "null argument");
  int64_t tr = std::max(a->rows, b->rows), tc = std::max(a->cols, b->cols);
  bool ok = (a->rows == tr || a->rows == 1) && (a->cols == tc || a->cols == 1) && (b->rows == tr || b->rows == 1) &&
            (b->cols == tc || b->cols == 1);
  if (!ok)
    return fail(UM_ERR_ARG, "Cannot broadcast shapes (%lld, %lld) and (%lld, %lld)", (i64)a->rows, (i64)a->cols,
                (i64)b->rows, (i64)b->cols);
  *rows = tr;
  *cols = tc;
  return UM_OK;
}

// ---- element get / set (slow path; Mat.apply / Mat.update, MatDOps.scala:38-68) ----------
int32_t um_get(const um_view* v, int64_t r, int64_t c, double* out) {
  UM_CTX();
  UM_VIEW(v);
  if (!out) return fail(UM_ERR_ARG, "null out");
  if (r < 0 || r >= v->rows || c < 0 || c >= v->cols)
    return fail(UM_ERR_ARG, "index (%lld, %lld) out of bounds for %lldx%lld", (i64)r, (i64)c, (i64)v->rows, (i64)v->cols);
  size_t es = dtype_size(v->dtype);
  const char* p = static_cast<const char*>(v->data) + (v->offset + r * v->rs + c * v->cs) * (int64_t)es;
  union {
    double d;
    float f;
    int32_t i;
    uint8_t b;
  } u;
  UM_CUDA(cudaMemcpyAsync(&u, p, es, cudaMemcpyDeviceToHost, ctx()->stream));
  UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  *out = v->dtype == UM_F64 ? u.d : v->dtype == UM_F32 ? (double)u.f : v->dtype == UM_I32 ? (double)u.i : (double)u.b;
  return UM_OK;
}

int32_t um_set(const um_view* v, int64_t r, int64_t c, double value) {
  UM_CTX();
  UM_VIEW(v);
  if (r < 0 || r >= v->rows || c < 0 || c >= v->cols)
    return fail(UM_ERR_ARG, "index (%lld, %lld) out of bounds for %lldx%lld", (i64)r, (i64)c, (i64)v->rows, (i64)v->cols);
  size_t es = dtype_size(v->dtype);
  char* p = static_cast<char*>(v->data) + (v->offset + r * v->rs + c * v->cs) * (int64_t)es;
  union {
    double d;
    float f;
    int32_t i;
    uint8_t b;
  } u;
  if (v->dtype == UM_F64) u.d = value;
  else if (v->dtype == UM_F32) u.f = (float)value;
  else if (v->dtype == UM_I32) u.i = (int32_t)value;
  else u.b = value != 0.0;
  UM_CUDA(cudaMemcpyAsync(p, &u, es, cudaMemcpyHostToDevice, ctx()->stream));
  UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  return UM_OK;
}

}  // extern "C"
