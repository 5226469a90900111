// comm.cu -- one process per GPU; the single exchange step of a column-sharded 3PRF fit.
// The reduction is an all-gather of every rank's packed partial block followed by a fixed rank-order sum on each
// GPU, so the result is bit-identical on every rank.
//
// Transport: the block is ~0.6 MB and the step is latency-bound, so the default path does not go through NCCL.
// Every rank maps every other rank's exchange buffer (CUDA IPC over NVLink / NVSwitch peer memory) at
// um_comm_init; k_p2p_allreduce stores the block straight into all peers' slots, bumps a per-source arrival counter
// with a system-scope atomic, waits for the counters and adds the slots in rank order.  One kernel, no proxy thread,
// no ring (measurements: DESIGN.md 4.4).
// NCCL (all-gather + the same sum) remains the transport when peer mapping is unavailable, the block exceeds the
// slot, or UM_COMM_P2P=0.
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace um {

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  double* gather = nullptr;  // world * capacity doubles
  i64 capacity = 0;
};
static Comm g_comm;

// ---- peer-memory exchange ----------------------------------------------------------------------------------
constexpr i64 P2P_CAP = i64(1) << 17;      // doubles per (parity, source rank) slot = 1 MiB; config 4 packs 73 809
constexpr int P2P_MAXW = 16;
constexpr int P2P_CTAS = 32;               // CTAs of the push kernel = arrivals per source per exchange
constexpr size_t P2P_FLAG_BYTES = 4096;    // arrival counters, one 128-byte line per source rank
struct P2PDev {                            // passed to the kernels by value
  char* peer[P2P_MAXW];                    // every rank's exchange block as mapped here (peer[rank] = the local one)
  int rank, world;
};
struct P2P {
  bool ok = false;
  char* local = nullptr;                   // [2 parities][world][P2P_CAP] doubles | flags
  P2PDev dev = {};
  unsigned epoch = 0;                      // exchanges issued so far
  int* err_host = nullptr;                 // mapped pinned word: a wait that gave up sets it
  int* err_dev = nullptr;
};
static P2P g_p2p;
static int p2p_setup();
static void p2p_teardown();
static size_t p2p_data_bytes(int world) { return (size_t)2 * world * P2P_CAP * sizeof(double); }
__device__ __forceinline__ size_t p2p_data_bytes_dev(int world) { return (size_t)2 * world * P2P_CAP * sizeof(double); }

bool comm_active() { return g_comm.comm != nullptr; }
int comm_world() { return g_comm.world; }

#define UM_NCCL(expr)                                                                            \
  do {                                                                                           \
    ncclResult_t _r = (expr);                                                                    \
    if (_r != ncclSuccess) return um::fail(UM_ERR_NCCL, "%s: %s", #expr, ncclGetErrorString(_r)); \
  } while (0)

__global__ void __launch_bounds__(256) k_rank_order_sum(const double* __restrict__ gathered, i64 count, int world,
                                                         double* __restrict__ out) {
  for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < count; i += (i64)gridDim.x * 256) {
    double s = gathered[i];
    for (int r = 1; r < world; ++r) s += gathered[(i64)r * count + i];
    out[i] = s;
  }
}

// One kernel per exchange.  Every CTA (1) stores its part of this rank's block into slot [parity][rank] of EVERY rank
// (16-byte stores over NVLink for the peers), (2) fences and posts one system-scope arrival on every destination's
// counter for this source, (3) waits until every source has delivered `expect` arrivals in total, (4) adds its part
// of the slots in rank order.  All P2P_CTAS CTAs are resident at once (32 <= SM count, the stream is otherwise idle),
// and every CTA posts before it waits, so no rank can wait on a store that is itself waiting.  The wait is bounded
// (~4 s): a peer that died must not leave this GPU spinning; the give-up is reported through *err.
__global__ void __launch_bounds__(256) k_p2p_allreduce(double* __restrict__ buf, i64 count, P2PDev d, int parity, unsigned expect,
                                                        int* err) {
  asm volatile("griddepcontrol.wait;" ::: "memory");   // no-op unless launched as a programmatic dependent (of the gather)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the finish kernel may become resident meanwhile
  const i64 nvec = count >> 1;
  const size_t slot = ((size_t)parity * d.world + d.rank) * P2P_CAP * sizeof(double);
  const double2* src = reinterpret_cast<const double2*>(buf);
  // contiguous part per CTA, so that step (4) re-reads what this CTA's peers' same-numbered CTAs wrote
  const i64 per = (nvec + gridDim.x - 1) / gridDim.x;
  const i64 v0 = min((i64)blockIdx.x * per, nvec), v1 = min(v0 + per, nvec);
  for (i64 i = v0 + threadIdx.x; i < v1; i += 256) {
    const double2 v = src[i];
    for (int r = 0; r < d.world; ++r) reinterpret_cast<double2*>(d.peer[r] + slot)[i] = v;
  }
  const bool tail = (count & 1) && blockIdx.x == gridDim.x - 1;
  if (tail && threadIdx.x == 0) {
    const double v = buf[count - 1];
    for (int r = 0; r < d.world; ++r) reinterpret_cast<double*>(d.peer[r] + slot)[count - 1] = v;
  }
  __threadfence_system();
  __syncthreads();
  const char* base = d.peer[d.rank];
  if (threadIdx.x < d.world) {
    unsigned* post = reinterpret_cast<unsigned*>(d.peer[threadIdx.x] + p2p_data_bytes_dev(d.world) + (size_t)d.rank * 128);
    __threadfence_system();
    atomicAdd_system(post, 1u);
    const unsigned* flag = reinterpret_cast<const unsigned*>(base + p2p_data_bytes_dev(d.world) + (size_t)threadIdx.x * 128);
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if ((int)(v - expect) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) { *err = 1; break; }
      __nanosleep(32);
    }
  }
  __syncthreads();
  const double* slots = reinterpret_cast<const double*>(base) + (size_t)parity * d.world * P2P_CAP;
  const i64 e0 = 2 * v0, e1 = tail ? count : 2 * v1;
  for (i64 i = e0 + threadIdx.x; i < e1; i += 256) {
    double s = __ldcg(slots + i);                       // L2: the peers' stores never passed through this SM's L1
    for (int r = 1; r < d.world; ++r) s += __ldcg(slots + (size_t)r * P2P_CAP + i);
    buf[i] = s;
  }
}

static int p2p_allreduce(double* buf, i64 count) {
  const unsigned e = ++g_p2p.epoch;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(P2P_CTAS, 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.stream = ctx()->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  UM_CUDA(cudaLaunchKernelEx(&cfg, k_p2p_allreduce, buf, count, g_p2p.dev, (int)(e & 1u), e * (unsigned)P2P_CTAS, g_p2p.err_dev));
  UM_LAUNCHED();
  return UM_OK;
}

// Called after a stream synchronisation: did a bounded wait give up?
int comm_check_error() {
  if (g_p2p.err_host && *g_p2p.err_host) return fail(UM_ERR_NCCL, "peer-memory exchange timed out waiting for another rank");
  return UM_OK;
}

int comm_allreduce_sum_f64_impl(double* buf, i64 count) {
  if (!g_comm.comm) return fail(UM_ERR_NCCL, "communicator not initialised");
  if (g_comm.world == 1 || count == 0) return UM_OK;
  // the transport choice must be the same on every rank: it depends only on what was agreed at um_comm_init and on
  // the (common) count, never on a rank-local property of the buffer
  if (g_p2p.ok && count <= P2P_CAP) {
    if (reinterpret_cast<uintptr_t>(buf) & 15) {  // a view at an odd element offset: bounce through an aligned block
      static thread_local double* bounce = nullptr;
      if (!bounce) UM_CUDA(cudaMalloc(&bounce, (size_t)P2P_CAP * sizeof(double)));
      cudaStream_t st = ctx()->stream;
      UM_CUDA(cudaMemcpyAsync(bounce, buf, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, st));
      const int s = p2p_allreduce(bounce, count);
      if (s != UM_OK) return s;
      UM_CUDA(cudaMemcpyAsync(buf, bounce, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, st));
      return UM_OK;
    }
    return p2p_allreduce(buf, count);
  }
  cudaStream_t st = ctx()->stream;
  if (count > g_comm.capacity) {
    if (g_comm.gather) {
      UM_CUDA(cudaStreamSynchronize(st));
      UM_CUDA(cudaFree(g_comm.gather));
      g_comm.gather = nullptr;
    }
    i64 cap = count < (1 << 16) ? (1 << 16) : count;
    UM_CUDA(cudaMalloc(&g_comm.gather, (size_t)cap * g_comm.world * sizeof(double)));
    g_comm.capacity = cap;
  }
  UM_NCCL(ncclAllGather(buf, g_comm.gather, (size_t)count, ncclDouble, g_comm.comm, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  i64 gx = (count + 255) / 256;
  if (gx > 1024) gx = 1024;
  k_rank_order_sum<<<(unsigned)gx, 256, 0, st>>>(g_comm.gather, count, g_comm.world, buf);
  UM_LAUNCHED();
  return UM_OK;
}

// Map every rank's exchange block into this process.  Any failure on any rank leaves ALL ranks on the NCCL transport
// (the decision is agreed with a min-reduction), it is never an error.
static int p2p_setup() {
  g_p2p = P2P{};
  const int world = g_comm.world, rank = g_comm.rank;
  const char* env = getenv("UM_COMM_P2P");
  int want = (world > 1 && world <= P2P_MAXW && !(env && env[0] == '0')) ? 1 : 0;
  if (world == 1) return UM_OK;
  cudaStream_t st = ctx()->stream;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  const size_t bytes = p2p_data_bytes(world) + P2P_FLAG_BYTES;
  if (want) {
    if (cudaMalloc(&g_p2p.local, bytes) != cudaSuccess || cudaMemsetAsync(g_p2p.local, 0, bytes, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess || cudaIpcGetMemHandle(&mine, g_p2p.local) != cudaSuccess ||
        cudaHostAlloc(&g_p2p.err_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&g_p2p.err_dev, g_p2p.err_host, 0) != cudaSuccess) {
      cudaGetLastError();
      want = 0;
    } else {
      *g_p2p.err_host = 0;
    }
  }
  // exchange the handles (and each rank's verdict) with the communicator that already exists
  struct Msg { cudaIpcMemHandle_t h; int ok; int pad[3]; };
  Msg* dmsg = nullptr;
  UM_CUDA(cudaMalloc(&dmsg, sizeof(Msg) * (size_t)(world + 1)));
  Msg m;
  m.h = mine; m.ok = want; m.pad[0] = m.pad[1] = m.pad[2] = 0;
  UM_CUDA(cudaMemcpyAsync(dmsg + world, &m, sizeof(Msg), cudaMemcpyHostToDevice, st));
  UM_NCCL(ncclAllGather(dmsg + world, dmsg, sizeof(Msg), ncclChar, g_comm.comm, st));
  Msg all[P2P_MAXW + 1];
  if (world <= P2P_MAXW) UM_CUDA(cudaMemcpyAsync(all, dmsg, sizeof(Msg) * (size_t)world, cudaMemcpyDeviceToHost, st));
  UM_CUDA(cudaStreamSynchronize(st));
  int ok = want;
  for (int r = 0; r < world && r < P2P_MAXW; ++r) ok = ok && all[r].ok;
  if (ok) {
    for (int r = 0; r < world; ++r) {
      if (r == rank) { g_p2p.dev.peer[r] = g_p2p.local; continue; }
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      g_p2p.dev.peer[r] = static_cast<char*>(ptr);
    }
  }
  // a rank whose mapping failed takes everybody back to NCCL
  int* dflag = reinterpret_cast<int*>(dmsg);
  UM_CUDA(cudaMemcpyAsync(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
  UM_NCCL(ncclAllReduce(dflag, dflag + 1, 1, ncclInt, ncclMin, g_comm.comm, st));
  int agreed = 0;
  UM_CUDA(cudaMemcpyAsync(&agreed, dflag + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  UM_CUDA(cudaStreamSynchronize(st));
  UM_CUDA(cudaFree(dmsg));
  g_p2p.dev.rank = rank;
  g_p2p.dev.world = world;
  g_p2p.ok = agreed != 0;
  if (!g_p2p.ok) p2p_teardown();
  return UM_OK;
}

static void p2p_teardown() {
  for (int r = 0; r < P2P_MAXW; ++r)
    if (g_p2p.dev.peer[r] && g_p2p.dev.peer[r] != g_p2p.local) cudaIpcCloseMemHandle(g_p2p.dev.peer[r]);
  if (g_p2p.local) cudaFree(g_p2p.local);
  if (g_p2p.err_host) cudaFreeHost(g_p2p.err_host);
  cudaGetLastError();
  g_p2p = P2P{};
}

}  // namespace um

using namespace um;

extern "C" {

int32_t um_comm_unique_id(uint8_t out[128]) {
  if (!out) return fail(UM_ERR_ARG, "null out");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  UM_NCCL(ncclGetUniqueId(&id));
  memcpy(out, &id, 128);
  return UM_OK;
}

int32_t um_comm_init(int32_t rank, int32_t world, const uint8_t idbytes[128]) {
  UM_CTX();
  UM_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
  if (g_comm.comm) return fail(UM_ERR_ARG, "communicator already initialised");
  ncclUniqueId id;
  memcpy(&id, idbytes, 128);
  UM_NCCL(ncclCommInitRank(&g_comm.comm, world, id, rank));
  g_comm.rank = rank;
  g_comm.world = world;
  return p2p_setup();
}

int32_t um_comm_destroy(void) {
  if (g_comm.comm) {
    if (g_p2p.ok) um_comm_barrier();   // no rank unmaps while a peer may still store into its block
    if (ctx()->stream) cudaStreamSynchronize(ctx()->stream);
    p2p_teardown();
    ncclCommDestroy(g_comm.comm);
    g_comm.comm = nullptr;
  }
  if (g_comm.gather) {
    cudaFree(g_comm.gather);
    g_comm.gather = nullptr;
    g_comm.capacity = 0;
  }
  g_comm.rank = 0;
  g_comm.world = 1;
  return UM_OK;
}

int32_t um_comm_info(int32_t* rank, int32_t* world) {
  if (rank) *rank = g_comm.rank;
  if (world) *world = g_comm.world;
  return UM_OK;
}

int32_t um_comm_transport(int32_t* p2p) {
  if (p2p) *p2p = g_p2p.ok ? 1 : 0;
  return UM_OK;
}

int32_t um_comm_allreduce_sum_f64(double* buf, int64_t count) {
  UM_CTX();
  UM_REQUIRE(buf || count == 0, "null buffer");
  return comm_allreduce_sum_f64_impl(buf, count);
}

int32_t um_comm_barrier(void) {
  UM_CTX();
  if (!g_comm.comm || g_comm.world == 1) return um_sync();
  static double* token = nullptr;
  if (!token) UM_CUDA(cudaMalloc(&token, 8 * 2));
  UM_NCCL(ncclAllReduce(token, token + 1, 1, ncclDouble, ncclSum, g_comm.comm, ctx()->stream));
  return um_sync();
}

}  // extern "C"
