// comm.cu -- one process per GPU; NCCL over NVLink 5 / NVSwitch for the single exchange step
// of a column-sharded 3PRF fit.  The reduction is an all-gather of
// every rank's packed partial block followed by a fixed rank-order sum on each GPU, so the
// result is bit-identical on every rank and independent of NCCL's algorithm choice.
#include <nccl.h>
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

int comm_allreduce_sum_f64_impl(double* buf, i64 count) {
  if (!g_comm.comm) return fail(UM_ERR_NCCL, "communicator not initialised");
  if (g_comm.world == 1 || count == 0) return UM_OK;
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
  return UM_OK;
}

int32_t um_comm_destroy(void) {
  if (g_comm.comm) {
    if (ctx()->stream) cudaStreamSynchronize(ctx()->stream);
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
