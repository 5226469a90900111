// l2_reuse.cu -- how much of a buffer re-read inside one kernel comes back from the L2 on B200?
// A persistent grid streams an S-byte buffer R times; pass r+1 re-reads what pass r read (either the
// same CTA->slice mapping, or rotated so that another SM re-reads it).  Effective GB/s per pass vs S.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_stream(const double2* __restrict__ p, size_t nvec, int passes, int rotate, double* sink) {
  const size_t per = nvec / gridDim.x;
  double acc = 0.0;
  for (int r = 0; r < passes; ++r) {
    const size_t slice = (blockIdx.x + (size_t)r * rotate) % gridDim.x;
    const double2* q = p + slice * per;
    for (size_t i = threadIdx.x; i + 3 * 256 < per; i += 4 * 256) {
      double2 a = q[i], b = q[i + 256], c = q[i + 512], d = q[i + 768];
      acc += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
    }
  }
  if (acc == 1.2345) sink[0] = acc;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t maxb = (size_t)1 << 30;
  double2* buf; cudaMalloc(&buf, maxb); cudaMemset(buf, 0, maxb);
  double* sink; cudaMalloc(&sink, 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = sms * 8;
  const size_t sizes_mb[] = {8, 16, 24, 32, 48, 64, 96, 128, 192, 256, 512};
  for (int rotate : {0, 37}) {
    printf("rotate=%d (0: same SM re-reads its slice; 37: a different SM re-reads it)\n", rotate);
    for (size_t mb : sizes_mb) {
      size_t nvec = mb * (1 << 20) / 16;
      nvec = nvec / (grid * 1024) * (grid * 1024);
      float ms1, msR;
      const int R = 16;
      k_stream<<<grid, 256>>>(buf, nvec, 1, rotate, sink);
      cudaDeviceSynchronize();
      cudaEventRecord(a); k_stream<<<grid, 256>>>(buf, nvec, 1, rotate, sink); cudaEventRecord(b); cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms1, a, b);
      cudaEventRecord(a); k_stream<<<grid, 256>>>(buf, nvec, R + 1, rotate, sink); cudaEventRecord(b); cudaEventSynchronize(b);
      cudaEventElapsedTime(&msR, a, b);
      double bytes = (double)nvec * 16;
      printf("  S=%4zu MiB  single pass %7.1f GB/s   re-read passes %7.1f GB/s\n", mb, bytes / (ms1 * 1e-3) / 1e9, bytes * R / ((msR - ms1) * 1e-3) / 1e9);
    }
  }
  return 0;
}
