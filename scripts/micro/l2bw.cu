// L2 vs DRAM bandwidth probe: read-only and copy over buffers of several sizes.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_read(const double2 *a, size_t n, int reps, double *out)
{
  double acc = 0;
  for (int rep = 0; rep < reps; ++rep)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      double2 v = a[i];
      acc += v.x + v.y;
    }
  if (acc == 1.2345) *out = acc;
}
__global__ void k_copy(const double2 *a, double2 *b, size_t n, int reps)
{
  for (int rep = 0; rep < reps; ++rep)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
int main()
{
  const size_t maxn = (size_t)1 << 27;  // 2 GiB of double2
  double2 *a, *b;
  double *out;
  cudaMalloc(&a, maxn * 16);
  cudaMalloc(&b, maxn * 16);
  cudaMalloc(&out, 8);
  cudaMemset(a, 0, maxn * 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int lg = 19; lg <= 27; lg += 1) {
    size_t n = (size_t)1 << lg;
    int reps = (int)((maxn * 4) / n);
    if (reps < 1) reps = 1;
    for (int mode = 0; mode < 2; ++mode) {
      for (int w = 0; w < 2; ++w) {
        if (w == 1) cudaEventRecord(e0);
        if (mode == 0) k_read<<<148 * 8, 512>>>(a, n, reps, out);
        else k_copy<<<148 * 8, 512>>>(a, b, n, reps);
        if (w == 1) cudaEventRecord(e1);
      }
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double bytes = (double)n * 16 * reps * (mode == 0 ? 1 : 2);
      printf("%s buffer %8.1f MiB: %8.1f GB/s\n", mode == 0 ? "read" : "copy", n * 16.0 / (1 << 20), bytes / ms / 1e6);
    }
  }
  return 0;
}
