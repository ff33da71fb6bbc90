// Micro-benchmark: the memory skeleton of a persistent, TMA-pipelined tiled pass on B200.
//   mode 0: contiguous tiles, 1-D bulk TMA load into a ring, y = 2x stored from registers   (write pass)
//   mode 1: strided tiles (runs of 128 B), 4-D tensor-map TMA load, y += x through
//           cp.reduce.async.bulk.tensor ... add (f64) from a shared-memory staging buffer     (accumulate pass)
//   mode 2: as mode 1 but read-modify-write from registers (ld.global old y, st.global)        (control)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_stream tma_stream.cu
// (cuTensorMapEncodeTiled comes from cudaGetDriverEntryPoint: no -lcuda)
// run:   ./tma_stream L T mode ctas_per_sm nbuf threads [p]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

typedef unsigned long long u64;
typedef unsigned int u32;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, u32 bytes, u64 *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, u64 *bar)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap *tm, const void *src, int c0, int c1, int c2, int c3)
{
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tm),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Args {
  const double2 *x;
  double2 *y;
  int T, R, nbuf, mode;
  int p;       // strided modes: position of the high window run (window = bits 0..B-1 and p..p+T-B-1)
  int B;       // log2 of the run length
  u64 ntiles;
};

// tile number -> (low outer, high outer) for the strided layout
__device__ __forceinline__ void strided_coords(const Args &a, u64 t, int &lo, int &hi)
{
  const int nlow = a.p - a.B;
  lo = (int)(t & ((1ull << nlow) - 1ull));
  hi = (int)(t >> nlow);
}

__global__ void k_stream(const Args a, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tile_elems = 1 << a.T;
  const u32 tile_bytes = (u32)tile_elems * 16u;
  double2 *ring = reinterpret_cast<double2 *>(smem);
  double2 *outb = ring + (size_t)a.nbuf * tile_elems;  // mode 1 only
  u64 *full = reinterpret_cast<u64 *>(smem + (size_t)(a.nbuf + (a.mode == 1 ? 1 : 0)) * tile_bytes);
  const int tid = threadIdx.x, NT = blockDim.x;
  if (tid == 0) {
    for (int b = 0; b < a.nbuf; ++b) mbar_init(&full[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](u64 t, int b) {
    mbar_expect_tx(&full[b], tile_bytes);
    if (a.mode == 0) {
      bulk_load_1d(ring + (size_t)b * tile_elems, a.x + t * tile_elems, tile_bytes, &full[b]);
    } else {
      int lo, hi;
      strided_coords(a, t, lo, hi);
      tma_load_4d(ring + (size_t)b * tile_elems, &tmx, 0, lo, 0, hi, &full[b]);
    }
  };

  const u64 stride = gridDim.x;
  u64 t = blockIdx.x;
  // prologue: fill the ring
  if (tid == 0) {
    u64 tt = t;
    for (int b = 0; b < a.nbuf && tt < a.ntiles; ++b, tt += stride) issue(tt, b);
  }
  int b = 0;
  u32 phase = 0;
  for (; t < a.ntiles; t += stride) {
    mbar_wait(&full[b], phase);
    const double2 *tile = ring + (size_t)b * tile_elems;
    if (a.mode == 0) {
      for (int r = 0; r < a.R; ++r) {
        const double2 v = tile[tid + r * NT];
        a.y[t * tile_elems + tid + r * NT] = make_double2(2.0 * v.x, 2.0 * v.y);
      }
    } else if (a.mode == 1) {
      // the previous tile's reduction must have finished reading the staging buffer
      if (tid == 0) bulk_wait_read0();
      __syncthreads();
      for (int r = 0; r < a.R; ++r) outb[tid + r * NT] = tile[tid + r * NT];
      fence_async();
    } else {
      int lo, hi;
      strided_coords(a, t, lo, hi);
      for (int r = 0; r < a.R; ++r) {
        const int l = tid + r * NT;
        const u64 idx = (u64)(l & ((1 << a.B) - 1)) | ((u64)lo << a.B) | ((u64)(l >> a.B) << a.p) | ((u64)hi << (a.p + a.T - a.B));
        const double2 v = tile[l], o = a.y[idx];
        a.y[idx] = make_double2(o.x + v.x, o.y + v.y);
      }
    }
    __syncthreads();  // everybody is done with ring[b] (and, mode 1, the staging buffer is complete)
    if (tid == 0) {
      if (a.mode == 1) {
        int lo, hi;
        strided_coords(a, t, lo, hi);
        tma_reduce_add_4d(&tmy, outb, 0, lo, 0, hi);
        bulk_commit();
      }
      const u64 tn = t + (u64)a.nbuf * stride;
      if (tn < a.ntiles) issue(tn, b);
    }
    if (++b == a.nbuf) {
      b = 0;
      phase ^= 1;
    }
  }
  if (tid == 0 && a.mode == 1) bulk_wait0();
}

__global__ void k_fill(double2 *v, u64 n, double s)
{
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    v[i] = make_double2(s * (double)(i % 1021), -s * (double)(i % 509));
}

__global__ void k_check(const double2 *x, const double2 *y, u64 n, double fx, double fy0, unsigned long long *bad)
{
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const double wr = fx * x[i].x + fy0 * (double)(i % 1021), wi = fx * x[i].y - fy0 * (double)(i % 509);
    if (y[i].x != wr || y[i].y != wi) atomicAdd(bad, 1ull);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
  const int L = argc > 1 ? atoi(argv[1]) : 26, T = argc > 2 ? atoi(argv[2]) : 11, mode = argc > 3 ? atoi(argv[3]) : 0;
  const int ctas = argc > 4 ? atoi(argv[4]) : 2, nbuf = argc > 5 ? atoi(argv[5]) : 3, threads = argc > 6 ? atoi(argv[6]) : 512;
  const int B = argc > 8 ? atoi(argv[8]) : 3;
  const int p = argc > 7 ? atoi(argv[7]) : L - (T - B);
  const u64 n = 1ull << L;
  double2 *x, *y;
  CK(cudaMalloc(&x, n * 16));
  CK(cudaMalloc(&y, n * 16));
  k_fill<<<1184, 256>>>(x, n, 1.0);
  k_fill<<<1184, 256>>>(y, n, 0.5);
  CK(cudaDeviceSynchronize());

  CUtensorMap tmx, tmy;
  memset(&tmx, 0, sizeof(tmx));
  memset(&tmy, 0, sizeof(tmy));
  if (mode != 0) {
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult st;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &st));
    const int h = T - B;
    // dims (innermost first): 16 doubles | 2^(p-3) rows below the window run | 2^h window rows | the rest
    cuuint64_t dims[4] = {2ull << B, 1ull << (p - B), 1ull << h, 1ull << (L - p - h)};
    cuuint64_t strides[3] = {16ull << B, (1ull << p) * 16, (1ull << (p + h)) * 16};
    cuuint32_t box[4] = {(cuuint32_t)(2u << B), 1, (cuuint32_t)(1u << h), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int k = 0; k < 2; ++k) {
      CUresult rc = enc(k ? &tmy : &tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, k ? (void *)y : (void *)x, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS) {
        printf("cuTensorMapEncodeTiled failed: %d\n", (int)rc);
        return 1;
      }
    }
  }
  Args a{x, y, T, (1 << T) / threads, nbuf, mode, p, B, n >> T};
  const size_t smem = (size_t)(nbuf + (mode == 1 ? 1 : 0)) * ((size_t)16 << T) + 64;
  CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = sms * ctas;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 5;
  k_stream<<<grid, threads, smem>>>(a, tmx, tmy);  // warm-up (also the run that is checked)
  CK(cudaDeviceSynchronize());
  unsigned long long *bad;
  CK(cudaMallocManaged(&bad, 8));
  *bad = 0;
  k_check<<<1184, 256>>>(x, y, n, mode == 0 ? 2.0 : 1.0, mode == 0 ? 0.0 : 0.5, bad);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) k_stream<<<grid, threads, smem>>>(a, tmx, tmy);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  const double bytes = (double)n * 16.0 * (mode == 0 ? 2.0 : 3.0);
  printf("L=%d T=%d mode=%d ctas/SM=%d nbuf=%d threads=%d p=%d B=%d smem=%zu: %.3f ms  %.0f GB/s  mismatches=%llu\n", L, T, mode, ctas,
         nbuf, threads, p, B, smem, ms, bytes / ms / 1e6, *bad);
  return 0;
}
