// Fused dot / norm / axpy kernels for the device-resident Krylov loops and the
// Vec surface.  These are HBM-streaming kernels: one 16-byte (complex128) load
// per element per operand, grid sized to a multiple of the SM count, partial
// sums reduced with warp shuffles.
#include "vecops.cuh"

#include <nccl.h>

#include <algorithm>

namespace dnm {

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// reduce NV per-thread accumulators over the block; thread 0 writes partials[blockIdx.x*NV + c]
template <int NV, bool MAX = false>
__device__ __forceinline__ void block_reduce_store(const double (&acc)[NV], double *partials)
{
  __shared__ double sh[NV][TPB / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const double v = MAX ? warp_max(acc[c]) : warp_sum(acc[c]);
    if (lane == 0) sh[c][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      double v = (lane < TPB / 32) ? sh[c][lane] : 0.0;
      v = MAX ? warp_max(v) : warp_sum(v);
      if (lane == 0) partials[(size_t)blockIdx.x * NV + c] = v;
    }
  }
}

// d_out[c] = reduce_b partials[b*nv + c]; one block per component c
__global__ void k_final_reduce(const double *__restrict__ partials, int nblocks, int nv, double *__restrict__ d_out,
                               int is_max, const int *__restrict__ active)
{
  if (active != nullptr && *active == 0) return;
  __shared__ double sh[TPB / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x;
  double v = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
    const double p = partials[(size_t)b * nv + c];
    v = is_max ? fmax(v, p) : v + p;
  }
  v = is_max ? warp_max(v) : warp_sum(v);
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = (lane < TPB / 32) ? sh[lane] : 0.0;
    t = is_max ? warp_max(t) : warp_sum(t);
    if (lane == 0) d_out[c] = t;
  }
}

__global__ void k_fill(cplx *__restrict__ v, int64_t n, cplx value)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = value;
}

__global__ void k_random_fill(cplx *__restrict__ v, int64_t n, int64_t offset, uint64_t seed)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (uint64_t)(i + offset) * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    uint64_t z2 = z * 0x9E3779B97F4A7C15ull + 0x7F4A7C15ull;
    z2 = (z2 ^ (z2 >> 30)) * 0xBF58476D1CE4E5B9ull;
    z2 ^= z2 >> 29;
    const double a = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    const double b = (double)(z2 >> 11) * (1.0 / 9007199254740992.0);
    v[i] = make_double2(2.0 * a - 1.0, 2.0 * b - 1.0);
  }
}

__global__ void k_scale(cplx *__restrict__ v, int64_t n, cplx a)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const cplx t = v[i];
    v[i] = make_double2(a.x * t.x - a.y * t.y, a.x * t.y + a.y * t.x);
  }
}

__global__ void k_axpby(cplx *__restrict__ y, const cplx *__restrict__ x, int64_t n, cplx a, cplx b)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const cplx xv = x[i], yv = y[i];
    y[i] = make_double2(a.x * xv.x - a.y * xv.y + b.x * yv.x - b.y * yv.y,
                        a.x * xv.y + a.y * xv.x + b.x * yv.y + b.y * yv.x);
  }
}

__global__ void k_dot(const cplx *__restrict__ x, const cplx *__restrict__ y, int64_t n, double *__restrict__ partials)
{
  double acc[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const cplx a = x[i], b = y[i];
    acc[0] += a.x * b.x + a.y * b.y;  // x * conj(y)
    acc[1] += a.y * b.x - a.x * b.y;
  }
  block_reduce_store<2>(acc, partials);
}

__global__ void k_sqnorm(const cplx *__restrict__ x, int64_t n, double *__restrict__ partials)
{
  double acc[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const cplx a = x[i];
    acc[0] += a.x * a.x + a.y * a.y;
  }
  block_reduce_store<1>(acc, partials);
}

template <bool MAX>
__global__ void k_absnorm(const cplx *__restrict__ x, int64_t n, double *__restrict__ partials)
{
  double acc[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const cplx a = x[i];
    const double m = hypot(a.x, a.y);
    acc[0] = MAX ? fmax(acc[0], m) : acc[0] + m;
  }
  block_reduce_store<1, MAX>(acc, partials);
}

// NV vectors against one w: each thread reads w[i] once and V_j[i] for every j.
template <int NV>
__global__ void __launch_bounds__(TPB) k_multi_dot(VecList vs, const cplx *__restrict__ w, int64_t n,
                                                   double *__restrict__ partials, const int *__restrict__ active)
{
  if (active != nullptr && *active == 0) return;
  double acc[2 * NV];
#pragma unroll
  for (int c = 0; c < 2 * NV; ++c) acc[c] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const cplx wv = w[i];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const cplx v = vs.p[j][i];
      acc[2 * j] += v.x * wv.x + v.y * wv.y;  // conj(v) * w
      acc[2 * j + 1] += v.x * wv.y - v.y * wv.x;
    }
  }
  block_reduce_store<2 * NV>(acc, partials);
}

template <int NV, bool WITH_NORM>
__global__ void __launch_bounds__(TPB) k_multi_axpy_sub(VecList vs, cplx *__restrict__ w, int64_t n,
                                                        const double *__restrict__ d_h, double *__restrict__ partials,
                                                        const int *__restrict__ active)
{
  if (active != nullptr && *active == 0) return;
  double hr[NV], hi[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    hr[j] = d_h[2 * j];
    hi[j] = d_h[2 * j + 1];
  }
  double acc[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    cplx wv = w[i];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const cplx v = vs.p[j][i];
      wv.x -= hr[j] * v.x - hi[j] * v.y;
      wv.y -= hr[j] * v.y + hi[j] * v.x;
    }
    w[i] = wv;
    if (WITH_NORM) acc[0] += wv.x * wv.x + wv.y * wv.y;
  }
  if (WITH_NORM) block_reduce_store<1>(acc, partials);
}

// out = sum_j c[j] V_j (first chunk overwrites, later chunks accumulate)
template <int NV>
__global__ void __launch_bounds__(TPB) k_multi_combine(VecList vs, cplx *out, int64_t n, const double *__restrict__ d_c,
                                                       int accumulate)
{
  double cr[NV], ci[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    cr[j] = d_c[2 * j];
    ci[j] = d_c[2 * j + 1];
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    cplx o = accumulate ? out[i] : make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const cplx v = vs.p[j][i];
      o.x += cr[j] * v.x - ci[j] * v.y;
      o.y += cr[j] * v.y + ci[j] * v.x;
    }
    out[i] = o;
  }
}

__global__ void k_scale_dev(cplx *dst, const cplx *src, int64_t n, const double *__restrict__ d_s, int reciprocal)
{
  double s = *d_s;
  if (reciprocal) s = (s != 0.0) ? 1.0 / s : 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const cplx t = src[i];
    dst[i] = make_double2(t.x * s, t.y * s);
  }
}

int stream_grid(int64_t n)
{
  const int64_t want = (n + TPB - 1) / TPB;
  const int64_t cap = (int64_t)G.sm_count * 8;
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

double *partials_for(int nblocks, int nv)
{
  const int64_t need = (int64_t)nblocks * nv;
  if (need > G.partial_capacity) {
    if (G.d_partials) DNM_CHECK_CUDA(cudaFree(G.d_partials));
    G.d_partials = nullptr;
    DNM_CHECK_CUDA(cudaMalloc(&G.d_partials, sizeof(double) * need));
    G.partial_capacity = need;
  }
  return G.d_partials;
}

void finish(int nblocks, int nv, double *d_out, bool is_max, const int *d_active = nullptr)
{
  k_final_reduce<<<nv, TPB, 0, G.stream>>>(G.d_partials, nblocks, nv, d_out, is_max ? 1 : 0, d_active);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
  if (is_max) allreduce_max_dev(d_out, nv);
  else allreduce_sum_dev(d_out, nv);
}

template <int NV>
void launch_multi_dot(const VecList &vs, const cplx *w, int64_t n, double *d_h, const int *d_active)
{
  const int g = reduce_grid(n);
  double *p = partials_for(g, 2 * NV);
  k_multi_dot<NV><<<g, TPB, 0, G.stream>>>(vs, w, n, p, d_active);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
  finish(g, 2 * NV, d_h, false, d_active);
}

template <int NV>
void launch_multi_axpy(const VecList &vs, cplx *w, int64_t n, const double *d_h, double *d_sq, const int *d_active)
{
  const int g = reduce_grid(n);
  if (d_sq) {
    double *p = partials_for(g, 1);
    k_multi_axpy_sub<NV, true><<<g, TPB, 0, G.stream>>>(vs, w, n, d_h, p, d_active);
    count_launch();
    DNM_CHECK_CUDA(cudaGetLastError());
    finish(g, 1, d_sq, false, d_active);
  } else {
    k_multi_axpy_sub<NV, false><<<g, TPB, 0, G.stream>>>(vs, w, n, d_h, nullptr, d_active);
    count_launch();
    DNM_CHECK_CUDA(cudaGetLastError());
  }
}

template <int NV>
void launch_multi_combine(const VecList &vs, cplx *out, int64_t n, const double *d_c, int accumulate)
{
  k_multi_combine<NV><<<stream_grid(n), TPB, 0, G.stream>>>(vs, out, n, d_c, accumulate);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
}

// run F<NV> for NV = vs.n in 1..MAX_FUSED
#define DNM_DISPATCH_NV(nv, CALL)                 \
  switch (nv) {                                   \
    case 1: { constexpr int NV = 1; CALL; } break;   \
    case 2: { constexpr int NV = 2; CALL; } break;   \
    case 3: { constexpr int NV = 3; CALL; } break;   \
    case 4: { constexpr int NV = 4; CALL; } break;   \
    case 5: { constexpr int NV = 5; CALL; } break;   \
    case 6: { constexpr int NV = 6; CALL; } break;   \
    case 7: { constexpr int NV = 7; CALL; } break;   \
    case 8: { constexpr int NV = 8; CALL; } break;   \
    case 9: { constexpr int NV = 9; CALL; } break;   \
    case 10: { constexpr int NV = 10; CALL; } break; \
    case 11: { constexpr int NV = 11; CALL; } break; \
    case 12: { constexpr int NV = 12; CALL; } break; \
    case 13: { constexpr int NV = 13; CALL; } break; \
    case 14: { constexpr int NV = 14; CALL; } break; \
    case 15: { constexpr int NV = 15; CALL; } break; \
    case 16: { constexpr int NV = 16; CALL; } break; \
    default: DNM_REQUIRE(false, DNM_ERR_INTERNAL, "bad fused vector count %d", (int)(nv)); \
  }

}  // namespace

int reduce_grid(int64_t n)
{
  const int64_t want = (n + TPB * 4 - 1) / (TPB * 4);
  const int64_t cap = (int64_t)G.sm_count * 4;
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

void vec_fill(cplx *v, int64_t n, cplx value)
{
  if (n == 0) return;
  k_fill<<<stream_grid(n), TPB, 0, G.stream>>>(v, n, value);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
}

void vec_random_fill(cplx *v, int64_t n, int64_t global_offset, uint64_t seed)
{
  if (n == 0) return;
  k_random_fill<<<stream_grid(n), TPB, 0, G.stream>>>(v, n, global_offset, seed);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
}

void vec_copy(cplx *dst, const cplx *src, int64_t n)
{
  if (n == 0 || dst == src) return;
  DNM_CHECK_CUDA(cudaMemcpyAsync(dst, src, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, G.stream));
}

void vec_scale(cplx *v, int64_t n, cplx a)
{
  if (n == 0) return;
  k_scale<<<stream_grid(n), TPB, 0, G.stream>>>(v, n, a);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
}

void vec_axpby(cplx *y, const cplx *x, int64_t n, cplx a, cplx b)
{
  if (n == 0) return;
  k_axpby<<<stream_grid(n), TPB, 0, G.stream>>>(y, x, n, a, b);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
}

void vec_dot_dev(const cplx *x, const cplx *y, int64_t n, double *d_out)
{
  const int g = reduce_grid(n);
  double *p = partials_for(g, 2);
  k_dot<<<g, TPB, 0, G.stream>>>(x, y, n, p);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
  finish(g, 2, d_out, false);
}

void vec_sqnorm_dev(const cplx *x, int64_t n, double *d_out)
{
  const int g = reduce_grid(n);
  double *p = partials_for(g, 1);
  k_sqnorm<<<g, TPB, 0, G.stream>>>(x, n, p);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
  finish(g, 1, d_out, false);
}

void vec_norm_other_dev(const cplx *x, int64_t n, int type, double *d_out)
{
  const int g = reduce_grid(n);
  double *p = partials_for(g, 1);
  if (type == 2) k_absnorm<true><<<g, TPB, 0, G.stream>>>(x, n, p);
  else k_absnorm<false><<<g, TPB, 0, G.stream>>>(x, n, p);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
  finish(g, 1, d_out, type == 2);
}

void multi_dot_dev(const VecList &vs, const cplx *w, int64_t n, double *d_h, const int *d_active)
{
  DNM_DISPATCH_NV(vs.n, launch_multi_dot<NV>(vs, w, n, d_h, d_active));
}

void multi_axpy_sub_dev(const VecList &vs, cplx *w, int64_t n, const double *d_h, double *d_sq, const int *d_active)
{
  DNM_DISPATCH_NV(vs.n, launch_multi_axpy<NV>(vs, w, n, d_h, d_sq, d_active));
}

void multi_combine_dev(const VecList &vs, cplx *out, int64_t n, const double *d_c)
{
  DNM_DISPATCH_NV(vs.n, launch_multi_combine<NV>(vs, out, n, d_c, 0));
}

void multi_combine_acc_dev(const VecList &vs, cplx *out, int64_t n, const double *d_c)
{
  DNM_DISPATCH_NV(vs.n, launch_multi_combine<NV>(vs, out, n, d_c, 1));
}

void vec_scale_dev(cplx *v, int64_t n, const double *d_s, bool reciprocal)
{
  if (n == 0) return;
  k_scale_dev<<<stream_grid(n), TPB, 0, G.stream>>>(v, v, n, d_s, reciprocal ? 1 : 0);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
}

void vec_scaled_copy_dev(cplx *dst, const cplx *src, int64_t n, const double *d_s, bool reciprocal)
{
  if (n == 0) return;
  k_scale_dev<<<stream_grid(n), TPB, 0, G.stream>>>(dst, src, n, d_s, reciprocal ? 1 : 0);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
}

void allreduce_sum_dev(double *d_buf, int count)
{
  if (G.nranks == 1) return;
  ncclResult_t r = ncclAllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, (ncclComm_t)G.nccl_comm, G.stream);
  DNM_REQUIRE(r == ncclSuccess, DNM_ERR_COMM, "ncclAllReduce(sum): %s", ncclGetErrorString(r));
}

void allreduce_max_dev(double *d_buf, int count)
{
  if (G.nranks == 1) return;
  ncclResult_t r = ncclAllReduce(d_buf, d_buf, count, ncclDouble, ncclMax, (ncclComm_t)G.nccl_comm, G.stream);
  DNM_REQUIRE(r == ncclSuccess, DNM_ERR_COMM, "ncclAllReduce(max): %s", ncclGetErrorString(r));
}

void fetch_doubles(const double *d_src, double *h_dst, int count)
{
  DNM_REQUIRE(count <= SCRATCH_DOUBLES, DNM_ERR_INTERNAL, "fetch too large");
  DNM_CHECK_CUDA(cudaMemcpyAsync(G.h_scratch, d_src, sizeof(double) * count, cudaMemcpyDeviceToHost, G.stream));
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  for (int i = 0; i < count; ++i) h_dst[i] = G.h_scratch[i];
}

}  // namespace dnm
