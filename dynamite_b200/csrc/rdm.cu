// Reduced density matrix  rho[a,b] = sum_tr psi(a,tr) * conj(psi(b,tr))
// (reference: _backend/bpetsc_template_1.c:15-165, host-serial there).
//
// Device formulation: psi is viewed as a (2^k x 2^(L-k)) matrix A whose entry
// (a, tr) is the amplitude of the state with kept bits a and traced bits tr
// (zero when that state is not in the subspace, :57-83), and rho = A A^H is
// accumulated as an FP64 complex rank-k update: 32x32 output tiles, the traced
// index streamed through shared memory in chunks, partial sums over slices of
// the traced range combined with FP64 atomics.  Sharded vectors: every rank
// takes a slice of the traced range, reads the amplitudes it needs from its
// peers over NVLink, and the 4^k partial matrices are summed with NCCL.
#include <algorithm>

#include "context.h"
#include "vecops.cuh"

namespace dnm {
namespace {

constexpr int TILE = 64;  // output tile edge; 16x16 threads, 4x4 complex accumulators each
constexpr int KC = 16;    // traced states staged per step

struct RdmParams {
  int L, k;
  i64 keep_mask;  // bits of the kept spins
  i64 tr_mask;    // the other bits below L
  i64 dim;        // 2^k
  i64 tr_begin, tr_end;
  i64 tr_per_slice;
  int ntiles;     // tiles per edge; blockIdx.x enumerates the upper triangle (a_tile <= b_tile)
  int nloc_bits;  // >= 0: vector is sharded, owner = idx >> nloc_bits
  const cplx *peer[MAX_RANKS];
};

// scatter the low bits of v into the set positions of mask (software pdep)
__device__ __forceinline__ i64 deposit(i64 v, i64 mask)
{
  i64 out = 0;
  while (mask) {
    const i64 low = mask & -mask;
    if (v & 1) out |= low;
    v >>= 1;
    mask ^= low;
  }
  return out;
}

template <class S>
__device__ __forceinline__ cplx amplitude(const S &sub, const RdmParams &P, i64 state)
{
  const i64 idx = sub.s2i(state);
  if (idx < 0) return make_double2(0.0, 0.0);
  if (P.nloc_bits < 0) return P.peer[0][idx];
  return P.peer[idx >> P.nloc_bits][idx & (((i64)1 << P.nloc_bits) - 1)];
}

// rho is Hermitian: only tiles on or above the diagonal are computed (the host mirrors them).
template <class S>
__global__ void __launch_bounds__(256) k_rdm(S sub, RdmParams P, double *__restrict__ rho)
{
  __shared__ double2 As[KC][TILE + 1];
  __shared__ double2 Bs[KC][TILE + 1];
  __shared__ i64 dep_a[TILE], dep_b[TILE], dep_t[KC];
  // upper-triangle tile pair from the linear block index
  int ta = 0, rem = blockIdx.x;
  while (rem >= P.ntiles - ta) {
    rem -= P.ntiles - ta;
    ++ta;
  }
  const int tb = ta + rem;
  const bool diag_tile = (ta == tb);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const i64 a0 = (i64)ta * TILE, b0 = (i64)tb * TILE;
  const i64 t_lo = P.tr_begin + (i64)blockIdx.z * P.tr_per_slice;
  const i64 t_hi = min(t_lo + P.tr_per_slice, P.tr_end);
  if (threadIdx.x < TILE) {
    dep_a[threadIdx.x] = (a0 + threadIdx.x < P.dim) ? deposit(a0 + threadIdx.x, P.keep_mask) : -1;
    dep_b[threadIdx.x] = (b0 + threadIdx.x < P.dim) ? deposit(b0 + threadIdx.x, P.keep_mask) : -1;
  }

  double acc[4][4][2] = {};
  for (i64 t0 = t_lo; t0 < t_hi; t0 += KC) {
    __syncthreads();  // previous step's reads are done (and dep_a/dep_b are visible)
    if (threadIdx.x < KC) dep_t[threadIdx.x] = (t0 + threadIdx.x < t_hi) ? deposit(t0 + threadIdx.x, P.tr_mask) : -1;
    __syncthreads();
    for (int e = threadIdx.x; e < TILE * KC; e += 256) {
      const int row = e & (TILE - 1), kk = e / TILE;
      const i64 tr = dep_t[kk];
      cplx va = make_double2(0.0, 0.0), vb = va;
      if (tr >= 0) {
        if (dep_a[row] >= 0) va = amplitude(sub, P, tr | dep_a[row]);
        if (!diag_tile && dep_b[row] >= 0) vb = amplitude(sub, P, tr | dep_b[row]);
      }
      As[kk][row] = va;
      if (!diag_tile) Bs[kk][row] = vb;
    }
    __syncthreads();
    const double2(*Bt)[TILE + 1] = diag_tile ? As : Bs;
#pragma unroll 4
    for (int kk = 0; kk < KC; ++kk) {
      double2 av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[i] = As[kk][4 * ty + i];
        bv[i] = Bt[kk][4 * tx + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j][0] += av[i].x * bv[j].x + av[i].y * bv[j].y;  // a * conj(b)
          acc[i][j][1] += av[i].y * bv[j].x - av[i].x * bv[j].y;
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const i64 a = a0 + 4 * ty + i, b = b0 + 4 * tx + j;
      if (a < P.dim && b < P.dim) {
        atomicAdd(&rho[2 * (a * P.dim + b)], acc[i][j][0]);
        atomicAdd(&rho[2 * (a * P.dim + b) + 1], acc[i][j][1]);
      }
    }
}

// fill the tiles below the diagonal: rho[b][a] = conj(rho[a][b]) (32x32 shared-memory transpose)
__global__ void __launch_bounds__(256) k_rdm_mirror(double2 *__restrict__ rho, i64 dim)
{
  __shared__ double2 t[32][33];
  const i64 bx = blockIdx.x, by = blockIdx.y;  // 32x32 blocks; source block (by, bx) with bx's TILE > by's TILE
  if ((bx * 32) / TILE <= (by * 32) / TILE) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const i64 a = by * 32 + r, b = bx * 32 + tx;
    if (a < dim && b < dim) t[r][tx] = rho[a * dim + b];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const i64 b = bx * 32 + r, a = by * 32 + tx;
    if (a < dim && b < dim) {
      const double2 v = t[tx][r];
      rho[b * dim + a] = make_double2(v.x, -v.y);
    }
  }
}

}  // namespace
}  // namespace dnm

using namespace dnm;

extern "C" int dnm_rdm(dnm_vec_t v, const dnm_subspace_t *sub, int64_t keep_size, const int64_t *keep, double *out)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v && sub && out && keep_size >= 0 && (keep_size == 0 || keep), DNM_ERR_ARG, "null pointer");
  HostSubspace h;
  h.copy_from(sub);
  const int L = (int)h.desc.L;
  DNM_REQUIRE(keep_size <= L && keep_size <= 15, DNM_ERR_ARG, "keep has %lld spins (limit min(L, 15))",
              (long long)keep_size);
  DNM_REQUIRE(v->global_n == h.dim, DNM_ERR_ARG, "vector length %lld does not match subspace dimension %lld",
              (long long)v->global_n, (long long)h.dim);
  i64 keep_mask = 0;
  for (int64_t i = 0; i < keep_size; ++i) {
    DNM_REQUIRE(keep[i] >= 0 && keep[i] < L, DNM_ERR_ARG, "keep index %lld out of range", (long long)keep[i]);
    DNM_REQUIRE(i == 0 || keep[i] > keep[i - 1], DNM_ERR_ARG, "keep array must be strictly increasing");
    keep_mask |= (i64)1 << keep[i];
  }
  RdmParams P{};
  P.L = L;
  P.k = (int)keep_size;
  P.keep_mask = keep_mask;
  P.tr_mask = (((i64)1 << L) - 1) & ~keep_mask;
  P.dim = (i64)1 << keep_size;
  const i64 tr_dim = (i64)1 << (L - keep_size);
  if (G.nranks > 1) {
    DNM_REQUIRE(h.desc.type == DNM_FULL || h.desc.type == DNM_PARITY, DNM_ERR_UNSUPPORTED,
                "sharded rdm is implemented for Full and Parity");
    int nb = 0;
    while (((i64)1 << nb) < v->local_n) ++nb;
    P.nloc_bits = nb;
    for (int p = 0; p < G.nranks; ++p) P.peer[p] = v->peer[p];
  } else {
    P.nloc_bits = -1;
    P.peer[0] = v->d;
  }
  // this rank's share of the traced range, then slices for occupancy
  const i64 per_rank = (tr_dim + G.nranks - 1) / G.nranks;
  P.tr_begin = std::min(tr_dim, per_rank * G.rank);
  P.tr_end = std::min(tr_dim, P.tr_begin + per_rank);
  const i64 tiles = (P.dim + TILE - 1) / TILE;
  P.ntiles = (int)tiles;
  const i64 tile_pairs = tiles * (tiles + 1) / 2;
  const i64 span = P.tr_end - P.tr_begin;
  i64 want_slices = std::max<i64>(1, ((i64)G.sm_count * 4) / tile_pairs);
  i64 slices = std::max<i64>(1, std::min<i64>(want_slices, (span + KC - 1) / KC));
  slices = std::min<i64>(slices, 65535);
  P.tr_per_slice = std::max<i64>(KC, ((span + slices - 1) / slices + KC - 1) / KC * KC);
  slices = std::max<i64>(1, (span + P.tr_per_slice - 1) / P.tr_per_slice);

  const size_t nout = 2 * (size_t)P.dim * P.dim;
  double *d_rho = nullptr;
  h.upload();
  try {
    DNM_CHECK_CUDA(cudaMalloc(&d_rho, sizeof(double) * nout));
    DNM_CHECK_CUDA(cudaMemsetAsync(d_rho, 0, sizeof(double) * nout, G.stream));
    if (G.nranks > 1) allreduce_sum_dev(G.d_scratch + SCRATCH_DOUBLES - 8, 1);  // peers' vectors are complete
    if (span > 0) {
      const dim3 grid((unsigned)tile_pairs, 1, (unsigned)slices);
      switch (h.desc.type) {
        case DNM_FULL: k_rdm<<<grid, 256, 0, G.stream>>>(h.full(), P, d_rho); break;
        case DNM_PARITY: k_rdm<<<grid, 256, 0, G.stream>>>(h.parity(), P, d_rho); break;
        case DNM_SPIN_CONSERVE: k_rdm<<<grid, 256, 0, G.stream>>>(h.spin_dev(), P, d_rho); break;
        case DNM_EXPLICIT: k_rdm<<<grid, 256, 0, G.stream>>>(h.explicit_dev(), P, d_rho); break;
      }
      count_launch();
      DNM_CHECK_CUDA(cudaGetLastError());
    }
    if (G.nranks > 1) {
      DNM_REQUIRE(nout < ((size_t)1 << 31), DNM_ERR_UNSUPPORTED, "reduced density matrix too large to all-reduce");
      allreduce_sum_dev(d_rho, (int)nout);
    }
    if (P.dim > TILE) {
      // tiles strictly below the diagonal were not computed
      const unsigned nb = (unsigned)((P.dim + 31) / 32);
      k_rdm_mirror<<<dim3(nb, nb), 256, 0, G.stream>>>(reinterpret_cast<double2 *>(d_rho), P.dim);
      count_launch();
      DNM_CHECK_CUDA(cudaGetLastError());
    }
    DNM_CHECK_CUDA(cudaMemcpyAsync(out, d_rho, sizeof(double) * nout, cudaMemcpyDeviceToHost, G.stream));
    DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  } catch (...) {
    if (d_rho) cudaFree(d_rho);
    h.release();
    throw;
  }
  cudaFree(d_rho);
  h.release();
  DNM_API_END
}
