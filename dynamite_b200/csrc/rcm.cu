// Auto-subspace construction on the device: the breadth-first search of the reference's
// compute_rcm (_backend/bsubspace.pyx:212-261) as a frontier expansion.
//
// The reference walks state_map as a queue: for state i (in order) and every unique mask j (in order)
// with a non-zero summed coefficient, the neighbour state ^ mask is appended unless it was seen.  The
// result therefore lists every reachable state in the order of its FIRST discovery, discoveries
// ordered by (i, j).  Here a chunk of queue entries is expanded by one kernel: every (i, j) candidate
// is inserted into an open-addressing hash table whose value is the smallest discovery key
// 1 + i * nmasks + j seen for that state (atomicMin; states found by earlier chunks keep their smaller
// keys), a second kernel flags the candidates that own their state's key, an exclusive scan turns
// the flags into queue positions and a scatter appends them -- the same order as the reference, bit
// for bit (tests/test_gpu_krylov.py::test_compute_rcm_device).  The serial host loop is what the
// reference's own docstring calls the scalability limit of Auto (subspaces.py:473-474).
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "context.h"

namespace dnm {
namespace {

typedef unsigned long long u64;
constexpr i64 EMPTY = -1;

struct RcmMsc {
  int nmasks;
  const i64 *mask;      // [nmasks] (runs of equal masks in the caller's term order)
  const int *first;     // [nmasks + 1] term range of each run
  const i64 *signs;     // [nterms]
  const double *coeffs; // [2 * nterms]
};

__device__ __forceinline__ u64 hash64(u64 x)
{
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

// insert `state` with discovery key `key` (keeps the minimum); table size is a power of two
__device__ __forceinline__ void table_insert(i64 *keys, u64 *vals, u64 cap_mask, i64 state, u64 key)
{
  u64 slot = hash64((u64)state) & cap_mask;
  for (;;) {
    const i64 k = *(volatile const i64 *)&keys[slot];
    if (k == state) break;
    if (k == EMPTY) {
      const u64 old = atomicCAS((u64 *)&keys[slot], (u64)EMPTY, (u64)state);
      if (old == (u64)EMPTY || old == (u64)state) break;
    }
    slot = (slot + 1) & cap_mask;
  }
  atomicMin(&vals[slot], key);
}

__device__ __forceinline__ u64 table_value(const i64 *keys, const u64 *vals, u64 cap_mask, i64 state)
{
  u64 slot = hash64((u64)state) & cap_mask;
  while (keys[slot] != state) slot = (slot + 1) & cap_mask;
  return vals[slot];
}

__global__ void k_rcm_init(i64 *keys, u64 *vals, u64 cap)
{
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < cap; i += (u64)gridDim.x * blockDim.x) {
    keys[i] = EMPTY;
    vals[i] = ~0ull;
  }
}

__global__ void k_rcm_seed(i64 *keys, u64 *vals, u64 cap_mask, i64 start) { table_insert(keys, vals, cap_mask, start, 0ull); }

// candidates of the queue entries [i0, i1): edge[c] = neighbour or EMPTY, c = (i - i0) * nmasks + j
__global__ void k_rcm_expand(RcmMsc M, const i64 *__restrict__ state_map, i64 i0, i64 i1, i64 *__restrict__ edge,
                             i64 *keys, u64 *vals, u64 cap_mask)
{
  const i64 ncand = (i1 - i0) * M.nmasks;
  for (i64 c = blockIdx.x * (i64)blockDim.x + threadIdx.x; c < ncand; c += (i64)gridDim.x * blockDim.x) {
    const i64 i = i0 + c / M.nmasks;
    const int j = (int)(c % M.nmasks);
    const i64 state = state_map[i];
    double tr = 0.0, ti = 0.0;
    for (int t = M.first[j]; t < M.first[j + 1]; ++t) {
      const double sg = (__popcll((u64)(state & M.signs[t])) & 1) ? -1.0 : 1.0;
      tr += sg * M.coeffs[2 * t];
      ti += sg * M.coeffs[2 * t + 1];
    }
    i64 e = EMPTY;
    if (tr != 0.0 || ti != 0.0) {
      e = state ^ M.mask[j];
      table_insert(keys, vals, cap_mask, e, 1ull + (u64)i * (u64)M.nmasks + (u64)j);
    }
    edge[c] = e;
  }
}

// flag[c] = 1 when candidate c is the first discovery of its state
__global__ void k_rcm_flag(int nmasks, i64 i0, i64 ncand, const i64 *__restrict__ edge, int *__restrict__ flag,
                           const i64 *keys, const u64 *vals, u64 cap_mask)
{
  for (i64 c = blockIdx.x * (i64)blockDim.x + threadIdx.x; c < ncand; c += (i64)gridDim.x * blockDim.x) {
    const i64 e = edge[c];
    int f = 0;
    if (e != EMPTY) {
      const u64 key = 1ull + (u64)(i0 + c / nmasks) * (u64)nmasks + (u64)(c % nmasks);
      f = table_value(keys, vals, cap_mask, e) == key ? 1 : 0;
    }
    flag[c] = f;
  }
}

__global__ void k_rcm_scatter(i64 ncand, const i64 *__restrict__ edge, const int *__restrict__ flag,
                              const int *__restrict__ pos, i64 *__restrict__ state_map, i64 filled, i64 max_states)
{
  for (i64 c = blockIdx.x * (i64)blockDim.x + threadIdx.x; c < ncand; c += (i64)gridDim.x * blockDim.x)
    if (flag[c] && filled + pos[c] < max_states) state_map[filled + pos[c]] = edge[c];
}

struct DevBuf {
  void *p = nullptr;
  ~DevBuf()
  {
    if (p) cudaFree(p);
  }
  template <class T>
  T *alloc(size_t n)
  {
    DNM_CHECK_CUDA(cudaMalloc(&p, sizeof(T) * std::max<size_t>(n, 1)));
    return (T *)p;
  }
};

}  // namespace
}  // namespace dnm

using namespace dnm;

extern "C" int dnm_compute_rcm_device(int64_t nterms, const int64_t *masks, const int64_t *signs, const double *coeffs,
                                      int64_t *state_map, int64_t max_states, int64_t start, int64_t L, int64_t *dim_out)
{
  DNM_API_BEGIN
  (void)L;
  require_init();
  DNM_REQUIRE(nterms >= 1 && masks && signs && coeffs && state_map && dim_out && max_states >= 1, DNM_ERR_ARG,
              "bad arguments to compute_rcm");
  // runs of equal masks, in the caller's term order (bsubspace.pyx:241-257)
  std::vector<i64> run_mask;
  std::vector<int> run_first;
  for (int64_t t = 0; t < nterms; ++t)
    if (t == 0 || masks[t] != masks[t - 1]) {
      run_mask.push_back(masks[t]);
      run_first.push_back((int)t);
    }
  run_first.push_back((int)nterms);
  const int nmasks = (int)run_mask.size();

  DevBuf b_mask, b_first, b_signs, b_coeffs, b_map, b_keys, b_vals, b_edge, b_flag, b_pos, b_tmp;
  i64 *d_mask = b_mask.alloc<i64>(nmasks);
  int *d_first = b_first.alloc<int>(nmasks + 1);
  i64 *d_signs = b_signs.alloc<i64>(nterms);
  double *d_coeffs = b_coeffs.alloc<double>(2 * nterms);
  i64 *d_map = b_map.alloc<i64>(max_states);
  DNM_CHECK_CUDA(cudaMemcpyAsync(d_mask, run_mask.data(), sizeof(i64) * nmasks, cudaMemcpyHostToDevice, G.stream));
  DNM_CHECK_CUDA(cudaMemcpyAsync(d_first, run_first.data(), sizeof(int) * (nmasks + 1), cudaMemcpyHostToDevice, G.stream));
  DNM_CHECK_CUDA(cudaMemcpyAsync(d_signs, signs, sizeof(i64) * nterms, cudaMemcpyHostToDevice, G.stream));
  DNM_CHECK_CUDA(cudaMemcpyAsync(d_coeffs, coeffs, sizeof(double) * 2 * nterms, cudaMemcpyHostToDevice, G.stream));
  DNM_CHECK_CUDA(cudaMemcpyAsync(d_map, &start, sizeof(i64), cudaMemcpyHostToDevice, G.stream));
  RcmMsc M{nmasks, d_mask, d_first, d_signs, d_coeffs};

  // queue entries expanded per step: bounded by the candidate buffers
  const i64 chunk_states = std::max<i64>(1, std::min<i64>(((i64)1 << 24) / nmasks, (i64)max_states));
  const i64 max_cand = chunk_states * nmasks;
  u64 cap = 1024;
  while (cap < 2ull * ((u64)max_states + (u64)max_cand)) cap <<= 1;  // (room for one step's overflow past max_states)
  i64 *d_keys = b_keys.alloc<i64>(cap);
  u64 *d_vals = b_vals.alloc<u64>(cap);
  const int grid = G.sm_count * 8;
  k_rcm_init<<<grid, 256, 0, G.stream>>>(d_keys, d_vals, cap);
  k_rcm_seed<<<1, 1, 0, G.stream>>>(d_keys, d_vals, cap - 1, (i64)start);
  count_launch(2);

  i64 *d_edge = b_edge.alloc<i64>(max_cand);
  int *d_flag = b_flag.alloc<int>(max_cand);
  int *d_pos = b_pos.alloc<int>(max_cand);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flag, d_pos, (int)max_cand, G.stream);
  void *d_tmp = b_tmp.alloc<char>(tmp_bytes);

  i64 filled = 1, done = 0;
  while (done < filled) {
    const i64 i0 = done, i1 = std::min(filled, done + chunk_states);
    const i64 ncand = (i1 - i0) * nmasks;
    k_rcm_expand<<<grid, 256, 0, G.stream>>>(M, d_map, i0, i1, d_edge, d_keys, d_vals, cap - 1);
    k_rcm_flag<<<grid, 256, 0, G.stream>>>(nmasks, i0, ncand, d_edge, d_flag, d_keys, d_vals, cap - 1);
    DNM_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flag, d_pos, (int)ncand, G.stream));
    k_rcm_scatter<<<grid, 256, 0, G.stream>>>(ncand, d_edge, d_flag, d_pos, d_map, filled, (i64)max_states);
    count_launch(4);
    int last_pos = 0, last_flag = 0;
    DNM_CHECK_CUDA(cudaMemcpyAsync(&last_pos, d_pos + (ncand - 1), sizeof(int), cudaMemcpyDeviceToHost, G.stream));
    DNM_CHECK_CUDA(cudaMemcpyAsync(&last_flag, d_flag + (ncand - 1), sizeof(int), cudaMemcpyDeviceToHost, G.stream));
    DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
    const i64 found = (i64)last_pos + last_flag;
    DNM_REQUIRE(filled + found <= max_states, DNM_ERR_ARG, "state_map size too small");
    filled += found;
    done = i1;
  }
  DNM_CHECK_CUDA(cudaMemcpyAsync(state_map, d_map, sizeof(i64) * filled, cudaMemcpyDeviceToHost, G.stream));
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  *dim_out = filled;
  DNM_API_END
}
