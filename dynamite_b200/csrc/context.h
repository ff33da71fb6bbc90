// Internal state shared by the translation units of libdynamite_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/dynamite_b200.h"
#include "subspace.cuh"

namespace dnm {

typedef double2 cplx;  // complex128 on the device

// ---- error plumbing: nothing throws across the C ABI -----------------------
void set_error(const char *fmt, ...);
extern thread_local int g_status;

struct Fail {
  int code;
};

#define DNM_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      dnm::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
      throw dnm::Fail{e_ == cudaErrorMemoryAllocation ? DNM_ERR_MEM : DNM_ERR_CUDA};      \
    }                                                                                     \
  } while (0)

#define DNM_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      dnm::set_error(__VA_ARGS__);    \
      throw dnm::Fail{code};          \
    }                                 \
  } while (0)

// wraps the body of every extern "C" entry point
#define DNM_API_BEGIN try {
#define DNM_API_END                                      \
  return DNM_OK;                                         \
  }                                                      \
  catch (const dnm::Fail &f) { return f.code; }          \
  catch (const std::bad_alloc &) {                       \
    dnm::set_error("host allocation failed");            \
    return DNM_ERR_MEM;                                  \
  }                                                      \
  catch (const std::exception &e) {                      \
    dnm::set_error("internal error: %s", e.what());      \
    return DNM_ERR_INTERNAL;                             \
  }

// ---- process-wide state ----------------------------------------------------
constexpr int MAX_RANKS = 16;
constexpr int SCRATCH_DOUBLES = 1 << 16;  // device + pinned host scratch (reductions, small matrices)

struct Globals {
  bool inited = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // side stream: remote (NVLink) passes overlap the local ones
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // copy streams and events of dnm_mat_mult_host_batch (created on first use)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_batch_start = nullptr, ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  double *d_scratch = nullptr;  // SCRATCH_DOUBLES
  double *h_scratch = nullptr;  // pinned, SCRATCH_DOUBLES
  double *d_partials = nullptr; // per-block partial sums
  int64_t partial_capacity = 0;
  int64_t launches = 0;
  // communicator
  int rank = 0, nranks = 1;
  void *nccl_comm = nullptr;  // ncclComm_t
};
extern Globals G;

void require_init();
// Krylov workspace vectors are recycled between solves (small problems are allocation-bound otherwise)
dnm_vec_t pool_acquire(int64_t global_n);
void pool_release(dnm_vec_t v);
void pool_clear();
int64_t pool_count(int64_t global_n);
inline void count_launch(int n = 1) { G.launches += n; }

// ---- objects behind the opaque handles --------------------------------------
}  // namespace dnm

struct dnm_vec_s {
  int64_t global_n = 0;
  int64_t local_n = 0;
  int64_t local_start = 0;
  dnm::cplx *d = nullptr;
  // peer-mapped device pointers of the same vector on the other ranks (multi-GPU)
  dnm::cplx *peer[dnm::MAX_RANKS] = {nullptr};
  bool owns = true;
};

namespace dnm {

// MSC terms in device layout.  Per mask the real-coefficient terms come first,
// then the imaginary ones (TERM_REAL of bpetsc_impl.h:34 resolved on the host).
struct MscDev {
  int nmasks = 0;
  int64_t nterms = 0;
  const i64 *masks = nullptr;      // [nmasks]
  const int *off_re = nullptr;     // [nmasks]   first real term
  const int *off_im = nullptr;     // [nmasks]   first imaginary term
  const int *off_end = nullptr;    // [nmasks]   one past the last term
  const i64 *signs = nullptr;      // [nterms]
  const double *coef = nullptr;    // [nterms]   the non-zero part of the coefficient
};

struct HostSubspace {
  dnm_subspace_t desc{};
  std::vector<i64> nck, state_map, rmap_idx, rmap_states;
  // device copies
  i64 *d_nck = nullptr, *d_state_map = nullptr, *d_rmap_idx = nullptr, *d_rmap_states = nullptr;
  i64 dim = 0;
  void copy_from(const dnm_subspace_t *s);  // deep copy + validation
  void upload();
  void release();
  SubFull full() const { return SubFull{desc.L}; }
  SubParity parity() const { return SubParity{desc.L, desc.space}; }
  SubSpinConserve spin_host() const { return SubSpinConserve{desc.L, desc.k, desc.ld_nchoosek, nck.data()}; }
  SubSpinConserve spin_dev() const { return SubSpinConserve{desc.L, desc.k, desc.ld_nchoosek, d_nck}; }
  SubExplicit explicit_host() const
  {
    return SubExplicit{desc.L, dim, state_map.data(), rmap_idx.empty() ? nullptr : rmap_idx.data(), rmap_states.data()};
  }
  SubExplicit explicit_dev() const { return SubExplicit{desc.L, dim, d_state_map, d_rmap_idx, d_rmap_states}; }
};

struct TiledPlan;  // matmult_tiled.cu

}  // namespace dnm

struct dnm_mat_s {
  // host copy of the MSC (as passed in)
  std::vector<dnm::i64> masks, mask_offsets, signs;
  std::vector<double> coeffs;  // interleaved complex
  int xparity = 0;
  dnm::HostSubspace left, right;
  int64_t M = 0, N = 0;              // global dims (after xparity halving)
  int64_t local_M = 0, local_N = 0;  // this rank's block
  // device MSC for the general kernel
  dnm::MscDev msc;
  std::vector<void *> owned;  // device allocations to free
  double *d_diag = nullptr;   // local_M doubles when precomputed
  double nrm = -1;
  bool same_explicit = false;  // Explicit -> the same Explicit space (state lists compare equal)
  int kernel_pref = 0;  // 0 auto, 1 general, 2 tiled
  int tile_bits = 0;    // 0 auto
  int tile_rows = 0;    // rows per thread in the tiled kernel: 0 auto, 8 or 16
  int pipeline = 0;     // pipelined persistent tiled kernel: 0 auto, 1 on, 2 off
  int jit = -1;         // operator-specialised (NVRTC) kernels for lean tiled passes: -1 auto, 0 off, 1 on
  int autotune = -1;    // time a few plan shapes at the first MatMult of a big matrix: -1 auto, 0 off, 1 on
  int tuned_shape = -1; // the shape the autotuner kept (index into TUNE_SHAPES), -1: none
  int far_bits = -1;    // outer positions a tiled pass may serve through the L2 (FAR masks): -1 auto
  int verbose = 0;
  dnm::TiledPlan *tiled = nullptr;
  int launches_per_mult = 0;
  int kernel_used = 0;
  dnm_vec_t work_x = nullptr, work_y = nullptr;  // device staging for dnm_mat_mult_host
  dnm_vec_t work_x2 = nullptr, work_y2 = nullptr;  // second pair: dnm_mat_mult_host_batch double-buffers
};
