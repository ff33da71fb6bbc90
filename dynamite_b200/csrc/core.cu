// Library state, error plumbing, communicator bootstrap, host-side subspace
// functions and the Vec surface of the C ABI.
#include <nccl.h>

#include <cstdarg>
#include <cstring>
#include <unordered_set>

#include "context.h"
#include "vecops.cuh"

namespace dnm {

Globals G;
thread_local int g_status = 0;
static thread_local char g_errbuf[1024] = "";

void set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_errbuf, sizeof(g_errbuf), fmt, ap);
  va_end(ap);
}

void require_init()
{
  DNM_REQUIRE(G.inited, DNM_ERR_CUDA,
              "dynamite_b200: no CUDA device bound (call dnm_init; this backend has no CPU fallback)");
}

// ---- workspace pool -------------------------------------------------------------
static std::vector<dnm_vec_t> g_pool;

int64_t pool_count(int64_t global_n)
{
  int64_t c = 0;
  for (dnm_vec_t v : g_pool) c += v->global_n == global_n;
  return c;
}

void pool_clear()
{
  std::vector<dnm_vec_t> old;
  old.swap(g_pool);
  for (dnm_vec_t v : old) dnm_vec_destroy(v);
}

dnm_vec_t pool_acquire(int64_t global_n)
{
  for (size_t i = 0; i < g_pool.size(); ++i)
    if (g_pool[i]->global_n == global_n) {
      dnm_vec_t v = g_pool[i];
      g_pool.erase(g_pool.begin() + i);
      return v;
    }
  if (!g_pool.empty()) pool_clear();  // a different problem size: give the memory back first
  dnm_vec_t v = nullptr;
  const int rc = dnm_vec_create(global_n, &v);
  if (rc) throw Fail{rc};
  return v;
}

void pool_release(dnm_vec_t v)
{
  if (!v) return;
  // keep at most a quarter of the device memory parked in the pool (the total is queried once:
  // cudaMemGetInfo costs milliseconds on some drivers and a Krylov solve releases dozens of vectors)
  static size_t t = 0;
  if (t == 0) {
    size_t f = 0;
    cudaMemGetInfo(&f, &t);
  }
  size_t pooled = 0;
  for (dnm_vec_t q : g_pool) pooled += sizeof(cplx) * (size_t)q->local_n;
  if (pooled + sizeof(cplx) * (size_t)v->local_n > t / 4 || g_pool.size() >= 256) {
    dnm_vec_destroy(v);
    return;
  }
  g_pool.push_back(v);
}

// ---- HostSubspace -----------------------------------------------------------

void HostSubspace::copy_from(const dnm_subspace_t *s)
{
  DNM_REQUIRE(s != nullptr, DNM_ERR_ARG, "null subspace descriptor");
  desc = *s;
  DNM_REQUIRE(s->L >= 1 && s->L <= 63, DNM_ERR_ARG, "L=%lld out of range [1,63]", (long long)s->L);
  switch (s->type) {
    case DNM_FULL:
      dim = (i64)1 << s->L;
      break;
    case DNM_PARITY:
      DNM_REQUIRE(s->space == 0 || s->space == 1, DNM_ERR_ARG, "Parity space must be 0 or 1");
      dim = (i64)1 << (s->L - 1);
      break;
    case DNM_SPIN_CONSERVE: {
      DNM_REQUIRE(s->k >= 0 && s->k <= s->L, DNM_ERR_ARG, "k must be between 0 and L");
      DNM_REQUIRE(s->nchoosek != nullptr && s->ld_nchoosek >= s->L + 1, DNM_ERR_ARG, "bad nchoosek table");
      nck.assign(s->nchoosek, s->nchoosek + (s->k + 1) * s->ld_nchoosek);
      desc.nchoosek = nck.data();
      dim = nck[s->k * s->ld_nchoosek + s->L];
      break;
    }
    case DNM_EXPLICIT: {
      DNM_REQUIRE(s->dim >= 1 && s->state_map && s->rmap_states, DNM_ERR_ARG, "bad Explicit subspace arrays");
      dim = s->dim;
      state_map.assign(s->state_map, s->state_map + dim);
      rmap_states.assign(s->rmap_states, s->rmap_states + dim);
      if (s->rmap_indices) rmap_idx.assign(s->rmap_indices, s->rmap_indices + dim);
      desc.state_map = state_map.data();
      desc.rmap_states = rmap_states.data();
      desc.rmap_indices = rmap_idx.empty() ? nullptr : rmap_idx.data();
      break;
    }
    default:
      DNM_REQUIRE(false, DNM_ERR_ARG, "invalid subspace type %d", (int)s->type);
  }
}

static i64 *upload_i64(const std::vector<i64> &h)
{
  if (h.empty()) return nullptr;
  i64 *d = nullptr;
  DNM_CHECK_CUDA(cudaMalloc(&d, sizeof(i64) * h.size()));
  DNM_CHECK_CUDA(cudaMemcpyAsync(d, h.data(), sizeof(i64) * h.size(), cudaMemcpyHostToDevice, G.stream));
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return d;
}

void HostSubspace::upload()
{
  release();
  d_nck = upload_i64(nck);
  d_state_map = upload_i64(state_map);
  d_rmap_idx = upload_i64(rmap_idx);
  d_rmap_states = upload_i64(rmap_states);
}

void HostSubspace::release()
{
  for (i64 **p : {&d_nck, &d_state_map, &d_rmap_idx, &d_rmap_states}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
}

}  // namespace dnm

using namespace dnm;

// ---- library / device --------------------------------------------------------

extern "C" const char *dnm_last_error(void) { return g_errbuf; }

extern "C" int dnm_device_count(int *count)
{
  DNM_API_BEGIN
  DNM_REQUIRE(count, DNM_ERR_ARG, "null pointer");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  DNM_API_END
}

extern "C" int dnm_init(int device)
{
  DNM_API_BEGIN
  if (G.inited) {
    DNM_REQUIRE(device == G.device, DNM_ERR_ARG, "already bound to device %d", G.device);
    return DNM_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    DNM_REQUIRE(false, DNM_ERR_CUDA, "no CUDA device available (%s); this backend has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  DNM_REQUIRE(device >= 0 && device < n, DNM_ERR_ARG, "device %d out of range (have %d)", device, n);
  DNM_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  DNM_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  DNM_REQUIRE(prop.major >= 10, DNM_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
              prop.major, prop.minor);
  G.device = device;
  G.sm_count = prop.multiProcessorCount;
  DNM_CHECK_CUDA(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
  {
    // the side stream outranks the main one so its few persistent CTAs get SM slots first
    int lo = 0, hi = 0;
    DNM_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    DNM_CHECK_CUDA(cudaStreamCreateWithPriority(&G.stream2, cudaStreamNonBlocking, hi));
  }
  DNM_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_fork, cudaEventDisableTiming));
  DNM_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_join, cudaEventDisableTiming));
  DNM_CHECK_CUDA(cudaEventCreate(&G.ev_start));
  DNM_CHECK_CUDA(cudaEventCreate(&G.ev_stop));
  DNM_CHECK_CUDA(cudaMalloc(&G.d_scratch, sizeof(double) * SCRATCH_DOUBLES));
  DNM_CHECK_CUDA(cudaMemset(G.d_scratch, 0, sizeof(double) * SCRATCH_DOUBLES));
  DNM_CHECK_CUDA(cudaHostAlloc(&G.h_scratch, sizeof(double) * SCRATCH_DOUBLES, cudaHostAllocDefault));
  G.launches = 0;
  G.inited = true;
  DNM_API_END
}

extern "C" int dnm_finalize(void)
{
  DNM_API_BEGIN
  if (!G.inited) return DNM_OK;
  cudaStreamSynchronize(G.stream);
  pool_clear();
  if (G.nccl_comm) {
    ncclCommDestroy((ncclComm_t)G.nccl_comm);
    G.nccl_comm = nullptr;
  }
  if (G.d_partials) cudaFree(G.d_partials);
  cudaFree(G.d_scratch);
  cudaFreeHost(G.h_scratch);
  cudaEventDestroy(G.ev_start);
  cudaEventDestroy(G.ev_stop);
  cudaEventDestroy(G.ev_fork);
  cudaEventDestroy(G.ev_join);
  cudaStreamDestroy(G.stream2);
  if (G.copy_in) cudaStreamDestroy(G.copy_in);
  if (G.copy_out) cudaStreamDestroy(G.copy_out);
  G.copy_in = G.copy_out = nullptr;
  if (G.ev_batch_start) cudaEventDestroy(G.ev_batch_start);
  G.ev_batch_start = nullptr;
  for (int b = 0; b < 2; ++b) {
    if (G.ev_in[b]) cudaEventDestroy(G.ev_in[b]);
    if (G.ev_cmp[b]) cudaEventDestroy(G.ev_cmp[b]);
    if (G.ev_out[b]) cudaEventDestroy(G.ev_out[b]);
    G.ev_in[b] = G.ev_cmp[b] = G.ev_out[b] = nullptr;
  }
  cudaStreamDestroy(G.stream);
  G = Globals();
  DNM_API_END
}

extern "C" int dnm_have_gpu(void) { return G.inited ? 1 : 0; }
extern "C" void *dnm_stream(void) { return (void *)G.stream; }

extern "C" int dnm_synchronize(void)
{
  DNM_API_BEGIN
  require_init();
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  DNM_API_END
}

extern "C" int dnm_timer_start(void)
{
  DNM_API_BEGIN
  require_init();
  DNM_CHECK_CUDA(cudaEventRecord(G.ev_start, G.stream));
  DNM_API_END
}

extern "C" int dnm_timer_stop(float *ms)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(ms, DNM_ERR_ARG, "null pointer");
  DNM_CHECK_CUDA(cudaEventRecord(G.ev_stop, G.stream));
  DNM_CHECK_CUDA(cudaEventSynchronize(G.ev_stop));
  DNM_CHECK_CUDA(cudaEventElapsedTime(ms, G.ev_start, G.ev_stop));
  DNM_API_END
}

extern "C" int dnm_mem_info(int64_t *free_bytes, int64_t *total_bytes)
{
  DNM_API_BEGIN
  require_init();
  size_t f = 0, t = 0;
  DNM_CHECK_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  DNM_API_END
}

extern "C" int64_t dnm_launch_count(int reset)
{
  const int64_t v = G.launches;
  if (reset) G.launches = 0;
  return v;
}

extern "C" int dnm_host_alloc(int64_t bytes, void **out)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(out && bytes > 0, DNM_ERR_ARG, "bad arguments to dnm_host_alloc");
  void *p = nullptr;
  DNM_CHECK_CUDA(cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault));
  *out = p;
  DNM_API_END
}

extern "C" int dnm_host_free(void *ptr)
{
  DNM_API_BEGIN
  if (ptr) DNM_CHECK_CUDA(cudaFreeHost(ptr));
  DNM_API_END
}

// ---- communicator --------------------------------------------------------------

extern "C" int dnm_comm_unique_id(char id[128])
{
  DNM_API_BEGIN
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId u;
  ncclResult_t r = ncclGetUniqueId(&u);
  DNM_REQUIRE(r == ncclSuccess, DNM_ERR_COMM, "ncclGetUniqueId: %s", ncclGetErrorString(r));
  memcpy(id, &u, 128);
  DNM_API_END
}

extern "C" int dnm_comm_init(int rank, int nranks, const char id[128])
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(nranks >= 1 && nranks <= MAX_RANKS && (nranks & (nranks - 1)) == 0, DNM_ERR_ARG,
              "number of ranks must be a power of 2 (<= %d), got %d", MAX_RANKS, nranks);
  DNM_REQUIRE(rank >= 0 && rank < nranks, DNM_ERR_ARG, "bad rank %d of %d", rank, nranks);
  DNM_REQUIRE(G.nccl_comm == nullptr, DNM_ERR_ARG, "communicator already initialised");
  if (nranks > 1) {
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t c;
    ncclResult_t r = ncclCommInitRank(&c, nranks, u, rank);
    DNM_REQUIRE(r == ncclSuccess, DNM_ERR_COMM, "ncclCommInitRank: %s", ncclGetErrorString(r));
    G.nccl_comm = c;
  }
  G.rank = rank;
  G.nranks = nranks;
  DNM_API_END
}

extern "C" int dnm_comm_rank(int *rank, int *nranks)
{
  if (rank) *rank = G.rank;
  if (nranks) *nranks = G.nranks;
  return DNM_OK;
}

extern "C" int dnm_comm_barrier(void)
{
  DNM_API_BEGIN
  require_init();
  if (G.nranks > 1) {
    allreduce_sum_dev(G.d_scratch + SCRATCH_DOUBLES - 8, 1);
    DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  }
  DNM_API_END
}

extern "C" int dnm_shard_plan(int n_index_bits, int nranks, int rank, int64_t nmasks, const int64_t *index_masks,
                              int32_t *partner, int64_t *local_masks)
{
  DNM_API_BEGIN
  DNM_REQUIRE(nranks >= 1 && (nranks & (nranks - 1)) == 0, DNM_ERR_ARG, "number of ranks must be a power of 2");
  DNM_REQUIRE(rank >= 0 && rank < nranks, DNM_ERR_ARG, "bad rank");
  DNM_REQUIRE(nmasks == 0 || (index_masks && partner && local_masks), DNM_ERR_ARG, "null pointer");
  int p = 0;
  while ((1 << p) < nranks) ++p;
  DNM_REQUIRE(n_index_bits > p && n_index_bits <= 62, DNM_ERR_ARG, "index space too small for %d ranks", nranks);
  const int nloc = n_index_bits - p;
  const int64_t lmask = ((int64_t)1 << nloc) - 1;
  for (int64_t k = 0; k < nmasks; ++k) {
    DNM_REQUIRE(index_masks[k] >= 0 && (index_masks[k] >> n_index_bits) == 0, DNM_ERR_ARG,
                "mask outside the index space");
    partner[k] = rank ^ (int32_t)(index_masks[k] >> nloc);
    local_masks[k] = index_masks[k] & lmask;
  }
  DNM_API_END
}

// ---- subspace index maps (host) ---------------------------------------------------

namespace {

template <class S>
void s2i_loop(const S &sub, int64_t n, const int64_t *states, int64_t *idxs)
{
  for (int64_t i = 0; i < n; ++i) idxs[i] = sub.s2i(states[i]);
}

template <class S>
void i2s_loop(const S &sub, int64_t n, const int64_t *idxs, int64_t *states)
{
  const i64 dim = sub.dim();
  for (int64_t i = 0; i < n; ++i) {
    DNM_REQUIRE(idxs[i] >= 0 && idxs[i] < dim, DNM_ERR_ARG,
                "Index %lld is out of bounds for subspace of dimension %lld.", (long long)idxs[i], (long long)dim);
    states[i] = sub.i2s(idxs[i]);
  }
}

}  // namespace

extern "C" int dnm_subspace_dim(const dnm_subspace_t *s, int64_t *dim)
{
  DNM_API_BEGIN
  DNM_REQUIRE(dim, DNM_ERR_ARG, "null pointer");
  HostSubspace h;
  h.copy_from(s);
  *dim = h.dim;
  DNM_API_END
}

extern "C" int dnm_subspace_s2i(const dnm_subspace_t *s, int64_t n, const int64_t *states, int64_t *idxs)
{
  DNM_API_BEGIN
  DNM_REQUIRE(s && (n == 0 || (states && idxs)), DNM_ERR_ARG, "null pointer");
  // work directly on the caller's arrays (borrowed for the call)
  switch (s->type) {
    case DNM_FULL: s2i_loop(SubFull{s->L}, n, states, idxs); break;
    case DNM_PARITY: s2i_loop(SubParity{s->L, s->space}, n, states, idxs); break;
    case DNM_SPIN_CONSERVE:
      s2i_loop(SubSpinConserve{s->L, s->k, s->ld_nchoosek, (const i64 *)s->nchoosek}, n, states, idxs);
      break;
    case DNM_EXPLICIT:
      s2i_loop(SubExplicit{s->L, s->dim, (const i64 *)s->state_map, (const i64 *)s->rmap_indices,
                           (const i64 *)s->rmap_states},
               n, states, idxs);
      break;
    default: DNM_REQUIRE(false, DNM_ERR_ARG, "invalid subspace type %d", (int)s->type);
  }
  DNM_API_END
}

extern "C" int dnm_subspace_i2s(const dnm_subspace_t *s, int64_t n, const int64_t *idxs, int64_t *states)
{
  DNM_API_BEGIN
  DNM_REQUIRE(s && (n == 0 || (states && idxs)), DNM_ERR_ARG, "null pointer");
  switch (s->type) {
    case DNM_FULL: i2s_loop(SubFull{s->L}, n, idxs, states); break;
    case DNM_PARITY: i2s_loop(SubParity{s->L, s->space}, n, idxs, states); break;
    case DNM_SPIN_CONSERVE:
      i2s_loop(SubSpinConserve{s->L, s->k, s->ld_nchoosek, (const i64 *)s->nchoosek}, n, idxs, states);
      break;
    case DNM_EXPLICIT:
      i2s_loop(SubExplicit{s->L, s->dim, (const i64 *)s->state_map, (const i64 *)s->rmap_indices,
                           (const i64 *)s->rmap_states},
               n, idxs, states);
      break;
    default: DNM_REQUIRE(false, DNM_ERR_ARG, "invalid subspace type %d", (int)s->type);
  }
  DNM_API_END
}

// Breadth-first closure of `start` under the operator's non-zero matrix
// elements (the Auto subspace, bsubspace.pyx:212-261).  Host code, as in the
// reference.
extern "C" int dnm_compute_rcm(int64_t nterms, const int64_t *masks, const int64_t *signs, const double *coeffs,
                               int64_t *state_map, int64_t max_states, int64_t start, int64_t L, int64_t *dim_out)
{
  DNM_API_BEGIN
  (void)L;
  DNM_REQUIRE(nterms >= 1 && masks && signs && coeffs && state_map && dim_out && max_states >= 1, DNM_ERR_ARG,
              "bad arguments to compute_rcm");
  std::unordered_set<int64_t> seen;
  seen.reserve((size_t)max_states);
  int64_t filled = 0;
  state_map[filled++] = start;
  seen.insert(start);
  for (int64_t i = 0; i < filled; ++i) {
    const int64_t state = state_map[i];
    double tr = 0, ti = 0;
    for (int64_t t = 0; t < nterms; ++t) {
      const double sg = parity64(state & signs[t]) ? -1.0 : 1.0;
      tr += sg * coeffs[2 * t];
      ti += sg * coeffs[2 * t + 1];
      if (t + 1 == nterms || masks[t + 1] != masks[t]) {
        if (tr != 0 || ti != 0) {
          const int64_t edge = state ^ masks[t];
          if (seen.insert(edge).second) {
            DNM_REQUIRE(filled < max_states, DNM_ERR_ARG, "state_map size too small");
            state_map[filled++] = edge;
          }
        }
        tr = ti = 0;
      }
    }
  }
  *dim_out = filled;
  DNM_API_END
}

// ---- Vec ----------------------------------------------------------------------------

namespace {

void share_with_peers(dnm_vec_s *v)
{
  // exchange one CUDA IPC handle per rank through NCCL and map every peer's copy
  cudaIpcMemHandle_t mine;
  DNM_CHECK_CUDA(cudaIpcGetMemHandle(&mine, v->d));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  char *stage = (char *)G.d_scratch;  // nranks*64 bytes
  DNM_CHECK_CUDA(cudaMemcpyAsync(stage + 64 * G.rank, &mine, 64, cudaMemcpyHostToDevice, G.stream));
  ncclResult_t r =
      ncclAllGather(stage + 64 * G.rank, stage, 64, ncclChar, (ncclComm_t)G.nccl_comm, G.stream);
  DNM_REQUIRE(r == ncclSuccess, DNM_ERR_COMM, "ncclAllGather(ipc handles): %s", ncclGetErrorString(r));
  std::vector<char> all(64 * G.nranks);
  DNM_CHECK_CUDA(cudaMemcpyAsync(all.data(), stage, all.size(), cudaMemcpyDeviceToHost, G.stream));
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  for (int p = 0; p < G.nranks; ++p) {
    if (p == G.rank) {
      v->peer[p] = v->d;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, all.data() + 64 * p, 64);
    void *ptr = nullptr;
    DNM_CHECK_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    v->peer[p] = (cplx *)ptr;
  }
}

void check_same_layout(dnm_vec_t a, dnm_vec_t b)
{
  DNM_REQUIRE(a && b, DNM_ERR_ARG, "null vector");
  DNM_REQUIRE(a->global_n == b->global_n && a->local_n == b->local_n, DNM_ERR_ARG,
              "vector sizes differ (%lld vs %lld)", (long long)a->global_n, (long long)b->global_n);
}

}  // namespace

extern "C" int dnm_vec_create(int64_t n, dnm_vec_t *out)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(out && n >= 1, DNM_ERR_ARG, "bad arguments to dnm_vec_create");
  DNM_REQUIRE(n % G.nranks == 0, DNM_ERR_ARG, "vector length %lld not divisible by %d ranks", (long long)n, G.nranks);
  dnm_vec_s *v = new dnm_vec_s();
  v->global_n = n;
  v->local_n = n / G.nranks;
  v->local_start = v->local_n * G.rank;
  // A shard that peers map through CUDA IPC gets an allocation of its own: cudaMalloc carves requests below
  // 2 MiB out of shared blocks, and an IPC handle names the whole block (two small vectors would then
  // share one handle and the first close would unmap the other).
  size_t bytes = sizeof(cplx) * (size_t)v->local_n;
  if (G.nranks > 1) bytes = std::max<size_t>(bytes, (size_t)2 << 20);
  try {
    if (cudaMalloc(&v->d, bytes) != cudaSuccess) {
      cudaGetLastError();
      v->d = nullptr;
      if (G.nranks == 1) pool_clear();  // parked workspace may be what is in the way
      DNM_CHECK_CUDA(cudaMalloc(&v->d, bytes));
    }
    DNM_CHECK_CUDA(cudaMemsetAsync(v->d, 0, sizeof(cplx) * v->local_n, G.stream));
    if (G.nranks > 1) share_with_peers(v);
  } catch (...) {
    if (v->d) cudaFree(v->d);
    delete v;
    throw;
  }
  *out = v;
  DNM_API_END
}

extern "C" int dnm_vec_destroy(dnm_vec_t v)
{
  DNM_API_BEGIN
  if (!v) return DNM_OK;
  if (G.inited) cudaStreamSynchronize(G.stream);
  if (G.nranks > 1) {
    // peers must have stopped reading before the allocation disappears
    allreduce_sum_dev(G.d_scratch + SCRATCH_DOUBLES - 8, 1);
    cudaStreamSynchronize(G.stream);
    for (int p = 0; p < G.nranks; ++p)
      if (p != G.rank && v->peer[p]) cudaIpcCloseMemHandle(v->peer[p]);
    allreduce_sum_dev(G.d_scratch + SCRATCH_DOUBLES - 8, 1);
    cudaStreamSynchronize(G.stream);
  }
  if (v->owns && v->d) cudaFree(v->d);
  delete v;
  DNM_API_END
}

extern "C" int dnm_vec_size(dnm_vec_t v, int64_t *global_n, int64_t *local_start, int64_t *local_end)
{
  DNM_API_BEGIN
  DNM_REQUIRE(v, DNM_ERR_ARG, "null vector");
  if (global_n) *global_n = v->global_n;
  if (local_start) *local_start = v->local_start;
  if (local_end) *local_end = v->local_start + v->local_n;
  DNM_API_END
}

extern "C" void *dnm_vec_device_ptr(dnm_vec_t v) { return v ? (void *)v->d : nullptr; }

extern "C" int dnm_vec_set_host(dnm_vec_t v, int64_t offset, int64_t count, const double *values)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v && values && offset >= 0 && count >= 0 && offset + count <= v->local_n, DNM_ERR_ARG,
              "bad range [%lld, %lld) for local size %lld", (long long)offset, (long long)(offset + count),
              (long long)(v ? v->local_n : 0));
  DNM_CHECK_CUDA(cudaMemcpyAsync(v->d + offset, values, sizeof(cplx) * count, cudaMemcpyHostToDevice, G.stream));
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  DNM_API_END
}

extern "C" int dnm_vec_get_host(dnm_vec_t v, int64_t offset, int64_t count, double *values)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v && values && offset >= 0 && count >= 0 && offset + count <= v->local_n, DNM_ERR_ARG,
              "bad range [%lld, %lld) for local size %lld", (long long)offset, (long long)(offset + count),
              (long long)(v ? v->local_n : 0));
  DNM_CHECK_CUDA(cudaMemcpyAsync(values, v->d + offset, sizeof(cplx) * count, cudaMemcpyDeviceToHost, G.stream));
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  DNM_API_END
}

// Any contiguous range of the GLOBAL vector, read through the peer mappings of the other ranks'
// shards (State.to_numpy on sharded vectors).  The caller synchronises the ranks first (Comm.barrier).
extern "C" int dnm_vec_get_host_global(dnm_vec_t v, int64_t offset, int64_t count, double *values)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v && values && offset >= 0 && count >= 0 && offset + count <= v->global_n, DNM_ERR_ARG,
              "bad range [%lld, %lld) for global size %lld", (long long)offset, (long long)(offset + count),
              (long long)(v ? v->global_n : 0));
  const int64_t per = v->local_n;
  int64_t done = 0;
  while (done < count) {
    const int64_t g = offset + done;
    const int r = (int)(g / per);
    const int64_t lo = g - (int64_t)r * per;
    const int64_t n = std::min<int64_t>(count - done, per - lo);
    const cplx *src = (G.nranks == 1 || r == G.rank) ? v->d : v->peer[r];
    DNM_REQUIRE(src != nullptr, DNM_ERR_COMM, "vector is not mapped on rank %d", r);
    DNM_CHECK_CUDA(cudaMemcpyAsync(values + 2 * done, src + lo, sizeof(cplx) * n, cudaMemcpyDeviceToHost, G.stream));
    done += n;
  }
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  DNM_API_END
}

namespace {
__global__ void k_scatter_set(cplx *v, int64_t count, const int64_t *idx, const cplx *vals, int add)
{
  // duplicates in idx are the caller's responsibility for add==0; add uses atomics
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    if (add) {
      atomicAdd(&v[idx[i]].x, vals[i].x);
      atomicAdd(&v[idx[i]].y, vals[i].y);
    } else {
      v[idx[i]] = vals[i];
    }
  }
}
__global__ void k_gather_get(const cplx *v, int64_t count, const int64_t *idx, cplx *vals)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    vals[i] = v[idx[i]];
}
}  // namespace

extern "C" int dnm_vec_set_values(dnm_vec_t v, int64_t count, const int64_t *local_idx, const double *values, int add)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v && (count == 0 || (local_idx && values)), DNM_ERR_ARG, "null pointer");
  if (count == 0) return DNM_OK;
  for (int64_t i = 0; i < count; ++i)
    DNM_REQUIRE(local_idx[i] >= 0 && local_idx[i] < v->local_n, DNM_ERR_ARG, "index %lld out of local range",
                (long long)local_idx[i]);
  int64_t *d_idx = nullptr;
  cplx *d_val = nullptr;
  DNM_CHECK_CUDA(cudaMalloc(&d_idx, sizeof(int64_t) * count));
  DNM_CHECK_CUDA(cudaMalloc(&d_val, sizeof(cplx) * count));
  DNM_CHECK_CUDA(cudaMemcpyAsync(d_idx, local_idx, sizeof(int64_t) * count, cudaMemcpyHostToDevice, G.stream));
  DNM_CHECK_CUDA(cudaMemcpyAsync(d_val, values, sizeof(cplx) * count, cudaMemcpyHostToDevice, G.stream));
  k_scatter_set<<<(int)std::min<int64_t>((count + 255) / 256, 1024), 256, 0, G.stream>>>(v->d, count, d_idx, d_val, add);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  cudaFree(d_idx);
  cudaFree(d_val);
  DNM_API_END
}

extern "C" int dnm_vec_get_values(dnm_vec_t v, int64_t count, const int64_t *local_idx, double *values)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v && (count == 0 || (local_idx && values)), DNM_ERR_ARG, "null pointer");
  if (count == 0) return DNM_OK;
  for (int64_t i = 0; i < count; ++i)
    DNM_REQUIRE(local_idx[i] >= 0 && local_idx[i] < v->local_n, DNM_ERR_ARG, "index %lld out of local range",
                (long long)local_idx[i]);
  int64_t *d_idx = nullptr;
  cplx *d_val = nullptr;
  DNM_CHECK_CUDA(cudaMalloc(&d_idx, sizeof(int64_t) * count));
  DNM_CHECK_CUDA(cudaMalloc(&d_val, sizeof(cplx) * count));
  DNM_CHECK_CUDA(cudaMemcpyAsync(d_idx, local_idx, sizeof(int64_t) * count, cudaMemcpyHostToDevice, G.stream));
  k_gather_get<<<(int)std::min<int64_t>((count + 255) / 256, 1024), 256, 0, G.stream>>>(v->d, count, d_idx, d_val);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
  DNM_CHECK_CUDA(cudaMemcpyAsync(values, d_val, sizeof(cplx) * count, cudaMemcpyDeviceToHost, G.stream));
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  cudaFree(d_idx);
  cudaFree(d_val);
  DNM_API_END
}

extern "C" int dnm_vec_set(dnm_vec_t v, double re, double im)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v, DNM_ERR_ARG, "null vector");
  vec_fill(v->d, v->local_n, make_double2(re, im));
  DNM_API_END
}

extern "C" int dnm_vec_set_random(dnm_vec_t v, uint64_t seed)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v, DNM_ERR_ARG, "null vector");
  vec_random_fill(v->d, v->local_n, v->local_start, seed);
  DNM_API_END
}

extern "C" int dnm_vec_copy(dnm_vec_t src, dnm_vec_t dst)
{
  DNM_API_BEGIN
  require_init();
  check_same_layout(src, dst);
  vec_copy(dst->d, src->d, src->local_n);
  DNM_API_END
}

extern "C" int dnm_vec_scale(dnm_vec_t v, double re, double im)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v, DNM_ERR_ARG, "null vector");
  vec_scale(v->d, v->local_n, make_double2(re, im));
  DNM_API_END
}

namespace {
__global__ void k_shift(cplx *__restrict__ v, int64_t n, cplx a)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    cplx t = v[i];
    t.x += a.x;
    t.y += a.y;
    v[i] = t;
  }
}
__global__ void k_sqrt_scalar(double *v) { *v = sqrt(*v); }
}  // namespace

// Vec.shift: v[i] += a for every entry (states.py:805)
extern "C" int dnm_vec_shift(dnm_vec_t v, double re, double im)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v, DNM_ERR_ARG, "null vector");
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((v->local_n + 255) / 256, (int64_t)G.sm_count * 8));
  k_shift<<<grid, 256, 0, G.stream>>>(v->d, v->local_n, make_double2(re, im));
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
  DNM_API_END
}

// Vec.normalize: v /= ||v||_2 with the norm kept on the device between the reduction and the scaling;
// returns the norm (a zero vector stays zero)
extern "C" int dnm_vec_normalize(dnm_vec_t v, double *norm_out)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v, DNM_ERR_ARG, "null vector");
  vec_sqnorm_dev(v->d, v->local_n, G.d_scratch);
  k_sqrt_scalar<<<1, 1, 0, G.stream>>>(G.d_scratch);
  count_launch();
  vec_scale_dev(v->d, v->local_n, G.d_scratch, true);
  double nrm = 0;
  fetch_doubles(G.d_scratch, &nrm, 1);
  if (norm_out) *norm_out = nrm;
  DNM_API_END
}

extern "C" int dnm_vec_axpby(dnm_vec_t y, double a_re, double a_im, double b_re, double b_im, dnm_vec_t x)
{
  DNM_API_BEGIN
  require_init();
  check_same_layout(x, y);
  vec_axpby(y->d, x->d, y->local_n, make_double2(a_re, a_im), make_double2(b_re, b_im));
  DNM_API_END
}

extern "C" int dnm_vec_dot(dnm_vec_t x, dnm_vec_t y, double out[2])
{
  DNM_API_BEGIN
  require_init();
  check_same_layout(x, y);
  DNM_REQUIRE(out, DNM_ERR_ARG, "null pointer");
  vec_dot_dev(x->d, y->d, x->local_n, G.d_scratch);
  fetch_doubles(G.d_scratch, out, 2);
  DNM_API_END
}

extern "C" int dnm_vec_norm(dnm_vec_t v, int type, double *out)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(v && out, DNM_ERR_ARG, "null pointer");
  DNM_REQUIRE(type >= 0 && type <= 2, DNM_ERR_ARG, "norm type must be 0 (2-norm), 1 (1-norm) or 2 (infinity)");
  if (type == 0) {
    vec_sqnorm_dev(v->d, v->local_n, G.d_scratch);
    fetch_doubles(G.d_scratch, out, 1);
    *out = sqrt(*out);
  } else {
    vec_norm_other_dev(v->d, v->local_n, type, G.d_scratch);
    fetch_doubles(G.d_scratch, out, 1);
  }
  DNM_API_END
}
