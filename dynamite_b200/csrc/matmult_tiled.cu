// Bit-window tiled MatMult for XOR-structured subspaces.
//
// For Full->Full and same-sector Parity->Parity (optionally under XParity) the
// product is   y[i] = sum_m D_m(i) * x[i ^ m]   over an index space of n bits,
// where D_m(i) = sum_t c_t (-1)^popcount(s_t & i) (row form; the reference's
// CPU fast path uses the same rewrite, bpetsc_template_2.c:844-857).
//
// Reading x once per mask from HBM costs (M+1)*N*16 bytes (the north-star
// traffic model).  Instead a CTA stages a TILE of 2^T amplitudes in shared
// memory: the tile is the set of indices that agree on all bits outside a
// WINDOW of T bit positions (the B lowest bits, for coalescing, plus T-B
// arbitrary higher bits).  Every mask whose set bits lie inside the window is
// then served from shared memory.  A host planner covers all masks with a
// small number of windows (passes); pass 0 writes y, later passes accumulate.
// HBM traffic drops from (M+1) to about 3*passes-1 vector sweeps.
//
// Sharding (one process per GPU): rank r owns the indices whose top p bits are
// r.  A mask with high part h reads x from rank r^h: the same kernel runs with
// its x pointer set to that peer's CUDA-IPC mapping, so remote amplitudes
// arrive by NVLink loads inside the MatMult kernel, tile by tile.
#include "matmult_tiled.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <memory>

#include "jit.h"
#include "tiled_kernel.cuh"
#include "vecops.cuh"

namespace dnm {

namespace {

using namespace tiled;

// stream the launch helpers below enqueue on (the side stream while remote passes are issued)
cudaStream_t g_launch_stream = nullptr;
// peer mappings of the current input vector, indexed by partner rank XOR (folded remote masks)
const cplx *g_peers[MAX_RANKS] = {nullptr};
inline cudaStream_t launch_stream() { return g_launch_stream ? g_launch_stream : G.stream; }

// Plain gather for masks no window can hold, and for index spaces smaller
// than one tile.  Terms in global index coordinates (sw/rb unused).
struct DirectParams {
  int nmasks;
  int accumulate;
  const i64 *mloc;  // [nmasks] local part of the mask
  const int *t_re, *t_im, *t_end;
  const i64 *so;  // full sign mask (global index coordinates)
  const double *cf;
  i64 rank_bits;
};

__global__ void __launch_bounds__(256)
    k_xor_direct(const DirectParams P, const cplx *__restrict__ x, cplx *__restrict__ y, const double *__restrict__ diag,
                 i64 nloc)
{
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < nloc; i += (i64)gridDim.x * blockDim.x) {
    const i64 ig = i | P.rank_bits;
    double ar = 0.0, ai = 0.0;
    if (diag != nullptr) {
      const cplx v = x[i];
      const double d = diag[i];
      ar = d * v.x;
      ai = d * v.y;
    }
    for (int mi = 0; mi < P.nmasks; ++mi) {
      double dr = 0.0, di = 0.0;
      int t = __ldg(&P.t_re[mi]);
      const int t1 = __ldg(&P.t_im[mi]), t2 = __ldg(&P.t_end[mi]);
      for (; t < t1; ++t) {
        const double c = __ldg(&P.cf[t]);
        dr += (__popcll((unsigned long long)(__ldg(&P.so[t]) & ig)) & 1) ? -c : c;
      }
      for (; t < t2; ++t) {
        const double c = __ldg(&P.cf[t]);
        di += (__popcll((unsigned long long)(__ldg(&P.so[t]) & ig)) & 1) ? -c : c;
      }
      const cplx v = x[i ^ __ldg(&P.mloc[mi])];
      ar += dr * v.x - di * v.y;
      ai += dr * v.y + di * v.x;
    }
    if (P.accumulate) {
      const cplx old = y[i];
      ar += old.x;
      ai += old.y;
    }
    y[i] = make_double2(ar, ai);
  }
}

// row-local helpers on the normalised terms (used when sharded)
__global__ void __launch_bounds__(256) k_xor_diag(const DirectParams P, double *__restrict__ diag, i64 nloc)
{
  const int t2 = P.t_end[0];
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < nloc; i += (i64)gridDim.x * blockDim.x) {
    const i64 ig = i | P.rank_bits;
    double v = 0.0;
    for (int t = 0; t < t2; ++t) {
      const double c = __ldg(&P.cf[t]);
      v += (__popcll((unsigned long long)(__ldg(&P.so[t]) & ig)) & 1) ? -c : c;
    }
    diag[i] = v;
  }
}

__global__ void __launch_bounds__(256) k_xor_norm(const DirectParams P, double *__restrict__ partials, i64 nloc)
{
  double best = 0.0;
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < nloc; i += (i64)gridDim.x * blockDim.x) {
    const i64 ig = i | P.rank_bits;
    double sum = 0.0, err = 0.0;
    for (int mi = 0; mi < P.nmasks; ++mi) {
      double dr = 0.0, di = 0.0;
      int t = __ldg(&P.t_re[mi]);
      const int t1 = __ldg(&P.t_im[mi]), t2 = __ldg(&P.t_end[mi]);
      for (; t < t1; ++t) {
        const double c = __ldg(&P.cf[t]);
        dr += (__popcll((unsigned long long)(__ldg(&P.so[t]) & ig)) & 1) ? -c : c;
      }
      for (; t < t2; ++t) {
        const double c = __ldg(&P.cf[t]);
        di += (__popcll((unsigned long long)(__ldg(&P.so[t]) & ig)) & 1) ? -c : c;
      }
      const double comp = __dsub_rn(hypot(dr, di), err);
      const double total = __dadd_rn(sum, comp);
      err = __dsub_rn(__dsub_rn(total, sum), comp);
      sum = total;
    }
    best = fmax(best, sum);
  }
  __shared__ double sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) best = fmax(best, sh[w]);
    partials[blockIdx.x] = best;
  }
}

__global__ void k_max_final(const double *__restrict__ partials, int n, double *__restrict__ out)
{
  double best = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) best = fmax(best, partials[i]);
  __shared__ double sh[256];
  sh[threadIdx.x] = best;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// ---- host side: normalisation and pass planning ------------------------------

struct NTerm {
  i64 sign;     // index-space sign mask
  double coef;  // non-zero part, row-form sign folded in
  bool imag;
};

struct NMask {
  i64 mask;  // index-space flip mask (global, includes rank bits)
  std::vector<NTerm> terms;
};

struct Pass {
  PassParams p{};
  SmallTables st{};
  bool small = false;
  int T = 0;
  int R = 16;
  int peer_xor = 0;  // x is read from rank ^ peer_xor
  int nterms = 0;
  int nmasks = 0;
  int nfar = 0;      // masks served through the L2 window (FAR groups)
  int nremote = 0;   // ... of which read another rank's shard (folded remote masks: generated kernels only)
  // what the pass was built from (so that the planner can rebuild it with the folded remote masks)
  std::vector<const NMask *> src_local, src_fold;
  size_t src_nfar = 0;
  i64 Fbits = 0;
  int Bbits = 0;
  // folded remote masks are evaluated only on the tiles whose index bit `filter_bit` equals `filter_val`
  // (the other half of the rows is served by another pass: the NVLink volume is spread over two passes)
  int filter_bit = -1, filter_val = 0;
  i64 wbits = 0;     // window bit positions
  std::vector<int> W;               // the same as a list: W[j] = index bit of window coordinate j
  const jit::Kernel *jk = nullptr;  // operator-specialised kernel of this pass (owned by TiledPlan::jit)
  std::vector<void *> owned;
};

struct Direct {
  DirectParams p{};
  int peer_xor = 0;
  std::vector<void *> owned;
};

}  // namespace

// A launch unit: one pass
struct Unit {
  std::vector<int> passes;  // indices into TiledPlan::passes
};

struct TiledPlan {
  int n = 0;      // index bits (global)
  int nloc = 0;   // index bits on this rank
  bool use_diag = false;
  std::vector<Pass> passes;
  std::vector<Unit> units;
  std::vector<Direct> directs;
  Direct all;  // every mask, for the row-local helpers (diag, norm)
  bool any_remote = false;
  bool overlap = false;      // (peer-load mode) remote passes accumulate into y_remote on the side stream
  cplx *y_remote = nullptr;  // [local rows], allocated on first use
  PassParams *d_batch = nullptr;  // all passes' parameters, when the whole product runs as one batched launch
  size_t batch_stage_bytes = 0;   // shared memory for the largest staged term table among them
  int batch_T = 0, batch_R = 0;
  jit::Module *jit = nullptr;  // generated kernels of the lean passes
  bool fold = false;           // remote masks are FAR groups of local passes (generated kernels only)
  bool fold_failed = false;    // ... but no lean pass could take them: plan again without folding
  bool dma = false;          // remote shards are staged by the copy engines while the local passes run
  cplx *stage[2] = {nullptr, nullptr};
  cudaEvent_t ev_staged[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  double cost = 0.0;  // estimated HBM sweeps per MatMult
  // which part of a partner's shard this rank reads at all (DMA staging copies only that):
  // kind 0 = everything, 1 = nothing (the group's passes are skipped too), 2 = [first, first+count)
  struct Need {
    int kind = 0;
    i64 first = 0, count = 0;
  };
  std::map<int, Need> need;
  ~TiledPlan()
  {
    for (auto &ps : passes)
      for (void *q : ps.owned) cudaFree(q);
    for (auto &d : directs)
      for (void *q : d.owned) cudaFree(q);
    for (void *q : all.owned) cudaFree(q);
    delete jit;
    if (y_remote) cudaFree(y_remote);
    if (d_batch) cudaFree(d_batch);
    for (int b = 0; b < 2; ++b) {
      if (stage[b]) cudaFree(stage[b]);
      if (ev_staged[b]) cudaEventDestroy(ev_staged[b]);
      if (ev_free[b]) cudaEventDestroy(ev_free[b]);
    }
  }
};

namespace {

int ilog2(i64 v)
{
  int n = 0;
  while (((i64)1 << n) < v) ++n;
  return n;
}

// dnm_jit_dryrun: plan and generate on the host only (no device allocations, no module load), so that
// the generator can be exercised -- and its output compiled by NVRTC -- on a machine without a GPU
bool g_plan_host_only = false;
struct DryRun {
  std::string src, log;
  size_t cubin_bytes = 0;
  int kernels = 0, remote_groups = 0, passes = 0, pipelined = 0;
  bool host = false;
} g_dry;

template <class Tv>
Tv *up(const std::vector<Tv> &h, std::vector<void *> &owned)
{
  if (g_plan_host_only) return nullptr;
  Tv *d = nullptr;
  const size_t bytes = sizeof(Tv) * std::max<size_t>(h.size(), 1);
  DNM_CHECK_CUDA(cudaMalloc(&d, bytes));
  owned.push_back(d);
  if (!h.empty()) DNM_CHECK_CUDA(cudaMemcpyAsync(d, h.data(), sizeof(Tv) * h.size(), cudaMemcpyHostToDevice, G.stream));
  return d;
}

// Rewrite the MSC terms in index space (see file header).
std::vector<NMask> normalise(const dnm_mat_s *A, int n)
{
  const bool par = A->left.desc.type == DNM_PARITY;
  const i64 space = A->left.desc.space;
  const i64 nmask = ((i64)1 << n) - 1;
  std::vector<NMask> out;
  for (size_t k = 0; k < A->masks.size(); ++k) {
    const i64 m = A->masks[k];
    if (par && parity64(m)) continue;  // leaves the sector (bpetsc_template_2.c:822-827)
    NMask nm;
    nm.mask = par ? (m >> 1) : m;
    DNM_REQUIRE((nm.mask & ~nmask) == 0, DNM_ERR_ARG,
                "mask 0x%llx flips a bit outside the %d-bit index space (XParity operators must be reduced first)",
                (unsigned long long)m, n);
    for (i64 t = A->mask_offsets[k]; t < A->mask_offsets[k + 1]; ++t) {
      const i64 s = A->signs[t];
      NTerm nt;
      nt.imag = parity64(m & s) != 0;
      double c = nt.imag ? A->coeffs[2 * t + 1] : A->coeffs[2 * t];
      i64 sp;
      if (par) {
        sp = s >> 1;
        if (s & 1) {  // bit 0 of the state is parity(idx)^space
          sp ^= nmask;
          if (space) c = -c;
        }
      } else {
        sp = s;
      }
      sp &= nmask;
      if (parity64(sp & nm.mask)) c = -c;  // column-state sign -> row-state sign
      nt.sign = sp;
      nt.coef = c;
      nm.terms.push_back(nt);
    }
    out.push_back(std::move(nm));
  }
  return out;
}

void fill_term_ranges(const std::vector<const NMask *> &masks, std::vector<int> &t_re, std::vector<int> &t_im,
                      std::vector<int> &t_end, std::vector<const NTerm *> &flat)
{
  for (const NMask *nm : masks) {
    t_re.push_back((int)flat.size());
    for (const NTerm &t : nm->terms)
      if (!t.imag) flat.push_back(&t);
    t_im.push_back((int)flat.size());
    for (const NTerm &t : nm->terms)
      if (t.imag) flat.push_back(&t);
    t_end.push_back((int)flat.size());
  }
}

Direct make_direct(const std::vector<const NMask *> &masks, int nloc, int accumulate)
{
  Direct d;
  std::vector<i64> mloc, so;
  std::vector<int> t_re, t_im, t_end;
  std::vector<double> cf;
  std::vector<const NTerm *> flat;
  const i64 lmask = ((i64)1 << nloc) - 1;
  for (const NMask *nm : masks) mloc.push_back(nm->mask & lmask);
  fill_term_ranges(masks, t_re, t_im, t_end, flat);
  for (const NTerm *t : flat) {
    so.push_back(t->sign);
    cf.push_back(t->coef);
  }
  d.p.nmasks = (int)masks.size();
  d.p.accumulate = accumulate;
  d.p.mloc = up(mloc, d.owned);
  d.p.t_re = up(t_re, d.owned);
  d.p.t_im = up(t_im, d.owned);
  d.p.t_end = up(t_end, d.owned);
  d.p.so = up(so, d.owned);
  d.p.cf = up(cf, d.owned);
  d.p.rank_bits = (i64)G.rank << nloc;
  return d;
}

// runs of consecutive outer positions, for tile_outer_bits
void fill_segments(PassParams &p)
{
  p.n_seg = 0;
  int k = 0;
  while (k < p.n_outer) {
    int e = k + 1;
    while (e < p.n_outer && p.outer_pos[e] == p.outer_pos[e - 1] + 1) ++e;
    if (p.n_seg == MAX_SEGS) {
      p.n_seg = -1;
      return;
    }
    p.seg_mask[p.n_seg] = (((unsigned long long)1 << (e - k)) - 1ull) << k;
    p.seg_shift[p.n_seg] = (signed char)((int)p.outer_pos[k] - k);
    ++p.n_seg;
    k = e;
  }
}

// `masks` = the masks inside the window followed by `nfar` FAR masks, which also flip outer
// positions inside `farbits` (the pass's L2 window: those positions become the lowest tile-number
// bits, so the tiles a far mask connects run at the same time and its operand is an L2 hit).
Pass make_pass(const std::vector<const NMask *> &masks, const std::vector<int> &W, int T, int R, int B, int nloc,
               int accumulate, size_t nfar = 0, i64 farbits = 0)
{
  Pass ps;
  ps.T = T;
  ps.R = R;
  int log_r = 0;
  while ((1 << log_r) < R) ++log_r;
  const int log_nt = T - log_r;
  const i64 lmask = ((i64)1 << nloc) - 1;
  i64 wbits = 0;
  for (int b : W) wbits |= (i64)1 << b;

  auto extract = [&](i64 v) {
    u32 o = 0;
    for (int b = 0; b < T; ++b)
      if ((v >> W[b]) & 1) o |= 1u << b;
    return o;
  };
  auto row_pattern = [&](u32 w) {
    u32 bits = 0;
    for (int r = 0; r < R; ++r)
      if (__builtin_parity((w >> log_nt) & (u32)r)) bits |= 1u << r;
    return bits;
  };

  // groups: (mask, real|imaginary); inside a group the terms whose sign mask has no
  // row bits come first
  std::vector<u32> lam, sw, rb, toff;
  std::vector<u16> t0, t1, t2, pat;
  std::vector<u8> kp, role;
  std::vector<u16> cls;  // [ngroups * 8]: PATH_WHT groups, end of each row class inside t1..t2
  std::vector<i64> so;
  std::vector<double> cf, tabs;
  std::vector<unsigned long long> rpat;
  // coefficient tables (PATH_TABLE) are only worth it, and only supported, in passes that are too
  // big for the constant-memory variant anyway
  size_t general_terms = 0, general_groups = 0;
  for (const NMask *nm : masks) {
    bool re = false, im = false;
    for (const NTerm &t : nm->terms) (t.imag ? im : re) = true;
    general_groups += (re ? 1 : 0) + (im ? 1 : 0);
    general_terms += nm->terms.size();
  }
  const bool allow_tables = R <= 8 && !(general_groups <= (size_t)SMALL_GROUPS && general_terms <= (size_t)SMALL_TERMS) &&
                            getenv("DNM_NO_TABLES") == nullptr;
  // small passes (tables in kernel-parameter memory) use the lean PATH_PAIR encoding wherever a
  // group has at most two distinct sign masks; it may add one padding entry per group
  const bool small_pass = general_groups <= (size_t)SMALL_GROUPS && general_terms + general_groups <= (size_t)SMALL_TERMS &&
                          getenv("DNM_NO_PAIR") == nullptr;
  bool any_table = false;
  // GF(2) basis (at most 6 vectors) of the sign masks of `all`; coords[t] = which basis vectors XOR
  // to term t's mask
  auto gf2_basis = [&](const std::vector<const NTerm *> &all, std::vector<i64> &basis, std::vector<u32> &coords) -> bool {
    std::vector<i64> reduced;     // eliminated forms of the basis vectors
    std::vector<u32> red_coord;   // their coordinates in `basis`
    coords.assign(all.size(), 0);
    for (size_t ti = 0; ti < all.size(); ++ti) {
      i64 v = all[ti]->sign;
      u32 c = 0;
      for (bool changed = true; changed;) {  // reduced[] is not kept in echelon order: iterate to a fixed point
        changed = false;
        for (size_t k = 0; k < reduced.size(); ++k) {
          const i64 top = (i64)1 << (63 - __builtin_clzll((unsigned long long)reduced[k]));
          if (v & top) {
            v ^= reduced[k];
            c ^= red_coord[k];
            changed = true;
          }
        }
      }
      if (v != 0) {
        if (basis.size() >= 6) return false;
        // new independent vector: the term's own mask joins the basis
        const u32 self = 1u << basis.size();
        basis.push_back(all[ti]->sign);
        reduced.push_back(v);
        red_coord.push_back(c ^ self);
        c = self;
      }
      coords[ti] = c;
    }
    return true;
  };
  // append the basis vectors as the group's "terms"; returns the per-row index bits (byte r = row group r)
  auto push_basis = [&](const std::vector<i64> &basis) -> unsigned long long {
    unsigned long long rp = 0;
    for (size_t k = 0; k < basis.size(); ++k) {
      const u32 w = extract(basis[k] & lmask);
      const u32 bits = row_pattern(w);
      sw.push_back(w);
      rb.push_back(bits);
      so.push_back(basis[k] & ~wbits);
      cf.push_back(0.0);
      for (int r = 0; r < R; ++r)
        if ((bits >> r) & 1u) rp |= (unsigned long long)1 << (8 * r + (int)k);
    }
    return rp;
  };
  std::vector<unsigned long long> gfar;  // per group: the local flip mask of a FAR group, else 0
  std::vector<u8> gisfar, gpeer;         // per group: FAR flag, partner rank xor (folded remote masks)
  int nremote = 0;
  for (size_t mi = 0; mi < masks.size(); ++mi) {
    const NMask *nm = masks[mi];
    gfar.resize(lam.size(), 0);
    gisfar.resize(lam.size(), 0);
    gpeer.resize(lam.size(), 0);
    const bool is_far = mi >= masks.size() - nfar;
    const int mpeer = (int)(nm->mask >> nloc);
    if (is_far && mpeer) ++nremote;
    const u32 l = extract(nm->mask & lmask);
    if (allow_tables && getenv("DNM_NO_CTABLE") == nullptr) {
      // masks with real AND imaginary terms: one joint basis, one table of complex coefficients,
      // one gather (SYK: 8 + 8 terms per mask, joint dimension 5)
      std::vector<const NTerm *> all;
      size_t nre = 0;
      for (const NTerm &t : nm->terms)
        if (!t.imag) all.push_back(&t);
      nre = all.size();
      for (const NTerm &t : nm->terms)
        if (t.imag) all.push_back(&t);
      std::vector<i64> basis;
      std::vector<u32> coords;
      if (nre > 0 && nre < all.size() && all.size() >= 4 && gf2_basis(all, basis, coords)) {
        const int d = (int)basis.size();
        lam.push_back(l);
        if (tabs.size() & 1) tabs.push_back(0.0);  // complex entries are read as 16-byte words
        toff.push_back((u32)tabs.size());
        t0.push_back((u16)sw.size());
        for (u32 p = 0; p < (1u << d); ++p) {
          double re = 0.0, im = 0.0;
          for (size_t ti = 0; ti < all.size(); ++ti)
            (ti < nre ? re : im) += (__builtin_parity(coords[ti] & p) ? -1.0 : 1.0) * all[ti]->coef;
          tabs.push_back(re);
          tabs.push_back(im);
        }
        rpat.push_back(push_basis(basis));
        t1.push_back((u16)sw.size());
        t2.push_back((u16)sw.size());
        pat.push_back(0);
        kp.push_back((u8)(PATH_CTABLE << 1));
        any_table = true;
        continue;
      }
    }
    for (int kind = 0; kind < 2; ++kind) {
      std::vector<const NTerm *> plain, rowdep;
      for (const NTerm &t : nm->terms) {
        if ((int)t.imag != kind) continue;
        (((extract(t.sign & lmask) >> log_nt) == 0) ? plain : rowdep).push_back(&t);
      }
      if (plain.empty() && rowdep.empty()) continue;
      lam.push_back(l);
      toff.push_back(0);
      rpat.push_back(0);
      if (small_pass) {
        // distinct sign masks of the group, coefficients of equal masks summed
        std::vector<std::pair<i64, double>> uniq;
        for (const NTerm &t : nm->terms) {
          if ((int)t.imag != kind) continue;
          bool found = false;
          for (auto &u : uniq)
            if (u.first == t.sign) {
              u.second += t.coef;
              found = true;
            }
          if (!found) uniq.emplace_back(t.sign, t.coef);
        }
        if (uniq.size() <= 2) {
          role.resize(sw.size(), 0);  // entries of earlier, ordinary groups
          if (sw.size() & 1) {  // the (c1+c2, c1-c2) scratch pair is read as one 16-byte word
            sw.push_back(0);
            rb.push_back(0);
            so.push_back(0);
            cf.push_back(0.0);
            role.push_back(0);
          }
          if (uniq.size() == 1) uniq.emplace_back(uniq[0].first, 0.0);
          const i64 s1 = uniq[0].first, s2 = uniq[1].first;
          t0.push_back((u16)sw.size());
          const u32 wa = extract(s1 & lmask), wb = extract((s1 ^ s2) & lmask);
          sw.push_back(wa);
          rb.push_back(row_pattern(wa));
          so.push_back(s1 & ~wbits);
          cf.push_back(uniq[0].second);
          role.push_back(1);
          sw.push_back(wb);
          rb.push_back(row_pattern(wb));
          so.push_back(s2 & ~wbits);
          cf.push_back(uniq[1].second);
          role.push_back(2);
          t1.push_back((u16)sw.size());
          t2.push_back((u16)sw.size());
          pat.push_back(0);
          kp.push_back((u8)(kind | (PATH_PAIR << 1)));
          gfar.push_back(is_far ? (unsigned long long)(nm->mask & lmask) : 0ull);
          gisfar.push_back(is_far ? 1 : 0);
          gpeer.push_back(is_far ? (u8)mpeer : 0);
          continue;
        }
      }
      t0.push_back((u16)sw.size());
      if (allow_tables && plain.size() + rowdep.size() >= 4) {
        std::vector<const NTerm *> all(plain);
        all.insert(all.end(), rowdep.begin(), rowdep.end());
        std::vector<i64> basis;
        std::vector<u32> coords;
        if (gf2_basis(all, basis, coords)) {
          const int d = (int)basis.size();
          toff.back() = (u32)tabs.size();
          for (u32 p = 0; p < (1u << d); ++p) {
            double acc = 0.0;
            for (size_t ti = 0; ti < all.size(); ++ti) acc += (__builtin_parity(coords[ti] & p) ? -1.0 : 1.0) * all[ti]->coef;
            tabs.push_back(acc);
          }
          rpat.back() = push_basis(basis);
          t1.push_back((u16)sw.size());
          t2.push_back((u16)sw.size());
          pat.push_back(0);
          kp.push_back((u8)(kind | (PATH_TABLE << 1)));
          any_table = true;
          continue;
        }
      }
      for (const NTerm *t : plain) {
        sw.push_back(extract(t->sign & lmask));
        rb.push_back(0);
        so.push_back(t->sign & ~wbits);  // outside the window, rank bits included
        cf.push_back(t->coef);
      }
      t1.push_back((u16)sw.size());
      // many row-dependent terms (SYK two-flip masks: ~36 per group, GF(2) rank > 6): order them by
      // row class rho = their sign bits on the row positions; the kernel sums each class into one
      // per-thread scalar and a size-R Walsh-Hadamard transform turns those into the row coefficients
      const bool use_wht = !small_pass && R == 8 && rowdep.size() >= 6 && getenv("DNM_NO_WHT") == nullptr;
      if (use_wht)
        std::stable_sort(rowdep.begin(), rowdep.end(), [&](const NTerm *a, const NTerm *b) {
          return (extract(a->sign & lmask) >> log_nt) < (extract(b->sign & lmask) >> log_nt);
        });
      bool one_pattern = true;
      u32 first_pat = 0;
      for (size_t k = 0; k < rowdep.size(); ++k) {
        const u32 w = extract(rowdep[k]->sign & lmask);
        const u32 bits = row_pattern(w);
        if (k == 0) first_pat = bits;
        else if (bits != first_pat) one_pattern = false;
        sw.push_back(w);
        rb.push_back(bits);
        so.push_back(rowdep[k]->sign & ~wbits);
        cf.push_back(rowdep[k]->coef);
      }
      t2.push_back((u16)sw.size());
      if (use_wht && !one_pattern) {
        cls.resize(lam.size() * 8, 0);
        u16 *c = &cls[(lam.size() - 1) * 8];
        size_t k = 0;
        for (int rho = 0; rho < 8; ++rho) {
          while (k < rowdep.size() && (int)(extract(rowdep[k]->sign & lmask) >> log_nt) <= rho) ++k;
          c[rho] = (u16)(t1.back() + k);
        }
      }
      const int path = rowdep.empty() ? PATH_SCALAR : (one_pattern ? PATH_TWO : (use_wht ? PATH_WHT : PATH_GENERAL));
      pat.push_back((u16)first_pat);
      kp.push_back((u8)(kind | (path << 1)));
    }
  }
  DNM_REQUIRE(sw.size() < 65535, DNM_ERR_UNSUPPORTED, "too many terms in one pass (%zu)", sw.size());
  gfar.resize(lam.size(), 0);
  gisfar.resize(lam.size(), 0);
  gpeer.resize(lam.size(), 0);

  std::vector<i64> rowoff((size_t)1 << (T - B));
  for (size_t h = 0; h < rowoff.size(); ++h) {
    i64 off = 0;
    for (int b = B; b < T; ++b)
      if ((h >> (b - B)) & 1) off |= (i64)1 << W[b];
    rowoff[h] = off;
  }
  role.resize(sw.size(), 0);
  // the same offsets as runs of consecutive window positions (opt-in arithmetic deposit in thread_base)
  ps.p.n_rseg = 0;
  if (getenv("DNM_ROWOFF_ARITH")) {
    int b = B;
    while (b < T) {
      int e = b + 1;
      while (e < T && W[e] == W[e - 1] + 1) ++e;
      if (ps.p.n_rseg == MAX_SEGS) {
        ps.p.n_rseg = 0;  // too many runs: keep the table
        break;
      }
      ps.p.rseg_mask[ps.p.n_rseg] = ((1u << (e - b)) - 1u) << (b - B);
      ps.p.rseg_shift[ps.p.n_rseg] = (unsigned char)(W[b] - (b - B));
      ++ps.p.n_rseg;
      b = e;
    }
  }
  ps.p.ngroups = (int)lam.size();
  ps.p.nterms = (int)sw.size();
  ps.p.B = B;
  ps.p.accumulate = accumulate;
  ps.p.n_outer = 0;
  ps.p.far_bits = popc64(farbits);
  for (int b = 0; b < nloc; ++b)
    if ((farbits >> b) & 1) ps.p.outer_pos[ps.p.n_outer++] = (unsigned char)b;
  for (int b = 0; b < nloc; ++b)
    if (!(((wbits | farbits) >> b) & 1)) ps.p.outer_pos[ps.p.n_outer++] = (unsigned char)b;
  fill_segments(ps.p);
  ps.p.lam = up(lam, ps.owned);
  ps.p.t0 = up(t0, ps.owned);
  ps.p.t1 = up(t1, ps.owned);
  ps.p.t2 = up(t2, ps.owned);
  ps.p.pat = up(pat, ps.owned);
  ps.p.kp = up(kp, ps.owned);
  cls.resize(lam.size() * 8, 0);
  ps.p.cls = up(cls, ps.owned);
  ps.p.tabs = up(tabs, ps.owned);
  ps.p.toff = up(toff, ps.owned);
  ps.p.rpat = up(rpat, ps.owned);
  ps.p.sw = up(sw, ps.owned);
  ps.p.rb = up(rb, ps.owned);
  ps.p.so = up(so, ps.owned);
  ps.p.cf = up(cf, ps.owned);
  ps.p.rowoff = up(rowoff, ps.owned);
  ps.p.rank_bits = (i64)G.rank << nloc;
  for (int r = 0; r < 16; ++r) ps.p.roff[r] = 0;
  for (int r = 0; r < R; ++r) ps.p.roff[r] = rowoff[((size_t)r << log_nt) >> B];
  ps.nmasks = (int)masks.size();
  ps.nterms = (int)sw.size();
  ps.wbits = wbits;
  ps.W = W;
  ps.nfar = (int)nfar;
  ps.nremote = nremote;
  // large passes stage csign/sw/rb (16 B per term) in shared memory when that still leaves two tiles per SM
  ps.p.staged = ((any_table || !(ps.p.ngroups <= SMALL_GROUPS && ps.nterms <= SMALL_TERMS)) && (size_t)ps.nterms * 16 <= 40 * 1024 &&
                 getenv("DNM_NO_STAGE") == nullptr) ? 1 : 0;
  ps.small = !any_table && ps.p.ngroups <= SMALL_GROUPS && ps.nterms <= SMALL_TERMS;
  if (ps.small) {
    for (int g = 0; g < ps.p.ngroups; ++g) {
      ps.st.lam[g] = lam[g];
      ps.st.pat[g] = pat[g];
      ps.st.t0[g] = (u8)t0[g];
      ps.st.t1[g] = (u8)t1[g];
      ps.st.t2[g] = (u8)t2[g];
      ps.st.kp[g] = kp[g];
    }
    for (int t = 0; t < ps.nterms; ++t) {
      ps.st.sw[t] = sw[t];
      ps.st.rb[t] = rb[t];
      ps.st.so[t] = so[t];
      ps.st.cf[t] = cf[t];
      ps.st.role[t] = role[t];
    }
    bool lean = R <= 8;
    for (int g = 0; g < ps.p.ngroups && lean; ++g) lean = (kp[g] >> 1) == PATH_PAIR && t0[g] == 2 * g;
    if (lean)
      for (int g = 0; g < ps.p.ngroups; ++g) {
        ps.st.gd[g] = make_uint4(lam[g], sw[2 * g], sw[2 * g + 1],
                                 (rb[2 * g] & 0xffu) | ((rb[2 * g + 1] & 0xffu) << 8) | ((u32)(kp[g] & 1) << 16) |
                                     (gisfar[g] ? (1u << 17) : 0u));
        ps.st.far[g] = gfar[g];
        ps.st.peer[g] = gpeer[g];
      }
    if (lean && getenv("DNM_NO_CPLX") == nullptr)
      for (int g = 0; g + 1 < ps.p.ngroups; ++g) {
        const bool plain_pair = (ps.st.gd[g].w & 0xffffu) == 0 && (ps.st.gd[g + 1].w & 0xffffu) == 0 && !gisfar[g] && !gisfar[g + 1];
        if (lam[g] == lam[g + 1] && !(kp[g] & 1) && (kp[g + 1] & 1) && plain_pair) {
          ps.st.gd[g].x |= 0x80000000u;
          ++g;
        }
      }
    ps.p.lean = lean ? 1 : 0;
    if (const char *e = getenv("DNM_RING_DEBUG")) ps.p.debug = atoi(e);
  }
  return ps;
}

// Greedy window cover of one partner group's masks.
// Folded remote operands: loaded straight into registers by default; DNM_REMOTE_STAGE=1 stages them with
// cp.async in shared-memory buffers (one exposed NVLink round trip per tile, but fewer resident CTAs:
// measured slower at 4 GPUs, 54 against 42 ms at L=32 MBL).
bool stage_remote_operands() { return getenv("DNM_REMOTE_STAGE") && atoi(getenv("DNM_REMOTE_STAGE")) != 0; }

// a mask the lean group loop can serve: at most two distinct sign masks per (real | imaginary) part
bool pair_eligible(const NMask *nm)
{
  for (int kind = 0; kind < 2; ++kind) {
    std::vector<i64> uniq;
    for (const NTerm &t : nm->terms) {
      if ((int)t.imag != kind) continue;
      if (std::find(uniq.begin(), uniq.end(), t.sign) == uniq.end()) uniq.push_back(t.sign);
    }
    if (uniq.size() > 2) return false;
  }
  return true;
}

// `fold` (may be null): masks of the REMOTE groups to attach to the first lean pass of this (local)
// group as FAR groups that read the partner's shard through its peer mapping; emptied when used.
void plan_group(TiledPlan &plan, std::vector<const NMask *> remaining, int peer_xor, int T, int R, int B,
                bool first_group, int verbose, int fmax = 0, std::vector<const NMask *> *fold = nullptr)
{
  const int nloc = plan.nloc;
  const i64 lmask = ((i64)1 << nloc) - 1;
  const i64 lowbits = ((i64)1 << B) - 1;
  bool wrote = !first_group;  // group 0's first pass overwrites y
  std::vector<const NMask *> leftovers;

  while (!remaining.empty() || !wrote) {
    i64 W = lowbits;
    int wsize = B;
    std::vector<const NMask *> chosen;
    std::vector<char> taken(remaining.size(), 0);
    for (;;) {
      int best = -1, best_new = 1 << 30;
      i64 best_bits = 0;
      for (size_t k = 0; k < remaining.size(); ++k) {
        if (taken[k]) continue;
        const i64 extra = (remaining[k]->mask & lmask) & ~W;
        const int nnew = popc64(extra);
        if (wsize + nnew > T) continue;
        if (nnew < best_new || (nnew == best_new && extra < best_bits)) {
          best = (int)k;
          best_new = nnew;
          best_bits = extra;
          if (nnew == 0) break;
        }
      }
      if (best < 0) break;
      taken[best] = 1;
      chosen.push_back(remaining[best]);
      W |= best_bits;
      wsize += best_new;
    }
    if (chosen.empty() && !remaining.empty() && wrote) {
      // nothing fits any window of this size: gather these straight from global memory
      leftovers = remaining;
      remaining.clear();
      break;
    }
    if (chosen.empty() && !remaining.empty() && !wrote) {
      // need a writing pass first; emit an empty one (y = diag*x or 0)
    }
    // pad the window to exactly T bits with the lowest unused positions
    for (int b = 0; b < nloc && wsize < T; ++b)
      if (!((W >> b) & 1)) {
        W |= (i64)1 << b;
        ++wsize;
      }
    std::vector<int> Wpos;
    for (int b = 0; b < nloc; ++b)
      if ((W >> b) & 1) Wpos.push_back(b);
    // masks are kept in ascending order inside a pass
    std::sort(chosen.begin(), chosen.end(), [](const NMask *a, const NMask *b) { return a->mask < b->mask; });
    // FAR masks: also flip up to fmax outer positions (the pass's L2 window); the operand comes from
    // global memory while the tiles of one far-bit block are in flight together (lean passes only)
    i64 F = 0;
    std::vector<const NMask *> farm;
    std::vector<char> taken_far(remaining.size(), 0);
    bool chosen_ok = R <= 8;
    for (const NMask *nm : chosen) chosen_ok = chosen_ok && pair_eligible(nm);
    if (fmax > 0 && chosen_ok) {
      size_t groups = 0;
      for (const NMask *nm : chosen) groups += 2;
      for (;;) {
        int best = -1, best_new = 1 << 30;
        i64 best_bits = 0;
        for (size_t k = 0; k < remaining.size(); ++k) {
          if (taken[k] || taken_far[k] || !pair_eligible(remaining[k])) continue;
          const i64 extra = (remaining[k]->mask & lmask) & ~(W | F);
          const int nnew = popc64(extra);
          if (popc64(F) + nnew > fmax) continue;
          if (nnew < best_new || (nnew == best_new && extra < best_bits)) {
            best = (int)k;
            best_new = nnew;
            best_bits = extra;
          }
        }
        if (best < 0 || groups + 2 > (size_t)SMALL_GROUPS || 2 * (groups + 2) > (size_t)SMALL_TERMS) break;
        taken_far[best] = 1;
        farm.push_back(remaining[best]);
        F |= best_bits;
        groups += 2;
      }
      std::sort(farm.begin(), farm.end(), [](const NMask *a, const NMask *b) { return a->mask < b->mask; });
    }
    std::vector<const NMask *> both(chosen);
    both.insert(both.end(), farm.begin(), farm.end());
    Pass ps = make_pass(both, Wpos, T, R, B, nloc, wrote ? 1 : 0, farm.size(), F);
    if (!farm.empty() && !(ps.small && ps.p.lean)) {
      // the far path only exists in the lean group loop: plan this pass without it
      for (void *q : ps.owned) cudaFree(q);
      farm.clear();
      std::fill(taken_far.begin(), taken_far.end(), 0);
      F = 0;
      ps = make_pass(chosen, Wpos, T, R, B, nloc, wrote ? 1 : 0);
    }
    for (size_t k = 0; k < remaining.size(); ++k) taken[k] = taken[k] || taken_far[k];
    if (fold && !fold->empty() && chosen_ok && ps.small && ps.p.lean) {
      // spread the NVLink volume over the passes: half of the remote masks here when another pass follows
      bool more_passes = false;
      for (size_t k = 0; k < remaining.size(); ++k) more_passes = more_passes || !taken[k];
      // spread the NVLink volume: half of the remote masks here when another pass follows (measured at 4
      // GPUs, L=32 MBL: 42 ms against 54 ms with every remote mask in the first pass)
      size_t take = more_passes ? (fold->size() + 1) / 2 : fold->size();
      // every folded group stages its operand in a shared-memory buffer of its own next to the tile (one per
      // mask, two when its real and imaginary parts cannot share the fetch)
      {
        auto buffers = [](const NMask *nm) {
          int nre = 0, nim = 0;
          std::vector<i64> sre, sim;
          for (const NTerm &t : nm->terms) {
            std::vector<i64> &v = t.imag ? sim : sre;
            if (std::find(v.begin(), v.end(), t.sign) == v.end()) v.push_back(t.sign);
            (t.imag ? nim : nre) += 1;
          }
          if (nre && nim && !(sre.size() == 1 && sim.size() == 1)) return 2;
          return 1;
        };
        const size_t room = stage_remote_operands() ? (size_t)(200 * 1024) / ((size_t)16 << T) - 1 : (size_t)64;
        size_t used = 0, fit = 0;
        while (fit < take && used + buffers((*fold)[fit]) <= room) used += buffers((*fold)[fit++]);
        take = fit;
      }
      while (take > 0 && ((size_t)ps.p.ngroups + 2 * take > (size_t)SMALL_GROUPS || (size_t)ps.nterms + 4 * take > (size_t)SMALL_TERMS))
        --take;
      if (take > 0) {
        std::vector<const NMask *> all3(chosen);
        all3.insert(all3.end(), farm.begin(), farm.end());
        all3.insert(all3.end(), fold->begin(), fold->begin() + take);
        Pass folded = make_pass(all3, Wpos, T, R, B, nloc, wrote ? 1 : 0, farm.size() + take, F);
        if (folded.small && folded.p.lean) {
          for (void *q : ps.owned) cudaFree(q);
          ps = std::move(folded);
          ps.src_fold.assign(fold->begin(), fold->begin() + take);
          fold->erase(fold->begin(), fold->begin() + take);
        } else {
          for (void *q : folded.owned) cudaFree(q);
        }
      }
    }
    ps.peer_xor = peer_xor;
    ps.src_local = chosen;
    ps.src_local.insert(ps.src_local.end(), farm.begin(), farm.end());
    ps.src_nfar = farm.size();
    ps.Fbits = F;
    ps.Bbits = B;
    if (verbose)
      fprintf(stderr, "[dnm] pass %zu: peer^%d window=0x%llx far=0x%llx masks=%d (%d far) terms=%d %s\n",
              plan.passes.size(), peer_xor, (unsigned long long)W, (unsigned long long)F, ps.nmasks, ps.nfar, ps.nterms,
              wrote ? "accumulate" : "write");
    plan.passes.push_back(std::move(ps));
    wrote = true;
    std::vector<const NMask *> rest;
    for (size_t k = 0; k < remaining.size(); ++k)
      if (!taken[k]) rest.push_back(remaining[k]);
    remaining.swap(rest);
  }
  if (!leftovers.empty()) {
    Direct d = make_direct(leftovers, nloc, 1);
    d.peer_xor = peer_xor;
    if (verbose) fprintf(stderr, "[dnm] direct gather: peer^%d masks=%d\n", peer_xor, d.p.nmasks);
    plan.directs.push_back(std::move(d));
  }
}

struct PlanInputs {
  std::vector<NMask> masks;
};

template <int T, int R, bool SMALL>
void launch_tiled_v(const Pass &ps, const cplx *x, cplx *y, const double *diag, i64 ntiles)
{
  static bool attr_set = false;
  const size_t smem = (sizeof(double2) << T) + (ps.p.staged ? (size_t)ps.nterms * 16 : 0);
  if (!attr_set) {
    DNM_CHECK_CUDA(cudaFuncSetAttribute(k_tiled<T, R, SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((sizeof(double2) << T) + 40 * 1024)));
    DNM_CHECK_CUDA(cudaFuncSetAttribute(k_tiled<T, R, SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set = true;
  }
  k_tiled<T, R, SMALL><<<(unsigned)ntiles, TileCfg<T, R>::NT, smem, launch_stream()>>>(ps.p, ps.st, x, y, diag);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
}

template <int T, int R>
void launch_tiled(const Pass &ps, const cplx *x, cplx *y, const double *diag, i64 ntiles)
{
  if (ps.small) launch_tiled_v<T, R, true>(ps, x, y, diag, ntiles);
  else launch_tiled_v<T, R, false>(ps, x, y, diag, ntiles);
}

// rows per thread for a tile size: 16 by default (8 on request) where the CTA stays >= 64 threads
int rows_for(int T, int want)
{
  if (T == 8) return 4;
  if (T == 9) return 8;
  return (want == 16) ? 16 : 8;  // 8 rows/thread (64 registers) doubles the resident warps; measured faster
}

void launch_pass(const Pass &ps, const cplx *x, cplx *y, const double *diag, i64 ntiles)
{
  if (ps.jk) {
    jit::launch(*ps.jk, (unsigned long long)ntiles, G.sm_count, launch_stream(), x, y, diag, (long long)ps.p.rank_bits,
                g_peers);
    count_launch();
    return;
  }
#define DNM_TILE_CASE(TT, RR) \
  if (ps.T == TT && ps.R == RR) return launch_tiled<TT, RR>(ps, x, y, diag, ntiles);
  DNM_TILE_CASE(8, 4)
  DNM_TILE_CASE(9, 8)
  DNM_TILE_CASE(10, 8)
  DNM_TILE_CASE(10, 16)
  DNM_TILE_CASE(11, 8)
  DNM_TILE_CASE(11, 16)
  DNM_TILE_CASE(12, 8)
  DNM_TILE_CASE(12, 16)
  DNM_TILE_CASE(13, 8)
  DNM_TILE_CASE(13, 16)
#undef DNM_TILE_CASE
  DNM_REQUIRE(false, DNM_ERR_INTERNAL, "no tiled kernel for T=%d R=%d", ps.T, ps.R);
}

int direct_grid(i64 rows)
{
  const i64 want = (rows + 255) / 256;
  const i64 cap = (i64)G.sm_count * 16;
  return (int)std::max<i64>(1, std::min(want, cap));
}

template <int T, int R>
void launch_batch_tr(const TiledPlan &plan, const cplx *x, cplx *y, const double *diag)
{
  static bool attr_set = false;
  const size_t smem = (sizeof(double2) << T) + plan.batch_stage_bytes;
  if (!attr_set) {
    DNM_CHECK_CUDA(cudaFuncSetAttribute(k_tiled_batch<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((sizeof(double2) << T) + 40 * 1024)));
    // table-driven passes read their coefficient tables through the L1: leave it what the CTAs
    // of one SM (1024 threads) do not need as shared memory
    const int ctas = std::max(1, 1024 / TileCfg<T, R>::NT);
    int pct = (int)((ctas * (smem + 2048) * 100 + 233471) / 233472);
    if (const char *e = getenv("DNM_BATCH_CARVEOUT")) pct = atoi(e);
    DNM_CHECK_CUDA(cudaFuncSetAttribute(k_tiled_batch<T, R>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        std::min(100, std::max(pct, 1))));
    attr_set = true;
  }
  const dim3 grid((unsigned)((long long)1 << (plan.nloc - T)), (unsigned)plan.passes.size());
  k_tiled_batch<T, R><<<grid, TileCfg<T, R>::NT, smem, G.stream>>>(plan.d_batch, x, y, diag, plan.use_diag ? 0 : -1);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
}

void launch_batch(const TiledPlan &plan, const cplx *x, cplx *y, const double *diag)
{
#define DNM_BATCH_CASE(TT, RR) \
  if (plan.batch_T == TT && plan.batch_R == RR) return launch_batch_tr<TT, RR>(plan, x, y, diag);
  DNM_BATCH_CASE(8, 4)
  DNM_BATCH_CASE(9, 8)
  DNM_BATCH_CASE(10, 8)
  DNM_BATCH_CASE(10, 16)
  DNM_BATCH_CASE(11, 8)
  DNM_BATCH_CASE(11, 16)
  DNM_BATCH_CASE(12, 8)
  DNM_BATCH_CASE(12, 16)
  DNM_BATCH_CASE(13, 8)
  DNM_BATCH_CASE(13, 16)
#undef DNM_BATCH_CASE
  DNM_REQUIRE(false, DNM_ERR_INTERNAL, "no batched kernel for T=%d R=%d", plan.batch_T, plan.batch_R);
}

// one launch unit per pass
void build_units(TiledPlan &plan, int)
{
  for (size_t i = 0; i < plan.passes.size(); ++i) {
    Unit u;
    u.passes.push_back((int)i);
    plan.units.push_back(std::move(u));
  }
}

// Which amplitudes of the partner's shard a remote group can touch on THIS rank.  A group whose
// terms come in pairs of equal magnitude (XX+YY: c1 = +-c2) has D(I) = sigma(s1.I) c1 (1 + eps sigma(sB.I)),
// zero on the half space parity(sB & I) != e; when sB only involves rank bits the whole group is
// either needed or not, when its only local bit is the top one the needed half is contiguous.
TiledPlan::Need stage_need(const std::vector<const NMask *> &masks, int nloc)
{
  TiledPlan::Need all;
  const i64 lmask = ((i64)1 << nloc) - 1;
  const i64 top = (i64)1 << (nloc - 1);
  const i64 rank_bits = (i64)G.rank << nloc;
  bool have = false, none = true;
  int half = -1;
  for (const NMask *nm : masks) {
    for (int kind = 0; kind < 2; ++kind) {
      std::vector<std::pair<i64, double>> uniq;
      for (const NTerm &t : nm->terms) {
        if ((int)t.imag != kind) continue;
        bool found = false;
        for (auto &u : uniq)
          if (u.first == t.sign) {
            u.second += t.coef;
            found = true;
          }
        if (!found) uniq.emplace_back(t.sign, t.coef);
      }
      if (uniq.empty()) continue;
      have = true;
      if (uniq.size() != 2 || std::fabs(uniq[0].second) != std::fabs(uniq[1].second) || uniq[0].second == 0.0) return all;
      const i64 sb = uniq[0].first ^ uniq[1].first;
      const int e = (uniq[0].second == uniq[1].second) ? 0 : 1;  // rows with parity(sb & I) == e survive
      const int v = e ^ parity64(sb & rank_bits);                // ... i.e. parity(sb & i) == v on this rank
      const i64 sl = sb & lmask;
      if (sl == 0) {
        if (v == 0) return all;  // every local row survives
        continue;                // no local row survives: this (mask, kind) needs nothing
      }
      if (sl != top) return all;
      // rows with top bit == v read partner amplitudes with top bit == v ^ (mask's top bit)
      const int h = v ^ (int)(((nm->mask & lmask) >> (nloc - 1)) & 1);
      if (half >= 0 && half != h) return all;
      half = h;
      none = false;
    }
  }
  if (!have) return all;
  TiledPlan::Need nd;
  if (none) {
    nd.kind = 1;
  } else {
    nd.kind = 2;
    nd.first = half ? top : 0;
    nd.count = top;
  }
  return nd;
}

// One candidate plan for a fixed tile size T and run length 2^B.
std::unique_ptr<TiledPlan> plan_with(dnm_mat_s *A, const std::vector<NMask> &masks, int T, int R, int B, int fmax_in,
                                     int verbose, bool allow_fold = false)
{
  std::unique_ptr<TiledPlan> plan(new TiledPlan());
  const int n = ilog2(A->M);
  const int p = ilog2(G.nranks);
  plan->n = n;
  plan->nloc = n - p;
  plan->use_diag = A->d_diag != nullptr;
  const int nloc = plan->nloc;

  std::vector<const NMask *> all;
  for (const NMask &nm : masks) all.push_back(&nm);
  plan->all = make_direct(all, nloc, 0);

  // group by partner rank
  std::map<int, std::vector<const NMask *>> groups;
  groups[0];  // group 0 always exists: its first pass writes y
  for (const NMask &nm : masks) {
    if (plan->use_diag && nm.mask == 0) continue;  // served from the cached diagonal
    groups[(int)(nm.mask >> nloc)].push_back(&nm);
  }

  if (T < 8) {
    // index space smaller than the smallest tile: plain gather only
    std::vector<const NMask *> g0 = groups[0];
    Direct d = make_direct(g0, nloc, 0);
    plan->directs.push_back(std::move(d));
    for (auto &kv : groups) {
      if (kv.first == 0) continue;
      Direct r = make_direct(kv.second, nloc, 1);
      r.peer_xor = kv.first;
      plan->directs.push_back(std::move(r));
      plan->any_remote = true;
    }
  } else {
    // With remote groups, their passes run on a side stream into a separate buffer while the
    // local passes run (NVLink-bound and HBM-bound work overlap); the first remote pass then
    // WRITES that buffer, exactly like the first local pass writes y.
    // Remote groups: default = DMA staging (copy engines pull the partner shard over NVLink into a
    // local buffer while the SMs run the local passes; the group's passes then run on local memory).
    // DNM_REMOTE=peer keeps the in-kernel NVLink loads, DNM_REMOTE=peer_overlap runs them on a side stream.
    const char *rmode = getenv("DNM_REMOTE");
    std::string mode = rmode ? rmode : (allow_fold ? "fold" : "dma");
    // fold: the remote masks join a local pass of generated code as FAR groups whose operand is the
    // partner's shard, read over NVLink inside that pass -- no staging buffers, no extra
    // read-modify-write sweep over y per remote group
    std::vector<const NMask *> fold;
    if (mode == "fold" && allow_fold && groups.size() > 1) {
      for (auto it = groups.begin(); it != groups.end();) {
        if (it->first == 0) {
          ++it;
          continue;
        }
        bool ok = true;
        for (const NMask *nm : it->second) ok = ok && pair_eligible(nm);
        if (ok) {
          fold.insert(fold.end(), it->second.begin(), it->second.end());
          it = groups.erase(it);
        } else {
          ++it;
        }
      }
      plan->fold = !fold.empty();
      if (plan->fold) plan->any_remote = true;
    }
    if (mode == "fold") mode = "dma";  // whatever could not be folded
    plan->dma = groups.size() > 1 && mode == "dma";
    plan->overlap = groups.size() > 1 && mode == "peer_overlap";
    bool first_local = true, first_remote = true;
    // L2 window (FAR masks)
    const int fmax = std::max(0, std::min(fmax_in, nloc - T));
    for (auto &kv : groups) {
      if (kv.first == 0) {
        plan_group(*plan, kv.second, kv.first, T, R, B, first_local, verbose, fmax, &fold);
        first_local = false;
      } else {
        const size_t before = plan->passes.size(), dbefore = plan->directs.size();
        plan_group(*plan, kv.second, kv.first, T, R, B, plan->overlap && first_remote, verbose, fmax);
        first_remote = false;
        plan->any_remote = true;
        // partial staging is only safe where a zero coefficient never touches the operand (lean passes)
        bool lean_only = plan->directs.size() == dbefore;
        for (size_t k = before; k < plan->passes.size(); ++k) lean_only = lean_only && plan->passes[k].small && plan->passes[k].p.lean;
        if (plan->dma && lean_only && getenv("DNM_FULL_STAGE") == nullptr) {
          plan->need[kv.first] = stage_need(kv.second, nloc);
          if (verbose)
            fprintf(stderr, "[dnm] rank %d partner^%d: staging kind %d first %lld count %lld\n", G.rank, kv.first,
                    plan->need[kv.first].kind, (long long)plan->need[kv.first].first, (long long)plan->need[kv.first].count);
        }
      }
    }
    if (plan->fold && !fold.empty()) plan->fold_failed = true;  // no lean pass could take the remote masks
    // Optional (DNM_FOLD_SPLIT=1): a second lean pass evaluates the same folded masks on the other half
    // of the rows (split by an index bit outside both windows, i.e. per tile), which spreads the NVLink
    // volume over two passes.  Measured at 2 GPUs (L=31): MBL 25.7 -> 26.0 ms, long_range 80 -> 90 ms --
    // the extra operand buffers cost more resident CTAs than the balance gains, so it is off.
    if (plan->fold && !plan->fold_failed && getenv("DNM_FOLD_SPLIT") != nullptr) {
      int ia = -1, ib = -1;
      for (size_t k = 0; k < plan->passes.size(); ++k)
        if (!plan->passes[k].src_fold.empty()) ia = (int)k;
      size_t holders = 0;
      for (const Pass &ps : plan->passes) holders += ps.src_fold.empty() ? 0 : 1;
      if (ia >= 0 && holders == 1) {
        const Pass &pa = plan->passes[ia];
        for (size_t k = 0; k < plan->passes.size() && ib < 0; ++k) {
          const Pass &pb = plan->passes[k];
          if ((int)k == ia || !(pb.small && pb.p.lean && pb.peer_xor == 0 && pb.R <= 8)) continue;
          if ((size_t)pb.p.ngroups + 2 * pa.src_fold.size() > (size_t)SMALL_GROUPS ||
              (size_t)pb.nterms + 4 * pa.src_fold.size() > (size_t)SMALL_TERMS)
            continue;
          ib = (int)k;
        }
        if (ib >= 0) {
          const i64 lmask = ((i64)1 << nloc) - 1;
          const i64 common = ~(plan->passes[ia].wbits | plan->passes[ib].wbits) & lmask;
          int q = -1;
          for (int b = nloc - 1; b >= 0 && q < 0; --b)
            if ((common >> b) & 1) q = b;
          if (q >= 0) {
            Pass &pb = plan->passes[ib];
            std::vector<const NMask *> all3(pb.src_local);
            all3.insert(all3.end(), pa.src_fold.begin(), pa.src_fold.end());
            Pass nb = make_pass(all3, pb.W, pb.T, pb.R, pb.Bbits, nloc, pb.p.accumulate, pb.src_nfar + pa.src_fold.size(), pb.Fbits);
            if (nb.small && nb.p.lean) {
              nb.peer_xor = 0;
              nb.src_local = pb.src_local;
              nb.src_fold = pa.src_fold;
              nb.src_nfar = pb.src_nfar;
              nb.Fbits = pb.Fbits;
              nb.Bbits = pb.Bbits;
              nb.filter_bit = q;
              nb.filter_val = 1;
              for (void *p : pb.owned) cudaFree(p);
              pb = std::move(nb);
              plan->passes[ia].filter_bit = q;
              plan->passes[ia].filter_val = 0;
              if (verbose)
                fprintf(stderr, "[dnm] folded remote masks split over passes %d and %d by index bit %d\n", ia, ib, q);
            } else {
              for (void *p : nb.owned) cudaFree(p);
            }
          }
        }
      }
    }
  }
  // cost in vector sweeps over HBM: a writing pass reads x and writes y (2), an
  // accumulating pass also re-reads y (3); a direct gather re-reads x once per mask.
  // pipeline: 1 = the ring kernel wherever it exists; 0 (auto) and 2 = one tile per CTA.  Measured on
  // B200 (profiles/r01_ring_experiment.md): the arithmetic phase is bound by shared-memory
  // bandwidth, not by the fetches the ring hides, and the ring kernel is 5-20 % slower.
  build_units(*plan, verbose);
  double cost = 0.0;
  for (const Unit &u : plan->units) {
    // a fused unit reads x once and writes y once (plus the old y when it accumulates)
    cost += plan->passes[u.passes.front()].p.accumulate ? 3.0 : 2.0;
  }
  for (const Direct &d : plan->directs) cost += (d.p.accumulate ? 2.0 : 1.0) + d.p.nmasks;
  // measured on B200: 128 KB tiles (one CTA per SM) and 64-byte runs are each a few % slower per sweep
  if (T >= 13) cost *= 1.08;
  if (B <= 2) cost *= 1.02;
  plan->cost = cost;
  return plan;
}

// shapes tried by the first-use autotuner (tile bits, run bits, L2-window positions); -1 in T = the
// generic-kernel default plan
struct TuneShape {
  int T, B, f;
};
const TuneShape TUNE_SHAPES[] = {{11, 4, 10}, {12, 4, 7}, {13, 4, 0}, {12, 4, 0}, {-1, 0, 0}};
constexpr int N_TUNE_SHAPES = 5;

TiledPlan *build_plan(dnm_mat_s *A, bool no_fold = false, int tune = -1)
{
  const int n = ilog2(A->M);
  const int nloc = n - ilog2(G.nranks);
  const std::vector<NMask> masks = normalise(A, n);  // only needed while planning: plans own device copies

  struct Cand {
    int T, B, f;
    bool for_jit;
  };
  std::vector<Cand> candidates;
  const char *env_b = getenv("DNM_TILE_RUN_BITS");
  int far_opt = A->far_bits;
  if (far_opt < 0 && getenv("DNM_FAR_BITS")) far_opt = atoi(getenv("DNM_FAR_BITS"));
  const bool jit_possible = A->jit != 0 && getenv("DNM_NO_JIT") == nullptr && (A->jit == 1 || nloc >= 22);
  // Generic kernel (profiles/r01_explore_tiles.log, r01_ring_experiment.md): up to 2^28 rows per GPU
  // 64 KB tiles (T=12, two CTAs per SM) win; beyond that 128 KB tiles (T=13, one pass fewer over HBM).
  {
    std::vector<int> Ts = A->tile_bits ? std::vector<int>{A->tile_bits}
                                       : (nloc >= 29 ? std::vector<int>{13} : std::vector<int>{12});
    std::vector<int> Bs = env_b ? std::vector<int>{atoi(env_b)} : std::vector<int>{3, 2};
    for (int T : Ts)
      for (int B : Bs) {
        T = std::min(T, nloc);
        B = std::max(0, std::min(B, T - 1));
        B = std::min(B, T - 4);  // a thread's own index bits (>= T-4 of them) must cover the run bits
        bool seen = false;
        for (const Cand &c : candidates) seen = seen || (c.T == T && c.B == B);
        if (!seen) candidates.push_back(Cand{T, B, std::max(far_opt, 0), false});
      }
  }
  // Generated kernels (profiles/r02_jit_*): 32 KB tiles (T=11, four CTAs per SM hide the L2 latency of
  // the FAR loads), an L2 window of up to 10 positions (two passes instead of three or four at L=30) and
  // 256-byte runs (128-byte runs reach 4.8 TB/s in an accumulating pass, 256-byte runs 6.2 TB/s:
  // scripts/micro/tma_stream.cu).
  if (jit_possible && !A->tile_bits && far_opt < 0 && nloc >= 15) {
    const int B = env_b ? std::max(0, std::min(atoi(env_b), 7)) : 4;
    candidates.push_back(Cand{11, B, std::min(10, nloc - 11), true});
  }
  if (tune >= 0 && TUNE_SHAPES[tune].T > 0) {
    // autotuner: exactly this shape, generated kernels wherever its passes are lean
    const TuneShape &ts = TUNE_SHAPES[tune];
    candidates.clear();
    const int T = std::min(ts.T, nloc);
    candidates.push_back(Cand{T, std::min(ts.B, T - 4), std::min(ts.f, nloc - T), false});
  } else if (tune >= 0) {
    // the default plan of the generic kernel (no generated code): drop the jit-oriented candidate
    if (candidates.size() > 1 && candidates.back().for_jit) candidates.pop_back();
  }
  std::unique_ptr<TiledPlan> best;
  Cand best_c{0, 0, 0, false};
  for (const Cand &c : candidates) {
    const bool fold_ok = jit_possible && G.nranks > 1 && !no_fold && !(tune >= 0 && TUNE_SHAPES[tune].T < 0);
    std::unique_ptr<TiledPlan> cand = plan_with(A, masks, c.T, rows_for(c.T, A->tile_rows), c.B, c.f, A->verbose, fold_ok);
    if (cand->fold_failed) cand = plan_with(A, masks, c.T, rows_for(c.T, A->tile_rows), c.B, c.f, A->verbose, false);
    if (A->verbose)
      fprintf(stderr, "[dnm] plan T=%d B=%d far<=%d: %zu passes + %zu direct, cost %.2f sweeps\n", c.T, c.B, c.f,
              cand->passes.size(), cand->directs.size(), cand->cost);
    if (c.for_jit) {
      // only worth it when every pass can run a generated kernel and no pass is added
      bool lean = cand->directs.empty();
      for (const Pass &ps : cand->passes) lean = lean && ps.small && ps.p.lean && ps.R == 8;
      if (!lean || (best && cand->passes.size() > best->passes.size())) continue;
      best = std::move(cand);
      best_c = c;
      continue;
    }
    if (!best || cand->cost < best->cost - 1e-9) {
      best = std::move(cand);
      best_c = c;
    }
  }
  // the batched launch below interprets its passes with the table-driven group loop, which has no FAR
  // path: small problems are planned without an L2 window
  auto batchable = [&](const TiledPlan &pl) {
    if (!(G.nranks == 1 && pl.directs.empty() && pl.passes.size() >= 3 && getenv("DNM_NO_BATCH") == nullptr)) return false;
    const Pass &p0 = pl.passes[0];
    bool same = true;
    for (const Pass &ps : pl.passes) same = same && ps.T == p0.T && ps.R == p0.R && ps.peer_xor == 0;
    return same && ((long long)1 << (pl.nloc - p0.T)) <= 4LL * G.sm_count && pl.passes.size() < 65536;
  };
  if (best && best_c.f > 0) {
    bool any_far = false;
    for (const Pass &ps : best->passes) any_far = any_far || ps.nfar > 0;
    if (any_far) {
      std::unique_ptr<TiledPlan> plain = plan_with(A, masks, best_c.T, rows_for(best_c.T, A->tile_rows), best_c.B, 0, 0);
      if (batchable(*plain) && A->jit != 1) best = std::move(plain);
    }
  }
  // Small problems (a pass has fewer tiles than the GPU has CTA slots) with several passes: run all
  // passes concurrently as one grid and combine in y with FP64 atomics.
  bool has_far = false;
  if (best)
    for (const Pass &ps : best->passes) has_far = has_far || ps.nfar > 0;
  if (best && !has_far && A->jit != 1 && !g_plan_host_only && batchable(*best)) {
    const Pass &p0 = best->passes[0];
    {
      std::vector<PassParams> all;
      for (const Pass &ps : best->passes) {
        PassParams q = ps.p;
        q.accumulate = 2;
        if (getenv("DNM_NO_BATCH_STAGE")) q.staged = 0;
        if (q.staged) best->batch_stage_bytes = std::max(best->batch_stage_bytes, (size_t)ps.nterms * 16);
        all.push_back(q);
      }
      DNM_CHECK_CUDA(cudaMalloc(&best->d_batch, sizeof(PassParams) * all.size()));
      DNM_CHECK_CUDA(cudaMemcpyAsync(best->d_batch, all.data(), sizeof(PassParams) * all.size(), cudaMemcpyHostToDevice,
                                     G.stream));
      best->batch_T = p0.T;
      best->batch_R = p0.R;
    }
  }
  // Operator-specialised kernels for the lean passes (jit.h): on by default for vectors that leave
  // the L2 (the generic kernel is kept for small problems, where compiling would dominate).
  const bool want_jit = !(tune >= 0 && TUNE_SHAPES[tune].T < 0) &&
                        (A->jit == 1 || (A->jit < 0 && getenv("DNM_NO_JIT") == nullptr && best && best->nloc >= 22));
  if (best && want_jit && !best->d_batch) {
    std::vector<jit::PassDesc> descs;
    std::vector<size_t> which;
    for (size_t k = 0; k < best->passes.size(); ++k) {
      const Pass &ps = best->passes[k];
      if (!(ps.small && ps.p.lean && ps.R == 8 && ps.T >= 9 && ps.p.accumulate != 2)) continue;
      jit::PassDesc d;
      d.p = &ps.p;
      d.st = &ps.st;
      d.T = ps.T;
      d.nloc = best->nloc;
      d.W = ps.W;
      d.stage_remote = stage_remote_operands();
      d.tma_stage = jit::tma_eligible(d) && (ps.peer_xor == 0 || best->dma) &&
                    !(getenv("DNM_JIT_TMA") && atoi(getenv("DNM_JIT_TMA")) == 0);
      // measured at L=30 MBL (same call): cp.async staging 18.4 ms, TMA staging 18.6 ms, TMA staging + reduce-add
      // epilogue 17.8 ms -> both on by default (DNM_JIT_TMA=0 / DNM_JIT_TMA_REDUCE=0 turn them off)
      d.tma_reduce = !(getenv("DNM_JIT_TMA_REDUCE") && atoi(getenv("DNM_JIT_TMA_REDUCE")) == 0);
      d.filter_bit = ps.filter_bit;
      d.filter_val = ps.filter_val;
      d.rows = 8;
      if (const char *e = getenv("DNM_JIT_ROWS")) d.rows = atoi(e) == 4 ? 4 : 8;
      d.nbuf = ps.p.accumulate == 1 ? 2 : 3;
      if (const char *e = getenv("DNM_JIT_NBUF")) d.nbuf = std::max(2, std::min(6, atoi(e)));
      // pipelined (persistent, TMA ring, reduce-add epilogue) wherever the tile is a TMA box and the ring fits;
      // remote operands that are read through peer mappings keep the classic kernel
      const size_t ring = (size_t)(d.nbuf + (ps.p.accumulate == 1 ? 1 : 0)) * ((size_t)16 << ps.T) + 2048;
      // Measured on B200 (profiles/r02_experiments.md §5): with 8-16 resident warps per SM the
      // persistent kernel cannot hide the L2 latency of the FAR loads and loses to one tile per CTA
      // (4 CTAs per SM at T=11); it stays opt-in (option pipeline = 1).
      d.pipelined = A->pipeline == 1 && jit::tma_eligible(d) && (ps.p.accumulate != 1 || jit::tma_reducible(d)) &&
                    ring <= 232448 && (ps.peer_xor == 0 || best->dma);
      descs.push_back(d);
      which.push_back(k);
    }
    if (!descs.empty()) {
      std::string log;
      const std::string src = jit::generate(descs);
      if (const char *dump = getenv("DNM_JIT_DUMP")) {
        // (autotuner trials go to <file>.shape<k>; dnm_mat_get_info "tuned_shape" names the one that was kept)
        const std::string path = tune >= 0 ? std::string(dump) + ".shape" + std::to_string(tune) : std::string(dump);
        if (FILE *f = fopen(path.c_str(), "w")) {
          fwrite(src.data(), 1, src.size(), f);
          fclose(f);
        }
      }
      if (g_plan_host_only) {
        g_dry.src = src;
        g_dry.host = jit::host_emulation();
        if (!g_dry.host) g_dry.cubin_bytes = jit::compile_cubin(src, log).size();  // (host source is for a C++ compiler)
        g_dry.log = log;
        g_dry.kernels = (int)descs.size();
        g_dry.passes = (int)best->passes.size();
        g_dry.remote_groups = g_dry.pipelined = 0;
        for (const Pass &ps : best->passes) g_dry.remote_groups += ps.nremote;
        for (const jit::PassDesc &d : descs) g_dry.pipelined += d.pipelined ? 1 : 0;
        return best.release();
      }
      best->jit = jit::compile(src, descs, log);
      if (!best->jit && best->fold && !no_fold) {
        // folded remote masks only exist in generated code: plan again the classic way
        if (A->verbose) fprintf(stderr, "[dnm] generated kernels unavailable (%s): planning without folded remote masks\n", log.c_str());
        best.reset();
        return build_plan(A, true, tune);
      }
      if (best->jit) {
        for (size_t k = 0; k < which.size(); ++k) best->passes[which[k]].jk = &best->jit->kernels[k];
        bool stranded = false;  // a pass with folded remote masks that did not get a generated kernel
        for (const Pass &ps : best->passes) stranded = stranded || (ps.nremote > 0 && !ps.jk);
        if (stranded && !no_fold) {
          best.reset();
          return build_plan(A, true, tune);
        }
        if (A->verbose) fprintf(stderr, "[dnm] %zu generated kernels (%zu bytes of source)\n", which.size(), src.size());
      } else {
        // NVRTC or the driver entry points missing is an environment matter (quiet unless asked); a source that
        // NVRTC REJECTS is a generator defect that costs the plan its fast kernels: never silent
        const bool env = log.find("could not be loaded") != std::string::npos || log.find("unavailable") != std::string::npos;
        if (A->verbose || A->jit == 1 || (!env && G.rank == 0 && getenv("DNM_QUIET") == nullptr))
          fprintf(stderr, "[dynamite_b200] generated kernels %s, using the generic tiled kernel: %.600s\n",
                  env ? "unavailable" : "REJECTED by the compiler (generator defect)", log.c_str());
      }
    }
  }
  if (!g_plan_host_only) DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return best.release();
}

void stream_barrier()
{
  if (G.nranks > 1) allreduce_sum_dev(G.d_scratch + SCRATCH_DOUBLES - 8, 1);
}

}  // namespace

bool tiled_supported(const dnm_mat_s *A)
{
  const int lt = A->left.desc.type, rt = A->right.desc.type;
  if (lt != rt) return false;
  if (lt == DNM_FULL) return true;
  if (lt == DNM_PARITY) return A->left.desc.space == A->right.desc.space;
  return false;
}

void tiled_free(dnm_mat_s *A)
{
  delete A->tiled;
  A->tiled = nullptr;
}

int tiled_jit_passes(dnm_mat_s *A)
{
  int n = 0;
  if (A->tiled)
    for (const Pass &ps : A->tiled->passes) n += ps.jk ? 1 : 0;
  return n;
}

int tiled_passes(dnm_mat_s *A) { return A->tiled ? (int)(A->tiled->passes.size() + A->tiled->directs.size()) : 0; }

namespace {
void run_plan(dnm_mat_s *A, TiledPlan &plan, dnm_vec_t xv, dnm_vec_t yv);

// First use of a big matrix: time a few plan shapes on the caller's own vectors and keep the
// fastest (which shape wins depends on the operator: XX+YY models skip half of the rows of every
// mask and like a wide L2 window, field-heavy models are bound by the operand fetches instead).
// Sharded: every rank runs the same trials and the slowest rank's time decides, so all ranks agree.
TiledPlan *autotune_plan(dnm_mat_s *A, dnm_vec_t xv, dnm_vec_t yv)
{
  const int nloc = ilog2(A->M) - ilog2(G.nranks);
  const bool jit_possible = A->jit != 0 && getenv("DNM_NO_JIT") == nullptr && (A->jit == 1 || nloc >= 22);
  const bool wanted = A->autotune == 1 || (A->autotune < 0 && getenv("DNM_NO_AUTOTUNE") == nullptr && nloc >= 24);
  if (!wanted || !jit_possible || A->tile_bits || A->far_bits >= 0 || getenv("DNM_FAR_BITS") || getenv("DNM_REMOTE"))
    return build_plan(A);
  std::unique_ptr<TiledPlan> best;
  float best_ms = 0.f;
  int best_shape = -1;
  cudaEvent_t e0, e1;
  DNM_CHECK_CUDA(cudaEventCreate(&e0));
  DNM_CHECK_CUDA(cudaEventCreate(&e1));
  for (int k = 0; k < N_TUNE_SHAPES; ++k) {
    if (TUNE_SHAPES[k].T < 0 && G.nranks > 1) continue;  // (its remote groups would allocate whole-shard staging buffers)
    if (TUNE_SHAPES[k].T > nloc - 1) continue;
    std::unique_ptr<TiledPlan> cand;
    try {
      cand.reset(build_plan(A, false, k));
    } catch (const Fail &) {
      continue;
    }
    if (!cand) continue;
    bool dup = false;  // identical to the best so far in every pass window: nothing new to learn
    if (best && best->passes.size() == cand->passes.size()) {
      dup = true;
      for (size_t i = 0; i < cand->passes.size(); ++i)
        dup = dup && cand->passes[i].wbits == best->passes[i].wbits && cand->passes[i].nfar == best->passes[i].nfar &&
              (cand->passes[i].jk != nullptr) == (best->passes[i].jk != nullptr);
    }
    if (dup) continue;
    float ms = 0.f;
    try {
      run_plan(A, *cand, xv, yv);  // warm-up (module load, first touch)
      DNM_CHECK_CUDA(cudaEventRecord(e0, G.stream));
      run_plan(A, *cand, xv, yv);
      run_plan(A, *cand, xv, yv);
      DNM_CHECK_CUDA(cudaEventRecord(e1, G.stream));
      DNM_CHECK_CUDA(cudaEventSynchronize(e1));
      DNM_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    } catch (const Fail &) {
      cudaGetLastError();
      continue;
    }
    if (G.nranks > 1) {
      double v = ms;
      DNM_CHECK_CUDA(cudaMemcpyAsync(G.d_scratch, &v, sizeof(double), cudaMemcpyHostToDevice, G.stream));
      allreduce_max_dev(G.d_scratch, 1);
      fetch_doubles(G.d_scratch, &v, 1);
      ms = (float)v;
    }
    if (A->verbose)
      fprintf(stderr, "[dnm] autotune shape T=%d B=%d far<=%d: %zu passes, %.3f ms per MatMult\n", TUNE_SHAPES[k].T,
              TUNE_SHAPES[k].B, TUNE_SHAPES[k].f, cand->passes.size(), ms / 2);
    if (!best || ms < best_ms) {
      best = std::move(cand);
      best_ms = ms;
      best_shape = k;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (!best) return build_plan(A);
  A->tuned_shape = best_shape;
  return best.release();
}
}  // namespace

void tiled_mult(dnm_mat_s *A, dnm_vec_t xv, dnm_vec_t yv)
{
  if (!A->tiled) A->tiled = autotune_plan(A, xv, yv);
  run_plan(A, *A->tiled, xv, yv);
}

namespace {
void run_plan(dnm_mat_s *A, TiledPlan &plan, dnm_vec_t xv, dnm_vec_t yv)
{
  const i64 nloc_rows = (i64)1 << plan.nloc;
  cplx *y = yv->d;
  int launches = 0;

  auto source = [&](int peer_xor) -> const cplx * {
    if (peer_xor == 0) return xv->d;
    const cplx *ptr = xv->peer[G.rank ^ peer_xor];
    DNM_REQUIRE(ptr != nullptr, DNM_ERR_COMM, "input vector is not mapped on peer rank %d", G.rank ^ peer_xor);
    return ptr;
  };

  if (plan.d_batch) {
    DNM_CHECK_CUDA(cudaMemsetAsync(y, 0, sizeof(cplx) * (size_t)nloc_rows, G.stream));
    launch_batch(plan, xv->d, y, plan.use_diag ? A->d_diag : nullptr);
    A->launches_per_mult = 1;
    return;
  }

  if (plan.any_remote) stream_barrier();  // every rank's x is complete before anyone pulls from it
  for (int h = 0; h < MAX_RANKS; ++h) g_peers[h] = nullptr;
  if (plan.fold)
    for (int h = 1; h < G.nranks; ++h) {
      g_peers[h] = xv->peer[G.rank ^ h];
      DNM_REQUIRE(g_peers[h] != nullptr, DNM_ERR_COMM, "input vector is not mapped on peer rank %d", G.rank ^ h);
    }

  if (plan.dma && plan.any_remote) {
    const size_t bytes = sizeof(cplx) * (size_t)nloc_rows;
    // distinct partners in pass order
    std::vector<int> partners;
    for (const Pass &ps : plan.passes)
      if (ps.peer_xor && std::find(partners.begin(), partners.end(), ps.peer_xor) == partners.end())
        partners.push_back(ps.peer_xor);
    for (const Direct &d : plan.directs)
      if (d.peer_xor && std::find(partners.begin(), partners.end(), d.peer_xor) == partners.end())
        partners.push_back(d.peer_xor);
    const int nbuf = partners.size() > 1 ? 2 : 1;
    for (int b = 0; b < nbuf; ++b) {
      if (!plan.stage[b]) DNM_CHECK_CUDA(cudaMalloc(&plan.stage[b], bytes));
      if (!plan.ev_staged[b]) DNM_CHECK_CUDA(cudaEventCreateWithFlags(&plan.ev_staged[b], cudaEventDisableTiming));
      if (!plan.ev_free[b]) DNM_CHECK_CUDA(cudaEventCreateWithFlags(&plan.ev_free[b], cudaEventDisableTiming));
    }
    DNM_CHECK_CUDA(cudaEventRecord(G.ev_fork, G.stream));
    DNM_CHECK_CUDA(cudaStreamWaitEvent(G.stream2, G.ev_fork, 0));
    // local passes on the SMs ...
    bool first = true;
    for (const Unit &u : plan.units) {
      const Pass &ps = plan.passes[u.passes.front()];
      if (ps.peer_xor != 0) continue;
      const double *diag = (first && plan.use_diag) ? A->d_diag : nullptr;
      launch_pass(ps, xv->d, y, diag, nloc_rows >> ps.T);
      first = false;
      ++launches;
    }
    for (const Direct &d : plan.directs) {
      if (d.peer_xor != 0) continue;
      const double *diag = (first && plan.use_diag) ? A->d_diag : nullptr;
      k_xor_direct<<<direct_grid(nloc_rows), 256, 0, G.stream>>>(d.p, xv->d, y, diag, nloc_rows);
      count_launch();
      DNM_CHECK_CUDA(cudaGetLastError());
      first = false;
      ++launches;
    }
    // ... while the copy engines stage the partner shards, two buffers deep.  Only the part of a
    // shard this rank can touch travels (plan.need): all of it, a contiguous half, or nothing.
    size_t used = 0;
    for (size_t gi = 0; gi < partners.size(); ++gi) {
      TiledPlan::Need nd;
      auto it = plan.need.find(partners[gi]);
      if (it != plan.need.end()) nd = it->second;
      if (nd.kind == 1) continue;  // every coefficient of this group vanishes on this rank
      const int b = (int)(used % nbuf);
      if ((int)used >= nbuf) DNM_CHECK_CUDA(cudaStreamWaitEvent(G.stream2, plan.ev_free[b], 0));
      ++used;
      const cplx *src = source(partners[gi]);
      if (nd.kind == 2 && getenv("DNM_POISON_STAGE"))  // tests: the part that does not travel must never be used
        DNM_CHECK_CUDA(cudaMemsetAsync(plan.stage[b], 0xff, bytes, G.stream2));
      if (nd.kind == 2)
        DNM_CHECK_CUDA(cudaMemcpyAsync(plan.stage[b] + nd.first, src + nd.first, sizeof(cplx) * (size_t)nd.count,
                                       cudaMemcpyDeviceToDevice, G.stream2));
      else
        DNM_CHECK_CUDA(cudaMemcpyAsync(plan.stage[b], src, bytes, cudaMemcpyDeviceToDevice, G.stream2));
      DNM_CHECK_CUDA(cudaEventRecord(plan.ev_staged[b], G.stream2));
      // the group's passes read the staged copy as ordinary local memory
      DNM_CHECK_CUDA(cudaStreamWaitEvent(G.stream, plan.ev_staged[b], 0));
      for (const Unit &u : plan.units) {
        const Pass &ps = plan.passes[u.passes.front()];
        if (ps.peer_xor != partners[gi]) continue;
        launch_pass(ps, plan.stage[b], y, nullptr, nloc_rows >> ps.T);
        ++launches;
      }
      for (const Direct &d : plan.directs) {
        if (d.peer_xor != partners[gi]) continue;
        k_xor_direct<<<direct_grid(nloc_rows), 256, 0, G.stream>>>(d.p, plan.stage[b], y, nullptr, nloc_rows);
        count_launch();
        DNM_CHECK_CUDA(cudaGetLastError());
        ++launches;
      }
      DNM_CHECK_CUDA(cudaEventRecord(plan.ev_free[b], G.stream));
    }
    stream_barrier();  // every peer has finished copying out of x
    A->launches_per_mult = launches;
    return;
  }

  const bool overlap = plan.overlap && plan.any_remote;
  cplx *yr = nullptr;
  if (overlap) {
    if (!plan.y_remote) DNM_CHECK_CUDA(cudaMalloc(&plan.y_remote, sizeof(cplx) * (size_t)nloc_rows));
    yr = plan.y_remote;
    // fork: the side stream starts once x is globally ready
    DNM_CHECK_CUDA(cudaEventRecord(G.ev_fork, G.stream));
    DNM_CHECK_CUDA(cudaStreamWaitEvent(G.stream2, G.ev_fork, 0));
    g_launch_stream = G.stream2;
    try {
      for (const Unit &u : plan.units) {
        const Pass &ps = plan.passes[u.passes.front()];
        if (ps.peer_xor == 0) continue;
        // a few resident CTAs per SM keep the NVLink busy and leave the rest of the SM to the local passes
        launch_pass(ps, source(ps.peer_xor), yr, nullptr, nloc_rows >> ps.T);
        ++launches;
      }
      for (const Direct &d : plan.directs) {
        if (d.peer_xor == 0) continue;
        k_xor_direct<<<direct_grid(nloc_rows), 256, 0, G.stream2>>>(d.p, source(d.peer_xor), yr, nullptr, nloc_rows);
        count_launch();
        DNM_CHECK_CUDA(cudaGetLastError());
        ++launches;
      }
    } catch (...) {
      g_launch_stream = nullptr;
      throw;
    }
    g_launch_stream = nullptr;
    DNM_CHECK_CUDA(cudaEventRecord(G.ev_join, G.stream2));
  }

  // local work (and, without overlap, the remote passes too) on the main stream
  bool first = true, joined = !overlap;
  for (size_t k = 0; k < plan.units.size(); ++k) {
    const Unit &u = plan.units[k];
    const Pass &ps = plan.passes[u.passes.front()];
    if (overlap && ps.peer_xor != 0) continue;
    const cplx *x = source(ps.peer_xor);
    const double *diag = (first && plan.use_diag) ? A->d_diag : nullptr;
    launch_pass(ps, x, y, diag, nloc_rows >> ps.T);
    first = false;
    ++launches;
  }
  for (const Direct &d : plan.directs) {
    if (overlap && d.peer_xor != 0) continue;
    const cplx *x = source(d.peer_xor);
    const double *diag = (first && plan.use_diag) ? A->d_diag : nullptr;
    k_xor_direct<<<direct_grid(nloc_rows), 256, 0, G.stream>>>(d.p, x, y, diag, nloc_rows);
    count_launch();
    DNM_CHECK_CUDA(cudaGetLastError());
    first = false;
    ++launches;
  }
  if (!joined) {
    DNM_CHECK_CUDA(cudaStreamWaitEvent(G.stream, G.ev_join, 0));
    vec_axpby(y, yr, nloc_rows, make_double2(1.0, 0.0), make_double2(1.0, 0.0));
    ++launches;
  }

  if (plan.any_remote) stream_barrier();  // peers are done reading x before it can change
  A->launches_per_mult = launches;
}
}  // namespace

void tiled_diag(dnm_mat_s *A, double *d_diag)
{
  if (!A->tiled) A->tiled = build_plan(A);
  const i64 rows = (i64)1 << A->tiled->nloc;
  k_xor_diag<<<direct_grid(rows), 256, 0, G.stream>>>(A->tiled->all.p, d_diag, rows);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
}

void tiled_norm(dnm_mat_s *A, double *d_out)
{
  if (!A->tiled) A->tiled = build_plan(A);
  const i64 rows = (i64)1 << A->tiled->nloc;
  const int grid = direct_grid(rows);
  double *d_part = nullptr;
  DNM_CHECK_CUDA(cudaMalloc(&d_part, sizeof(double) * grid));
  k_xor_norm<<<grid, 256, 0, G.stream>>>(A->tiled->all.p, d_part, rows);
  count_launch();
  k_max_final<<<1, 256, 0, G.stream>>>(d_part, grid, d_out);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
  allreduce_max_dev(d_out, 1);
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  cudaFree(d_part);
}

}  // namespace dnm

// Host-only dry run of the planner + generator + NVRTC (no GPU needed): plans the MatMult of a
// Full/Parity operator as rank `rank` of `nranks` would, generates the pass kernels and compiles them
// to a cubin for sm_100a.  tune_shape: -1 = the default heuristics, >= 0 = one autotuner shape.
extern "C" int dnm_jit_set_host_emulation(int on)
{
  DNM_API_BEGIN
  dnm::jit::set_host_emulation(on != 0);
  DNM_API_END
}

extern "C" int dnm_jit_dryrun(int64_t nmasks, const int64_t *masks, const int64_t *mask_offsets, const int64_t *signs,
                              const double *coeffs, const dnm_subspace_t *sub, int nranks, int rank, int tile_bits,
                              int far_bits, int pipeline, int tune_shape, char *src_out, int64_t src_cap, int64_t *src_len,
                              int64_t *cubin_bytes, int *n_kernels, int *n_passes, int *n_remote_groups, int *n_pipelined)
{
  DNM_API_BEGIN
  using namespace dnm;
  DNM_REQUIRE(nmasks >= 1 && masks && mask_offsets && signs && coeffs && sub, DNM_ERR_ARG, "null or empty arguments");
  DNM_REQUIRE(nranks >= 1 && nranks <= MAX_RANKS && (nranks & (nranks - 1)) == 0 && rank >= 0 && rank < nranks, DNM_ERR_ARG,
              "bad rank layout");
  DNM_REQUIRE(tune_shape < N_TUNE_SHAPES, DNM_ERR_ARG, "tune_shape out of range");
  const int64_t nterms = mask_offsets[nmasks];
  dnm_mat_s A;
  A.masks.assign(masks, masks + nmasks);
  A.mask_offsets.assign(mask_offsets, mask_offsets + nmasks + 1);
  A.signs.assign(signs, signs + nterms);
  A.coeffs.assign(coeffs, coeffs + 2 * nterms);
  A.left.copy_from(sub);
  A.right.copy_from(sub);
  A.M = A.N = A.left.dim;
  DNM_REQUIRE(tiled_supported(&A), DNM_ERR_UNSUPPORTED, "the tiled MatMult needs a Full or Parity subspace");
  A.local_M = A.local_N = A.M / nranks;
  // a cached diagonal is assumed (as benchmark.py builds its matrices): only its presence matters to the planner
  A.d_diag = A.masks[0] == 0 ? reinterpret_cast<double *>(0x10) : nullptr;
  A.jit = 1;
  A.tile_bits = tile_bits;
  A.far_bits = far_bits;
  A.pipeline = pipeline;
  const int save_n = G.nranks, save_r = G.rank;
  G.nranks = nranks;
  G.rank = rank;
  g_plan_host_only = true;
  g_dry = DryRun();
  TiledPlan *plan = nullptr;
  try {
    plan = build_plan(&A, false, tune_shape);
  } catch (...) {
    g_plan_host_only = false;
    G.nranks = save_n;
    G.rank = save_r;
    A.d_diag = nullptr;
    throw;
  }
  g_plan_host_only = false;
  G.nranks = save_n;
  G.rank = save_r;
  A.d_diag = nullptr;
  delete plan;
  if (src_len) *src_len = (int64_t)g_dry.src.size();
  if (src_out && src_cap > 0) {
    const size_t n = std::min<size_t>((size_t)src_cap - 1, g_dry.src.size());
    memcpy(src_out, g_dry.src.data(), n);
    src_out[n] = 0;
  }
  if (cubin_bytes) *cubin_bytes = (int64_t)g_dry.cubin_bytes;
  if (n_kernels) *n_kernels = g_dry.kernels;
  if (n_passes) *n_passes = g_dry.passes;
  if (n_remote_groups) *n_remote_groups = g_dry.remote_groups;
  if (n_pipelined) *n_pipelined = g_dry.pipelined;
  DNM_REQUIRE(g_dry.kernels == 0 || g_dry.host || g_dry.cubin_bytes > 0, DNM_ERR_INTERNAL, "NVRTC rejected the generated source: %s",
              g_dry.log.c_str());
  DNM_API_END
}

