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
#include <map>
#include <memory>

#include "vecops.cuh"

namespace dnm {

namespace {

typedef unsigned int u32;

constexpr int SMALL_MASKS = 32;   // passes up to this size keep their tables in kernel-parameter
constexpr int SMALL_TERMS = 96;   // (constant) memory; bigger ones read them from global memory

struct PassParams {
  int nmasks;
  int B;           // log2 of the contiguous run length
  int n_outer;     // number of index bits outside the window
  int accumulate;  // 0: y = ..., 1: y += ...
  const u32 *lam;  // [nmasks] mask in window coordinates
  const int *t_re; // [nmasks] first real term
  const int *t_im; // [nmasks] first imaginary term
  const int *t_end;
  const u32 *sw;   // [nterms] sign bits inside the window (window coordinates)
  const u32 *rb;   // [nterms] bit r = parity((sw >> LOG_NT) & r)
  const i64 *so;   // [nterms] sign bits outside the window (global index coordinates)
  const double *cf;
  const i64 *rowoff;  // [2^(T-B)] offset of each contiguous run
  i64 rank_bits;      // global index bits contributed by the rank
  i64 roff[16];       // offset contributed by the r-th row group of a thread
  unsigned char outer_pos[48];
};

// the same tables by value, for small passes
struct SmallTables {
  u32 lam[SMALL_MASKS];
  unsigned short t_re[SMALL_MASKS], t_im[SMALL_MASKS], t_end[SMALL_MASKS];
  u32 sw[SMALL_TERMS];
  u32 rb[SMALL_TERMS];
  i64 so[SMALL_TERMS];
  double cf[SMALL_TERMS];
};

template <int T>
struct TileCfg {
  static constexpr int R = (T >= 10) ? 16 : (T == 9 ? 8 : 4);
  static constexpr int NT = (1 << T) / R;
  static constexpr int LOG_NT = (T >= 10) ? T - 4 : (T == 9 ? 6 : 6);
  static constexpr int MINB = (T >= 13) ? 1 : 2;
};

__device__ __forceinline__ double flip_if(int hi, int lo, u32 bits, int r)
{
  // negate when bit r of `bits` is set: xor into the IEEE sign bit
  return __hiloint2double(hi ^ (int)((bits << (31 - r)) & 0x80000000u), lo);
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ void cp_async_wait_all()
{
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// One CTA = one tile of 2^T amplitudes.  Thread `tid` owns the R rows
// l = tid + r*NT of the tile (window coordinates).
template <int T, bool SMALL>
__global__ void __launch_bounds__(TileCfg<T>::NT, TileCfg<T>::MINB)
    k_tiled(const __grid_constant__ PassParams P, const __grid_constant__ SmallTables S,
            const cplx *__restrict__ x, cplx *__restrict__ y, const double *__restrict__ diag)
{
  constexpr int R = TileCfg<T>::R;
  constexpr int NT = TileCfg<T>::NT;
  constexpr int LOG_NT = TileCfg<T>::LOG_NT;
  extern __shared__ double2 tile[];
  const int tid = threadIdx.x;

  // scatter the tile number into the bit positions outside the window
  i64 base_g = 0;
  {
    const unsigned long long b = blockIdx.x;
    for (int k = 0; k < P.n_outer; ++k) base_g |= (i64)((b >> k) & 1ull) << P.outer_pos[k];
  }
  const i64 outer_g = base_g | P.rank_bits;  // sign-relevant bits shared by the whole tile
  // this thread's part of the address: its rows differ only by the uniform P.roff[r]
  base_g |= __ldg(&P.rowoff[tid >> P.B]) | (i64)(tid & ((1 << P.B) - 1));

  // stage the tile: runs of 2^B contiguous amplitudes, asynchronous 16-byte copies
#pragma unroll
  for (int r = 0; r < R; ++r) cp_async16(&tile[tid + r * NT], &x[base_g | P.roff[r]]);
  cp_async_wait_all();
  __syncthreads();

  double ar[R], ai[R];
  if (diag != nullptr) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double d = __ldg(&diag[base_g | P.roff[r]]);
      const double2 v = tile[tid + r * NT];
      ar[r] = d * v.x;
      ai[r] = d * v.y;
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) ar[r] = ai[r] = 0.0;
  }

  for (int mi = 0; mi < P.nmasks; ++mi) {
    const u32 lam = SMALL ? S.lam[mi] : __ldg(&P.lam[mi]);
    const int base = tid ^ (int)(lam & (NT - 1));
    const int hi_l = (int)(lam >> LOG_NT);
    const int t0 = SMALL ? (int)S.t_re[mi] : __ldg(&P.t_re[mi]);
    const int t1 = SMALL ? (int)S.t_im[mi] : __ldg(&P.t_im[mi]);
    const int t2 = SMALL ? (int)S.t_end[mi] : __ldg(&P.t_end[mi]);
#pragma unroll 1
    for (int kind = 0; kind < 2; ++kind) {
      const int ta = kind ? t1 : t0, tb = kind ? t2 : t1;
      if (ta == tb) continue;
      double d[R];
      for (int t = ta; t < tb; ++t) {
        const double c = SMALL ? S.cf[t] : __ldg(&P.cf[t]);
        const i64 so = SMALL ? S.so[t] : __ldg(&P.so[t]);
        const u32 sw = SMALL ? S.sw[t] : __ldg(&P.sw[t]);
        const u32 rb = SMALL ? S.rb[t] : __ldg(&P.rb[t]);
        const int p = (__popcll((unsigned long long)(so & outer_g)) ^ __popc(sw & (u32)tid)) & 1;
        const u32 bits = rb ^ (u32)(-p);
        const int chi = __double2hiint(c), clo = __double2loint(c);
        if (t == ta) {
#pragma unroll
          for (int r = 0; r < R; ++r) d[r] = flip_if(chi, clo, bits, r);
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) d[r] += flip_if(chi, clo, bits, r);
        }
      }
      if (kind == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (d[r] != 0.0) {
            const double2 v = tile[base + ((r ^ hi_l) << LOG_NT)];
            ar[r] += d[r] * v.x;
            ai[r] += d[r] * v.y;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (d[r] != 0.0) {
            const double2 v = tile[base + ((r ^ hi_l) << LOG_NT)];
            ar[r] -= d[r] * v.y;
            ai[r] += d[r] * v.x;
          }
        }
      }
    }
  }

  if (P.accumulate) {
    // reuse the tile buffer to fetch the previous pass's y with full memory-level parallelism
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) cp_async16(&tile[tid + r * NT], &y[base_g | P.roff[r]]);
    cp_async_wait_all();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double2 old = tile[tid + r * NT];  // written by this thread's own copies
      y[base_g | P.roff[r]] = make_double2(ar[r] + old.x, ai[r] + old.y);
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) y[base_g | P.roff[r]] = make_double2(ar[r], ai[r]);
  }
}

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
  int peer_xor = 0;  // x is read from rank ^ peer_xor
  int nterms = 0;
  std::vector<void *> owned;
};

struct Direct {
  DirectParams p{};
  int peer_xor = 0;
  std::vector<void *> owned;
};

}  // namespace

struct TiledPlan {
  int n = 0;      // index bits (global)
  int nloc = 0;   // index bits on this rank
  bool use_diag = false;
  std::vector<Pass> passes;
  std::vector<Direct> directs;
  Direct all;  // every mask, for the row-local helpers (diag, norm)
  bool any_remote = false;
  ~TiledPlan()
  {
    for (auto &ps : passes)
      for (void *q : ps.owned) cudaFree(q);
    for (auto &d : directs)
      for (void *q : d.owned) cudaFree(q);
    for (void *q : all.owned) cudaFree(q);
  }
};

namespace {

int ilog2(i64 v)
{
  int n = 0;
  while (((i64)1 << n) < v) ++n;
  return n;
}

template <class Tv>
Tv *up(const std::vector<Tv> &h, std::vector<void *> &owned)
{
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

Pass make_pass(const std::vector<const NMask *> &masks, const std::vector<int> &W, int T, int B, int nloc,
               int accumulate)
{
  Pass ps;
  ps.T = T;
  const int log_nt = (T >= 10) ? T - 4 : 6;
  const int R = (1 << T) >> log_nt;
  const i64 lmask = ((i64)1 << nloc) - 1;
  i64 wbits = 0;
  for (int b : W) wbits |= (i64)1 << b;

  auto extract = [&](i64 v) {
    u32 o = 0;
    for (int b = 0; b < T; ++b)
      if ((v >> W[b]) & 1) o |= 1u << b;
    return o;
  };

  std::vector<u32> lam, sw, rb;
  std::vector<int> t_re, t_im, t_end;
  std::vector<i64> so;
  std::vector<double> cf;
  std::vector<const NTerm *> flat;
  for (const NMask *nm : masks) lam.push_back(extract(nm->mask & lmask));
  fill_term_ranges(masks, t_re, t_im, t_end, flat);
  for (const NTerm *t : flat) {
    const u32 w = extract(t->sign & lmask);
    sw.push_back(w);
    u32 bits = 0;
    for (int r = 0; r < R; ++r)
      if (__builtin_parity((w >> log_nt) & (u32)r)) bits |= 1u << r;
    rb.push_back(bits);
    so.push_back(t->sign & ~wbits);  // outside the window, rank bits included
    cf.push_back(t->coef);
  }
  std::vector<i64> rowoff((size_t)1 << (T - B));
  for (size_t h = 0; h < rowoff.size(); ++h) {
    i64 off = 0;
    for (int b = B; b < T; ++b)
      if ((h >> (b - B)) & 1) off |= (i64)1 << W[b];
    rowoff[h] = off;
  }
  ps.p.nmasks = (int)masks.size();
  ps.p.B = B;
  ps.p.accumulate = accumulate;
  ps.p.n_outer = 0;
  for (int b = 0; b < nloc; ++b)
    if (!((wbits >> b) & 1)) ps.p.outer_pos[ps.p.n_outer++] = (unsigned char)b;
  ps.p.lam = up(lam, ps.owned);
  ps.p.t_re = up(t_re, ps.owned);
  ps.p.t_im = up(t_im, ps.owned);
  ps.p.t_end = up(t_end, ps.owned);
  ps.p.sw = up(sw, ps.owned);
  ps.p.rb = up(rb, ps.owned);
  ps.p.so = up(so, ps.owned);
  ps.p.cf = up(cf, ps.owned);
  ps.p.rowoff = up(rowoff, ps.owned);
  ps.p.rank_bits = (i64)G.rank << nloc;
  for (int r = 0; r < 16; ++r) ps.p.roff[r] = 0;
  for (int r = 0; r < R; ++r) ps.p.roff[r] = rowoff[((size_t)r << log_nt) >> B];
  ps.nterms = (int)flat.size();
  ps.small = (int)masks.size() <= SMALL_MASKS && ps.nterms <= SMALL_TERMS;
  if (ps.small) {
    for (size_t k = 0; k < masks.size(); ++k) {
      ps.st.lam[k] = lam[k];
      ps.st.t_re[k] = (unsigned short)t_re[k];
      ps.st.t_im[k] = (unsigned short)t_im[k];
      ps.st.t_end[k] = (unsigned short)t_end[k];
    }
    for (int t = 0; t < ps.nterms; ++t) {
      ps.st.sw[t] = sw[t];
      ps.st.rb[t] = rb[t];
      ps.st.so[t] = so[t];
      ps.st.cf[t] = cf[t];
    }
  }
  return ps;
}

// Greedy window cover of one partner group's masks.
void plan_group(TiledPlan &plan, std::vector<const NMask *> remaining, int peer_xor, int T, int B, bool first_group,
                int verbose)
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
    Pass ps = make_pass(chosen, Wpos, T, B, nloc, wrote ? 1 : 0);
    ps.peer_xor = peer_xor;
    if (verbose)
      fprintf(stderr, "[dnm] pass %zu: peer^%d window=0x%llx masks=%d terms=%d %s\n", plan.passes.size(), peer_xor,
              (unsigned long long)W, ps.p.nmasks, ps.nterms, wrote ? "accumulate" : "write");
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

template <int T, bool SMALL>
void launch_tiled_v(const Pass &ps, const cplx *x, cplx *y, const double *diag, i64 ntiles)
{
  static bool attr_set = false;
  const size_t smem = sizeof(double2) << T;
  if (!attr_set) {
    DNM_CHECK_CUDA(cudaFuncSetAttribute(k_tiled<T, SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DNM_CHECK_CUDA(cudaFuncSetAttribute(k_tiled<T, SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set = true;
  }
  k_tiled<T, SMALL><<<(unsigned)ntiles, TileCfg<T>::NT, smem, G.stream>>>(ps.p, ps.st, x, y, diag);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
}

template <int T>
void launch_tiled(const Pass &ps, const cplx *x, cplx *y, const double *diag, i64 ntiles)
{
  if (ps.small) launch_tiled_v<T, true>(ps, x, y, diag, ntiles);
  else launch_tiled_v<T, false>(ps, x, y, diag, ntiles);
}

int direct_grid(i64 rows)
{
  const i64 want = (rows + 255) / 256;
  const i64 cap = (i64)G.sm_count * 16;
  return (int)std::max<i64>(1, std::min(want, cap));
}

TiledPlan *build_plan(dnm_mat_s *A)
{
  std::unique_ptr<TiledPlan> plan(new TiledPlan());
  const int n = ilog2(A->M);
  int p = ilog2(G.nranks);
  plan->n = n;
  plan->nloc = n - p;
  plan->use_diag = A->d_diag != nullptr;
  const int nloc = plan->nloc;

  auto masks = std::make_shared<std::vector<NMask>>(normalise(A, n));
  // NMask storage must outlive the planning only (everything is uploaded)
  std::vector<const NMask *> all;
  for (const NMask &nm : *masks) all.push_back(&nm);
  plan->all = make_direct(all, nloc, 0);

  // group by partner rank
  std::map<int, std::vector<const NMask *>> groups;
  groups[0];  // group 0 always exists: its first pass writes y
  for (const NMask &nm : *masks) {
    if (plan->use_diag && nm.mask == 0) continue;  // served from the cached diagonal
    groups[(int)(nm.mask >> nloc)].push_back(&nm);
  }

  int T = A->tile_bits ? A->tile_bits : 12;
  T = std::min(T, nloc);
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
    int B = std::min(3, T - 1);
    if (const char *e = getenv("DNM_TILE_RUN_BITS")) B = std::max(0, std::min(atoi(e), T - 1));
    B = std::min(B, (T >= 10) ? T - 4 : 6);  // a thread's own index bits must cover the run bits
    bool first = true;
    for (auto &kv : groups) {
      plan_group(*plan, kv.second, kv.first, T, B, first, A->verbose);
      first = false;
      if (kv.first != 0) plan->any_remote = true;
    }
  }
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return plan.release();
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

int tiled_passes(dnm_mat_s *A) { return A->tiled ? (int)(A->tiled->passes.size() + A->tiled->directs.size()) : 0; }

void tiled_mult(dnm_mat_s *A, dnm_vec_t xv, dnm_vec_t yv)
{
  if (!A->tiled) A->tiled = build_plan(A);
  TiledPlan &plan = *A->tiled;
  const i64 nloc_rows = (i64)1 << plan.nloc;
  cplx *y = yv->d;
  int launches = 0;

  auto source = [&](int peer_xor) -> const cplx * {
    if (peer_xor == 0) return xv->d;
    const cplx *ptr = xv->peer[G.rank ^ peer_xor];
    DNM_REQUIRE(ptr != nullptr, DNM_ERR_COMM, "input vector is not mapped on peer rank %d", G.rank ^ peer_xor);
    return ptr;
  };

  if (plan.any_remote) stream_barrier();  // every rank's x is complete before anyone pulls from it

  bool first = true;
  for (const Pass &ps : plan.passes) {
    const cplx *x = source(ps.peer_xor);
    const double *diag = (first && plan.use_diag) ? A->d_diag : nullptr;
    const i64 ntiles = nloc_rows >> ps.T;
    switch (ps.T) {
      case 8: launch_tiled<8>(ps, x, y, diag, ntiles); break;
      case 9: launch_tiled<9>(ps, x, y, diag, ntiles); break;
      case 10: launch_tiled<10>(ps, x, y, diag, ntiles); break;
      case 11: launch_tiled<11>(ps, x, y, diag, ntiles); break;
      case 12: launch_tiled<12>(ps, x, y, diag, ntiles); break;
      case 13: launch_tiled<13>(ps, x, y, diag, ntiles); break;
      default: DNM_REQUIRE(false, DNM_ERR_INTERNAL, "no tiled kernel for T=%d", ps.T);
    }
    first = false;
    ++launches;
  }
  for (const Direct &d : plan.directs) {
    const cplx *x = source(d.peer_xor);
    const double *diag = (first && plan.use_diag) ? A->d_diag : nullptr;
    k_xor_direct<<<direct_grid(nloc_rows), 256, 0, G.stream>>>(d.p, x, y, diag, nloc_rows);
    count_launch();
    DNM_CHECK_CUDA(cudaGetLastError());
    first = false;
    ++launches;
  }

  if (plan.any_remote) stream_barrier();  // peers are done reading x before it can change
  A->launches_per_mult = launches;
}

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
