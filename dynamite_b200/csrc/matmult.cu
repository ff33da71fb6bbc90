// Shell matrix: creation, the general (any subspace pair) gather MatMult,
// precomputed diagonal, infinity norm and CheckConserves.
//
// Reference semantics (paths relative to /root/reference/src/dynamite/_backend/):
//   y[row] = sum_masks ( sum_terms +-coeff ) * x[S2I_R(I2S_L(row) ^ mask)]
//   with the sign (-1)^popcount(sign & bra) evaluated on the COLUMN state
//   (msc_tools.py:63-80, bcuda_template_2.cu:229-268).
// This file is the B200 general-pair path: one thread per output row, 16-byte
// gathers of x, term tables read through warp-uniform (broadcast) loads.  The
// Full/Parity same-sector case is normally taken by the tiled kernel in
// matmult_tiled.cu instead.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <type_traits>

#include "context.h"
#include "matmult_tiled.h"
#include "vecops.cuh"

namespace dnm {

namespace {

constexpr int TPB = 256;

// +-c according to the parity of v: flip the IEEE sign bit
__device__ __forceinline__ double signed_coef(double c, i64 v)
{
  const long long bits = __double_as_longlong(c) ^ ((long long)(__popcll((unsigned long long)v) & 1) << 63);
  return __longlong_as_double(bits);
}

// matrix element of mask `mi` on column state `bra` -> (re, im)
__device__ __forceinline__ void mask_element(const MscDev &msc, int mi, i64 bra, double &cr, double &ci)
{
  cr = 0.0;
  ci = 0.0;
  int t = __ldg(&msc.off_re[mi]);
  const int t_im = __ldg(&msc.off_im[mi]);
  const int t_end = __ldg(&msc.off_end[mi]);
  for (; t < t_im; ++t) cr += signed_coef(__ldg(&msc.coef[t]), bra & __ldg(&msc.signs[t]));
  for (; t < t_end; ++t) ci += signed_coef(__ldg(&msc.coef[t]), bra & __ldg(&msc.signs[t]));
}

template <class LS, class RS>
__global__ void __launch_bounds__(TPB)
    k_mult_general(LS ls, RS rs, MscDev msc, const double *__restrict__ diag, const cplx *__restrict__ x,
                   cplx *__restrict__ y, i64 M)
{
  for (i64 row = blockIdx.x * (i64)blockDim.x + threadIdx.x; row < M; row += (i64)gridDim.x * blockDim.x) {
    const i64 ket = ls.i2s(row);
    double ar = 0.0, ai = 0.0;
    int mi = 0;
    if (diag != nullptr) {  // bcuda_template_2.cu:232-238
      const cplx xv = x[row];
      const double d = diag[row];
      ar = d * xv.x;
      ai = d * xv.y;
      mi = 1;
    }
    for (; mi < msc.nmasks; ++mi) {
      const i64 bra = ket ^ __ldg(&msc.masks[mi]);
      const i64 col = rs.s2i(bra);
      if (col < 0) continue;  // outside the right subspace
      double cr, ci;
      mask_element(msc, mi, bra, cr, ci);
      const cplx xv = x[col];
      ar += cr * xv.x - ci * xv.y;
      ai += cr * xv.y + ci * xv.x;
    }
    y[row] = make_double2(ar, ai);
  }
}

// Explicit / Auto -> the same sorted Explicit space.  The reference looks every column up with a
// binary search over the whole sorted state list (bsubspace_impl.h:306-331; its own TODO at
// bcuda_impl.cu:154-179 asks for shared memory): log2(dim) dependent global loads per (row, mask).
// Here a CTA owns a block of EX_ROWS consecutive rows, whose kets (ascending) share every bit above
// the highest bit in which the first and the last differ.  For a mask m all their images ket ^ m
// therefore lie in ONE interval [P, P + 2^h) of state values: two threads bracket that interval in
// the sorted list once per (block, mask) -- all masks in parallel -- the CTA stages that slice of the
// list in shared memory with coalesced loads, and every row finishes its search there (typically
// 8-11 shared-memory steps instead of 24 global ones at dim 10^7).  Bit-exact with S2I.
constexpr int EX_ROWS_DEFAULT = 256;
constexpr int EX_CAP = 4096;       // staged slice, states
constexpr int EX_MAX_MASKS = 256;  // bracket table, masks

__device__ __forceinline__ i64 lower_bound_dev(const i64 *__restrict__ a, i64 lo, i64 len, i64 v)
{
  while (len > 0) {
    const i64 half = len >> 1;
    if (__ldg(&a[lo + half]) < v) {
      lo += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  return lo;
}

template <int EX_ROWS>
__global__ void __launch_bounds__(EX_ROWS)
    k_mult_explicit(SubExplicit sub, MscDev msc, const double *__restrict__ diag, const cplx *__restrict__ x,
                    cplx *__restrict__ y, i64 M)
{
  // the staged slice: 64-bit states, or (whenever the varying bits fit) their low 32 bits -- half the
  // shared-memory traffic and bank conflicts of the search, twice the capacity
  __shared__ i64 s_slice[EX_CAP];
  unsigned int *s_slice32 = reinterpret_cast<unsigned int *>(s_slice);
  __shared__ i64 s_lo[EX_MAX_MASKS], s_hi[EX_MAX_MASKS];
  __shared__ i64 s_self;  // first row whose ket has the block's prefix
  const int tid = threadIdx.x;
  for (i64 row0 = (i64)blockIdx.x * EX_ROWS; row0 < M; row0 += (i64)gridDim.x * EX_ROWS) {
    const i64 row = row0 + tid;
    const bool ok = row < M;
    const i64 last = min(row0 + (i64)EX_ROWS, M) - 1;
    const i64 ket = ok ? __ldg(&sub.state_map[row]) : 0;
    const i64 s_min = __ldg(&sub.state_map[row0]), s_max = __ldg(&sub.state_map[last]);
    const i64 diff = s_min ^ s_max;
    const int h = diff ? 64 - __clzll(diff) : 0;  // bits [0, h) vary inside the block
    const i64 lowmask = (h >= 63) ? (i64)0x7fffffffffffffffll : (((i64)1 << h) - 1);
    const bool narrow = h <= 32;
    const int cap = narrow ? 2 * EX_CAP : EX_CAP;
    // bracket the image interval of every mask (2 searches per mask, all in parallel)
    __syncthreads();  // the previous block's tables are no longer in use
    for (int k = tid; k < 2 * msc.nmasks + 1; k += EX_ROWS) {
      if (k == 2 * msc.nmasks) {
        s_self = lower_bound_dev(sub.rmap_states, 0, sub.n, s_min & ~lowmask);
        continue;
      }
      const int mi = k >> 1;
      const i64 P = (s_min ^ __ldg(&msc.masks[mi])) & ~lowmask;
      if (k & 1) {
        // first state >= P + 2^h (the whole list when the interval reaches the top)
        s_hi[mi] = (h >= 63) ? sub.n : lower_bound_dev(sub.rmap_states, 0, sub.n, P + lowmask + 1);
      } else {
        s_lo[mi] = lower_bound_dev(sub.rmap_states, 0, sub.n, P);
      }
    }
    double ar = 0.0, ai = 0.0;
    int mi = 0;
    if (diag != nullptr) {
      if (ok) {
        const cplx xv = x[row];
        const double d = diag[row];
        ar = d * xv.x;
        ai = d * xv.y;
      }
      mi = 1;
    }
    __syncthreads();
    const i64 self = s_self;
    i64 staged_lo = -1, staged_len = -1;  // the interval whose slice is in shared memory (block-uniform)
    for (; mi < msc.nmasks; ++mi) {
      const i64 c_lo = s_lo[mi], len = s_hi[mi] - c_lo;
      const i64 m = __ldg(&msc.masks[mi]);
      const i64 bra = ket ^ m;
      i64 col = -1;
      bool need_search = len > 0;
      if (len > 0 && (m & lowmask) == 0) {
        // the mask only flips bits ABOVE the varying ones: the images keep the order of the kets.  When
        // the image interval has the same shape as the block's own (every structured space: the low-bit
        // patterns allowed next to the new prefix are the ones allowed next to the old) the column is the
        // same offset into it -- one coalesced probe verifies that; any mismatch falls back to the search
        const i64 g = c_lo + (row - self);
        const bool hit = ok && g < c_lo + len && __ldg(&sub.rmap_states[g]) == bra;
        if (hit) col = g;
        need_search = __syncthreads_or(ok && !hit) != 0;
      }
      if (need_search) {  // (block-uniform: every thread takes part in the staging barriers)
        if (len <= cap) {
          // every mask that only flips varying bits maps the block into its OWN interval: that slice is
          // staged once and reused
          if (c_lo != staged_lo || len != staged_len) {
            __syncthreads();  // everybody is done with the previous slice
            if (narrow) {
              for (i64 i = tid; i < len; i += EX_ROWS)
                s_slice32[i] = (unsigned int)(__ldg(&sub.rmap_states[c_lo + i]) & lowmask);
            } else {
              for (i64 i = tid; i < len; i += EX_ROWS) s_slice[i] = __ldg(&sub.rmap_states[c_lo + i]);
            }
            __syncthreads();
            staged_lo = c_lo;
            staged_len = len;
          }
          // branch-free lower bound: the same number of steps in every thread
          int n = (int)len, lo = 0;
          if (col >= 0) {
            // already placed by the probe
          } else if (narrow) {
            const unsigned int key = (unsigned int)(bra & lowmask);
            while (n > 1) {
              const int half = n >> 1;
              lo = (s_slice32[lo + half - 1] < key) ? lo + half : lo;
              n -= half;
            }
            lo += (s_slice32[lo] < key) ? 1 : 0;
            if (lo < (int)len && s_slice32[lo] == key) col = c_lo + lo;
          } else {
            while (n > 1) {
              const int half = n >> 1;
              lo = (s_slice[lo + half - 1] < bra) ? lo + half : lo;
              n -= half;
            }
            lo += (s_slice[lo] < bra) ? 1 : 0;
            if (lo < (int)len && s_slice[lo] == bra) col = c_lo + lo;
          }
        } else if (col < 0) {
          const i64 pos = lower_bound_dev(sub.rmap_states, c_lo, len, bra);
          if (pos < c_lo + len && __ldg(&sub.rmap_states[pos]) == bra) col = pos;
        }
      }
      if (!ok || col < 0) continue;  // outside the subspace
      double cr, ci;
      mask_element(msc, mi, bra, cr, ci);
      const cplx xv = x[col];
      ar += cr * xv.x - ci * xv.y;
      ai += cr * xv.y + ci * xv.x;
    }
    if (ok) y[row] = make_double2(ar, ai);
  }
}

// SpinConserve -> the same SpinConserve sector (the BASELINE eigsolve config): the row index IS
// the rank of the ket, so the column is row + rank_delta over the few bits the mask spans, and
// the binomial table is staged in shared memory.  Bit-exact with S2I (tests/test_gpu_matmult.py).
// NARROW: chains of at most 31 spins run the integer work in 32-bit registers (states, ranks and
// binomials all fit), which halves the instruction count of this issue-bound kernel.
template <bool NARROW>
__global__ void __launch_bounds__(TPB)
    k_mult_spinconserve(SubSpinConserve sub, MscDev msc, const double *__restrict__ diag, const cplx *__restrict__ x,
                        cplx *__restrict__ y, i64 M)
{
  typedef typename std::conditional<NARROW, unsigned int, unsigned long long>::type state_t;
  typedef typename std::conditional<NARROW, int, i64>::type rank_t;
  extern __shared__ unsigned char s_raw[];
  rank_t *tab = reinterpret_cast<rank_t *>(s_raw);
  const int ld = (int)sub.ld, k = (int)sub.k, L = (int)sub.L;
  for (int i = threadIdx.x; i < (k + 1) * ld; i += blockDim.x) tab[i] = (rank_t)sub.nck[i];
  __syncthreads();

  auto popc = [](state_t v) -> int { return NARROW ? __popc((unsigned int)v) : __popcll((unsigned long long)v); };
  auto ctz = [](state_t v) -> int { return NARROW ? __ffs((int)v) - 1 : __ffsll((long long)v) - 1; };
  auto clz = [](state_t v) -> int { return NARROW ? __clz((int)v) : __clzll((long long)v); };
  constexpr int TOP = NARROW ? 31 : 63;

  for (i64 row = blockIdx.x * (i64)blockDim.x + threadIdx.x; row < M; row += (i64)gridDim.x * blockDim.x) {
    // unrank (bsubspace_impl.h:210-228)
    state_t ket = 0;
    {
      rank_t idx = (rank_t)row;
      int kk = k;
      for (int n = L; n > 0; --n) {
        ket <<= 1;
        const rank_t c = (kk > n - 1) ? 0 : tab[kk * ld + n - 1];
        if (idx >= c) {
          idx -= c;
          --kk;
          ket |= 1;
        }
      }
    }
    double ar = 0.0, ai = 0.0;
    int mi = 0;
    if (diag != nullptr) {
      const cplx xv = x[row];
      const double d = diag[row];
      ar = d * xv.x;
      ai = d * xv.y;
      mi = 1;
    }
    for (; mi < msc.nmasks; ++mi) {
      const state_t mask = (state_t)__ldg(&msc.masks[mi]);
      const state_t bra = ket ^ mask;
      if (popc(bra) != k) continue;  // leaves the sector
      rank_t delta = 0;
      if (mask != 0) {
        // only the set bits inside the span of the mask change their (position, ordinal) pair
        const int lo = ctz(mask), hi = TOP - clz(mask);
        const state_t below_mask = ((state_t)1 << lo) - 1;
        const state_t span = (hi == TOP ? ~(state_t)0 : (((state_t)1 << (hi + 1)) - 1)) & ~below_mask;
        const int below = popc(ket & below_mask);
        state_t b = bra & span;
        int j = below;
        while (b) {
          const int n = ctz(b);
          ++j;
          if (j <= n) delta += tab[j * ld + n];
          b &= b - 1;
        }
        state_t a = ket & span;
        j = below;
        while (a) {
          const int n = ctz(a);
          ++j;
          if (j <= n) delta -= tab[j * ld + n];
          a &= a - 1;
        }
      }
      double cr, ci;
      mask_element(msc, mi, (i64)bra, cr, ci);
      const cplx xv = x[row + (i64)delta];
      ar += cr * xv.x - ci * xv.y;
      ai += cr * xv.y + ci * xv.x;
    }
    y[row] = make_double2(ar, ai);
  }
}

// bcuda_template_1.cu:29-66
template <class S>
__global__ void __launch_bounds__(TPB) k_diag(S sub, MscDev msc, double *__restrict__ diag, i64 M)
{
  const int t_end = msc.off_end[0];
  for (i64 row = blockIdx.x * (i64)blockDim.x + threadIdx.x; row < M; row += (i64)gridDim.x * blockDim.x) {
    const i64 state = sub.i2s(row);
    double v = 0.0;
    for (int t = 0; t < t_end; ++t) v += signed_coef(__ldg(&msc.coef[t]), state & __ldg(&msc.signs[t]));
    diag[row] = v;
  }
}

// bcuda_template_2.cu:331-403 : max over rows of sum_masks |element| (Kahan, as the CPU version :945-966)
template <class LS, class RS>
__global__ void __launch_bounds__(TPB) k_norm(LS ls, RS rs, MscDev msc, double *__restrict__ partials, i64 M)
{
  double best = 0.0;
  for (i64 row = blockIdx.x * (i64)blockDim.x + threadIdx.x; row < M; row += (i64)gridDim.x * blockDim.x) {
    const i64 ket = ls.i2s(row);
    double sum = 0.0, err = 0.0;
    for (int mi = 0; mi < msc.nmasks; ++mi) {
      const i64 bra = ket ^ __ldg(&msc.masks[mi]);
      if (rs.s2i(bra) < 0) continue;
      double cr, ci;
      mask_element(msc, mi, bra, cr, ci);
      const double comp = __dsub_rn(hypot(cr, ci), err);
      const double total = __dadd_rn(sum, comp);
      err = __dsub_rn(__dsub_rn(total, sum), comp);
      sum = total;
    }
    best = fmax(best, sum);
  }
  __shared__ double sh[TPB / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < TPB / 32; ++w) best = fmax(best, sh[w]);
    partials[blockIdx.x] = best;
  }
}

__global__ void k_max_partials(const double *__restrict__ partials, int n, double *__restrict__ out)
{
  double best = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) best = fmax(best, partials[i]);
  __shared__ double sh[TPB];
  sh[threadIdx.x] = best;
  __syncthreads();
  for (int s = TPB / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// bpetsc_template_2.c:990-1056.  Unsplit complex coefficients (the operator
// need not be Hermitian here); flag[0] is set to 1 when a non-zero element
// leaves the left subspace.
struct ConsDev {
  int nmasks;
  const i64 *masks;
  const i64 *offsets;  // nmasks+1
  const i64 *signs;
  const double *cre, *cim;
};

template <class LS, class RS>
__global__ void __launch_bounds__(TPB) k_check_conserves(LS ls, RS rs, ConsDev c, int *__restrict__ flag, i64 N)
{
  for (i64 col = blockIdx.x * (i64)blockDim.x + threadIdx.x; col < N; col += (i64)gridDim.x * blockDim.x) {
    if (*(volatile int *)flag) return;
    const i64 bra = rs.i2s(col);
    for (int mi = 0; mi < c.nmasks; ++mi) {
      const i64 ket = bra ^ __ldg(&c.masks[mi]);
      if (ls.s2i(ket) >= 0) continue;
      double vr = 0.0, vi = 0.0;
      for (i64 t = __ldg(&c.offsets[mi]); t < __ldg(&c.offsets[mi + 1]); ++t) {
        const i64 sv = bra & __ldg(&c.signs[t]);
        vr += signed_coef(__ldg(&c.cre[t]), sv);
        vi += signed_coef(__ldg(&c.cim[t]), sv);
      }
      if (vr != 0.0 || vi != 0.0) {
        *flag = 1;
        return;
      }
    }
  }
}

template <class S>
__global__ void k_s2i(S sub, i64 n, const i64 *__restrict__ states, i64 *__restrict__ idxs)
{
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    idxs[i] = sub.s2i(states[i]);
}

template <class S>
__global__ void k_i2s(S sub, i64 n, const i64 *__restrict__ idxs, i64 *__restrict__ states)
{
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    states[i] = sub.i2s(idxs[i]);
}

int row_grid(i64 rows)
{
  const i64 want = (rows + TPB - 1) / TPB;
  const i64 cap = (i64)G.sm_count * 16;  // 16 resident CTAs of 256 threads would be 2x oversubscribed: fine for tails
  return (int)std::max<i64>(1, std::min(want, cap));
}

// call f(device-side subspace struct) for the runtime type
template <class F>
void with_sub(const HostSubspace &s, F &&f)
{
  switch (s.desc.type) {
    case DNM_FULL: f(s.full()); break;
    case DNM_PARITY: f(s.parity()); break;
    case DNM_SPIN_CONSERVE: f(s.spin_dev()); break;
    case DNM_EXPLICIT: f(s.explicit_dev()); break;
    default: DNM_REQUIRE(false, DNM_ERR_ARG, "invalid subspace type");
  }
}

template <class T>
T *upload(const std::vector<T> &h, std::vector<void *> *owned)
{
  if (h.empty()) return nullptr;
  T *d = nullptr;
  DNM_CHECK_CUDA(cudaMalloc(&d, sizeof(T) * h.size()));
  if (owned) owned->push_back(d);
  DNM_CHECK_CUDA(cudaMemcpyAsync(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, G.stream));
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return d;
}

void validate_msc(int64_t nmasks, const int64_t *masks, const int64_t *offs, const int64_t *signs, const double *coeffs)
{
  DNM_REQUIRE(nmasks >= 1 && masks && offs && signs && coeffs, DNM_ERR_ARG, "empty or null MSC arrays");
  DNM_REQUIRE(offs[0] == 0, DNM_ERR_ARG, "mask_offsets[0] must be 0");
  for (int64_t m = 0; m < nmasks; ++m) {
    DNM_REQUIRE(offs[m + 1] > offs[m], DNM_ERR_ARG, "mask_offsets must be strictly increasing");
    if (m) DNM_REQUIRE(masks[m] > masks[m - 1], DNM_ERR_ARG, "msc must be sorted first");
  }
  DNM_REQUIRE(offs[nmasks] < ((int64_t)1 << 31), DNM_ERR_ARG, "too many terms");
}

}  // namespace

void general_mult(dnm_mat_s *A, const cplx *x, cplx *y)
{
  const i64 M = A->M;
  const dnm_subspace_t &l = A->left.desc, &r = A->right.desc;
  if (l.type == DNM_SPIN_CONSERVE && r.type == DNM_SPIN_CONSERVE && l.L == r.L && l.k == r.k && l.L < 63 &&
      getenv("DNM_NO_SC_KERNEL") == nullptr) {
    const SubSpinConserve sub = A->right.spin_dev();
    const size_t smem = sizeof(i64) * (size_t)((sub.k + 1) * sub.ld);
    if (smem <= 48 * 1024) {
      if (l.L <= 31) k_mult_spinconserve<true><<<row_grid(M), TPB, smem, G.stream>>>(sub, A->msc, A->d_diag, x, y, M);
      else k_mult_spinconserve<false><<<row_grid(M), TPB, smem, G.stream>>>(sub, A->msc, A->d_diag, x, y, M);
      count_launch();
      DNM_CHECK_CUDA(cudaGetLastError());
      return;
    }
  }
  if (l.type == DNM_EXPLICIT && r.type == DNM_EXPLICIT && A->same_explicit && A->left.rmap_idx.empty() &&
      A->msc.nmasks <= EX_MAX_MASKS && getenv("DNM_NO_EXPLICIT_KERNEL") == nullptr) {
    int rows = EX_ROWS_DEFAULT;
    if (const char *e = getenv("DNM_EX_ROWS")) rows = atoi(e);
    const i64 blocks = (M + rows - 1) / rows;
    const int grid = (int)std::max<i64>(1, std::min<i64>(blocks, (i64)G.sm_count * 16));
    if (rows == 128) k_mult_explicit<128><<<grid, 128, 0, G.stream>>>(A->right.explicit_dev(), A->msc, A->d_diag, x, y, M);
    else if (rows == 512) k_mult_explicit<512><<<grid, 512, 0, G.stream>>>(A->right.explicit_dev(), A->msc, A->d_diag, x, y, M);
    else k_mult_explicit<256><<<grid, 256, 0, G.stream>>>(A->right.explicit_dev(), A->msc, A->d_diag, x, y, M);
    count_launch();
    DNM_CHECK_CUDA(cudaGetLastError());
    return;
  }
  with_sub(A->left, [&](auto ls) {
    with_sub(A->right, [&](auto rs) {
      k_mult_general<<<row_grid(M), TPB, 0, G.stream>>>(ls, rs, A->msc, A->d_diag, x, y, M);
    });
  });
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
}

}  // namespace dnm

using namespace dnm;

extern "C" int dnm_mat_create(int64_t nmasks, const int64_t *masks, const int64_t *mask_offsets, const int64_t *signs,
                              const double *coeffs, const dnm_subspace_t *left, const dnm_subspace_t *right, int xparity,
                              dnm_mat_t *out)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(out, DNM_ERR_ARG, "null output handle");
  validate_msc(nmasks, masks, mask_offsets, signs, coeffs);
  const int64_t nterms = mask_offsets[nmasks];

  std::unique_ptr<dnm_mat_s> A(new dnm_mat_s());
  A->masks.assign(masks, masks + nmasks);
  A->mask_offsets.assign(mask_offsets, mask_offsets + nmasks + 1);
  A->signs.assign(signs, signs + nterms);
  A->coeffs.assign(coeffs, coeffs + 2 * nterms);
  A->xparity = xparity ? 1 : 0;
  A->left.copy_from(left);
  A->right.copy_from(right);
  DNM_REQUIRE(A->left.desc.L == A->right.desc.L, DNM_ERR_ARG, "left and right subspaces have different L");
  A->M = A->left.dim;
  A->N = A->right.dim;
  A->same_explicit = A->left.desc.type == DNM_EXPLICIT && A->right.desc.type == DNM_EXPLICIT &&
                     A->left.state_map == A->right.state_map;
  if (xparity) {  // bcuda_template_2.cu:19-22
    DNM_REQUIRE(A->M % 2 == 0 && A->N % 2 == 0, DNM_ERR_ARG, "XParity needs even parent dimensions");
    A->M /= 2;
    A->N /= 2;
  }

  // Hermiticity term by term (msc_tools.py:94-118), then split each mask's
  // terms into a real-coefficient run followed by an imaginary-coefficient run.
  std::vector<i64> h_signs(nterms);
  std::vector<double> h_coef(nterms);
  std::vector<int> off_re(nmasks), off_im(nmasks), off_end(nmasks);
  int64_t w = 0;
  for (int64_t m = 0; m < nmasks; ++m) {
    off_re[m] = (int)w;
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) off_im[m] = (int)w;
      for (int64_t t = mask_offsets[m]; t < mask_offsets[m + 1]; ++t) {
        const int imag_term = parity64(masks[m] & signs[t]);
        const double re = coeffs[2 * t], im = coeffs[2 * t + 1];
        if (pass == 0) {
          DNM_REQUIRE(imag_term ? (re == 0.0) : (im == 0.0), DNM_ERR_ARG,
                      "Building non-Hermitian matrices currently not supported.");
        }
        if (imag_term != pass) continue;
        h_signs[w] = signs[t];
        h_coef[w] = imag_term ? im : re;
        ++w;
      }
    }
    off_end[m] = (int)w;
  }

  A->left.upload();
  A->right.upload();
  A->msc.nmasks = (int)nmasks;
  A->msc.nterms = nterms;
  A->msc.masks = upload(A->masks, &A->owned);
  A->msc.off_re = upload(off_re, &A->owned);
  A->msc.off_im = upload(off_im, &A->owned);
  A->msc.off_end = upload(off_end, &A->owned);
  A->msc.signs = upload(h_signs, &A->owned);
  A->msc.coef = upload(h_coef, &A->owned);

  if (G.nranks > 1) {
    DNM_REQUIRE(tiled_supported(A.get()), DNM_ERR_UNSUPPORTED,
                "multi-GPU sharding is implemented for Full->Full and same-sector Parity->Parity only "
                "(other subspaces fit one GPU: run them as replicas)");
    DNM_REQUIRE(A->M % G.nranks == 0 && (A->M / G.nranks) >= 2, DNM_ERR_UNSUPPORTED,
                "dimension %lld too small to shard over %d ranks", (long long)A->M, G.nranks);
  }
  A->local_M = A->M / G.nranks;
  A->local_N = A->N / G.nranks;
  *out = A.release();
  DNM_API_END
}

extern "C" int dnm_mat_destroy(dnm_mat_t A)
{
  DNM_API_BEGIN
  if (!A) return DNM_OK;
  if (G.inited) cudaStreamSynchronize(G.stream);
  tiled_free(A);
  if (A->work_x) dnm_vec_destroy(A->work_x);
  if (A->work_y) dnm_vec_destroy(A->work_y);
  if (A->work_x2) dnm_vec_destroy(A->work_x2);
  if (A->work_y2) dnm_vec_destroy(A->work_y2);
  for (void *p : A->owned) cudaFree(p);
  if (A->d_diag) cudaFree(A->d_diag);
  A->left.release();
  A->right.release();
  delete A;
  DNM_API_END
}

extern "C" int dnm_mat_size(dnm_mat_t A, int64_t *M, int64_t *N)
{
  DNM_API_BEGIN
  DNM_REQUIRE(A, DNM_ERR_ARG, "null matrix");
  if (M) *M = A->M;
  if (N) *N = A->N;
  DNM_API_END
}

extern "C" int dnm_mat_precompute_diagonal(dnm_mat_t A)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(A, DNM_ERR_ARG, "null matrix");
  if (A->masks[0] != 0) return DNM_OK;  // no diagonal: leave diag unset (bpetsc_template_1.c:178-181)
  if (A->d_diag) return DNM_OK;
  DNM_REQUIRE(A->M == A->N && A->left.desc.type == A->right.desc.type, DNM_ERR_ARG,
              "precompute_diagonal needs identical left and right subspaces");
  DNM_CHECK_CUDA(cudaMalloc(&A->d_diag, sizeof(double) * A->local_M));
  const i64 M = A->local_M;
  if (G.nranks > 1) {
    tiled_diag(A, A->d_diag);
  } else {
    with_sub(A->right, [&](auto sub) { k_diag<<<row_grid(M), TPB, 0, G.stream>>>(sub, A->msc, A->d_diag, M); });
    count_launch();
    DNM_CHECK_CUDA(cudaGetLastError());
  }
  tiled_free(A);  // the pass plan depends on whether the diagonal is cached
  DNM_API_END
}

extern "C" int dnm_mat_mult(dnm_mat_t A, dnm_vec_t x, dnm_vec_t y)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(A && x && y, DNM_ERR_ARG, "null handle");
  DNM_REQUIRE(x->global_n == A->N, DNM_ERR_ARG, "input vector has length %lld, matrix has %lld columns",
              (long long)x->global_n, (long long)A->N);
  DNM_REQUIRE(y->global_n == A->M, DNM_ERR_ARG, "output vector has length %lld, matrix has %lld rows",
              (long long)y->global_n, (long long)A->M);
  DNM_REQUIRE(x != y && x->d != y->d, DNM_ERR_ARG, "MatMult cannot be done in place");
  const bool want_tiled = A->kernel_pref != 1 && tiled_supported(A);
  if (want_tiled) {
    tiled_mult(A, x, y);
    A->kernel_used = 2;
  } else {
    DNM_REQUIRE(G.nranks == 1, DNM_ERR_UNSUPPORTED, "general kernel is single-GPU only");
    general_mult(A, x->d, y->d);
    A->kernel_used = 1;
    A->launches_per_mult = 1;
  }
  DNM_API_END
}

extern "C" int dnm_mat_mult_host(dnm_mat_t A, const double *x_host, double *y_host)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(A && x_host && y_host, DNM_ERR_ARG, "null pointer");
  DNM_REQUIRE(G.nranks == 1, DNM_ERR_UNSUPPORTED, "dnm_mat_mult_host is single-rank");
  if (!A->work_x) {
    int rc = dnm_vec_create(A->N, &A->work_x);
    if (rc) return rc;
  }
  if (!A->work_y) {
    int rc = dnm_vec_create(A->M, &A->work_y);
    if (rc) return rc;
  }
  DNM_CHECK_CUDA(cudaMemcpyAsync(A->work_x->d, x_host, sizeof(cplx) * A->N, cudaMemcpyHostToDevice, G.stream));
  int rc = dnm_mat_mult(A, A->work_x, A->work_y);
  if (rc) return rc;
  DNM_CHECK_CUDA(cudaMemcpyAsync(y_host, A->work_y->d, sizeof(cplx) * A->M, cudaMemcpyDeviceToHost, G.stream));
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  DNM_API_END
}

// A stream of products with HOST buffers, software-pipelined over the PCIe link: while product k is
// evaluated, the input of product k+1 travels to the device on one copy stream and the result of
// product k-1 travels back on another (the link is full duplex), two device buffers deep on either
// side.  One product costs max(H2D, D2H) instead of H2D + MatMult + D2H.
extern "C" int dnm_mat_mult_host_batch(dnm_mat_t A, int64_t count, const double *const *x_hosts, double *const *y_hosts)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(A && count >= 0 && (count == 0 || (x_hosts && y_hosts)), DNM_ERR_ARG, "bad arguments to dnm_mat_mult_host_batch");
  DNM_REQUIRE(G.nranks == 1, DNM_ERR_UNSUPPORTED, "dnm_mat_mult_host_batch is single-rank");
  for (int64_t k = 0; k < count; ++k) DNM_REQUIRE(x_hosts[k] && y_hosts[k], DNM_ERR_ARG, "null host buffer (product %lld)", (long long)k);
  if (count == 0) return DNM_OK;
  dnm_vec_t *xb[2] = {&A->work_x, &A->work_x2}, *yb[2] = {&A->work_y, &A->work_y2};
  const int nb = count > 1 ? 2 : 1;
  for (int b = 0; b < nb; ++b) {
    if (!*xb[b]) {
      int rc = dnm_vec_create(A->N, xb[b]);
      if (rc) return rc;
    }
    if (!*yb[b]) {
      int rc = dnm_vec_create(A->M, yb[b]);
      if (rc) return rc;
    }
  }
  if (!G.copy_in) {
    DNM_CHECK_CUDA(cudaStreamCreateWithFlags(&G.copy_in, cudaStreamNonBlocking));
    DNM_CHECK_CUDA(cudaStreamCreateWithFlags(&G.copy_out, cudaStreamNonBlocking));
    DNM_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_batch_start, cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) {
      DNM_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_in[b], cudaEventDisableTiming));
      DNM_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_cmp[b], cudaEventDisableTiming));
      DNM_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_out[b], cudaEventDisableTiming));
    }
  }
  const size_t bytes_x = sizeof(cplx) * (size_t)A->N, bytes_y = sizeof(cplx) * (size_t)A->M;
  // whatever happens, no copy may still be in flight on the caller's buffers when this returns
  struct Drain {
    ~Drain()
    {
      cudaStreamSynchronize(G.copy_in);
      cudaStreamSynchronize(G.stream);
      cudaStreamSynchronize(G.copy_out);
    }
  } drain;
  // earlier work on the library stream may still use the staging vectors
  DNM_CHECK_CUDA(cudaEventRecord(G.ev_batch_start, G.stream));
  DNM_CHECK_CUDA(cudaStreamWaitEvent(G.copy_in, G.ev_batch_start, 0));
  DNM_CHECK_CUDA(cudaStreamWaitEvent(G.copy_out, G.ev_batch_start, 0));
  auto send = [&](int64_t k) {
    const int b = (int)(k & 1);
    if (k >= 2) DNM_CHECK_CUDA(cudaStreamWaitEvent(G.copy_in, G.ev_cmp[b], 0));  // product k-2 has consumed this buffer
    DNM_CHECK_CUDA(cudaMemcpyAsync((*xb[b])->d, x_hosts[k], bytes_x, cudaMemcpyHostToDevice, G.copy_in));
    DNM_CHECK_CUDA(cudaEventRecord(G.ev_in[b], G.copy_in));
  };
  send(0);
  for (int64_t k = 0; k < count; ++k) {
    const int b = (int)(k & 1);
    if (k + 1 < count) send(k + 1);
    DNM_CHECK_CUDA(cudaStreamWaitEvent(G.stream, G.ev_in[b], 0));
    if (k >= 2) DNM_CHECK_CUDA(cudaStreamWaitEvent(G.stream, G.ev_out[b], 0));  // result k-2 has left this buffer
    int rc = dnm_mat_mult(A, *xb[b], *yb[b]);
    if (rc) return rc;
    DNM_CHECK_CUDA(cudaEventRecord(G.ev_cmp[b], G.stream));
    DNM_CHECK_CUDA(cudaStreamWaitEvent(G.copy_out, G.ev_cmp[b], 0));
    DNM_CHECK_CUDA(cudaMemcpyAsync(y_hosts[k], (*yb[b])->d, bytes_y, cudaMemcpyDeviceToHost, G.copy_out));
    DNM_CHECK_CUDA(cudaEventRecord(G.ev_out[b], G.copy_out));
  }
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.copy_out));
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  DNM_API_END
}

extern "C" int dnm_mat_norm_inf(dnm_mat_t A, double *nrm)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(A && nrm, DNM_ERR_ARG, "null pointer");
  if (A->nrm >= 0) {  // cached (bcuda_template_2.cu:289-296)
    *nrm = A->nrm;
    return DNM_OK;
  }
  double *d_part = nullptr;
  if (G.nranks > 1) {
    tiled_norm(A, G.d_scratch);
  } else {
    const i64 M = A->M;
    const int grid = row_grid(M);
    DNM_CHECK_CUDA(cudaMalloc(&d_part, sizeof(double) * grid));
    with_sub(A->left, [&](auto ls) {
      with_sub(A->right, [&](auto rs) { k_norm<<<grid, TPB, 0, G.stream>>>(ls, rs, A->msc, d_part, M); });
    });
    count_launch();
    DNM_CHECK_CUDA(cudaGetLastError());
    k_max_partials<<<1, TPB, 0, G.stream>>>(d_part, grid, G.d_scratch);
    count_launch();
    DNM_CHECK_CUDA(cudaGetLastError());
  }
  double v = 0;
  fetch_doubles(G.d_scratch, &v, 1);
  if (d_part) cudaFree(d_part);
  A->nrm = v;
  *nrm = v;
  DNM_API_END
}

extern "C" int dnm_mat_set_option(dnm_mat_t A, const char *key, int64_t value)
{
  DNM_API_BEGIN
  DNM_REQUIRE(A && key, DNM_ERR_ARG, "null pointer");
  if (!strcmp(key, "kernel")) {
    DNM_REQUIRE(value >= 0 && value <= 2, DNM_ERR_ARG, "kernel must be 0 (auto), 1 (general) or 2 (tiled)");
    DNM_REQUIRE(value != 2 || tiled_supported(A), DNM_ERR_UNSUPPORTED,
                "tiled kernel needs Full->Full or same-sector Parity->Parity");
    A->kernel_pref = (int)value;
  } else if (!strcmp(key, "tile_bits")) {
    DNM_REQUIRE(value == 0 || (value >= 8 && value <= 13), DNM_ERR_ARG, "tile_bits must be 0 (auto) or in [8,13]");
    A->tile_bits = (int)value;
    tiled_free(A);
  } else if (!strcmp(key, "tile_rows")) {
    DNM_REQUIRE(value == 0 || value == 8 || value == 16, DNM_ERR_ARG, "tile_rows must be 0 (auto), 8 or 16");
    A->tile_rows = (int)value;
    tiled_free(A);
  } else if (!strcmp(key, "pipeline")) {
    DNM_REQUIRE(value >= 0 && value <= 2, DNM_ERR_ARG, "pipeline must be 0 (auto), 1 (on) or 2 (off)");
    A->pipeline = (int)value;
    tiled_free(A);
  } else if (!strcmp(key, "jit")) {
    DNM_REQUIRE(value >= -1 && value <= 1, DNM_ERR_ARG, "jit must be -1 (auto), 0 (off) or 1 (on)");
    A->jit = (int)value;
    tiled_free(A);
  } else if (!strcmp(key, "autotune")) {
    DNM_REQUIRE(value >= -1 && value <= 1, DNM_ERR_ARG, "autotune must be -1 (auto), 0 (off) or 1 (on)");
    A->autotune = (int)value;
    tiled_free(A);
  } else if (!strcmp(key, "far_bits")) {
    DNM_REQUIRE(value >= -1 && value <= 16, DNM_ERR_ARG, "far_bits must be -1 (auto) or in [0,16]");
    A->far_bits = (int)value;
    tiled_free(A);
  } else if (!strcmp(key, "verbose")) {
    A->verbose = (int)value;
  } else {
    DNM_REQUIRE(false, DNM_ERR_ARG, "unknown option '%s'", key);
  }
  DNM_API_END
}

extern "C" int dnm_mat_get_info(dnm_mat_t A, const char *key, double *value)
{
  DNM_API_BEGIN
  DNM_REQUIRE(A && key && value, DNM_ERR_ARG, "null pointer");
  // number of masks whose image is not identically outside the subspace is
  // not known here; the traffic model counts every unique mask
  const double nloc = (double)A->local_N;
  if (!strcmp(key, "kernel")) *value = A->kernel_used;
  else if (!strcmp(key, "passes")) *value = tiled_passes(A);
  else if (!strcmp(key, "unique_masks")) *value = (double)A->masks.size();
  else if (!strcmp(key, "nterms")) *value = (double)A->signs.size();
  else if (!strcmp(key, "model_bytes")) *value = ((double)A->masks.size() + 1.0) * nloc * 16.0;
  else if (!strcmp(key, "compulsory_bytes")) *value = 2.0 * nloc * 16.0 + (A->d_diag ? 8.0 * nloc : 0.0);
  else if (!strcmp(key, "launches_per_mult")) *value = A->launches_per_mult;
  else if (!strcmp(key, "has_diag")) *value = A->d_diag ? 1.0 : 0.0;
  else if (!strcmp(key, "jit_passes")) *value = tiled_jit_passes(A);
  else if (!strcmp(key, "tuned_shape")) *value = A->tuned_shape;
  else DNM_REQUIRE(false, DNM_ERR_ARG, "unknown info key '%s'", key);
  DNM_API_END
}

extern "C" int dnm_check_conserves(int64_t nmasks, const int64_t *masks, const int64_t *mask_offsets,
                                   const int64_t *signs, const double *coeffs, const dnm_subspace_t *left,
                                   const dnm_subspace_t *right, int xparity, int *result)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(result, DNM_ERR_ARG, "null pointer");
  validate_msc(nmasks, masks, mask_offsets, signs, coeffs);
  const int64_t nterms = mask_offsets[nmasks];
  HostSubspace L, R;
  L.copy_from(left);
  R.copy_from(right);
  L.upload();
  R.upload();
  std::vector<void *> owned;
  std::vector<i64> hm(masks, masks + nmasks), ho(mask_offsets, mask_offsets + nmasks + 1), hs(signs, signs + nterms);
  std::vector<double> cre(nterms), cim(nterms);
  for (int64_t t = 0; t < nterms; ++t) {
    cre[t] = coeffs[2 * t];
    cim[t] = coeffs[2 * t + 1];
  }
  ConsDev c;
  c.nmasks = (int)nmasks;
  int *d_flag = nullptr;
  int h_flag = 0;
  try {
    c.masks = upload(hm, &owned);
    c.offsets = upload(ho, &owned);
    c.signs = upload(hs, &owned);
    c.cre = upload(cre, &owned);
    c.cim = upload(cim, &owned);
    DNM_CHECK_CUDA(cudaMalloc(&d_flag, sizeof(int)));
    owned.push_back(d_flag);
    DNM_CHECK_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), G.stream));
    i64 N = R.dim;
    if (xparity) N /= 2;
    // rows are split over ranks like the reference's PetscLayout (:1011-1015)
    with_sub(L, [&](auto ls) {
      with_sub(R, [&](auto rs) { k_check_conserves<<<row_grid(N), TPB, 0, G.stream>>>(ls, rs, c, d_flag, N); });
    });
    count_launch();
    DNM_CHECK_CUDA(cudaGetLastError());
    DNM_CHECK_CUDA(cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, G.stream));
    DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  } catch (...) {
    for (void *p : owned) cudaFree(p);
    L.release();
    R.release();
    throw;
  }
  for (void *p : owned) cudaFree(p);
  L.release();
  R.release();
  *result = h_flag ? 0 : 1;
  DNM_API_END
}

namespace {
template <bool S2I>
void device_map(const dnm_subspace_t *s, int64_t n, const int64_t *in, int64_t *out)
{
  HostSubspace h;
  h.copy_from(s);
  h.upload();
  i64 *d_in = nullptr, *d_out = nullptr;
  try {
    if (!S2I)
      for (int64_t i = 0; i < n; ++i)
        DNM_REQUIRE(in[i] >= 0 && in[i] < h.dim, DNM_ERR_ARG,
                    "Index %lld is out of bounds for subspace of dimension %lld.", (long long)in[i], (long long)h.dim);
    DNM_CHECK_CUDA(cudaMalloc(&d_in, sizeof(i64) * std::max<int64_t>(n, 1)));
    DNM_CHECK_CUDA(cudaMalloc(&d_out, sizeof(i64) * std::max<int64_t>(n, 1)));
    DNM_CHECK_CUDA(cudaMemcpyAsync(d_in, in, sizeof(i64) * n, cudaMemcpyHostToDevice, G.stream));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 4096));
    with_sub(h, [&](auto sub) {
      if (S2I) k_s2i<<<grid, 256, 0, G.stream>>>(sub, n, d_in, d_out);
      else k_i2s<<<grid, 256, 0, G.stream>>>(sub, n, d_in, d_out);
    });
    count_launch();
    DNM_CHECK_CUDA(cudaGetLastError());
    DNM_CHECK_CUDA(cudaMemcpyAsync(out, d_out, sizeof(i64) * n, cudaMemcpyDeviceToHost, G.stream));
    DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  } catch (...) {
    if (d_in) cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    h.release();
    throw;
  }
  cudaFree(d_in);
  cudaFree(d_out);
  h.release();
}
}  // namespace

extern "C" int dnm_subspace_s2i_device(const dnm_subspace_t *s, int64_t n, const int64_t *states, int64_t *idxs)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(s && (n == 0 || (states && idxs)), DNM_ERR_ARG, "null pointer");
  if (n) device_map<true>(s, n, states, idxs);
  DNM_API_END
}

extern "C" int dnm_subspace_i2s_device(const dnm_subspace_t *s, int64_t n, const int64_t *idxs, int64_t *states)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(s && (n == 0 || (states && idxs)), DNM_ERR_ARG, "null pointer");
  if (n) device_map<false>(s, n, idxs, states);
  DNM_API_END
}
