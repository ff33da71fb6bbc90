// The window-tiled MatMult kernel (device side).  See matmult_tiled.cu for the
// formulation and the host planner.
//
// One CTA owns one tile of 2^T amplitudes held in shared memory; thread `tid`
// owns the R rows  l = tid + r*NT  (window coordinates), so the low LOG_NT
// window bits of a row are the thread id and the high log2(R) bits are r.
//
// Work is organised in GROUPS: one (mask, real|imaginary) pair with its terms.
// For row l the coefficient is  D(l) = sum_t c_t * s_tile(t) * s_tid(t) * s_r(t)
// with three sign factors: tile-uniform (bits outside the window), per-thread
// (window bits covered by tid) and per-row (window bits covered by r).  Terms
// whose sign mask has no r bits contribute a per-thread scalar; most physical
// Hamiltonians have at most ONE distinct r-pattern per group, which gives two
// per-thread scalars (Dp, Dm) and a uniform 16-bit row pattern -- no per-row
// integer work at all, only LDS.128 + 2 DFMA per (row, mask).  Anything else
// takes the general path (per-term, per-row sign application).
#pragma once

#include "context.h"

namespace dnm {
namespace tiled {

typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

constexpr int SMALL_GROUPS = 48;  // passes up to this size keep their tables in kernel-parameter
constexpr int SMALL_TERMS = 96;   // (constant) memory; bigger ones read them from global memory

enum { PATH_SCALAR = 0, PATH_TWO = 1, PATH_GENERAL = 2 };

struct PassParams {
  int ngroups;
  int nterms;
  int B;           // log2 of the contiguous run length
  int n_outer;     // number of index bits outside the window
  int accumulate;  // 0: y = ..., 1: y += ...
  // group tables [ngroups]
  const u32 *lam;  // mask in window coordinates
  const u16 *t0;   // first term (terms t0..t1 have no r bits in their sign mask)
  const u16 *t1;   // first term with r bits
  const u16 *t2;   // one past the last term
  const u16 *pat;  // PATH_TWO: bit r = sign of the r-dependent terms on row group r
  const u8 *kp;    // bit 0: imaginary coefficients, bits 1..2: path
  // term tables [nterms]
  const u32 *sw;   // sign bits inside the window (window coordinates)
  const u32 *rb;   // bit r = parity((sw >> LOG_NT) & r)
  const i64 *so;   // sign bits outside the window (global index coordinates)
  const double *cf;
  const i64 *rowoff;  // [2^(T-B)] offset of each contiguous run
  i64 rank_bits;      // global index bits contributed by the rank
  i64 roff[16];       // offset contributed by the r-th row group of a thread
  unsigned char outer_pos[48];
};

// the same tables by value, for small passes (terms indices fit u8)
struct SmallTables {
  u32 lam[SMALL_GROUPS];
  u16 pat[SMALL_GROUPS];
  u8 t0[SMALL_GROUPS], t1[SMALL_GROUPS], t2[SMALL_GROUPS], kp[SMALL_GROUPS];
  u32 sw[SMALL_TERMS];
  u32 rb[SMALL_TERMS];
  i64 so[SMALL_TERMS];
  double cf[SMALL_TERMS];
};

constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v / 2); }

template <int T, int R>
struct TileCfg {
  static constexpr int NT = (1 << T) / R;
  static constexpr int LOG_NT = T - ilog2c(R);
  static constexpr int REGS = (R >= 16) ? 128 : 64;
  static constexpr int MINB_RAW = 65536 / (NT * REGS);
  static constexpr int MINB = MINB_RAW < 1 ? 1 : (MINB_RAW > 8 ? 8 : MINB_RAW);
};

__device__ __forceinline__ double flip_sign(double c, int p)
{
  // negate when p (0/1) is set: xor into the IEEE sign bit
  return __hiloint2double(__double2hiint(c) ^ (p << 31), __double2loint(c));
}

__device__ __forceinline__ double flip_if(int hi, int lo, u32 bits, int r)
{
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

// acc[r] += (IMAG ? i*d : d) * tile[row r ^ lam]   for the rows selected by `rows` (uniform bit mask)
template <int R, int LOG_NT, bool IMAG, bool ALL>
__device__ __forceinline__ void gather_rows(double (&ar)[R], double (&ai)[R], const double2 *tile, int base, int hi_l,
                                            double d, u32 rows)
{
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (ALL || ((rows >> r) & 1u)) {
      const double2 v = tile[base + ((r ^ hi_l) << LOG_NT)];
      if (IMAG) {
        ar[r] -= d * v.y;
        ai[r] += d * v.x;
      } else {
        ar[r] += d * v.x;
        ai[r] += d * v.y;
      }
    }
  }
}

template <int T, int R, bool SMALL>
__global__ void __launch_bounds__(TileCfg<T, R>::NT, TileCfg<T, R>::MINB)
    k_tiled(const __grid_constant__ PassParams P, const __grid_constant__ SmallTables S,
            const cplx *__restrict__ x, cplx *__restrict__ y, const double *__restrict__ diag)
{
  constexpr int NT = TileCfg<T, R>::NT;
  constexpr int LOG_NT = TileCfg<T, R>::LOG_NT;
  constexpr u32 ALLROWS = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
  extern __shared__ double2 tile[];
  __shared__ double csign[SMALL ? SMALL_TERMS : 1];
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
  if (SMALL) {
    // coefficient with the tile-uniform sign applied, once per tile
    for (int t = tid; t < P.nterms; t += NT)
      csign[t] = flip_sign(S.cf[t], __popcll((unsigned long long)(S.so[t] & outer_g)) & 1);
  }
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

  // coefficient of term t with tile and thread signs applied
  auto term = [&](int t) -> double {
    if (SMALL) return flip_sign(csign[t], __popc(S.sw[t] & (u32)tid) & 1);
    const int p = (__popcll((unsigned long long)(__ldg(&P.so[t]) & outer_g)) ^ __popc(__ldg(&P.sw[t]) & (u32)tid)) & 1;
    return flip_sign(__ldg(&P.cf[t]), p);
  };

#pragma unroll 1
  for (int g = 0; g < P.ngroups; ++g) {
    const u32 lam = SMALL ? S.lam[g] : __ldg(&P.lam[g]);
    const int base = tid ^ (int)(lam & (NT - 1));
    const int hi_l = (int)(lam >> LOG_NT);
    const int t0 = SMALL ? (int)S.t0[g] : (int)__ldg(&P.t0[g]);
    const int t1 = SMALL ? (int)S.t1[g] : (int)__ldg(&P.t1[g]);
    const int t2 = SMALL ? (int)S.t2[g] : (int)__ldg(&P.t2[g]);
    const int kp = SMALL ? (int)S.kp[g] : (int)__ldg(&P.kp[g]);
    const bool imag = kp & 1;
    const int path = kp >> 1;

    double c0 = 0.0;  // terms without r bits: one scalar per thread
    for (int t = t0; t < t1; ++t) c0 += term(t);

    if (path == PATH_SCALAR) {
      if (c0 != 0.0) {
        if (imag) gather_rows<R, LOG_NT, true, true>(ar, ai, tile, base, hi_l, c0, ALLROWS);
        else gather_rows<R, LOG_NT, false, true>(ar, ai, tile, base, hi_l, c0, ALLROWS);
      }
    } else if (path == PATH_TWO) {
      double ch = 0.0;  // terms sharing the single r pattern `pat`
      for (int t = t1; t < t2; ++t) ch += term(t);
      const u32 pat = SMALL ? (u32)S.pat[g] : (u32)__ldg(&P.pat[g]);
      const double dp = c0 + ch, dm = c0 - ch;
      if (dp != 0.0) {
        if (imag) gather_rows<R, LOG_NT, true, false>(ar, ai, tile, base, hi_l, dp, ~pat & ALLROWS);
        else gather_rows<R, LOG_NT, false, false>(ar, ai, tile, base, hi_l, dp, ~pat & ALLROWS);
      }
      if (dm != 0.0) {
        if (imag) gather_rows<R, LOG_NT, true, false>(ar, ai, tile, base, hi_l, dm, pat);
        else gather_rows<R, LOG_NT, false, false>(ar, ai, tile, base, hi_l, dm, pat);
      }
    } else {
      double d[R];
#pragma unroll
      for (int r = 0; r < R; ++r) d[r] = c0;
      for (int t = t1; t < t2; ++t) {
        const double c = term(t);
        const u32 bits = SMALL ? S.rb[t] : __ldg(&P.rb[t]);
        const int chi = __double2hiint(c), clo = __double2loint(c);
#pragma unroll
        for (int r = 0; r < R; ++r) d[r] += flip_if(chi, clo, bits, r);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (d[r] != 0.0) {
          const double2 v = tile[base + ((r ^ hi_l) << LOG_NT)];
          if (imag) {
            ar[r] -= d[r] * v.y;
            ai[r] += d[r] * v.x;
          } else {
            ar[r] += d[r] * v.x;
            ai[r] += d[r] * v.y;
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

}  // namespace tiled
}  // namespace dnm
