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

constexpr int MAX_SEGS = 12;

enum { PATH_SCALAR = 0, PATH_TWO = 1, PATH_GENERAL = 2, PATH_TABLE = 3, PATH_PAIR = 4, PATH_CTABLE = 5, PATH_WHT = 6 };

struct PassParams {
  int ngroups;
  int nterms;
  int B;           // log2 of the contiguous run length
  int n_outer;     // number of index bits outside the window
  int accumulate;  // 0: y = ..., 1: y += ... (read-modify-write), 2: y += ... with FP64 atomics (batched passes)
  int staged;      // large pass whose per-term tables are staged in shared memory behind the tile
  // group tables [ngroups]
  const u32 *lam;  // mask in window coordinates
  const u16 *t0;   // first term (terms t0..t1 have no r bits in their sign mask)
  const u16 *t1;   // first term with r bits
  const u16 *t2;   // one past the last term
  const u16 *pat;  // PATH_TWO: bit r = sign of the r-dependent terms on row group r
  const u8 *kp;    // bit 0: imaginary coefficients, bits 1..3: path
  // PATH_PAIR groups (at most two distinct sign masks s1, s2 -- every nearest-neighbour model):
  //   D(l) = sigma(s1 & l) * (c1 + c2 * sigma((s1^s2) & l)),   sigma(v) = (-1)^popcount(v)
  // stored as two "terms" at an even index t0: sw/rb hold the window bits / row patterns of s1 and
  // of s1^s2, so/cf the outside-window bits and coefficients of the two terms; per tile the
  // scratch holds (c1' + c2', c1' - c2') with the tile-uniform signs folded in.

  // PATH_TABLE groups (many terms per mask): the terms' sign masks span a GF(2) space of dimension
  // d <= 6; the group's "terms" t0..t1 are the d basis vectors and the coefficient of a row is
  // tabs[toff + index], index bit k = parity(basis_k & row)
  // PATH_CTABLE: the same for a mask with real and imaginary terms -- joint basis, table entries
  // are complex (re, im) pairs at the even offset toff, one gather serves both parts
  const double *tabs;
  const u16 *cls;                  // [ngroups * 8] PATH_WHT: end of row class rho among the terms t1..t2
  const u32 *toff;                 // [ngroups]
  const unsigned long long *rpat;  // [ngroups] byte r = index bits contributed by row group r
  // term tables [nterms]
  const u32 *sw;   // sign bits inside the window (window coordinates)
  const u32 *rb;   // bit r = parity((sw >> LOG_NT) & r)
  const i64 *so;   // sign bits outside the window (global index coordinates)
  const double *cf;
  const i64 *rowoff;  // [2^(T-B)] offset of each contiguous run
  i64 rank_bits;      // global index bits contributed by the rank
  i64 roff[16];       // offset contributed by the r-th row group of a thread
  unsigned char outer_pos[48];
  // the same deposit as runs of consecutive positions: outer = OR_j (tile_id & seg_mask[j]) << seg_shift[j]
  // (n_seg < 0: too many runs, use outer_pos bit by bit)
  int n_seg;
  signed char seg_shift[MAX_SEGS];  // negative: shift right (far positions above lower outer ones)
  unsigned long long seg_mask[MAX_SEGS];
  // every group is a PATH_PAIR group at t0 = 2g with row patterns of at most 8 bits: the kernels
  // run the lean group loop on SmallTables::gd
  int lean;
  // opt-in (DNM_ROWOFF_ARITH, untuned): rowoff[h] deposited arithmetically from runs of consecutive
  // window positions, like the tile number, instead of the dependent global load (n_rseg > 0)
  int n_rseg;
  unsigned char rseg_shift[MAX_SEGS];
  unsigned int rseg_mask[MAX_SEGS];
  int far_bits;  // lowest tile-number bits = the pass's L2 window (FAR groups flip only those outer positions)
  int debug;     // timing experiments only (DNM_TILE_DEBUG)
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
  u8 role[SMALL_TERMS];  // 0: ordinary term, 1 / 2: first / second entry of a PATH_PAIR group
  // lean passes: one 16-byte descriptor per group
  //   x = lam, y = window bits of s1, z = window bits of s1^s2, w = pat(s1) | pat(s1^s2) << 8 | imag << 16
  // x bit 31: the next group is the imaginary group of the same mask and neither has a row
  // pattern -- the two share one gather (X and Y fields, hopping with complex amplitudes)
  // w bit 17: FAR group -- the mask also flips index bits outside the window (inside the pass's
  // "L2 window", PassParams::far_bits); far[g] is the whole local flip mask in index coordinates
  // and the operand is read from global memory (an L2 hit: the tiles of one far-bit block run
  // concurrently), not from the tile
  uint4 gd[SMALL_GROUPS];
  unsigned long long far[SMALL_GROUPS];
  // FAR group whose operand lives on another rank (sharded vectors): x is read from rank ^ peer[g]
  // through its peer mapping, inside the pass (generated kernels only)
  unsigned char peer[SMALL_GROUPS];
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

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }

template <int N>
__device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// acc[r] += (IMAG ? i*d_r : d_r) * (row r ^ HI of the tile, this thread's partner column `col`).
// HI is a template parameter, so in a contiguous tile every LDS has an immediate offset and there
// is no per-row address arithmetic.  RING (ring kernel): rows live in pairs (quarters of the
// tile) at element offsets qoff[0..R/2) of the ring buffer.
// one gather serving a real and an imaginary group of the same mask: acc += (cr + i*ci) * x
template <int R, int LOG_NT, int HI>
__device__ __forceinline__ void gather_fixed_cplx(double (&ar)[R], double (&ai)[R], const double2 *col,
                                                  double cr, double ci)
{
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const double2 v = col[(r ^ HI) << LOG_NT];
    ar[r] += cr * v.x;
    ar[r] -= ci * v.y;
    ai[r] += cr * v.y;
    ai[r] += ci * v.x;
  }
}

template <int R, int LOG_NT>
__device__ __forceinline__ void gather_switch_cplx(double (&ar)[R], double (&ai)[R], const double2 *col, int hi_l, double cr, double ci)
{
#define DNM_HI_CASE(H) \
  case H:               \
    if (H < R) gather_fixed_cplx<R, LOG_NT, (H < R ? H : 0)>(ar, ai, col, cr, ci); \
    break;
  switch (hi_l) {
    DNM_HI_CASE(0) DNM_HI_CASE(1) DNM_HI_CASE(2) DNM_HI_CASE(3) DNM_HI_CASE(4) DNM_HI_CASE(5) DNM_HI_CASE(6)
    DNM_HI_CASE(7)
    default: break;
  }
#undef DNM_HI_CASE
}

// acc[r] += tab[q ^ rowbits(r)] * x(row r ^ HI): per-row complex coefficients read from a table
template <int R, int LOG_NT, int HI>
__device__ __forceinline__ void gather_fixed_ctab(double (&ar)[R], double (&ai)[R], const double2 *col,
                                                  const double2 *tab, u32 q, unsigned long long rp)
{
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const double2 c = __ldg(&tab[q ^ (u32)((rp >> (8 * (r & 7))) & 0xffull)]);
    const double2 v = col[(r ^ HI) << LOG_NT];
    ar[r] += c.x * v.x;
    ar[r] -= c.y * v.y;
    ai[r] += c.x * v.y;
    ai[r] += c.y * v.x;
  }
}

template <int R, int LOG_NT>
__device__ __forceinline__ void gather_switch_ctab(double (&ar)[R], double (&ai)[R], const double2 *col, int hi_l, const double2 *tab, u32 q,
                                                   unsigned long long rp)
{
#define DNM_HI_CASE(H) \
  case H:               \
    if (H < R) gather_fixed_ctab<R, LOG_NT, (H < R ? H : 0)>(ar, ai, col, tab, q, rp); \
    break;
  switch (hi_l) {
    DNM_HI_CASE(0) DNM_HI_CASE(1) DNM_HI_CASE(2) DNM_HI_CASE(3) DNM_HI_CASE(4) DNM_HI_CASE(5) DNM_HI_CASE(6)
    DNM_HI_CASE(7)
    default: break;
  }
#undef DNM_HI_CASE
}

template <int R, int LOG_NT, int HI, bool IMAG, bool SCALAR>
__device__ __forceinline__ void gather_fixed(double (&ar)[R], double (&ai)[R], const double2 *col,
                                             double d0, const double (&d)[R])
{
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const double c = SCALAR ? d0 : d[r];
    // rows with a zero coefficient (flip-flop terms: half of them) are not fetched; shared-memory
    // bandwidth is what bounds the arithmetic phase
    if (!SCALAR && c == 0.0) continue;
    const double2 v = col[(r ^ HI) << LOG_NT];
    if (IMAG) {
      ar[r] -= c * v.y;
      ai[r] += c * v.x;
    } else {
      ar[r] += c * v.x;
      ai[r] += c * v.y;
    }
  }
}

template <int R, int LOG_NT, bool IMAG, bool SCALAR>
__device__ __forceinline__ void gather_switch(double (&ar)[R], double (&ai)[R], const double2 *col,
                                              int hi_l, double d0, const double (&d)[R])
{
#define DNM_HI_CASE(H) \
  case H:               \
    if (H < R) gather_fixed<R, LOG_NT, (H < R ? H : 0), IMAG, SCALAR>(ar, ai, col, d0, d); \
    break;
  switch (hi_l) {
    DNM_HI_CASE(0) DNM_HI_CASE(1) DNM_HI_CASE(2) DNM_HI_CASE(3) DNM_HI_CASE(4) DNM_HI_CASE(5) DNM_HI_CASE(6)
    DNM_HI_CASE(7) DNM_HI_CASE(8) DNM_HI_CASE(9) DNM_HI_CASE(10) DNM_HI_CASE(11) DNM_HI_CASE(12) DNM_HI_CASE(13)
    DNM_HI_CASE(14) DNM_HI_CASE(15)
    default: break;
  }
#undef DNM_HI_CASE
}

// One tile of one pass.  `tile` is the CTA's 2^T-entry shared buffer, `csign` its per-term scratch.
// global index bits fixed by the tile number (the positions outside the window)
__device__ __forceinline__ i64 tile_outer_bits(const PassParams &P, unsigned long long tile_id)
{
  i64 g = 0;
  if (P.n_seg >= 0) {
    for (int j = 0; j < P.n_seg; ++j) {
      const unsigned long long v = tile_id & P.seg_mask[j];
      const int sh = P.seg_shift[j];
      g |= (i64)(sh >= 0 ? (v << sh) : (v >> (-sh)));
    }
  } else {
    for (int k = 0; k < P.n_outer; ++k) g |= (i64)((tile_id >> k) & 1ull) << P.outer_pos[k];
  }
  return g;
}

// this thread's part of the address: its rows differ only by the uniform P.roff[r]
__device__ __forceinline__ i64 thread_base(const PassParams &P, i64 outer)
{
  const int tid = threadIdx.x;
  if (P.n_rseg > 0) {
    const unsigned h = (unsigned)tid >> P.B;
    i64 off = 0;
    for (int j = 0; j < P.n_rseg; ++j) off |= (i64)(h & P.rseg_mask[j]) << P.rseg_shift[j];
    return outer | off | (i64)(tid & ((1 << P.B) - 1));
  }
  return outer | __ldg(&P.rowoff[tid >> P.B]) | (i64)(tid & ((1 << P.B) - 1));
}

// pull the 128-byte lines of a tile of `v` towards the L2 (one request per line)
template <int R>
__device__ __forceinline__ void prefetch_tile_l2(const PassParams &P, const cplx *v, i64 base_g)
{
  if ((threadIdx.x & 7) == 0 || P.B < 3) {
#pragma unroll
    for (int r = 0; r < R; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(v + (base_g | P.roff[r])));
  }
}

// Per-tile coefficient scratch: every term's coefficient with the tile-uniform sign applied (and,
// for staged passes, the window sign bits) -- written once per tile, read by every thread.
template <int NT, bool SMALL>
__device__ __forceinline__ void stage_tables(const PassParams &P, const SmallTables &S, i64 outer_g, double *csign)
{
  const int tid = threadIdx.x;
  uint2 *swrb = reinterpret_cast<uint2 *>(csign + P.nterms);  // staged passes only
  if (SMALL) {
    for (int t = tid; t < P.nterms; t += NT) {
      const int role = S.role[t];
      const double c = flip_sign(S.cf[t], __popcll((unsigned long long)(S.so[t] & outer_g)) & 1);
      if (role == 0) {
        csign[t] = c;
      } else {  // (c1 + c2, c1 - c2) of a pair group
        const int u = (role == 1) ? t + 1 : t - 1;
        const double o = flip_sign(S.cf[u], __popcll((unsigned long long)(S.so[u] & outer_g)) & 1);
        csign[t] = (role == 1) ? c + o : o - c;
      }
    }
  } else if (P.staged) {
    // many-term passes (SYK): the same, plus the window sign bits, staged once per tile so the
    // per-thread term loop reads shared memory instead of four global tables
    for (int t = tid; t < P.nterms; t += NT) {
      csign[t] = flip_sign(__ldg(&P.cf[t]), __popcll((unsigned long long)(__ldg(&P.so[t]) & outer_g)) & 1);
      swrb[t] = make_uint2(__ldg(&P.sw[t]), __ldg(&P.rb[t]));
    }
  }
}

// Stage one tile of x in shared memory (asynchronous 16-byte copies, runs of 2^B amplitudes) and
// the per-tile coefficient scratch; returns after the data is visible to the whole CTA.
template <int NT, int R, bool SMALL>
__device__ __forceinline__ void stage_tile(const PassParams &P, const SmallTables &S, i64 outer_g, i64 base_g,
                                           const cplx *__restrict__ x, double2 *tile, double *csign)
{
  const int tid = threadIdx.x;
#pragma unroll
  for (int r = 0; r < R; ++r) cp_async16(&tile[tid + r * NT], &x[base_g | P.roff[r]]);
  stage_tables<NT, SMALL>(P, S, outer_g, csign);
  cp_async_wait_all();
  __syncthreads();
}

// acc[r] += (IMAG ? i*d_r : d_r) * x[row r ^ m] for a FAR mask m: coalesced 16-byte loads from
// global memory (L2), four rows in flight at a time
__device__ __forceinline__ double2 ld_far(const cplx *p)
{
  double2 v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

template <int R, bool SCALAR>
__device__ __forceinline__ void gather_far(double (&ar)[R], double (&ai)[R], const cplx *__restrict__ xg, i64 base_g,
                                           const PassParams &P, i64 m, bool imag, double d0, const double (&d)[R])
{
  constexpr int CH = R < 4 ? R : 4;
#pragma unroll
  for (int h = 0; h < R; h += CH) {
    double2 v[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const double c = SCALAR ? d0 : d[h + k];
      v[k] = make_double2(0.0, 0.0);
      if (SCALAR || c != 0.0) v[k] = ld_far(xg + ((base_g | P.roff[h + k]) ^ m));
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const double c = SCALAR ? d0 : d[h + k];
      if (imag) {
        ar[h + k] -= c * v[k].y;
        ai[h + k] += c * v[k].x;
      } else {
        ar[h + k] += c * v[k].x;
        ai[h + k] += c * v[k].y;
      }
    }
  }
}

// The lean group loop (P.lean): every group has at most two distinct sign masks, so
//   D(l) = sigma(s1 & l) * (c1 + c2 * sigma((s1^s2) & l))
// and a group costs one 16-byte descriptor (kernel-parameter memory, uniform), one 16-byte
// scratch read (c1+c2, c1-c2 with the tile signs folded in), two popcounts and the gather.
template <int R, int LOG_NT>
__device__ __forceinline__ void process_groups_lean(const PassParams &P, const SmallTables &S, const double2 *tile,
                                                    const double *csign, double (&ar)[R], double (&ai)[R],
                                                    const cplx *__restrict__ xg, i64 base_g)
{
  constexpr int NT = 1 << LOG_NT;
  const int tid = threadIdx.x;
  const double2 *cpm = reinterpret_cast<const double2 *>(csign);
#pragma unroll 1
  for (int g = 0; g < P.ngroups; ++g) {
    const uint4 gd = S.gd[g];
    const double2 cc = cpm[g];
    const double2 *col = tile + (tid ^ (int)(gd.x & (NT - 1)));
    const int hi = (int)((gd.x & 0x7fffffffu) >> LOG_NT);
    const int pa = __popc(gd.y & (u32)tid) & 1, pb = __popc(gd.z & (u32)tid) & 1;
    const bool imag = ((gd.w >> 16) & 1u) != 0;
    const bool far = ((gd.w >> 17) & 1u) != 0;
    if ((gd.w & 0xffffu) == 0) {
      const double c = flip_sign(pb ? cc.y : cc.x, pa);
      if (far) {
        const double none[R] = {};
        if (c != 0.0) gather_far<R, true>(ar, ai, xg, base_g, P, (i64)S.far[g], imag, c, none);
        continue;
      }
      if (gd.x >> 31) {
        // real and imaginary group of one mask: one fetch, a complex coefficient
        ++g;
        const uint4 gi = S.gd[g];
        const double2 ci2 = cpm[g];
        const int qa = __popc(gi.y & (u32)tid) & 1, qb = __popc(gi.z & (u32)tid) & 1;
        const double ci = flip_sign(qb ? ci2.y : ci2.x, qa);
        if (c != 0.0 || ci != 0.0) gather_switch_cplx<R, LOG_NT>(ar, ai, col, hi, c, ci);
        continue;
      }
      if (c != 0.0) {
        const double none[R] = {};
        if (imag) gather_switch<R, LOG_NT, true, true>(ar, ai, col, hi, c, none);
        else gather_switch<R, LOG_NT, false, true>(ar, ai, col, hi, c, none);
      }
    } else {
      const u32 qa = (gd.w & 0xffu) ^ (0u - (u32)pa), qb = ((gd.w >> 8) & 0xffu) ^ (0u - (u32)pb);
      double d[R];
      bool any = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const double c = ((qb >> r) & 1u) ? cc.y : cc.x;
        d[r] = flip_if(__double2hiint(c), __double2loint(c), qa, r);
        any = any || (c != 0.0);
      }
      if (any && far) {
        gather_far<R, false>(ar, ai, xg, base_g, P, (i64)S.far[g], imag, 0.0, d);
      } else if (any) {
        if (imag) gather_switch<R, LOG_NT, true, false>(ar, ai, col, hi, 0.0, d);
        else gather_switch<R, LOG_NT, false, false>(ar, ai, col, hi, 0.0, d);
      }
    }
  }
}

// Accumulate every group of the pass into this thread's R rows  l = tid + r*NT.
// Contiguous tile: `tile` is the 2^T-entry buffer.  RING mode (ring kernel): the tile is four
// quarters (row pairs) at element offsets qoff[0..3] of `tile`.
template <int R, int LOG_NT, bool SMALL>
__device__ __forceinline__ void process_groups(const PassParams &P, const SmallTables &S, const double2 *tile,
                                               const double *csign, i64 outer_g, double (&ar)[R], double (&ai)[R],
                                               const cplx *__restrict__ xg, i64 base_g)
{
  if (SMALL && R <= 8 && P.lean) return process_groups_lean<R, LOG_NT>(P, S, tile, csign, ar, ai, xg, base_g);
  constexpr int NT = 1 << LOG_NT;
  constexpr int rbase = 0;
  constexpr int RH = R;
  const int tid = threadIdx.x;
  const uint2 *swrb = reinterpret_cast<const uint2 *>(csign + P.nterms);  // staged passes only

  // coefficient of term t with tile and thread signs applied
  auto term = [&](int t) -> double {
    if (SMALL) return flip_sign(csign[t], __popc(S.sw[t] & (u32)tid) & 1);
    if (P.staged) return flip_sign(csign[t], __popc(swrb[t].x & (u32)tid) & 1);
    const int p = (__popcll((unsigned long long)(__ldg(&P.so[t]) & outer_g)) ^ __popc(__ldg(&P.sw[t]) & (u32)tid)) & 1;
    return flip_sign(__ldg(&P.cf[t]), p);
  };

#pragma unroll 1
  for (int g = 0; g < P.ngroups; ++g) {
    const u32 lam = SMALL ? S.lam[g] : __ldg(&P.lam[g]);
    const int base = tid ^ (int)(lam & (NT - 1));
    const int hi_l = (int)(lam >> LOG_NT);
    const int t0 = SMALL ? (int)S.t0[g] : (int)__ldg(&P.t0[g]);
    const int kp = SMALL ? (int)S.kp[g] : (int)__ldg(&P.kp[g]);
    const bool imag = kp & 1;
    const int path = kp >> 1;
    // row r gathers from row (r ^ hi_l) of column `base`
    const double2 *col = tile + base;
    const int hi_lo = hi_l;

    if (path == PATH_PAIR) {
      // the lean path: no term loops, one 16-byte scratch read, two popcounts
      const u32 swa = SMALL ? S.sw[t0] : __ldg(&P.sw[t0]);
      const u32 swb = SMALL ? S.sw[t0 + 1] : __ldg(&P.sw[t0 + 1]);
      const u32 pata = ((SMALL ? S.rb[t0] : __ldg(&P.rb[t0])) >> rbase) & ((1u << RH) - 1);
      const u32 patb = ((SMALL ? S.rb[t0 + 1] : __ldg(&P.rb[t0 + 1])) >> rbase) & ((1u << RH) - 1);
      double2 cc;  // (c1 + c2, c1 - c2), tile signs applied
      if (SMALL) {
        cc = *reinterpret_cast<const double2 *>(csign + t0);
      } else {
        const double c1 = flip_sign(__ldg(&P.cf[t0]), __popcll((unsigned long long)(__ldg(&P.so[t0]) & outer_g)) & 1);
        const double c2 =
            flip_sign(__ldg(&P.cf[t0 + 1]), __popcll((unsigned long long)(__ldg(&P.so[t0 + 1]) & outer_g)) & 1);
        cc = make_double2(c1 + c2, c1 - c2);
      }
      const int pa = __popc(swa & (u32)tid) & 1, pb = __popc(swb & (u32)tid) & 1;
      if ((pata | patb) == 0) {
        const double c = flip_sign(pb ? cc.y : cc.x, pa);
        if (c != 0.0) {
          const double none[RH] = {};
          if (imag) gather_switch<RH, LOG_NT, true, true>(ar, ai, col, hi_lo, c, none);
          else gather_switch<RH, LOG_NT, false, true>(ar, ai, col, hi_lo, c, none);
        }
      } else {
        const u32 qa = pata ^ (0u - (u32)pa), qb = patb ^ (0u - (u32)pb);
        double d[RH];
        bool any = false;
#pragma unroll
        for (int r = 0; r < RH; ++r) {
          const double c = ((qb >> r) & 1u) ? cc.y : cc.x;
          d[r] = flip_if(__double2hiint(c), __double2loint(c), qa, r);
          any = any || (c != 0.0);
        }
        if (any) {
          if (imag) gather_switch<RH, LOG_NT, true, false>(ar, ai, col, hi_lo, 0.0, d);
          else gather_switch<RH, LOG_NT, false, false>(ar, ai, col, hi_lo, 0.0, d);
        }
      }
      continue;
    }

    const int t1 = SMALL ? (int)S.t1[g] : (int)__ldg(&P.t1[g]);
    const int t2 = SMALL ? (int)S.t2[g] : (int)__ldg(&P.t2[g]);
    if constexpr (!SMALL && R <= 8) if (path == PATH_CTABLE) {
      // index of this thread into the group's table of complex coefficients: one parity per basis vector
      u32 q = 0;
      for (int t = t0; t < t1; ++t) {
        u32 p;
        if (P.staged) p = (u32)(__double2hiint(csign[t]) >> 31) ^ (u32)__popc(swrb[t].x & (u32)tid);
        else p = (u32)__popcll((unsigned long long)(__ldg(&P.so[t]) & outer_g)) ^ (u32)__popc(__ldg(&P.sw[t]) & (u32)tid);
        q |= (p & 1u) << (t - t0);
      }
      const double2 *tab = reinterpret_cast<const double2 *>(P.tabs + __ldg(&P.toff[g]));
      gather_switch_ctab<R, LOG_NT>(ar, ai, col, hi_lo, tab, q, __ldg(&P.rpat[g]));
      continue;
    }
    double c0 = 0.0;  // terms without r bits: one scalar per thread
    if (SMALL || path != PATH_TABLE)
      for (int t = t0; t < t1; ++t) c0 += term(t);

    if (path == PATH_SCALAR) {
      if (c0 != 0.0) {
        const double none[RH] = {};
        if (imag) gather_switch<RH, LOG_NT, true, true>(ar, ai, col, hi_lo, c0, none);
        else gather_switch<RH, LOG_NT, false, true>(ar, ai, col, hi_lo, c0, none);
      }
    } else {
      double d[RH];
      bool any;
      if (path == PATH_TWO) {
        double ch = 0.0;  // terms sharing the single r pattern `pat`
        for (int t = t1; t < t2; ++t) ch += term(t);
        const u32 pat = (SMALL ? (u32)S.pat[g] : (u32)__ldg(&P.pat[g])) >> rbase;
        const double dp = c0 + ch, dm = c0 - ch;
#pragma unroll
        for (int r = 0; r < RH; ++r) d[r] = ((pat >> r) & 1u) ? dm : dp;
        any = (dp != 0.0) || (dm != 0.0);
      } else if (!SMALL && R == 8 && path == PATH_WHT) {
        // E[rho] = sum of the class-rho terms (per-thread scalars); d[r] = sum_rho (-1)^popc(rho & r) E[rho]
        double E[8];
        int t = t1;
#pragma unroll
        for (int rho = 0; rho < 8; ++rho) {
          const int te = (int)__ldg(&P.cls[g * 8 + rho]);
          double e = (rho == 0) ? c0 : 0.0;
          for (; t < te; ++t) e += term(t);
          E[rho] = e;
        }
#pragma unroll
        for (int h = 1; h < 8; h <<= 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (!(i & h)) {
              const double a = E[i], b = E[i | h];
              E[i] = a + b;
              E[i | h] = a - b;
            }
          }
        }
        any = false;
#pragma unroll
        for (int r = 0; r < RH; ++r) {
          d[r] = E[r & 7];
          any = any || (d[r] != 0.0);
        }
      } else if (!SMALL && path == PATH_TABLE) {
        // index of this thread's rows into the group's coefficient table: one parity per basis vector
        u32 q = 0;
        for (int t = t0; t < t1; ++t) {
          u32 p;
          if (P.staged) p = (u32)(__double2hiint(csign[t]) >> 31) ^ (u32)__popc(swrb[t].x & (u32)tid);
          else p = (u32)__popcll((unsigned long long)(__ldg(&P.so[t]) & outer_g)) ^ (u32)__popc(__ldg(&P.sw[t]) & (u32)tid);
          q |= (p & 1u) << (t - t0);
        }
        const unsigned long long rp = __ldg(&P.rpat[g]);
        const double *tab = P.tabs + __ldg(&P.toff[g]);
        any = false;
#pragma unroll
        for (int r = 0; r < RH; ++r) {
          d[r] = __ldg(&tab[q ^ (u32)((rp >> (8 * ((rbase + r) & 7))) & 0xffull)]);
          any = any || (d[r] != 0.0);
        }
      } else {
#pragma unroll
        for (int r = 0; r < RH; ++r) d[r] = c0;
        for (int t = t1; t < t2; ++t) {
          const double c = term(t);
          const u32 bits = (SMALL ? S.rb[t] : (P.staged ? swrb[t].y : __ldg(&P.rb[t]))) >> rbase;
          const int chi = __double2hiint(c), clo = __double2loint(c);
#pragma unroll
          for (int r = 0; r < RH; ++r) d[r] += flip_if(chi, clo, bits, r);
        }
        any = false;
#pragma unroll
        for (int r = 0; r < RH; ++r) any = any || (d[r] != 0.0);
      }
      if (any) {
        if (imag) gather_switch<RH, LOG_NT, true, false>(ar, ai, col, hi_lo, 0.0, d);
        else gather_switch<RH, LOG_NT, false, false>(ar, ai, col, hi_lo, 0.0, d);
      }
    }
  }
}

// One tile of one pass.  `tile` is the CTA's 2^T-entry shared buffer, `csign` its per-term scratch.
template <int T, int R, bool SMALL>
__device__ __forceinline__ void tile_body(const PassParams &P, const SmallTables &S, i64 outer, i64 base_g,
                                          const cplx *__restrict__ x, cplx *__restrict__ y,
                                          const double *__restrict__ diag, double2 *tile, double *csign)
{
  constexpr int NT = TileCfg<T, R>::NT;
  constexpr int LOG_NT = TileCfg<T, R>::LOG_NT;
  const int tid = threadIdx.x;
  const i64 outer_g = outer | P.rank_bits;  // sign-relevant bits shared by the whole tile
  stage_tile<NT, R, SMALL>(P, S, outer_g, base_g, x, tile, csign);

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

  process_groups<R, LOG_NT, SMALL>(P, S, tile, csign, outer_g, ar, ai, x, base_g);

  if (P.accumulate == 2) {
    // passes of a small problem run concurrently in one grid: combine in the L2 with FP64 atomics
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double *dst = reinterpret_cast<double *>(y + (base_g | P.roff[r]));
      atomicAdd(dst, ar[r]);
      atomicAdd(dst + 1, ai[r]);
    }
  } else if (P.accumulate) {
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

// one pass, one tile per CTA
template <int T, int R, bool SMALL>
__global__ void __launch_bounds__(TileCfg<T, R>::NT, TileCfg<T, R>::MINB)
    k_tiled(const __grid_constant__ PassParams P, const __grid_constant__ SmallTables S,
            const cplx *__restrict__ x, cplx *__restrict__ y, const double *__restrict__ diag)
{
  extern __shared__ double2 tile[];
  __shared__ __align__(16) double csign_small[SMALL ? SMALL_TERMS : 2];
  // staged large passes keep their term tables in dynamic shared memory right behind the tile
  double *csign = SMALL ? csign_small : reinterpret_cast<double *>(tile + (1 << T));
  const i64 outer = tile_outer_bits(P, blockIdx.x);
  tile_body<T, R, SMALL>(P, S, outer, thread_base(P, outer), x, y, diag, tile, csign);
}

// every pass of a small problem in ONE launch (blockIdx.y = pass): a pass of an L2-resident vector
// has too few tiles to fill the GPU, and the passes only interact through y, which is combined
// with atomics.  Tables come from global memory (PassParams array on the device).
template <int T, int R>
__global__ void __launch_bounds__(TileCfg<T, R>::NT, TileCfg<T, R>::MINB)
    k_tiled_batch(const PassParams *__restrict__ Ps, const cplx *__restrict__ x, cplx *__restrict__ y,
                  const double *__restrict__ diag, int diag_pass)
{
  extern __shared__ double2 tile[];
  const PassParams &P = Ps[blockIdx.y];
  const SmallTables &unused = *reinterpret_cast<const SmallTables *>(Ps);  // never read when SMALL == false
  double *csign = reinterpret_cast<double *>(tile + (1 << T));
  const i64 outer = tile_outer_bits(P, blockIdx.x);
  tile_body<T, R, false>(P, unused, outer, thread_base(P, outer), x, y, ((int)blockIdx.y == diag_pass) ? diag : nullptr,
                         tile, csign);
}

}  // namespace tiled
}  // namespace dnm
