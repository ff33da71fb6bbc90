// Subspace rank/unrank maps, shared by host code and device kernels.
//
// Bit-exact contract: /root/reference/src/dynamite/_backend/bsubspace_impl.h:57-361
// (device copies in bcuda_impl.cu:18-183).  Each struct is a POD that can be
// passed by value to a kernel; pointer members point to host memory when used
// from host code and to device memory inside kernels.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define DNM_HD __host__ __device__ __forceinline__
#else
#define DNM_HD inline
#endif

namespace dnm {

typedef int64_t i64;

DNM_HD int popc64(i64 v)
{
#if defined(__CUDA_ARCH__)
  return __popcll((unsigned long long)v);
#else
  return __builtin_popcountll((unsigned long long)v);
#endif
}

DNM_HD int parity64(i64 v) { return popc64(v) & 1; }

// index of lowest set bit; v != 0
DNM_HD int ctz64(i64 v)
{
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll((unsigned long long)v);
#endif
}

// leading zero bits; v != 0
DNM_HD int clz64(i64 v)
{
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return __builtin_clzll((unsigned long long)v);
#endif
}

struct SubFull {
  i64 L;
  DNM_HD i64 dim() const { return (i64)1 << L; }
  DNM_HD i64 i2s(i64 idx) const { return idx; }
  DNM_HD i64 s2i(i64 state) const { return state; }
};

struct SubParity {
  i64 L;
  i64 space;
  DNM_HD i64 dim() const { return (i64)1 << (L - 1); }
  // the lowest bit is whatever makes the popcount parity equal `space`
  DNM_HD i64 i2s(i64 idx) const { return (idx << 1) | (i64)(parity64(idx) ^ (int)space); }
  DNM_HD i64 s2i(i64 state) const { return (parity64(state) == (int)space) ? (state >> 1) : (i64)-1; }
};

struct SubSpinConserve {
  i64 L;
  i64 k;
  i64 ld;          // row length of nck
  const i64 *nck;  // nck[kk*ld + n] = C(n, kk), (k+1) x (L+1)
  DNM_HD i64 dim() const { return nck[k * ld + L]; }

  // combinatorial number system: sum over set bits (position n, ordinal j) of C(n, j)
  DNM_HD i64 rank_nocheck(i64 state) const
  {
    i64 idx = 0;
    i64 j = 0;
    while (state) {
      const int n = ctz64(state);
      ++j;
      if (j <= n) idx += nck[j * ld + n];
      state &= state - 1;
    }
    return idx;
  }
  DNM_HD i64 s2i(i64 state) const
  {
    if (popc64(state) != (int)k) return -1;
    return rank_nocheck(state);
  }
  // rank(bra) - rank(ket) for bra = ket ^ mask with equal popcounts (mask != 0): only the set
  // bits inside the span of the mask change their (position, ordinal) pair, so the sum runs
  // over those few bits instead of all k.  Exact integer arithmetic: same result as s2i(bra).
  DNM_HD i64 rank_delta(i64 ket, i64 bra, i64 mask) const
  {
    const int lo = ctz64(mask);
    const int hi = 63 - clz64(mask);
    const i64 span = (hi == 63 ? (i64)-1 : (((i64)1 << (hi + 1)) - 1)) & ~(((i64)1 << lo) - 1);
    const i64 below = popc64(ket & (((i64)1 << lo) - 1));
    i64 delta = 0;
    i64 b = bra & span, j = below;
    while (b) {
      const int n = ctz64(b);
      ++j;
      if (j <= n) delta += nck[j * ld + n];
      b &= b - 1;
    }
    i64 a = ket & span;
    j = below;
    while (a) {
      const int n = ctz64(a);
      ++j;
      if (j <= n) delta -= nck[j * ld + n];
      a &= a - 1;
    }
    return delta;
  }
  DNM_HD i64 i2s(i64 idx) const
  {
    i64 state = 0;
    i64 kk = k;
    for (i64 n = L; n > 0; --n) {
      state <<= 1;
      const i64 c = (kk > n - 1) ? 0 : nck[kk * ld + n - 1];
      if (idx >= c) {
        idx -= c;
        --kk;
        state |= 1;
      }
    }
    return state;
  }
};

struct SubExplicit {
  i64 L;
  i64 n;                  // dimension
  const i64 *state_map;   // idx -> state
  const i64 *rmap_idx;    // nullptr when state_map is sorted
  const i64 *rmap_states; // sorted states
  DNM_HD i64 dim() const { return n; }
  DNM_HD i64 i2s(i64 idx) const { return state_map[idx]; }
  DNM_HD i64 s2i(i64 state) const
  {
    // lower-bound search, then one equality test
    i64 lo = 0, len = n;
    while (len > 0) {
      const i64 half = len >> 1;
      const i64 v = rmap_states[lo + half];
      if (v < state) {
        lo += half + 1;
        len -= half + 1;
      } else {
        len = half;
      }
    }
    if (lo >= n || rmap_states[lo] != state) return -1;
    return rmap_idx ? rmap_idx[lo] : lo;
  }
};

}  // namespace dnm
