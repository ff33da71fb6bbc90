// Chebyshev propagator  y = exp(i s A) x  for a Hermitian A with spectrum inside [-a, a].
//
// Why it exists.  dynamite evolves states with SLEPc's expokit (computations.py:89-112): sub-steps of an
// m-dimensional Krylov approximation, m = 30 by default.  On one B200 an L=30 state is 16 GiB, so the
// basis that fits is m = 6..8 and expokit degenerates into hundreds of tiny sub-steps (590 MatMults for
// BASELINE C3, t*||H|| = 50).  The Jacobi-Anger expansion
//
//     exp(i z x) = J_0(z) + 2 sum_{k>=1} i^k J_k(z) T_k(x),        x in [-1, 1],  z = s*a,
//
// needs THREE work vectors whatever its degree, no inner products, no host round trips, and its
// coefficients fall off super-exponentially once k > |z|: about |z| + 8 |z|^(1/3) + 10 MatMults for
// fifteen digits (87 for z = 50).  The bound a = ||A||_inf >= rho(A) is the one the reference already
// computes for its own step-size estimate (computations.py:185-194, MatNorm).
//
// No CUDA in this header: the recurrence is written against an `Ops` concept so that the SAME code
// runs on device vectors (krylov.cu) and, in the CPU tests, on std::vector with a dense matrix
// (tests/test_chebyshev_host.py compiles tests/cheb_host.cpp against this file).
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

namespace dnm {
namespace cheb {

// J_0(x) .. J_nmax(x), x >= 0, by Miller's backward recurrence J_{k-1} = (2k/x) J_k - J_{k+1} started
// well above max(nmax, x) and normalised with J_0 + 2 (J_2 + J_4 + ...) = 1.
inline std::vector<double> bessel_j(int nmax, double x)
{
  std::vector<double> out((size_t)nmax + 1, 0.0);
  if (x == 0.0) {
    out[0] = 1.0;
    return out;
  }
  const int top = std::max(nmax, (int)std::ceil(x));
  const int start = top + 40 + (int)std::ceil(6.0 * std::cbrt((double)top + 1.0));
  std::vector<long double> j((size_t)start + 2, 0.0L);
  j[(size_t)start] = 1e-300L;
  long double sum = 0.0L;  // J_0 + 2 * (even orders)
  for (int k = start; k >= 1; --k) {
    j[(size_t)k - 1] = (2.0L * k / (long double)x) * j[(size_t)k] - j[(size_t)k + 1];
    if (std::fabs(j[(size_t)k - 1]) > 1e250L) {
      for (int q = k - 1; q <= start; ++q) j[(size_t)q] *= 1e-250L;
    }
  }
  sum = j[0];
  for (int k = 2; k <= start; k += 2) sum += 2.0L * j[(size_t)k];
  for (int k = 0; k <= nmax; ++k) out[(size_t)k] = (double)(j[(size_t)k] / sum);
  return out;
}

struct Plan {
  std::vector<std::complex<double>> c;  // y = sum_k c[k] T_k(A / a) x
  double tail = 0.0;                    // bound of what was cut off: sum_{k > K} 2 |J_k(z)|
  double z = 0.0;
};

// Coefficients of exp(i s A), spec(A) in [-a, a], cut where the remaining terms sum to less than eps.
// max_terms < 0: no limit.  Returns an empty plan when max_terms is too small to reach eps.
inline Plan plan(double s, double a, double eps, long long max_terms = -1)
{
  Plan p;
  const double z = s * a, az = std::fabs(z);
  p.z = z;
  eps = std::max(eps, 1e-16);
  // terms beyond |z| + c |z|^(1/3) decay like the Airy function; this many are always enough for 1e-16
  int nmax = (int)std::ceil(az + 12.0 * std::cbrt(az + 1.0) + 40.0);
  std::vector<double> J;
  int K = 0;
  for (;;) {
    J = bessel_j(nmax, az);
    // smallest K with 2 * sum_{k > K} |J_k| <= eps (the terms above nmax are far below the last one kept)
    double tail = 0.0;
    K = nmax;
    while (K > 0 && tail + 2.0 * std::fabs(J[(size_t)K]) <= eps) {
      tail += 2.0 * std::fabs(J[(size_t)K]);
      --K;
    }
    p.tail = tail;
    if (K < nmax - 2) break;  // the cut is inside the computed range
    nmax *= 2;
  }
  if (max_terms >= 0 && (long long)K + 1 > max_terms) return Plan();
  p.c.resize((size_t)K + 1);
  const std::complex<double> iu(0.0, 1.0);
  const std::complex<double> ipow[4] = {{1.0, 0.0}, iu, {-1.0, 0.0}, -iu};
  for (int k = 0; k <= K; ++k) {
    // i^k J_k(z), with J_k(-|z|) = (-1)^k J_k(|z|)
    double jk = J[(size_t)k];
    if (z < 0.0 && (k & 1)) jk = -jk;
    p.c[(size_t)k] = (k == 0 ? 1.0 : 2.0) * jk * ipow[k & 3];
  }
  return p;
}

// The three-term recurrence on three work vectors 0, 1, 2.  Ops provides
//   load(dst)              v[dst] <- x
//   mult(src, dst)         v[dst] <- A v[src]
//   scale(dst, r)          v[dst] <- r v[dst]                     (r real)
//   axpby(dst, a, src, b)  v[dst] <- a v[src] + b v[dst]          (a, b real)
//   y_set(c, src)          y <- c v[src]                          (c complex)
//   y_add(c, src)          y <- y + c v[src]
// Returns the number of MatMults.
template <class Ops>
long long apply(Ops &ops, const Plan &p, double a)
{
  const long long K = (long long)p.c.size() - 1;
  int w0 = 0, w1 = 1;
  const int tmp = 2;
  ops.load(w0);  // T_0 x
  ops.y_set(p.c[0], w0);
  if (K < 1) return 0;
  ops.mult(w0, w1);  // T_1 x = (A / a) x
  ops.scale(w1, 1.0 / a);
  ops.y_add(p.c[1], w1);
  for (long long k = 2; k <= K; ++k) {
    ops.mult(w1, tmp);
    ops.axpby(w0, 2.0 / a, tmp, -1.0);  // T_k x = 2 (A / a) T_{k-1} x - T_{k-2} x, over T_{k-2} x
    ops.y_add(p.c[(size_t)k], w0);
    std::swap(w0, w1);
  }
  return K;
}

}  // namespace cheb
}  // namespace dnm
