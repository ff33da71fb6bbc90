// CPU instantiation of csrc/chebyshev.h for tests/test_chebyshev_host.py: the recurrence and the
// coefficient code that krylov.cu runs on device vectors, here on std::vector with a dense matrix.
// Test infrastructure only (compiled by the test with g++; never part of the library).
#include <complex>
#include <limits>
#include <vector>

#include "chebyshev.h"

typedef std::complex<double> cd;

namespace {
struct HostOps {
  int n;
  const cd *A;  // n x n, row-major
  const cd *x;
  cd *y;
  std::vector<cd> v[3];
  long long mults = 0, passes = 0;
  HostOps(int n_, const cd *A_, const cd *x_, cd *y_) : n(n_), A(A_), x(x_), y(y_)
  {
    // work vectors and y start as NaN: nothing may be read before it is written
    const double q = std::numeric_limits<double>::quiet_NaN();
    for (auto &w : v) w.assign((size_t)n, cd(q, q));
    for (int i = 0; i < n; ++i) y[i] = cd(q, q);
  }
  void load(int dst) { v[dst].assign(x, x + n); }
  void mult(int src, int dst)
  {
    for (int i = 0; i < n; ++i) {
      cd acc = 0;
      for (int j = 0; j < n; ++j) acc += A[(size_t)i * n + j] * v[src][(size_t)j];
      v[dst][(size_t)i] = acc;
    }
    ++mults;
  }
  void scale(int dst, double r)
  {
    for (auto &e : v[dst]) e *= r;
  }
  void axpby(int dst, double a, int src, double b)
  {
    for (int i = 0; i < n; ++i) v[dst][(size_t)i] = a * v[src][(size_t)i] + b * v[dst][(size_t)i];
  }
  void y_set(cd c, int src)
  {
    for (int i = 0; i < n; ++i) y[i] = c * v[src][(size_t)i];
  }
  void y_add(cd c, int src)
  {
    for (int i = 0; i < n; ++i) y[i] += c * v[src][(size_t)i];
  }
};
}  // namespace

extern "C" void cheb_bessel(int nmax, double x, double *out)
{
  const std::vector<double> j = dnm::cheb::bessel_j(nmax, x);
  for (int k = 0; k <= nmax; ++k) out[k] = j[(size_t)k];
}

// returns the number of coefficients (0: max_terms too small), writes min(count, cap) of them
extern "C" long long cheb_plan(double s, double a, double eps, long long max_terms, double *c_out, long long cap, double *tail)
{
  const dnm::cheb::Plan p = dnm::cheb::plan(s, a, eps, max_terms);
  for (long long k = 0; k < (long long)p.c.size() && k < cap; ++k) {
    c_out[2 * k] = p.c[(size_t)k].real();
    c_out[2 * k + 1] = p.c[(size_t)k].imag();
  }
  if (tail) *tail = p.tail;
  return (long long)p.c.size();
}

// y = exp(i s A) x ; returns the number of MatMults, -1 when max_terms is too small
extern "C" long long cheb_apply_dense(int n, const double *A, const double *x, double *y, double s, double a, double eps,
                                      long long max_terms)
{
  const dnm::cheb::Plan p = dnm::cheb::plan(s, a, eps, max_terms);
  if (p.c.empty()) return -1;
  HostOps ops(n, (const cd *)A, (const cd *)x, (cd *)y);
  const long long k = dnm::cheb::apply(ops, p, a);
  return k == ops.mults ? k : -2;
}
