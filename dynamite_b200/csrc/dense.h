// Small dense host-side linear algebra for the projected problems of the
// Krylov loops: complex matrix exponential (Pade-13 scaling and squaring,
// Higham 2005) and a cyclic-Jacobi real symmetric eigensolver.  Sizes are
// (ncv+2)^2 at most, so clarity wins over speed.
#pragma once

#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

namespace dnm {
namespace dense {

typedef std::complex<double> cd;

// column-major n x n
struct CMat {
  int n;
  std::vector<cd> a;
  explicit CMat(int n_ = 0) : n(n_), a((size_t)n_ * n_, cd(0, 0)) {}
  cd &operator()(int i, int j) { return a[(size_t)j * n + i]; }
  const cd &operator()(int i, int j) const { return a[(size_t)j * n + i]; }
};

inline CMat eye(int n)
{
  CMat I(n);
  for (int i = 0; i < n; ++i) I(i, i) = 1.0;
  return I;
}

inline CMat mul(const CMat &A, const CMat &B)
{
  const int n = A.n;
  CMat C(n);
  for (int j = 0; j < n; ++j)
    for (int k = 0; k < n; ++k) {
      const cd b = B(k, j);
      if (b == cd(0, 0)) continue;
      for (int i = 0; i < n; ++i) C(i, j) += A(i, k) * b;
    }
  return C;
}

inline CMat axpby(cd alpha, const CMat &A, cd beta, const CMat &B)
{
  CMat C(A.n);
  for (size_t i = 0; i < C.a.size(); ++i) C.a[i] = alpha * A.a[i] + beta * B.a[i];
  return C;
}

inline double norm1(const CMat &A)
{
  double best = 0;
  for (int j = 0; j < A.n; ++j) {
    double s = 0;
    for (int i = 0; i < A.n; ++i) s += std::abs(A(i, j));
    best = std::max(best, s);
  }
  return best;
}

// solve A X = B in place (B overwritten by X); LU with partial pivoting.  Returns false if singular.
inline bool solve(CMat A, CMat &B)
{
  const int n = A.n;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::abs(A(k, k));
    for (int i = k + 1; i < n; ++i)
      if (std::abs(A(i, k)) > best) {
        best = std::abs(A(i, k));
        piv = i;
      }
    if (best == 0.0) return false;
    if (piv != k)
      for (int j = 0; j < n; ++j) {
        std::swap(A(k, j), A(piv, j));
        std::swap(B(k, j), B(piv, j));
      }
    const cd inv = 1.0 / A(k, k);
    for (int i = k + 1; i < n; ++i) {
      const cd f = A(i, k) * inv;
      if (f == cd(0, 0)) continue;
      for (int j = k + 1; j < n; ++j) A(i, j) -= f * A(k, j);
      for (int j = 0; j < n; ++j) B(i, j) -= f * B(k, j);
    }
  }
  for (int j = 0; j < n; ++j)
    for (int i = n - 1; i >= 0; --i) {
      cd s = B(i, j);
      for (int k = i + 1; k < n; ++k) s -= A(i, k) * B(k, j);
      B(i, j) = s / A(i, i);
    }
  return true;
}

// exp(A), Pade [13/13] with scaling and squaring
inline CMat expm(const CMat &A_in)
{
  static const double b[14] = {64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0,
                               129060195264000.0,   10559470521600.0,    670442572800.0,     33522128640.0,
                               1323241920.0,        40840800.0,          960960.0,           16380.0,
                               182.0,               1.0};
  const double theta13 = 5.371920351148152;
  const int n = A_in.n;
  CMat A = A_in;
  int s = 0;
  const double nrm = norm1(A);
  if (nrm > theta13) {
    s = (int)std::ceil(std::log2(nrm / theta13));
    const double f = std::ldexp(1.0, -s);
    for (cd &v : A.a) v *= f;
  }
  const CMat I = eye(n);
  const CMat A2 = mul(A, A), A4 = mul(A2, A2), A6 = mul(A4, A2);
  CMat t1(n), t2(n), U(n), V(n);
  for (size_t i = 0; i < t1.a.size(); ++i) {
    t1.a[i] = b[13] * A6.a[i] + b[11] * A4.a[i] + b[9] * A2.a[i];
    t2.a[i] = b[12] * A6.a[i] + b[10] * A4.a[i] + b[8] * A2.a[i];
  }
  CMat u_in = mul(A6, t1), v_in = mul(A6, t2);
  for (size_t i = 0; i < u_in.a.size(); ++i) {
    u_in.a[i] += b[7] * A6.a[i] + b[5] * A4.a[i] + b[3] * A2.a[i] + b[1] * I.a[i];
    v_in.a[i] += b[6] * A6.a[i] + b[4] * A4.a[i] + b[2] * A2.a[i] + b[0] * I.a[i];
  }
  U = mul(A, u_in);
  V = v_in;
  CMat P = axpby(1.0, V, 1.0, U);   // V + U
  CMat Q = axpby(1.0, V, -1.0, U);  // V - U
  solve(Q, P);                      // P <- Q^{-1} P
  for (int k = 0; k < s; ++k) P = mul(P, P);
  return P;
}

// Real symmetric eigenproblem by cyclic Jacobi: S (n x n, column-major, symmetric)
// -> w (ascending) and eigenvectors in the columns of Z.
inline void sym_eig(int n, std::vector<double> S, std::vector<double> &w, std::vector<double> &Z)
{
  Z.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) Z[(size_t)i * n + i] = 1.0;
  auto at = [&](std::vector<double> &M, int i, int j) -> double & { return M[(size_t)j * n + i]; };
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0, diag = 0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) (i == j ? diag : off) += at(S, i, j) * at(S, i, j);
    if (off <= 1e-32 * std::max(diag, 1e-300)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = at(S, p, q);
        if (apq == 0.0) continue;
        const double app = at(S, p, p), aqq = at(S, q, q);
        const double tau = (aqq - app) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double skp = at(S, k, p), skq = at(S, k, q);
          at(S, k, p) = c * skp - s * skq;
          at(S, k, q) = s * skp + c * skq;
        }
        for (int k = 0; k < n; ++k) {
          const double spk = at(S, p, k), sqk = at(S, q, k);
          at(S, p, k) = c * spk - s * sqk;
          at(S, q, k) = s * spk + c * sqk;
        }
        for (int k = 0; k < n; ++k) {
          const double zkp = at(Z, k, p), zkq = at(Z, k, q);
          at(Z, k, p) = c * zkp - s * zkq;
          at(Z, k, q) = s * zkp + c * zkq;
        }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return at(S, a, a) < at(S, b, b); });
  w.resize(n);
  std::vector<double> Zs((size_t)n * n);
  for (int j = 0; j < n; ++j) {
    w[j] = at(S, order[j], order[j]);
    for (int i = 0; i < n; ++i) Zs[(size_t)j * n + i] = Z[(size_t)order[j] * n + i];
  }
  Z.swap(Zs);
}

}  // namespace dense
}  // namespace dnm
