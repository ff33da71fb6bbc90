// Device-resident Krylov consumers of the shell MatMult:
//
//   dnm_evolve   : y = exp(scale*A) x by the EXPOKIT sub-stepped Arnoldi scheme
//                  (Sidje, ACM TOMS 24 (1998), algorithm 3.2 with the augmented
//                  (m+2)x(m+2) Hessenberg matrix), which is what dynamite gets from
//                  SLEPc's MFN type "expokit" (computations.py:89-112).  SLEPc 3.20.2
//                  itself is not under /root/reference; the step-size controller
//                  below restates the published algorithm with SLEPc's constants
//                  (gamma=0.9, delta=1.2, mxrej=10, t rounded to 2 digits) and the
//                  initial-step formula that IS visible in the reference
//                  (computations.py:511-520).
//   dnm_eigsolve : thick-restart Lanczos == Krylov-Schur for Hermitian problems
//                  (Stewart 2001; Wu & Simon 2000), SLEPc's default EPS for HEP
//                  (computations.py:208-287), restart keeping 50% of the basis.
//
// The state vectors and the whole Krylov basis stay in HBM.  One Arnoldi/Lanczos
// column is: MatMult, one fused multi-dot, one fused multi-axpy that also
// returns the new norm, a conditional re-orthogonalisation (classical
// Gram-Schmidt with the DGKS "if needed" test, eta = 1/sqrt(2), decided ON THE
// DEVICE so there is no host round trip inside the column loop), and a scale.
// Only the small projected matrix crosses to the host, once per sub-step /
// restart, for the dense expm / eigensolve.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>

#include "chebyshev.h"
#include "context.h"
#include "dense.h"
#include "vecops.cuh"

namespace dnm {

namespace {

// ---- scalar bookkeeping kernels (one thread) ----------------------------------------

// after the first Gram-Schmidt pass on column j:
//   Hcol[0..nh) = h ;  refine = ||w||^2 < 0.5 * (||w||^2 + sum |h|^2)
__global__ void k_col_first(const double *__restrict__ d_h, int nh, const double *__restrict__ d_sq,
                            double *__restrict__ Hcol, int *__restrict__ d_refine)
{
  double hh = 0.0;
  for (int i = 0; i < nh; ++i) {
    Hcol[2 * i] = d_h[2 * i];
    Hcol[2 * i + 1] = d_h[2 * i + 1];
    hh += d_h[2 * i] * d_h[2 * i] + d_h[2 * i + 1] * d_h[2 * i + 1];
  }
  const double sq = *d_sq;
  *d_refine = (sq < 0.5 * (sq + hh)) ? 1 : 0;
}

// after the (conditional) second pass: fold h2 into the column, decide linear
// dependence, store the norm as the sub-diagonal entry and as the scale factor
__global__ void k_col_second(const double *__restrict__ d_h2, int nh, const double *__restrict__ d_sq,
                             const double *__restrict__ d_sq2, const int *__restrict__ d_refine,
                             double *__restrict__ Hcol, double *__restrict__ d_norm, double *__restrict__ d_lindep)
{
  double nrm2 = *d_sq;
  double lindep = 0.0;
  if (*d_refine) {
    double hh = 0.0;
    for (int i = 0; i < nh; ++i) {
      Hcol[2 * i] += d_h2[2 * i];
      Hcol[2 * i + 1] += d_h2[2 * i + 1];
      hh += d_h2[2 * i] * d_h2[2 * i] + d_h2[2 * i + 1] * d_h2[2 * i + 1];
    }
    nrm2 = *d_sq2;
    if (nrm2 < 0.5 * (nrm2 + hh)) lindep = 1.0;  // still shrinking: numerically in the span
  }
  double nrm = sqrt(nrm2);
  if (!(nrm > 0.0)) lindep = 1.0;
  Hcol[2 * nh] = nrm;  // H[j+1, j]
  Hcol[2 * nh + 1] = 0.0;
  *d_lindep = lindep;
  *d_norm = (lindep != 0.0) ? 0.0 : nrm;  // scale-by-reciprocal treats 0 as "zero the vector"
}

__global__ void k_sqrt_inplace(double *v) { *v = sqrt(*v); }

// Lanczos column j of a Hermitian operator: h = (<v_{j-1}, w>, <v_j, w>) (only <v_0, w> when j = 0),
// sq = ||w - h.V||^2.  Writes the tridiagonal column, the scale factor and the breakdown flag.
__global__ void k_lanczos_col(const double *__restrict__ d_h, int j, const double *__restrict__ d_sq,
                              double *__restrict__ Hcol, double *__restrict__ d_norm, double *__restrict__ d_lindep)
{
  const int nh = j == 0 ? 1 : 2;
  double hh = 0.0;
  for (int i = 0; i < nh; ++i) {
    const int row = j + 1 - nh + i;
    Hcol[2 * row] = d_h[2 * i];
    Hcol[2 * row + 1] = d_h[2 * i + 1];
    hh += d_h[2 * i] * d_h[2 * i] + d_h[2 * i + 1] * d_h[2 * i + 1];
  }
  const double sq = *d_sq;
  const double nrm = sqrt(sq);
  // invariant subspace: what is left of A v_j is rounding noise of its components along the basis
  const double lindep = (!(nrm > 0.0) || sq <= 1e-28 * (sq + hh)) ? 1.0 : 0.0;
  Hcol[2 * (j + 1)] = nrm;  // H[j+1, j]
  Hcol[2 * (j + 1) + 1] = 0.0;
  *d_lindep = lindep;
  *d_norm = (lindep != 0.0) ? 0.0 : nrm;
}

// V[:, first : first+nout] <- V[:, first : first+nin] * Q   (Q real, nin x nout, column-major)
// in place: every row block is staged in shared memory before anything is written.
constexpr int ROT_ROWS = 64;
__global__ void __launch_bounds__(ROT_ROWS)
    k_rotate_basis(cplx *const *__restrict__ vptr, int first, int nin, int nout, const double *__restrict__ Q, int64_t n)
{
  extern __shared__ double2 stage[];  // [nin][ROT_ROWS]
  for (int64_t row0 = (int64_t)blockIdx.x * ROT_ROWS; row0 < n; row0 += (int64_t)gridDim.x * ROT_ROWS) {
    const int64_t row = row0 + threadIdx.x;
    const bool ok = row < n;
    for (int j = 0; j < nin; ++j) stage[j * ROT_ROWS + threadIdx.x] = ok ? vptr[first + j][row] : make_double2(0, 0);
    // each thread only re-reads what it wrote: no barrier needed
    for (int i = 0; i < nout; ++i) {
      double ar = 0.0, ai = 0.0;
      for (int j = 0; j < nin; ++j) {
        const double q = __ldg(&Q[(size_t)i * nin + j]);
        const double2 v = stage[j * ROT_ROWS + threadIdx.x];
        ar += q * v.x;
        ai += q * v.y;
      }
      if (ok) vptr[first + i][row] = make_double2(ar, ai);
    }
  }
}

int stream_blocks(int64_t n, int tpb) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + tpb - 1) / tpb, (int64_t)G.sm_count * 8)); }

// ---- the Krylov basis ------------------------------------------------------------------

struct Basis {
  std::vector<dnm_vec_t> owned;   // vectors created here
  std::vector<dnm_vec_t> v;       // the basis, v[0] may be borrowed
  int64_t nloc = 0;
  ~Basis()
  {
    for (dnm_vec_t q : owned) pool_release(q);
  }
  cplx *ptr(int j) const { return v[j]->d; }
};

// device scalars used by one column (carved out of G.d_scratch)
struct ColScratch {
  double *h, *h2, *sq, *sq2, *norm;
  int *refine;
};

ColScratch col_scratch(int maxcols)
{
  ColScratch s;
  double *base = G.d_scratch + 4096;  // leave the front for ad-hoc reductions
  s.h = base;
  s.h2 = base + 2 * maxcols;
  s.sq = base + 4 * maxcols;
  s.sq2 = s.sq + 1;
  s.norm = s.sq + 2;
  s.refine = (int *)(s.sq + 4);
  return s;
}

void fused_dots(const Basis &B, int ncols, const cplx *w, double *d_h, const int *active)
{
  for (int c0 = 0; c0 < ncols; c0 += MAX_FUSED) {
    VecList vs;
    vs.n = std::min(MAX_FUSED, ncols - c0);
    for (int k = 0; k < vs.n; ++k) vs.p[k] = B.ptr(c0 + k);
    multi_dot_dev(vs, w, B.nloc, d_h + 2 * c0, active);
  }
}

void fused_axpys(const Basis &B, int ncols, cplx *w, const double *d_h, double *d_sq, const int *active)
{
  for (int c0 = 0; c0 < ncols; c0 += MAX_FUSED) {
    VecList vs;
    vs.n = std::min(MAX_FUSED, ncols - c0);
    for (int k = 0; k < vs.n; ++k) vs.p[k] = B.ptr(c0 + k);
    const bool last = c0 + MAX_FUSED >= ncols;
    multi_axpy_sub_dev(vs, w, B.nloc, d_h + 2 * c0, last ? d_sq : nullptr, active);
  }
}

// One Arnoldi column:  v[j+1] = orth(A v[j]) / norm ;  H[0..j+1, j] written to d_H (column-major, ld)
void arnoldi_column(dnm_mat_t A, Basis &B, int j, double *d_H, int ld, double *d_lindep, const ColScratch &S)
{
  int rc = dnm_mat_mult(A, B.v[j], B.v[j + 1]);
  if (rc) throw Fail{rc};
  cplx *w = B.ptr(j + 1);
  const int nh = j + 1;
  double *Hcol = d_H + 2 * (size_t)ld * j;
  fused_dots(B, nh, w, S.h, nullptr);
  fused_axpys(B, nh, w, S.h, S.sq, nullptr);
  k_col_first<<<1, 1, 0, G.stream>>>(S.h, nh, S.sq, Hcol, S.refine);
  fused_dots(B, nh, w, S.h2, S.refine);
  fused_axpys(B, nh, w, S.h2, S.sq2, S.refine);
  k_col_second<<<1, 1, 0, G.stream>>>(S.h2, nh, S.sq, S.sq2, S.refine, Hcol, S.norm, d_lindep + j);
  count_launch(2);
  DNM_CHECK_CUDA(cudaGetLastError());
  vec_scale_dev(w, B.nloc, S.norm, true);
}

// One Lanczos column (Hermitian A, which dnm_mat_create guarantees): the three-term recurrence with
// the two coefficients taken as computed inner products (local re-orthogonalisation against
// v[j-1] and v[j]) -- 2 basis vectors per dot / axpy instead of j+1, the same H (tridiagonal) layout.
void lanczos_column(dnm_mat_t A, Basis &B, int j, double *d_H, int ld, double *d_lindep, const ColScratch &S)
{
  int rc = dnm_mat_mult(A, B.v[j], B.v[j + 1]);
  if (rc) throw Fail{rc};
  cplx *w = B.ptr(j + 1);
  double *Hcol = d_H + 2 * (size_t)ld * j;
  VecList vs;
  vs.n = 0;
  if (j > 0) vs.p[vs.n++] = B.ptr(j - 1);
  vs.p[vs.n++] = B.ptr(j);
  multi_dot_dev(vs, w, B.nloc, S.h);
  multi_axpy_sub_dev(vs, w, B.nloc, S.h, S.sq);
  k_lanczos_col<<<1, 1, 0, G.stream>>>(S.h, j, S.sq, Hcol, S.norm, d_lindep + j);
  count_launch();
  DNM_CHECK_CUDA(cudaGetLastError());
  vec_scale_dev(w, B.nloc, S.norm, true);
}

int64_t vector_budget(int64_t local_n, int64_t global_n)
{
  size_t f = 0, t = 0;
  DNM_CHECK_CUDA(cudaMemGetInfo(&f, &t));
  const int64_t reserve = (int64_t)512 << 20;
  // workspace already parked in the pool counts as available
  int64_t fit = std::max<int64_t>(0, ((int64_t)f - reserve) / (int64_t)(sizeof(cplx) * local_n)) + pool_count(global_n);
  if (G.nranks > 1) {
    // every rank must derive the same Krylov dimension (the same number of MatMults, reductions and
    // collective vector creations): agree on the smallest budget
    double v = -(double)fit;
    DNM_CHECK_CUDA(cudaMemcpyAsync(G.d_scratch, &v, sizeof(double), cudaMemcpyHostToDevice, G.stream));
    allreduce_max_dev(G.d_scratch, 1);
    fetch_doubles(G.d_scratch, &v, 1);
    fit = (int64_t)(-v);
  }
  return fit;
}

// wall-clock phase accounting, printed when DNM_TRACE is set
struct PhaseTimer {
  const char *names[8] = {"setup", "issue", "wait", "dense", "combine", "teardown", "other", ""};
  double acc[8] = {0};
  std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
  bool on = getenv("DNM_TRACE") != nullptr;
  void mark(int phase)
  {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    acc[phase] += std::chrono::duration<double>(now - last).count();
    last = now;
  }
  void report(const char *what)
  {
    if (!on) return;
    fprintf(stderr, "[dnm trace] %s:", what);
    for (int i = 0; i < 7; ++i) fprintf(stderr, " %s=%.1fms", names[i], acc[i] * 1e3);
    fprintf(stderr, "\n");
  }
};

// csrc/chebyshev.h on device vectors: every operation is one of the library's own vector primitives
// (the ones dnm_vec_copy / dnm_vec_scale / dnm_vec_axpby expose) or the MatMult
struct DeviceChebOps {
  dnm_mat_t A;
  dnm_vec_t x, y;
  dnm_vec_t w[3];
  int64_t nloc;
  void load(int dst) { vec_copy(w[dst]->d, x->d, nloc); }
  void mult(int src, int dst)
  {
    const int rc = dnm_mat_mult(A, w[src], w[dst]);
    if (rc) throw Fail{rc};
  }
  void scale(int dst, double r) { vec_scale(w[dst]->d, nloc, make_double2(r, 0.0)); }
  void axpby(int dst, double a, int src, double b)
  {
    vec_axpby(w[dst]->d, w[src]->d, nloc, make_double2(a, 0.0), make_double2(b, 0.0));
  }
  void y_set(std::complex<double> c, int src)
  {
    vec_copy(y->d, w[src]->d, nloc);
    vec_scale(y->d, nloc, make_double2(c.real(), c.imag()));
  }
  void y_add(std::complex<double> c, int src)
  {
    vec_axpby(y->d, w[src]->d, nloc, make_double2(c.real(), c.imag()), make_double2(1.0, 0.0));
  }
};

double round2(double t)
{
  // round up to two significant digits, as expokit does with its step sizes
  if (!(t > 0.0) || !std::isfinite(t)) return t;
  const double s = std::pow(10.0, std::floor(std::log10(t)) - 1.0);
  return std::ceil(t / s) * s;
}

}  // namespace
}  // namespace dnm

using namespace dnm;
typedef std::complex<double> cd;

// orthogonalisation requested through dnm_evolve_algo for the next dnm_evolve: -1 = default
static int g_evolve_algo = -1;
// what the most recent dnm_evolve ran: 0 expokit/Lanczos, 1 expokit/Arnoldi, 2 Chebyshev (-1: none yet)
static int g_evolve_last = -1;

extern "C" int dnm_evolve_last_algo(void) { return g_evolve_last; }

extern "C" int dnm_evolve(dnm_mat_t A, dnm_vec_t x, dnm_vec_t y, double scale_re, double scale_im, double tol, int ncv,
                          int max_it, int *reason_out, int *its_out, int *matmults_out)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(A && x && y, DNM_ERR_ARG, "null handle");
  DNM_REQUIRE(A->M == A->N, DNM_ERR_ARG, "evolve needs a square matrix");
  DNM_REQUIRE(x->global_n == A->N && y->global_n == A->N, DNM_ERR_ARG, "vector and matrix sizes differ");
  DNM_REQUIRE(x != y && x->d != y->d, DNM_ERR_ARG, "evolve cannot be done in place");
  const int64_t N = A->N, nloc = x->local_n;
  const cd tscale(scale_re, scale_im);
  int reason = 0, its = 0, matmults = 0;

  if (tol <= 0) tol = 1e-7;                        // SLEPc MFN default
  int m = ncv > 0 ? ncv : (int)std::min<int64_t>(30, N);  // SLEPc MFN default
  m = (int)std::min<int64_t>(m, N);
  // basis: y doubles as v[0]; m+1 more vectors (v[1..m] and the A*v[m] probe)
  const int64_t fit = vector_budget(nloc, N);

  // Chebyshev propagator (csrc/chebyshev.h) instead of expokit: on request (algo 2 / DNM_EVOLVE_CHEB=1),
  // and by default when the Krylov basis would have to be cut down to fit device memory -- there
  // expokit degenerates into hundreds of short sub-steps while the expansion needs three work vectors
  // whatever its degree.  Real-time evolution only (scale purely imaginary: the propagator is unitary).
  {
    const bool applicable = scale_re == 0.0 && scale_im != 0.0 && fit >= 3;
    bool use_cheb = false;
    if (g_evolve_algo == 2) {
      DNM_REQUIRE(scale_re == 0.0, DNM_ERR_UNSUPPORTED, "the Chebyshev propagator serves real-time evolution only");
      DNM_REQUIRE(fit >= 3, DNM_ERR_MEM, "not enough device memory for three work vectors");
      use_cheb = scale_im != 0.0;
    } else if (g_evolve_algo < 0 && applicable) {
      const char *e = getenv("DNM_EVOLVE_CHEB");
      use_cheb = e ? atoi(e) != 0 : (m + 1 > fit);
    }
    if (use_cheb) {
      double anorm = 0, sq = 0;
      int rc = dnm_mat_norm_inf(A, &anorm);
      if (rc) return rc;
      vec_sqnorm_dev(x->d, nloc, G.d_scratch);
      fetch_doubles(G.d_scratch, &sq, 1);
      if (sq == 0.0 || anorm == 0.0) {  // zero vector stays zero; zero operator is the identity map
        vec_copy(y->d, x->d, nloc);
        if (reason_out) *reason_out = DNM_CONVERGED_TOL;
        if (its_out) *its_out = 0;
        if (matmults_out) *matmults_out = 0;
        return DNM_OK;
      }
      // rho(A) <= ||A||_inf; the margin covers the rounding of the norm's own summation
      const double a = anorm * (1.0 + 1e-9);
      const double eps = std::min(1e-14, tol * 1e-3);
      const cheb::Plan plan = cheb::plan(scale_im, a, eps, max_it > 0 ? (long long)max_it * (m + 1) : -1);
      if (plan.c.empty()) {
        if (reason_out) *reason_out = DNM_DIVERGED_ITS;
        if (its_out) *its_out = 0;
        if (matmults_out) *matmults_out = 0;
        return DNM_OK;
      }
      if (G.rank == 0 && getenv("DNM_QUIET") == nullptr && m + 1 > fit)
        fprintf(stderr,
                "[dynamite_b200] evolve: a Krylov basis of %d vectors does not fit device memory (room for %lld): "
                "Chebyshev propagator, %zu MatMults\n",
                m + 1, (long long)fit, plan.c.size() - 1);
      double sq_out = 0;
      {
        Basis W;  // three work vectors, back to the pool when the block ends
        W.nloc = nloc;
        for (int j = 0; j < 3; ++j) {
          dnm_vec_t q = pool_acquire(N);
          W.owned.push_back(q);
        }
        DeviceChebOps ops{A, x, y, {W.owned[0], W.owned[1], W.owned[2]}, nloc};
        matmults = (int)cheb::apply(ops, plan, a);
        vec_sqnorm_dev(y->d, nloc, G.d_scratch);
        fetch_doubles(G.d_scratch, &sq_out, 1);
      }
      // exp(i s A) is unitary: a norm that moved means the spectrum left [-a, a] or the recurrence broke;
      // then the sub-stepped scheme below recomputes y from x
      const double drift = std::fabs(std::sqrt(sq_out / sq) - 1.0);
      if (std::isfinite(drift) && drift <= 1e-9) {
        g_evolve_last = 2;
        if (reason_out) *reason_out = DNM_CONVERGED_TOL;
        if (its_out) *its_out = 1;
        if (matmults_out) *matmults_out = matmults;
        return DNM_OK;
      }
      if (G.rank == 0)
        fprintf(stderr, "[dynamite_b200] evolve: Chebyshev propagator changed the norm by %.3e, falling back to expokit\n", drift);
      DNM_REQUIRE(g_evolve_algo != 2, DNM_ERR_INTERNAL, "Chebyshev propagator lost unitarity (norm drift %.3e)", drift);
      matmults = 0;
    }
  }

  if (m + 1 > fit) {
    DNM_REQUIRE(fit >= 3, DNM_ERR_MEM, "not enough device memory for a Krylov basis (room for %lld vectors)",
                (long long)fit);
    if (G.rank == 0 && getenv("DNM_QUIET") == nullptr)
      fprintf(stderr, "[dynamite_b200] evolve: Krylov dimension reduced from %d to %d to fit device memory\n", m, (int)fit - 1);
    m = (int)fit - 1;
  }
  DNM_REQUIRE(m >= 1, DNM_ERR_ARG, "Krylov dimension must be at least 1");
  if (max_it <= 0) max_it = (int)std::min<int64_t>(2147483647, std::max<int64_t>(100, 2 * N / m));  // SLEPc default

  const double t_out = std::abs(tscale);
  if (t_out == 0.0) {
    vec_copy(y->d, x->d, nloc);
    if (reason_out) *reason_out = DNM_CONVERGED_TOL;
    if (its_out) *its_out = 0;
    if (matmults_out) *matmults_out = 0;
    return DNM_OK;
  }
  const cd sgn = tscale / t_out;

  PhaseTimer trace;
  Basis B;
  B.nloc = nloc;
  B.v.push_back(y);
  for (int j = 0; j < m + 1; ++j) {
    dnm_vec_t q = pool_acquire(N);
    B.owned.push_back(q);
    B.v.push_back(q);
  }

  const int ld = m + 2;
  double *d_H = nullptr, *d_lindep = nullptr;
  DNM_CHECK_CUDA(cudaMalloc(&d_H, sizeof(double) * (2 * (size_t)ld * ld + ld + 2 * ld)));
  std::unique_ptr<double, void (*)(double *)> d_H_guard(d_H, [](double *p) { cudaFree(p); });
  d_lindep = d_H + 2 * (size_t)ld * ld;
  double *d_coef = d_lindep + ld;  // 2*ld doubles: beta*F for the final combination
  const size_t hbytes = sizeof(double) * (2 * (size_t)ld * ld + ld);
  std::vector<double> h_H(2 * (size_t)ld * ld + ld);
  const ColScratch S = col_scratch(ld);
  DNM_REQUIRE(4096 + hbytes / sizeof(double) <= (size_t)SCRATCH_DOUBLES, DNM_ERR_ARG, "ncv=%d is too large", m);

  double anorm = 0;
  {
    int rc = dnm_mat_norm_inf(A, &anorm);
    if (rc) return rc;
  }
  const double rndoff = anorm * 2.220446049250313e-16;
  (void)rndoff;

  // beta = ||x||, w = x (held in y == v[0])
  vec_copy(y->d, x->d, nloc);
  vec_sqnorm_dev(y->d, nloc, G.d_scratch);
  double beta = 0;
  fetch_doubles(G.d_scratch, &beta, 1);
  beta = std::sqrt(beta);
  if (beta == 0.0 || anorm == 0.0) {
    // zero vector stays zero; zero operator is the identity map
    if (reason_out) *reason_out = DNM_CONVERGED_TOL;
    if (its_out) *its_out = 0;
    if (matmults_out) *matmults_out = 0;
    return DNM_OK;
  }

  const double gamma = 0.9, delta = 1.2;
  const int mxrej = 10;
  double xm = 1.0 / m;
  const double fact = std::pow((m + 1) / 2.72, m + 1) * std::sqrt(2.0 * M_PI * (m + 1));
  double t_new = round2((1.0 / anorm) * std::pow((fact * tol) / (4.0 * beta * anorm), xm));
  double t_now = 0.0;

  // Hermitian operator: Lanczos recurrence (DNM_EVOLVE_ORTH=full keeps the Arnoldi column with full
  // classical Gram-Schmidt + DGKS refinement)
  const char *orth_env = getenv("DNM_EVOLVE_ORTH");
  const bool lanczos = g_evolve_algo >= 0 ? g_evolve_algo == 0 : !(orth_env && !strcmp(orth_env, "full"));
  g_evolve_last = lanczos ? 0 : 1;
  // CUDA-graph capture of the basis construction: opt-in (DNM_EVOLVE_GRAPH=1).  Measured on C1 (L=20,
  // scripts/explore_c1.py): once pool_release stopped calling cudaMemGetInfo per vector the plain launch
  // sequence takes 4.7 ms per evolve, while capturing + instantiating a graph per call costs 5-70 ms.
  bool use_graph = G.nranks == 1 && nloc <= ((int64_t)1 << 24) && getenv("DNM_EVOLVE_GRAPH") != nullptr &&
                   getenv("DNM_NO_GRAPH") == nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_failed = false;
  struct GraphGuard {
    cudaGraphExec_t &g;
    ~GraphGuard()
    {
      if (g) cudaGraphExecDestroy(g);
    }
  } graph_guard{graph_exec};
  if (use_graph) {
    // the plan (and any generated kernels) must exist before a capture starts
    int rc = dnm_mat_mult(A, B.v[0], B.v[1]);
    if (rc) return rc;
  }

  trace.mark(0);
  while (reason == 0) {
    ++its;
    if (!std::isfinite(t_new)) t_new = 1e300;
    double t_step = std::min(t_out - t_now, t_new);

    // v[0] = w / beta  (w lives in v[0])
    vec_scale(B.ptr(0), nloc, make_double2(1.0 / beta, 0.0));
    auto build_basis = [&]() {
      DNM_CHECK_CUDA(cudaMemsetAsync(d_H, 0, hbytes, G.stream));
      for (int j = 0; j < m; ++j) {
        if (lanczos) lanczos_column(A, B, j, d_H, ld, d_lindep, S);
        else arnoldi_column(A, B, j, d_H, ld, d_lindep, S);
      }
      // probe vector A*v[m] for the error estimate (harmless if the basis broke down earlier)
      int rc = dnm_mat_mult(A, B.v[m], B.v[m + 1]);
      if (rc) throw Fail{rc};
      vec_sqnorm_dev(B.ptr(m + 1), nloc, S.sq);
      // pinned staging (pageable targets cannot be captured); the front of h_scratch belongs to fetch_doubles
      DNM_CHECK_CUDA(cudaMemcpyAsync(G.h_scratch + 4096, d_H, hbytes, cudaMemcpyDeviceToHost, G.stream));
    };
    // (opt-in) the whole basis construction of a sub-step is a fixed launch sequence: captured once per
    // call as a CUDA graph
    if (use_graph && !graph_exec && !graph_failed) {
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(G.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        bool ok = true;
        try {
          build_basis();
        } catch (...) {
          ok = false;
        }
        if (cudaStreamEndCapture(G.stream, &graph) != cudaSuccess || !ok || !graph) ok = false;
        if (ok && cudaGraphInstantiate(&graph_exec, graph, 0) != cudaSuccess) ok = false;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {
          const cudaError_t why = cudaGetLastError();
          if (trace.on) fprintf(stderr, "[dnm trace] evolve: graph capture failed (%s), launching directly\n", cudaGetErrorString(why));
          graph_exec = nullptr;
          graph_failed = true;
        }
      } else {
        cudaGetLastError();
        graph_failed = true;
      }
    }
    if (graph_exec) {
      DNM_CHECK_CUDA(cudaGraphLaunch(graph_exec, G.stream));
    } else {
      build_basis();
    }
    matmults += m + 1;
    double avnorm2 = 0;
    trace.mark(1);
    fetch_doubles(S.sq, &avnorm2, 1);  // also synchronises the copy above
    std::copy(G.h_scratch + 4096, G.h_scratch + 4096 + hbytes / sizeof(double), h_H.begin());
    trace.mark(2);
    const double avnorm = std::sqrt(avnorm2);

    // happy breakdown: the Krylov space became invariant after mb vectors
    int mb = m, k1 = 2;
    for (int j = 0; j < m; ++j)
      if (h_H[2 * (size_t)ld * ld + j] != 0.0) {
        mb = j + 1;
        k1 = 0;
        t_step = t_out - t_now;
        break;
      }

    auto Hentry = [&](int i, int j) { return cd(h_H[2 * ((size_t)ld * j + i)], h_H[2 * ((size_t)ld * j + i) + 1]); };
    std::vector<cd> F;
    double err_loc = 0;
    int ireject = 0;
    const int mx = mb + k1;
    for (;;) {
      dense::CMat Hs(mx);
      for (int j = 0; j < mb; ++j)
        for (int i = 0; i < std::min(mb, j + 2); ++i) Hs(i, j) = sgn * t_step * Hentry(i, j);
      if (k1) {
        Hs(m, m - 1) = sgn * t_step * Hentry(m, m - 1);
        Hs(m + 1, m) = sgn * t_step;  // the unit entry of the augmented matrix
      }
      const dense::CMat E = dense::expm(Hs);
      F.assign(mx, cd(0, 0));
      for (int i = 0; i < mx; ++i) F[i] = E(i, 0);
      if (k1 == 0) {
        err_loc = tol;
        break;
      }
      const double p1 = std::abs(beta * F[m]);
      const double p2 = std::abs(beta * F[m + 1] * avnorm);
      if (p1 > 10.0 * p2) {
        err_loc = p2;
        xm = 1.0 / m;
      } else if (p1 > p2) {
        err_loc = (p1 * p2) / (p1 - p2);
        xm = 1.0 / m;
      } else {
        err_loc = p1;
        xm = 1.0 / (m > 1 ? m - 1 : 1);
      }
      if (err_loc <= delta * t_step * tol || ireject >= mxrej) break;
      t_step = round2(gamma * t_step * std::pow(t_step * tol / err_loc, xm));
      ++ireject;
    }

    trace.mark(3);
    // w = V[:, 0:mx'] * (beta F), in place on v[0]
    const int ncomb = mb + std::max(0, k1 - 1);
    std::vector<double> coef(2 * ncomb);
    for (int j = 0; j < ncomb; ++j) {
      const cd c = beta * F[j];
      coef[2 * j] = c.real();
      coef[2 * j + 1] = c.imag();
    }
    DNM_CHECK_CUDA(cudaMemcpyAsync(d_coef, coef.data(), sizeof(double) * 2 * ncomb, cudaMemcpyHostToDevice, G.stream));
    for (int c0 = 0; c0 < ncomb; c0 += MAX_FUSED) {
      VecList vs;
      vs.n = std::min(MAX_FUSED, ncomb - c0);
      for (int k = 0; k < vs.n; ++k) vs.p[k] = B.ptr(c0 + k);
      if (c0 == 0) multi_combine_dev(vs, B.ptr(0), nloc, d_coef);
      else multi_combine_acc_dev(vs, B.ptr(0), nloc, d_coef + 2 * c0);
    }
    vec_sqnorm_dev(B.ptr(0), nloc, G.d_scratch);
    fetch_doubles(G.d_scratch, &beta, 1);  // also makes `coef` safe to free
    beta = std::sqrt(beta);
    trace.mark(4);

    t_now += t_step;
    if (t_now >= t_out * (1.0 - 1e-15)) {
      reason = DNM_CONVERGED_TOL;
    } else {
      t_new = round2(gamma * t_step * std::pow((t_step * tol) / std::max(err_loc, 1e-300), xm));
      if (its >= max_it) reason = DNM_DIVERGED_ITS;
      if (beta == 0.0) reason = DNM_CONVERGED_TOL;  // decayed to exactly zero
    }
  }
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  trace.mark(6);
  trace.report("evolve");
  if (reason_out) *reason_out = reason;
  if (its_out) *its_out = its;
  if (matmults_out) *matmults_out = matmults;
  DNM_API_END
}

extern "C" int dnm_evolve_algo(dnm_mat_t A, dnm_vec_t x, dnm_vec_t y, double scale_re, double scale_im, double tol, int ncv,
                               int max_it, int algo, int *reason_out, int *its_out, int *matmults_out)
{
  if (algo < -1 || algo > 2) {
    set_error("algo must be -1 (default), 0 (expokit, Lanczos recurrence), 1 (expokit, full Arnoldi) or 2 (Chebyshev)");
    return DNM_ERR_ARG;
  }
  g_evolve_algo = algo;
  const int rc = dnm_evolve(A, x, y, scale_re, scale_im, tol, ncv, max_it, reason_out, its_out, matmults_out);
  g_evolve_algo = -1;
  return rc;
}

extern "C" int dnm_eigsolve(dnm_mat_t A, int nev, int which, double tol, int max_it, int ncv, uint64_t seed,
                            int max_pairs, int *nconv_out, double *evals, double *errest, dnm_vec_t *evecs,
                            int *reason_out, int *its_out, int *matmults_out)
{
  DNM_API_BEGIN
  require_init();
  DNM_REQUIRE(A && nconv_out && evals, DNM_ERR_ARG, "null pointer");
  DNM_REQUIRE(A->M == A->N, DNM_ERR_ARG, "eigsolve needs a square matrix");
  DNM_REQUIRE(nev >= 1, DNM_ERR_ARG, "nev must be at least 1");
  DNM_REQUIRE(which >= 0 && which <= 2, DNM_ERR_ARG, "which must be 0 (lowest), 1 (highest) or 2 (exterior)");
  const int64_t N = A->N, nloc = A->local_N;
  DNM_REQUIRE(nev <= N, DNM_ERR_ARG, "nev=%d exceeds the dimension %lld", nev, (long long)N);
  if (tol <= 0) tol = 1e-8;  // SLEPc EPS default
  // SLEPc defaults: ncv = max(2*nev, nev+15), capped by the dimension
  if (ncv <= 0) ncv = std::max(2 * nev, nev + 15);
  ncv = (int)std::min<int64_t>(ncv, N);
  const int64_t fit = vector_budget(nloc, N);
  if (ncv + 1 > fit) {
    DNM_REQUIRE(fit >= nev + 3, DNM_ERR_MEM, "not enough device memory for the Lanczos basis (room for %lld vectors)",
                (long long)fit);
    if (G.rank == 0 && getenv("DNM_QUIET") == nullptr)
      fprintf(stderr, "[dynamite_b200] eigsolve: ncv reduced from %d to %d to fit device memory\n", ncv, (int)fit - 1);
    ncv = (int)fit - 1;
  }
  DNM_REQUIRE(ncv >= nev && ncv <= 512, DNM_ERR_ARG, "bad ncv=%d for nev=%d", ncv, nev);
  if (max_it <= 0) max_it = (int)std::min<int64_t>(2147483647, std::max<int64_t>(100, 2 * N / ncv));

  Basis B;
  B.nloc = nloc;
  for (int j = 0; j < ncv + 1; ++j) {
    dnm_vec_t q = pool_acquire(N);
    B.owned.push_back(q);
    B.v.push_back(q);
  }
  const int ld = ncv + 1;
  double *d_H = nullptr;
  DNM_CHECK_CUDA(cudaMalloc(&d_H, sizeof(double) * (2 * (size_t)ld * ld + ld + (size_t)ld * ld) + sizeof(cplx *) * ld));
  std::unique_ptr<double, void (*)(double *)> guard(d_H, [](double *p) { cudaFree(p); });
  double *d_lindep = d_H + 2 * (size_t)ld * ld;
  double *d_Q = d_lindep + ld;
  cplx **d_vptr = (cplx **)(d_Q + (size_t)ld * ld);
  {
    std::vector<cplx *> ptrs(ld);
    for (int j = 0; j < ld; ++j) ptrs[j] = B.ptr(j);
    DNM_CHECK_CUDA(cudaMemcpyAsync(d_vptr, ptrs.data(), sizeof(cplx *) * ld, cudaMemcpyHostToDevice, G.stream));
    DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  }
  const size_t hbytes = sizeof(double) * (2 * (size_t)ld * ld + ld);
  std::vector<double> h_H(2 * (size_t)ld * ld + ld);
  const ColScratch S = col_scratch(ld);
  DNM_CHECK_CUDA(cudaFuncSetAttribute(k_rotate_basis, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

  auto random_unit = [&](int col, int northo) {
    // random vector, orthogonalised (twice) against v[0..northo), normalised
    vec_random_fill(B.ptr(col), nloc, (int64_t)G.rank * nloc, seed + 7919ull * (uint64_t)col);
    for (int rep = 0; rep < 2 && northo > 0; ++rep) {
      fused_dots(B, northo, B.ptr(col), S.h, nullptr);
      fused_axpys(B, northo, B.ptr(col), S.h, nullptr, nullptr);
    }
    vec_sqnorm_dev(B.ptr(col), nloc, S.norm);
    k_sqrt_inplace<<<1, 1, 0, G.stream>>>(S.norm);
    count_launch();
    vec_scale_dev(B.ptr(col), nloc, S.norm, true);
  };

  // ordering of Ritz values for `which`
  auto better = [&](double a, double b) {
    if (which == 0) return a < b;
    if (which == 1) return a > b;
    return std::fabs(a) > std::fabs(b);
  };

  std::vector<double> theta(ncv, 0.0);  // diagonal of the kept part (locked + restarted Ritz values)
  std::vector<double> spike(ncv, 0.0);  // coupling of kept Ritz vectors to the first new Lanczos vector
  std::vector<double> resid(ncv, 0.0);
  int nconv = 0, l = 0, its = 0, matmults = 0, reason = 0;
  PhaseTimer trace;
  random_unit(0, 0);
  trace.mark(0);

  while (reason == 0) {
    ++its;
    int nv = ncv;
    const int k0 = nconv + l;
    DNM_CHECK_CUDA(cudaMemsetAsync(d_H, 0, hbytes, G.stream));
    for (int j = k0; j < nv; ++j) {
      arnoldi_column(A, B, j, d_H, ld, d_lindep, S);
      ++matmults;
    }
    DNM_CHECK_CUDA(cudaMemcpyAsync(h_H.data(), d_H, hbytes, cudaMemcpyDeviceToHost, G.stream));
    trace.mark(1);
    DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
    trace.mark(2);
    auto Hre = [&](int i, int j) { return h_H[2 * ((size_t)ld * j + i)]; };
    bool breakdown = false;
    for (int j = k0; j < nv; ++j)
      if (h_H[2 * (size_t)ld * ld + j] != 0.0) {
        nv = j + 1;  // v[j+1] is not usable: the basis is invariant
        breakdown = true;
        break;
      }
    const double beta = breakdown ? 0.0 : Hre(nv, nv - 1);

    // projected matrix on the active block [nconv, nv): arrowhead + tridiagonal, real symmetric
    const int na = nv - nconv;
    std::vector<double> T((size_t)na * na, 0.0);
    auto Tat = [&](int i, int j) -> double & { return T[(size_t)j * na + i]; };
    for (int i = nconv; i < k0; ++i) {
      Tat(i - nconv, i - nconv) = theta[i];
      Tat(i - nconv, k0 - nconv) = Tat(k0 - nconv, i - nconv) = spike[i];
    }
    for (int j = k0; j < nv; ++j) {
      Tat(j - nconv, j - nconv) = Hre(j, j);
      if (j + 1 < nv) Tat(j + 1 - nconv, j - nconv) = Tat(j - nconv, j + 1 - nconv) = Hre(j + 1, j);
    }
    std::vector<double> w, Z;
    dense::sym_eig(na, T, w, Z);
    std::vector<int> order(na);
    for (int i = 0; i < na; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return better(w[a], w[b]); });

    // converged = leading Ritz pairs with |beta * last component| <= tol * |theta|
    int k = 0;
    std::vector<double> res(na);
    for (int i = 0; i < na; ++i) res[i] = std::fabs(beta * Z[(size_t)order[i] * na + (na - 1)]);
    while (k < na) {
      const double th = std::fabs(w[order[k]]);
      const double rel = (th > res[k]) ? res[k] / th : res[k];
      if (rel > tol) break;
      ++k;
    }
    const int nconv_new = nconv + k;
    bool done = false;
    if (nconv_new >= nev) {
      reason = DNM_CONVERGED_TOL;
      done = true;
    } else if (its >= max_it) {
      reason = DNM_DIVERGED_ITS;
      done = true;
    }
    int lnew = 0;
    if (!done) {
      if (breakdown) {
        lnew = 0;  // restart from a fresh random vector below
      } else {
        lnew = std::max(1, (int)((nv - nconv_new) * 0.5));
        lnew = std::min(lnew, nv - nconv_new - 1 > 0 ? nv - nconv_new - 1 : 1);
      }
    }
    const int nkeep = done ? k : k + lnew;

    trace.mark(3);
    // rotate the basis: V[:, nconv : nconv+nkeep] = V[:, nconv:nv] * Z[:, order[0:nkeep]]
    if (nkeep > 0) {
      std::vector<double> Q((size_t)na * nkeep);
      for (int i = 0; i < nkeep; ++i)
        for (int j = 0; j < na; ++j) Q[(size_t)i * na + j] = Z[(size_t)order[i] * na + j];
      DNM_CHECK_CUDA(cudaMemcpyAsync(d_Q, Q.data(), sizeof(double) * Q.size(), cudaMemcpyHostToDevice, G.stream));
      const size_t smem = sizeof(double2) * (size_t)na * ROT_ROWS;
      DNM_REQUIRE(smem <= 200 * 1024, DNM_ERR_UNSUPPORTED, "ncv=%d too large for the basis rotation kernel", ncv);
      k_rotate_basis<<<stream_blocks(nloc, ROT_ROWS), ROT_ROWS, smem, G.stream>>>(d_vptr, nconv, na, nkeep, d_Q, nloc);
      count_launch();
      DNM_CHECK_CUDA(cudaGetLastError());
      DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));  // Q is a host temporary
    }
    trace.mark(4);
    for (int i = 0; i < nkeep; ++i) {
      theta[nconv + i] = w[order[i]];
      spike[nconv + i] = (i < k) ? 0.0 : beta * Z[(size_t)order[i] * na + (na - 1)];
      resid[nconv + i] = res[i];
    }
    if (done) {
      nconv = nconv_new;
      break;
    }
    // next Lanczos vector goes right after the kept block
    if (breakdown) {
      random_unit(nconv_new, nconv_new);
    } else {
      vec_copy(B.ptr(nconv_new + lnew), B.ptr(nv), nloc);
    }
    nconv = nconv_new;
    l = lnew;
  }

  // results: locked pairs in the order they were accepted, re-sorted by `which`
  std::vector<int> perm(nconv);
  for (int i = 0; i < nconv; ++i) perm[i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return better(theta[a], theta[b]); });
  const int nret = std::min(nconv, max_pairs);
  for (int i = 0; i < nret; ++i) {
    evals[i] = theta[perm[i]];
    if (errest) errest[i] = resid[perm[i]];
    if (evecs) {
      DNM_REQUIRE(evecs[i] && evecs[i]->global_n == N, DNM_ERR_ARG, "eigenvector slot %d has the wrong size", i);
      vec_copy(evecs[i]->d, B.ptr(perm[i]), nloc);
    }
  }
  DNM_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  trace.mark(6);
  trace.report("eigsolve (combine = basis rotation)");
  *nconv_out = nret;
  if (reason_out) *reason_out = reason;
  if (its_out) *its_out = its;
  if (matmults_out) *matmults_out = matmults;
  DNM_API_END
}
