// Device vector primitives used by the Vec surface and the Krylov loops.
// All launches go to G.stream; reductions are deterministic two-stage sums
// (per-block partials, then one block), and in multi-rank mode are completed
// by an NCCL all-reduce on the same stream.
#pragma once

#include "context.h"

namespace dnm {

constexpr int MAX_FUSED = 16;  // vectors handled per fused multi-dot / multi-axpy launch

struct VecList {
  const cplx *p[MAX_FUSED];
  int n;
};

int reduce_grid(int64_t n);

void vec_fill(cplx *v, int64_t n, cplx value);
// uniform [-1,1]^2 per entry from a counter-based hash of (global index, seed)
void vec_random_fill(cplx *v, int64_t n, int64_t global_offset, uint64_t seed);
void vec_copy(cplx *dst, const cplx *src, int64_t n);
// v *= a  (a on host)
void vec_scale(cplx *v, int64_t n, cplx a);
// y = a*x + b*y  (a, b on host)
void vec_axpby(cplx *y, const cplx *x, int64_t n, cplx a, cplx b);

// d_out[0..1] = sum_i x_i * conj(y_i)    (global over ranks)
void vec_dot_dev(const cplx *x, const cplx *y, int64_t n, double *d_out);
// d_out[0] = sum_i |x_i|^2   (global over ranks; NOT square-rooted)
void vec_sqnorm_dev(const cplx *x, int64_t n, double *d_out);
// d_out[0] = sum |x_i| (type 1) or max |x_i| (type 2), global
void vec_norm_other_dev(const cplx *x, int64_t n, int type, double *d_out);

// h[j] = sum_i conj(V_j[i]) * w[i], j < vs.n           (global over ranks)
// d_h: vs.n complex (interleaved) on the device
// d_active (optional, device int): when it points to 0 the launch is a no-op
void multi_dot_dev(const VecList &vs, const cplx *w, int64_t n, double *d_h, const int *d_active = nullptr);
// w -= sum_j h[j] * V_j ;  d_sq[0] = ||w_new||^2 if d_sq != nullptr (global)
void multi_axpy_sub_dev(const VecList &vs, cplx *w, int64_t n, const double *d_h, double *d_sq,
                        const int *d_active = nullptr);
// out = sum_j c[j] * V_j   (c on device, complex); out may alias V_0
void multi_combine_dev(const VecList &vs, cplx *out, int64_t n, const double *d_c);
// out += sum_j c[j] * V_j
void multi_combine_acc_dev(const VecList &vs, cplx *out, int64_t n, const double *d_c);
// v *= (*d_s)  or  v *= 1/(*d_s)  with a real device scalar
void vec_scale_dev(cplx *v, int64_t n, const double *d_s, bool reciprocal);
// dst = src * (1 / *d_s)
void vec_scaled_copy_dev(cplx *dst, const cplx *src, int64_t n, const double *d_s, bool reciprocal);

// sum `count` doubles across ranks in place on the device (no-op for one rank)
void allreduce_sum_dev(double *d_buf, int count);
void allreduce_max_dev(double *d_buf, int count);
// copy `count` doubles from device scratch to host and wait for them
void fetch_doubles(const double *d_src, double *h_dst, int count);

}  // namespace dnm
