/*
 * dnm_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see dnm_oracle.h).
 *
 * Every function cites the reference lines it restates; paths are relative to
 * /root/reference/src/dynamite/_backend/.  Integers are int64, scalars are
 * complex128 stored as interleaved (re, im) doubles.
 */
#include "dnm_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

static inline int par64(int64_t v) { return __builtin_parityll((unsigned long long)v); }
static inline int pop64(int64_t v) { return __builtin_popcountll((unsigned long long)v); }
static inline int ctz64(int64_t v) { return __builtin_ctzll((unsigned long long)v); }

/* ------------------------------------------------------------------ */
/* index maps: bsubspace_impl.h:57-361                                 */
/* ------------------------------------------------------------------ */

int64_t orc_dim(const orc_subspace *s)
{
  switch (s->type) {
  case ORC_FULL: /* :57-59 */
    return (int64_t)1 << s->L;
  case ORC_PARITY: /* :112-114 */
    return (int64_t)1 << (s->L - 1);
  case ORC_SPIN_CONSERVE: /* :187-189 */
    return s->nchoosek[s->k * s->ld_nchoosek + s->L];
  case ORC_EXPLICIT: /* :302-304 */
    return s->dim;
  }
  return -1;
}

/* combinatorial number system rank, :191-202 (the k<=n guard is at :197) */
static int64_t rank_spinconserve(const orc_subspace *s, int64_t state)
{
  int64_t idx = 0, ones = 0;
  while (state) {
    int64_t n = ctz64(state);
    ++ones;
    if (ones <= n) idx += s->nchoosek[ones * s->ld_nchoosek + n];
    state &= state - 1;
  }
  return idx;
}

int64_t orc_s2i(const orc_subspace *s, int64_t state)
{
  switch (s->type) {
  case ORC_FULL: /* :61-63 */
    return state;
  case ORC_PARITY: /* :116-123 */
    return (par64(state) == s->space) ? (state >> 1) : -1;
  case ORC_SPIN_CONSERVE: /* :204-208 */
    if (pop64(state) != s->k) return -1;
    return rank_spinconserve(s, state);
  case ORC_EXPLICIT: { /* :306-331 binary search over the sorted states */
    int64_t lo = 0, hi = s->dim - 1;
    while (lo <= hi) {
      int64_t mid = (lo + hi) / 2;
      int64_t v = s->rmap_states[mid];
      if (v == state) return s->rmap_indices ? s->rmap_indices[mid] : mid;
      if (v < state) lo = mid + 1;
      else hi = mid - 1;
    }
    return -1;
  }
  }
  return -1;
}

int64_t orc_i2s(const orc_subspace *s, int64_t idx)
{
  switch (s->type) {
  case ORC_FULL: /* :69-74 */
    return idx;
  case ORC_PARITY: /* :129-134 */
    return (idx << 1) | (par64(idx) ^ s->space);
  case ORC_SPIN_CONSERVE: { /* :210-228 greedy unrank from the top bit down */
    int64_t state = 0, k = s->k;
    for (int64_t n = s->L; n > 0; --n) {
      state <<= 1;
      int64_t c = (k > n - 1) ? 0 : s->nchoosek[k * s->ld_nchoosek + n - 1];
      if (idx >= c) {
        idx -= c;
        --k;
        state |= 1;
      }
    }
    return state;
  }
  case ORC_EXPLICIT: /* :333-338 */
    return s->state_map[idx];
  }
  return -1;
}

/* NextState: :76-83, :136-143, :230-245, :340-347.  The reference builds the
 * SpinConserve low-bit fill with a 32-bit literal (:242, undefined for to>31);
 * the oracle uses 64-bit arithmetic, which is the evident intent. */
int64_t orc_next_state(const orc_subspace *s, int64_t prev, int64_t idx)
{
  if (s->type != ORC_SPIN_CONSERVE) return orc_i2s(s, idx);
  int tz = ctz64(prev);
  prev >>= tz;
  ++prev;
  int to = ctz64(prev);
  prev >>= to;
  prev <<= to + tz;
  prev |= ((int64_t)1 << (to - 1)) - 1;
  return prev;
}

void orc_s2i_array(const orc_subspace *s, int64_t n, const int64_t *states, int64_t *idxs)
{
  for (int64_t i = 0; i < n; ++i) idxs[i] = orc_s2i(s, states[i]);
}

void orc_i2s_array(const orc_subspace *s, int64_t n, const int64_t *idxs, int64_t *states)
{
  for (int64_t i = 0; i < n; ++i) states[i] = orc_i2s(s, idxs[i]);
}

/* ------------------------------------------------------------------ */
/* matrix element of one mask on column state `bra`                    */
/* bpetsc_template_2.c:399-407 with real_coeffs per :286-290 and       */
/* TERM_REAL per bpetsc_impl.h:34                                      */
/* ------------------------------------------------------------------ */
static inline double real_coeff(const orc_msc *msc, int64_t t)
{
  double re = msc->coeffs[2 * t];
  return (re != 0) ? re : msc->coeffs[2 * t + 1];
}

static inline cplx mask_element(const orc_msc *msc, int64_t mi, int64_t bra)
{
  cplx v = 0;
  int64_t m = msc->masks[mi];
  for (int64_t t = msc->mask_offsets[mi]; t < msc->mask_offsets[mi + 1]; ++t) {
    int64_t sg = msc->signs[t];
    double sign = 1 - 2 * par64(bra & sg);
    if (!par64(m & sg)) v += sign * real_coeff(msc, t);
    else v += I * sign * real_coeff(msc, t);
  }
  return v;
}

/* bpetsc_template_2.c:371-412 (single rank); dims per :227-230 */
int orc_matmult(const orc_msc *msc, const orc_subspace *left, const orc_subspace *right,
                int xparity, const double *diag, const double *x_, double *y_)
{
  const cplx *x = (const cplx *)x_;
  cplx *y = (cplx *)y_;
  int64_t M = orc_dim(left);
  if (xparity) M /= 2;
  int64_t ket = 0;
  for (int64_t row = 0; row < M; ++row) {
    ket = (row == 0) ? orc_i2s(left, 0) : orc_next_state(left, ket, row);
    cplx acc = 0;
    int64_t mi = 0;
    if (diag) {
      acc += diag[row] * x[row];
      mi = 1;
    }
    for (; mi < msc->nmasks; ++mi) {
      int64_t bra = ket ^ msc->masks[mi];
      int64_t col = orc_s2i(right, bra);
      if (col == -1) continue;
      acc += mask_element(msc, mi, bra) * x[col];
    }
    y[row] = acc;
  }
  return 0;
}

/* bpetsc_template_1.c:169-202 */
int orc_precompute_diag(const orc_msc *msc, const orc_subspace *sub, int xparity, double *diag)
{
  if (msc->nmasks == 0 || msc->masks[0] != 0) return 1; /* no diagonal */
  int64_t M = orc_dim(sub);
  if (xparity) M /= 2;
  for (int64_t row = 0; row < M; ++row) {
    int64_t state = orc_i2s(sub, row);
    double v = 0;
    for (int64_t t = 0; t < msc->mask_offsets[1]; ++t) {
      double sign = 1 - 2 * par64(state & msc->signs[t]);
      v += sign * real_coeff(msc, t);
    }
    diag[row] = v;
  }
  return 0;
}

/* same, rows [row_first, row_last) only (bench.py's bounded CPU sample) */
int orc_precompute_diag_range(const orc_msc *msc, const orc_subspace *sub, int64_t row_first,
                              int64_t row_last, double *diag)
{
  if (msc->nmasks == 0 || msc->masks[0] != 0) return 1;
  for (int64_t row = row_first; row < row_last; ++row) {
    int64_t state = orc_i2s(sub, row);
    double v = 0;
    for (int64_t t = 0; t < msc->mask_offsets[1]; ++t) {
      double sign = 1 - 2 * par64(state & msc->signs[t]);
      v += sign * real_coeff(msc, t);
    }
    diag[row] = v;
  }
  return 0;
}

/* bpetsc_template_2.c:906-981: max row sum of |element|, Kahan-summed */
int orc_norm_inf(const orc_msc *msc, const orc_subspace *left, const orc_subspace *right,
                 int xparity, double *nrm)
{
  int64_t M = orc_dim(left);
  if (xparity) M /= 2;
  double best = 0;
  for (int64_t row = 0; row < M; ++row) {
    int64_t ket = orc_i2s(left, row);
    double sum = 0, err = 0;
    for (int64_t mi = 0; mi < msc->nmasks; ++mi) {
      int64_t bra = ket ^ msc->masks[mi];
      if (orc_s2i(right, bra) == -1) continue;
      double comp = cabs(mask_element(msc, mi, bra)) - err;
      double total = sum + comp;
      err = (total - sum) - comp;
      sum = total;
    }
    if (sum > best) best = sum;
  }
  *nrm = best;
  return 0;
}

/* bpetsc_template_2.c:990-1056 (uses the complex coeffs directly, :1036-1039) */
int orc_check_conserves(const orc_msc *msc, const orc_subspace *left, const orc_subspace *right,
                        int xparity, int *result)
{
  int64_t N = orc_dim(right);
  if (xparity) N /= 2;
  *result = 1;
  for (int64_t col = 0; col < N; ++col) {
    int64_t bra = orc_i2s(right, col);
    for (int64_t mi = 0; mi < msc->nmasks; ++mi) {
      int64_t ket = bra ^ msc->masks[mi];
      if (orc_s2i(left, ket) != -1) continue;
      cplx v = 0;
      for (int64_t t = msc->mask_offsets[mi]; t < msc->mask_offsets[mi + 1]; ++t) {
        double sign = 1 - 2 * par64(bra & msc->signs[t]);
        v += sign * (msc->coeffs[2 * t] + I * msc->coeffs[2 * t + 1]);
      }
      if (v != 0) {
        *result = 0;
        return 0;
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* rdm: bpetsc_template_1.c:15-165                                     */
/* ------------------------------------------------------------------ */

/* :29-55 interleave kept and traced bits back into a full state */
static int64_t weave(int64_t keep_state, int64_t tr_state, const int64_t *keep,
                     int64_t keep_size, int64_t L)
{
  int64_t out = 0, ki = 0, ti = 0;
  for (int64_t b = 0; b < L; ++b) {
    int64_t bit;
    if (ki < keep_size && keep[ki] == b) {
      bit = (keep_state >> ki) & 1;
      ++ki;
    } else {
      bit = (tr_state >> ti) & 1;
      ++ti;
    }
    out |= bit << b;
  }
  return out;
}

int orc_rdm(const double *x_, const orc_subspace *sub, int64_t keep_size, const int64_t *keep,
            int64_t rtn_dim, double *rtn_)
{
  const cplx *x = (const cplx *)x_;
  cplx *rtn = (cplx *)rtn_;
  for (int64_t i = 1; i < keep_size; ++i)
    if (keep[i] <= keep[i - 1]) return 1; /* :109-113 */

  int64_t kdim = (int64_t)1 << keep_size;
  int64_t *which = malloc(sizeof(int64_t) * kdim);
  cplx *amp = malloc(sizeof(cplx) * kdim);
  memset(rtn, 0, sizeof(cplx) * rtn_dim * rtn_dim);

  int64_t tr_dim = (int64_t)1 << (sub->L - keep_size);
  for (int64_t tr = 0; tr < tr_dim; ++tr) {
    int64_t filled = 0;
    for (int64_t ks = 0; ks < kdim; ++ks) { /* :57-83 */
      int64_t idx = orc_s2i(sub, weave(ks, tr, keep, keep_size, sub->L));
      if (idx == -1) continue;
      which[filled] = ks;
      amp[filled] = x[idx];
      ++filled;
    }
    for (int64_t i = 0; i < filled; ++i) /* :144-154 */
      for (int64_t j = 0; j < filled; ++j)
        rtn[which[i] * rtn_dim + which[j]] += amp[i] * conj(amp[j]);
  }
  free(which);
  free(amp);
  return 0;
}

/* ------------------------------------------------------------------ */
/* compute_rcm: bsubspace.pyx:212-261 (BFS over the operator's graph)  */
/* ------------------------------------------------------------------ */

typedef struct {
  int64_t *slot;
  int64_t cap; /* power of two */
} seen_set;

static int seen_insert(seen_set *s, int64_t v)
{
  uint64_t h = (uint64_t)v * 0x9E3779B97F4A7C15ull;
  int64_t i = (int64_t)(h >> 11) & (s->cap - 1);
  while (s->slot[i] != -1) {
    if (s->slot[i] == v) return 0;
    i = (i + 1) & (s->cap - 1);
  }
  s->slot[i] = v;
  return 1;
}

int64_t orc_compute_rcm(int64_t nterms, const int64_t *masks, const int64_t *signs,
                        const double *coeffs, int64_t *state_map, int64_t max_states,
                        int64_t start, int64_t L)
{
  (void)L;
  seen_set seen;
  seen.cap = 16;
  while (seen.cap < 2 * max_states) seen.cap <<= 1;
  seen.slot = malloc(sizeof(int64_t) * seen.cap);
  for (int64_t i = 0; i < seen.cap; ++i) seen.slot[i] = -1;

  int64_t filled = 0;
  state_map[filled++] = start;
  seen_insert(&seen, start);

  for (int64_t i = 0; i < max_states && i < filled; ++i) {
    int64_t state = state_map[i];
    cplx tot = 0;
    for (int64_t t = 0; t < nterms; ++t) {
      int sg = par64(state & signs[t]);
      tot += (1 - 2 * sg) * (coeffs[2 * t] + I * coeffs[2 * t + 1]);
      if (t + 1 == nterms || masks[t + 1] != masks[t]) {
        if (tot != 0) {
          int64_t edge = state ^ masks[t];
          if (seen_insert(&seen, edge)) {
            if (filled >= max_states) {
              free(seen.slot);
              return -1; /* 'state_map size too small' */
            }
            state_map[filled++] = edge;
          }
        }
        tot = 0;
      }
    }
  }
  free(seen.slot);
  return filled;
}

/* ------------------------------------------------------------------ */
/* Fast Full/Parity product: bpetsc_template_2.c:563-889               */
/*                                                                     */
/* Restated for one address space: rows are processed in blocks of     */
/* 2048 (VECSET_CACHE_SIZE :526); per mask the per-row coefficient is  */
/* accumulated term by term with a 64x64 sign table on the low six     */
/* index bits and a hoisted parity of the high bits (sum_term          */
/* :637-683), then multiplied into x in contiguous runs of 2^ctz(mask) */
/* (do_cache_product :598-635).  Signs are evaluated on the ROW index, */
/* so each coefficient carries the factor (-1)^parity(mask & sign)     */
/* (:844-846).  MPI ranks of the reference become pthreads over        */
/* blocks; the VecSetValues/VecAssembly hand-off (:866-873) becomes a  */
/* direct store, which only favours this baseline.                     */
/* ------------------------------------------------------------------ */
#define BLK 2048
#define LKP 64

#include <pthread.h>

typedef struct {
  const orc_msc *msc;
  const orc_subspace *sub;
  const double *diag;
  const cplx *x;
  cplx *y;
  const double *tab_plain, *tab_par;
  int64_t blk_begin, blk_end; /* block indices handled by this thread */
} fast_job;

static void *fast_worker(void *arg)
{
  const fast_job *job = (const fast_job *)arg;
  const orc_msc *msc = job->msc;
  const int is_parity = (job->sub->type == ORC_PARITY);
  const cplx *x = job->x;
  cplx *coef = malloc(sizeof(cplx) * BLK);
  cplx *vals = malloc(sizeof(cplx) * BLK);
  const int64_t hi_mask = ~(int64_t)(LKP - 1);

  for (int64_t blk = job->blk_begin; blk < job->blk_end; ++blk) {
    const int64_t b0 = blk * BLK;
    int64_t mi = 0;
    if (job->diag) { /* :810-816 */
      for (int64_t i = 0; i < BLK; ++i) vals[i] = job->diag[b0 + i] * x[b0 + i];
      mi = 1;
    } else {
      memset(vals, 0, sizeof(cplx) * BLK);
    }

    for (; mi < msc->nmasks; ++mi) {
      int64_t mask = msc->masks[mi];
      if (is_parity && par64(mask)) continue; /* :822-827 */
      int64_t m = is_parity ? (mask >> 1) : mask; /* S2I_nocheck */
      memset(coef, 0, sizeof(cplx) * BLK);

      for (int64_t t = msc->mask_offsets[mi]; t < msc->mask_offsets[mi + 1]; ++t) {
        int64_t sign = msc->signs[t];
        int64_t s = is_parity ? (sign >> 1) : sign;
        int msp = par64(mask & sign);
        double c = (msp ? -1.0 : 1.0) * real_coeff(msc, t); /* :844-846 */
        int chk = is_parity && (sign & 1);                   /* :849-856 */
        const double *row = (chk ? job->tab_par : job->tab_plain) + (s & (LKP - 1)) * LKP;
        for (int64_t i = 0; i < BLK; i += LKP) { /* sum_term :657-666 */
          int flip = par64((b0 + i) & hi_mask & s);
          if (chk) flip ^= par64((b0 + i) & hi_mask);
          double tc = flip ? -c : c;
          if (!msp)
            for (int j = 0; j < LKP; ++j) coef[i + j] += row[j] * tc;
          else
            for (int j = 0; j < LKP; ++j) coef[i + j] += I * (row[j] * tc);
        }
      }

      /* do_cache_product :598-635 */
      int64_t run = (m == 0) ? BLK : ((int64_t)1 << ctz64(m));
      if (run > BLK) run = BLK;
      for (int64_t i = 0; i < BLK; i += run) {
        const cplx *src = x + ((b0 + i) ^ m);
        for (int64_t j = 0; j < run; ++j) vals[i + j] += coef[i + j] * src[j];
      }
    }
    memcpy(job->y + b0, vals, sizeof(cplx) * BLK);
  }
  free(coef);
  free(vals);
  return NULL;
}

/* blocks [blk_first, blk_last) of 2048 rows only: the bounded sample bench.py times.
 * y must still have room for the whole vector (only the sampled rows are written). */
int orc_matmult_fast_range(const orc_msc *msc, const orc_subspace *sub, const double *diag,
                           const double *x_, double *y_, int64_t blk_first, int64_t blk_last,
                           int nthreads)
{
  if (sub->type != ORC_FULL && sub->type != ORC_PARITY) return -1;
  const int64_t N = orc_dim(sub);
  if (N < BLK) return -1; /* reference falls back to the general path (:549) */
  if (blk_last > N / BLK) blk_last = N / BLK;
  if (blk_first < 0 || blk_first >= blk_last) return -1;
  if (nthreads < 1) nthreads = 1;
  const int64_t nblk = blk_last - blk_first;
  if (nthreads > nblk) nthreads = (int)nblk;

  /* sign tables, :575-596 */
  double *tab_plain = malloc(sizeof(double) * LKP * LKP);
  double *tab_par = malloc(sizeof(double) * LKP * LKP);
  for (int i = 0; i < LKP; ++i)
    for (int j = 0; j < LKP; ++j) {
      int p = par64(i & j);
      tab_plain[i * LKP + j] = p ? -1.0 : 1.0;
      int q = p ^ par64(j) ^ (int)sub->space;
      tab_par[i * LKP + j] = q ? -1.0 : 1.0;
    }

  pthread_t *tid = malloc(sizeof(pthread_t) * nthreads);
  fast_job *jobs = malloc(sizeof(fast_job) * nthreads);
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = (fast_job){msc, sub, diag, (const cplx *)x_, (cplx *)y_, tab_plain, tab_par,
                         blk_first + nblk * t / nthreads, blk_first + nblk * (t + 1) / nthreads};
    if (t > 0) pthread_create(&tid[t], NULL, fast_worker, &jobs[t]);
  }
  fast_worker(&jobs[0]);
  for (int t = 1; t < nthreads; ++t) pthread_join(tid[t], NULL);
  free(tid);
  free(jobs);
  free(tab_plain);
  free(tab_par);
  return nthreads;
}

int orc_matmult_fast(const orc_msc *msc, const orc_subspace *sub, const double *diag,
                     const double *x_, double *y_, int nthreads)
{
  if (sub->type != ORC_FULL && sub->type != ORC_PARITY) return -1;
  return orc_matmult_fast_range(msc, sub, diag, x_, y_, 0, orc_dim(sub) / BLK, nthreads);
}
