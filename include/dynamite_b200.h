/*
 * dynamite_b200.h -- C ABI of the B200-native backend for dynamite's matrix-free
 * MSC shell path (MatMult, its Krylov consumers, rdm, subspace index maps).
 *
 * This is the drop-in boundary: every entry point below replaces one piece of
 * dynamite's Cython/C backend or of the petsc4py/slepc4py surface that
 * dynamite's Python layer calls for this path.  Reference citations are
 * relative to /root/reference/src/dynamite/ .
 *
 * Conventions
 *   - every function returns 0 on success and a non-zero dnm_status otherwise
 *     (the PetscErrorCode role, _backend/bpetsc.pyx:135-136); the message is
 *     available from dnm_last_error().  Nothing throws across the boundary.
 *   - indices/states are int64 (PETSC_USE_64BIT_INDICES), scalars complex128
 *     passed as interleaved (re, im) doubles (PETSC_USE_COMPLEX).
 *   - caller arrays are borrowed for the duration of the call only; the
 *     library deep-copies what it keeps (as _backend/bpetsc_template_2.c:275-297).
 *   - one process drives one GPU (the reference's rank<->device mapping,
 *     _backend/bcuda_template_2.cu:64-67); all device work of a process is
 *     issued on one library-owned stream (dnm_stream()).
 *   - there is NO CPU fallback: every compute entry point fails with
 *     DNM_ERR_CUDA if no sm_100-class device is usable.  The subspace index
 *     maps and compute_rcm are host functions in the reference as well
 *     (_backend/bsubspace.pyx:144-261) and are host functions here.
 */
#ifndef DYNAMITE_B200_H
#define DYNAMITE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  DNM_OK = 0,
  DNM_ERR_ARG = 1,        /* bad argument (PETSC_ERR_ARG_*) */
  DNM_ERR_CUDA = 2,       /* CUDA runtime / no device */
  DNM_ERR_MEM = 3,        /* allocation failure */
  DNM_ERR_UNSUPPORTED = 4,
  DNM_ERR_COMM = 5,       /* NCCL / IPC failure */
  DNM_ERR_INTERNAL = 6
} dnm_status;

/* _backend/bsubspace_impl.h:17-23 (same numeric values) */
typedef enum {
  DNM_FULL = 0,
  DNM_PARITY = 1,
  DNM_EXPLICIT = 2,
  DNM_SPIN_CONSERVE = 3
} dnm_subspace_type;

/* One descriptor for the four data_* structs of _backend/bsubspace_impl.h
 * (:39-42, :95-99, :161-167, :265-272).  Unused fields are ignored. */
typedef struct {
  int32_t type;                /* dnm_subspace_type */
  int64_t L;
  int64_t space;               /* Parity: 0 even / 1 odd */
  int64_t k;                   /* SpinConserve: number of set bits */
  int64_t ld_nchoosek;         /* SpinConserve: row length of nchoosek[(k+1) x ld] */
  const int64_t *nchoosek;     /* SpinConserve: nchoosek[kk*ld + n] = C(n, kk) */
  int64_t dim;                 /* Explicit */
  const int64_t *state_map;    /* Explicit: idx -> state */
  const int64_t *rmap_indices; /* Explicit: NULL if state_map is sorted */
  const int64_t *rmap_states;  /* Explicit: sorted states */
} dnm_subspace_t;

typedef struct dnm_vec_s *dnm_vec_t;
typedef struct dnm_mat_s *dnm_mat_t;

/* ---- library / device -------------------------------------------------- */

/* Bind this process to CUDA device `device` and create the library stream.
 * Role of config.initialize(gpu=True) (__init__.py:51-157). */
int dnm_init(int device);
int dnm_finalize(void);
const char *dnm_last_error(void);
/* 1 if a usable CUDA device is bound.  _backend/bbuild.pyx have_gpu_shell(). */
int dnm_have_gpu(void);
int dnm_device_count(int *count);
/* cudaStream_t of the library stream, as void*. */
void *dnm_stream(void);
int dnm_synchronize(void);
/* CUDA-event stopwatch on the library stream (for bench.py). */
int dnm_timer_start(void);
int dnm_timer_stop(float *milliseconds);
/* free / total device memory in bytes */
int dnm_mem_info(int64_t *free_bytes, int64_t *total_bytes);
/* number of kernels this library has launched since dnm_init / last reset */
int64_t dnm_launch_count(int reset);
/* page-locked host buffers for callers that stage vectors on the host */
int dnm_host_alloc(int64_t bytes, void **out);
int dnm_host_free(void *ptr);

/* ---- multi-GPU (one process per GPU) ----------------------------------- */

/* NCCL bootstrap.  Rank 0 calls dnm_comm_unique_id and ships the 128 bytes to
 * the other ranks by any host channel (torch.distributed in bench.py); then
 * every rank calls dnm_comm_init.  Replaces PETSC_COMM_WORLD for this path.
 * nranks must be a power of two (the reference's fast path has the same
 * requirement, _backend/bpetsc_template_2.c:542-546). */
int dnm_comm_unique_id(char id[128]);
int dnm_comm_init(int rank, int nranks, const char id[128]);
int dnm_comm_rank(int *rank, int *nranks);
/* Vectors created after dnm_comm_init are CUDA-IPC shared with every peer rank
 * (handles travel over NCCL inside dnm_vec_create, which is then collective),
 * so MatMult kernels can load remote amplitudes over NVLink directly. */
int dnm_comm_barrier(void);
/* Host-only description of the sharding (no device needed; used by the multi-rank CPU tests).
 * The index space has n_index_bits bits (L for Full, L-1 for Parity; one less under XParity) and
 * rank r of nranks = 2^p owns the indices whose top p bits equal r.  For index-space flip mask
 * index_masks[k], a row owned by `rank` gathers from rank partner[k] = rank ^ (mask >> n_local)
 * at local index (i ^ local_masks[k]).  This is the reference's power-of-two block layout
 * (_backend/bpetsc_template_2.c:768-797: proc_start_idx = proc_mask & (proc_me ^ m)). */
int dnm_shard_plan(int n_index_bits, int nranks, int rank, int64_t nmasks, const int64_t *index_masks,
                   int32_t *partner, int64_t *local_masks);

/* ---- subspace index maps (host) ---------------------------------------- */

/* Dim_*            _backend/bsubspace_impl.h:57,112,187,302 ; bsubspace.pyx:144-154 */
int dnm_subspace_dim(const dnm_subspace_t *s, int64_t *dim);
/* S2I_*_array      bsubspace_impl.h:85,145,248,349 ; bsubspace.pyx:182-206.  -1 = not in subspace */
int dnm_subspace_s2i(const dnm_subspace_t *s, int64_t n, const int64_t *states, int64_t *idxs);
/* I2S_*_array      bsubspace_impl.h:89,152,256,357 ; bsubspace.pyx:158-180.
 * DNM_ERR_ARG if an index is outside [0, dim) (the PetscAssert of :70,130,211,334). */
int dnm_subspace_i2s(const dnm_subspace_t *s, int64_t n, const int64_t *idxs, int64_t *states);
/* compute_rcm      bsubspace.pyx:212-261.  masks/signs/coeffs are per-term arrays
 * sorted by mask.  DNM_ERR_ARG with message 'state_map size too small' on overflow. */
int dnm_compute_rcm(int64_t nterms, const int64_t *masks, const int64_t *signs,
                    const double *coeffs, int64_t *state_map, int64_t max_states,
                    int64_t start, int64_t L, int64_t *dim_out);
/* The same search on the device (frontier expansion with a hash table of first-discovery keys,
 * csrc/rcm.cu): identical output, element for element.  Needs dnm_init. */
int dnm_compute_rcm_device(int64_t nterms, const int64_t *masks, const int64_t *signs,
                           const double *coeffs, int64_t *state_map, int64_t max_states,
                           int64_t start, int64_t L, int64_t *dim_out);
/* The same maps evaluated by the device functions the kernels use (for the
 * bit-exact device-vs-oracle parity tests).  Host arrays in, host arrays out. */
int dnm_subspace_s2i_device(const dnm_subspace_t *s, int64_t n, const int64_t *states, int64_t *idxs);
int dnm_subspace_i2s_device(const dnm_subspace_t *s, int64_t n, const int64_t *idxs, int64_t *states);

/* ---- state vectors (the petsc4py Vec surface of SURVEY.md 3.5) ---------- */

/* Vector of global length n, block-distributed over the ranks of the
 * communicator (n/nranks each; single rank: all local). */
int dnm_vec_create(int64_t n, dnm_vec_t *out);
int dnm_vec_destroy(dnm_vec_t v);
int dnm_vec_size(dnm_vec_t v, int64_t *global_n, int64_t *local_start, int64_t *local_end);
void *dnm_vec_device_ptr(dnm_vec_t v);
/* host <-> device, offsets/counts in elements of the LOCAL block */
int dnm_vec_set_host(dnm_vec_t v, int64_t offset, int64_t count, const double *values);
int dnm_vec_get_host(dnm_vec_t v, int64_t offset, int64_t count, double *values);
/* the same for any range of the GLOBAL vector: other ranks' blocks are read through their peer
 * mappings (the ranks must be synchronised by the caller; every rank may call it) */
int dnm_vec_get_host_global(dnm_vec_t v, int64_t offset, int64_t count, double *values);
/* scattered local writes / reads (Vec.setValues / vec[idxs]) */
int dnm_vec_set_values(dnm_vec_t v, int64_t count, const int64_t *local_idx, const double *values, int add);
int dnm_vec_get_values(dnm_vec_t v, int64_t count, const int64_t *local_idx, double *values);
int dnm_vec_set(dnm_vec_t v, double re, double im);                 /* Vec.set */
/* counter-based pseudo-random fill on the device (uniform in [-1,1] per component); synthetic
 * benchmark inputs.  State.set_random keeps numpy's host stream for parity with the reference. */
int dnm_vec_set_random(dnm_vec_t v, uint64_t seed);
int dnm_vec_copy(dnm_vec_t src, dnm_vec_t dst);                     /* Vec.copy */
int dnm_vec_scale(dnm_vec_t v, double re, double im);               /* Vec.scale */
int dnm_vec_shift(dnm_vec_t v, double re, double im);               /* Vec.shift: v[i] += a (states.py:805) */
int dnm_vec_normalize(dnm_vec_t v, double *norm_out);               /* Vec.normalize, returns the norm */
/* y = a*x + b*y                                                       Vec.axpby */
int dnm_vec_axpby(dnm_vec_t y, double a_re, double a_im, double b_re, double b_im, dnm_vec_t x);
/* out = sum_i x_i * conj(y_i)   (PETSc VecDot(x, y))                  Vec.dot */
int dnm_vec_dot(dnm_vec_t x, dnm_vec_t y, double out[2]);
/* type: 0 = 2-norm, 1 = 1-norm, 2 = infinity norm                     Vec.norm */
int dnm_vec_norm(dnm_vec_t v, int type, double *out);

/* ---- shell matrix ------------------------------------------------------- */

/* BuildMat / BuildGPUShell  _backend/bpetsc_impl.c:168-252, bcuda_template_2.cu:4-108,
 * called from bpetsc.build_mat (_backend/bpetsc.pyx:78-138).
 * masks: unique ascending flip masks [nmasks]; mask_offsets [nmasks+1] into
 * signs/coeffs (operators.py:653-669); coeffs complex [nterms].
 * The matrix must be Hermitian term by term (msc_tools.py:94-118), as
 * build_mat enforces (operators.py:605-606): DNM_ERR_ARG otherwise. */
int dnm_mat_create(int64_t nmasks, const int64_t *masks, const int64_t *mask_offsets,
                   const int64_t *signs, const double *coeffs,
                   const dnm_subspace_t *left, const dnm_subspace_t *right,
                   int xparity, dnm_mat_t *out);
/* PrecomputeDiagonal  _backend/bcuda_template_1.cu:4-66 (bpetsc.pyx:141-147).
 * No-op when masks[0] != 0. */
int dnm_mat_precompute_diagonal(dnm_mat_t A);
/* MATOP_MULT  _backend/bcuda_template_2.cu:141-273.  y = A x, device resident,
 * asynchronous on the library stream. */
int dnm_mat_mult(dnm_mat_t A, dnm_vec_t x, dnm_vec_t y);
/* Same product with HOST buffers: H2D of x, multiply, D2H of y, synchronous
 * (what Mat.mult costs a caller whose Vec lives on the host).  The two device
 * work vectors are created on first use and kept until dnm_mat_destroy. */
int dnm_mat_mult_host(dnm_mat_t A, const double *x_host, double *y_host);
/* `count` products y_k = A x_k with HOST buffers, pipelined over the PCIe link: the H2D copy of x_(k+1)
 * and the D2H copy of y_(k-1) run on their own streams while product k is evaluated (two device
 * buffers deep per direction, kept until dnm_mat_destroy).  Pinned buffers (dnm_host_alloc) are needed
 * for the copies to overlap; the x_k (and the y_k) may alias one another, an x_k must not alias a y_j.
 * Synchronous: every y_k is complete on return.  Single rank. */
int dnm_mat_mult_host_batch(dnm_mat_t A, int64_t count, const double *const *x_hosts,
                            double *const *y_hosts);
/* MATOP_NORM (NORM_INFINITY only)  _backend/bcuda_template_2.cu:275-403; cached. */
int dnm_mat_norm_inf(dnm_mat_t A, double *nrm);
int dnm_mat_size(dnm_mat_t A, int64_t *M, int64_t *N);
/* MATOP_DESTROY  _backend/bcuda_template_2.cu:110-139 */
int dnm_mat_destroy(dnm_mat_t A);
/* Tuning / introspection knobs.  keys:
 *   "kernel"    0 auto, 1 general gather, 2 tiled window
 *   "tile_bits" 0 auto, 8..13 (log2 of the tile held in shared memory)
 *   "tile_rows" 0 auto, 8, 16 (rows per thread of the generic tiled kernel)
 *   "far_bits"  -1 auto, 0..16: index positions outside the window a tiled pass may serve through
 *               the L2 (FAR masks: operands read from global memory while the tiles of one
 *               far-bit block are in flight together)
 *   "jit"       -1 auto (on for >= 2^22 rows per GPU), 0 off, 1 on: operator-specialised pass
 *               kernels generated as CUDA source and compiled with NVRTC for sm_100a
 *   "pipeline"  generated kernels: 1 = persistent CTAs with a TMA (cp.async.bulk[.tensor]) +
 *               mbarrier ring and a cp.reduce.async.bulk.tensor add epilogue; 0 / 2 = one tile
 *               per CTA (default: measured faster)
 *   "autotune"  -1 auto (on for >= 2^24 rows per GPU), 0 off, 1 on: the first MatMult times a
 *               few plan shapes on the caller's vectors and keeps the fastest
 *   "verbose"   plan / autotune messages on stderr */
int dnm_mat_set_option(dnm_mat_t A, const char *key, int64_t value);
/* keys: "kernel", "passes", "jit_passes" (passes running generated kernels), "tuned_shape",
 * "unique_masks", "nterms", "model_bytes", "compulsory_bytes", "launches_per_mult", "has_diag" */
int dnm_mat_get_info(dnm_mat_t A, const char *key, double *value);
/* Diagnostics, no GPU needed: plan the tiled MatMult of a Full/Parity operator as rank `rank` of
 * `nranks` would, generate the operator-specialised pass kernels (csrc/jit.cu) and compile them with
 * NVRTC to an sm_100a cubin.  tune_shape -1 = default heuristics, >= 0 = one autotuner shape.  The
 * generated source is copied to src_out (src_cap bytes, NUL-terminated) when given. */
int dnm_jit_dryrun(int64_t nmasks, const int64_t *masks, const int64_t *mask_offsets,
                   const int64_t *signs, const double *coeffs, const dnm_subspace_t *sub, int nranks,
                   int rank, int tile_bits, int far_bits, int pipeline, int tune_shape, char *src_out,
                   int64_t src_cap, int64_t *src_len, int64_t *cubin_bytes, int *n_kernels,
                   int *n_passes, int *n_remote_groups, int *n_pipelined);
/* Tests only: while on, dnm_jit_dryrun emits the SAME pass kernels for a C++ compiler -- a prelude
 * maps the CUDA vocabulary to the host (one OS thread per CUDA thread, TMA boxes as copy loops) and
 * an `extern "C" dnm_emu_mult` harness walks the passes -- and skips NVRTC (cubin_bytes = 0).  The
 * product never sets it: tests/test_jit_emulation.py checks what the generated code computes
 * against the oracle on machines without a GPU. */
int dnm_jit_set_host_emulation(int on);
/* CheckConserves  _backend/bpetsc_template_2.c:990-1056 (bpetsc.pyx:150-193) */
int dnm_check_conserves(int64_t nmasks, const int64_t *masks, const int64_t *mask_offsets,
                        const int64_t *signs, const double *coeffs,
                        const dnm_subspace_t *left, const dnm_subspace_t *right,
                        int xparity, int *result);

/* ---- Krylov consumers ---------------------------------------------------- */

/* reasons mirror SLEPc's MFNConvergedReason / EPSConvergedReason values */
enum {
  DNM_CONVERGED_TOL = 1,
  DNM_CONVERGED_ITS = 2,       /* MFN only */
  DNM_DIVERGED_ITS = -1,
  DNM_DIVERGED_BREAKDOWN = -2,
  DNM_DIVERGED_SYMMETRY_LOST = -3
};

/* y = exp(scale * A) x with the expokit sub-stepped Arnoldi scheme, i.e. what
 * computations.evolve gets from SLEPc.MFN type 'expokit' with FN exp scaled
 * by -i*t (computations.py:89-112).  tol<=0, ncv<=0, max_it<=0 select the
 * SLEPc defaults (1e-7, min(30,N), max(100, 2N/ncv)); ncv is additionally
 * capped by free device memory (see dnm_evolve_algo, algo -1, for what happens
 * then).  Outputs may be NULL. */
int dnm_evolve(dnm_mat_t A, dnm_vec_t x, dnm_vec_t y, double scale_re, double scale_im,
               double tol, int ncv, int max_it, int *reason, int *its, int *matmults);
/* The same with the algorithm spelled out: algo 0 = expokit sub-stepping with the Lanczos recurrence
 * (the operator is Hermitian), 1 = expokit sub-stepping with full Arnoldi orthogonalisation (MFN type
 * "krylov" is served by it), 2 = Chebyshev propagator (csrc/chebyshev.h: Jacobi-Anger expansion on
 * [-||A||_inf, ||A||_inf], three work vectors, no inner products; real-time evolution only, i.e.
 * scale_re == 0), -1 = what dnm_evolve does: expokit/Lanczos, unless the Krylov basis would have to be
 * cut down to fit device memory and the evolution is in real time -- then Chebyshev
 * (DNM_EVOLVE_CHEB=0/1 in the environment forces the choice).  A Chebyshev result whose norm moved
 * by more than 1e-9 is discarded and recomputed by expokit (an error under algo 2). */
int dnm_evolve_algo(dnm_mat_t A, dnm_vec_t x, dnm_vec_t y, double scale_re, double scale_im,
                    double tol, int ncv, int max_it, int algo, int *reason, int *its, int *matmults);
/* Which algorithm the most recent dnm_evolve / dnm_evolve_algo ran (0, 1, 2 as above; -1 before the first). */
int dnm_evolve_last_algo(void);

/* Hermitian eigensolve by thick-restart Lanczos (Krylov-Schur), i.e. what
 * computations.eigsolve gets from SLEPc.EPS HEP with the default solver
 * (computations.py:208-287).  which: 0 lowest (SMALLEST_REAL), 1 highest
 * (LARGEST_REAL), 2 exterior (LARGEST_MAGNITUDE).  On return *nconv pairs are
 * stored: evals[i] and, if evecs != NULL, evecs[i] (pre-created vectors, at
 * least `nev` of them; pairs beyond the capacity `max_pairs` are dropped). */
int dnm_eigsolve(dnm_mat_t A, int nev, int which, double tol, int max_it, int ncv,
                 uint64_t seed, int max_pairs, int *nconv, double *evals, double *errest,
                 dnm_vec_t *evecs, int *reason, int *its, int *matmults);

/* ReducedDensityMatrix  _backend/bpetsc_template_1.c:87-165 (bpetsc.pyx:245-276).
 * out: host, (2^keep_size)^2 complex row-major, valid on every rank. */
int dnm_rdm(dnm_vec_t v, const dnm_subspace_t *sub, int64_t keep_size, const int64_t *keep,
            double *out);

#ifdef __cplusplus
}
#endif
#endif
