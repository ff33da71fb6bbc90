// Operator-specialised MatMult kernels: CUDA source generated per tiled pass when the plan is built,
// compiled for sm_100a with NVRTC and launched through the driver API.
//
// The generic window-tiled kernel (tiled_kernel.cuh) interprets group descriptors: on B200 it issues
// ~105 warp instructions per (warp, mask) of which 18 are the LDS/LDG + DFMA that do the work
// (profiles/r02_generic_T12f7_*).  For a lean pass (every mask has at most two distinct sign masks:
// all nearest-neighbour, Ising, long-range and field models) every quantity the interpreter decodes
// -- window coordinates of the mask, which sign bits fall on tile / thread / row positions, which
// rows vanish, the coefficients themselves -- is known when the plan is built, so the pass is emitted
// as straight-line code with immediates.
//
// Two kernel shapes:
//  * classic: one tile per CTA, cp.async staging (any window);
//  * pipelined: persistent CTAs, the tiles arrive through TMA (cp.async.bulk[.tensor]) in a ring of
//    shared-memory buffers signalled by mbarriers while the previous tile is being evaluated; an
//    accumulating pass adds its result tile into y with cp.reduce.async.bulk.tensor (.add, f64) from
//    a staging buffer, so the old y never enters the SM.  Needs a window that is a box of a <= 5-D
//    tensor (runs of consecutive bit positions).
#pragma once

#include <string>
#include <vector>

#include "tiled_kernel.cuh"

namespace dnm {
namespace jit {

constexpr int MAX_TMA_RANK = 5;

struct Kernel {
  void *func = nullptr;  // CUfunction
  size_t smem = 0;
  int threads = 0;
  // pipelined kernels
  bool pipelined = false;
  bool tma_args = false;  // one tile per CTA, but the tile is staged by TMA (tensor-map kernel signature)
  int ctas_per_sm = 1;
  int rank = 0;  // tensor rank of the tile box, 0: contiguous tile (1-D bulk copy, no tensor map)
  unsigned long long dims[MAX_TMA_RANK] = {0}, strides[MAX_TMA_RANK] = {0};  // strides in bytes, [0] unused
  unsigned box[MAX_TMA_RANK] = {0};
  bool reduce = false;  // accumulating pass: also needs the tensor map of y
  unsigned long long uid = 0;  // distinguishes kernels in the tensor-map cache
};

struct Module {
  void *mod = nullptr;  // CUmodule
  std::vector<Kernel> kernels;
  ~Module();
};

// what the generator needs to know about one lean pass
struct PassDesc {
  const tiled::PassParams *p = nullptr;
  const tiled::SmallTables *st = nullptr;
  int T = 0;
  int nloc = 0;        // index bits on this rank
  std::vector<int> W;  // window positions, W[j] = index bit of window coordinate j (ascending)
  // folded remote masks only on the tiles whose index bit filter_bit equals filter_val (-1: all tiles)
  int filter_bit = -1, filter_val = 0;
  // folded remote operands: false = loaded straight into registers inside the group, true = staged by
  // cp.async in shared-memory buffers of their own (classic kernels)
  bool stage_remote = false;
  // one tile per CTA: the tile is staged by one elected thread with cp.async.bulk[.tensor] + mbarrier
  // instead of eight cp.async per thread; tma_reduce: accumulating passes add their result tile into y
  // with cp.reduce.async.bulk.tensor instead of fetching the old y
  bool tma_stage = false, tma_reduce = false;
  // shape of the generated kernel
  bool pipelined = false;
  int rows = 8;  // rows per thread (4 or 8)
  int nbuf = 2;  // ring depth of a pipelined kernel
};

// can this window be fetched as one TMA box?
bool tma_eligible(const PassDesc &pd);
// ... and y can take a tensor reduce-add of the result tile (the contiguous tile has no tensor map)
bool tma_reducible(const PassDesc &pd);

// tests: emit the same kernel bodies for a C++ compiler (CUDA vocabulary emulated by the prelude)
void set_host_emulation(bool on);
bool host_emulation();

// CUDA source of one kernel per pass, named dnm_jit_p<k>
std::string generate(const std::vector<PassDesc> &passes);

// nullptr (and `log` filled) when NVRTC or the driver entry points are not available or the
// compilation fails; the caller then keeps the generic kernel
Module *compile(const std::string &src, const std::vector<PassDesc> &passes, std::string &log);

// compile only (no GPU needed): cubin bytes, empty on failure
std::vector<char> compile_cubin(const std::string &src, std::string &log);

// peers[h] = mapping of the input vector on rank (this ^ h), for passes with folded remote masks
void launch(const Kernel &k, unsigned long long ntiles, int sm_count, cudaStream_t stream, const cplx *x, cplx *y,
            const double *diag, long long rank_bits, const cplx *const *peers);

}  // namespace jit
}  // namespace dnm
