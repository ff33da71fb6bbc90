// Operator-specialised MatMult kernels: CUDA source generated per tiled pass at build time of the
// plan, compiled for sm_100a with NVRTC and launched through the driver API.
//
// The generic window-tiled kernel (tiled_kernel.cuh) interprets group descriptors: on B200 it issues
// ~105 warp instructions per (warp, mask) of which 18 are the LDS/LDG + DFMA that do the work
// (profiles/r02_far_T12f7_*).  For a lean pass (every mask has at most two distinct sign masks:
// all nearest-neighbour, Ising, long-range and field models) every quantity the interpreter decodes
// -- window coordinates of the mask, which sign bits fall on tile / thread / row positions, which
// rows vanish, the coefficients themselves -- is known when the plan is built, so the pass is emitted
// as straight-line code with immediates.
#pragma once

#include <string>
#include <vector>

#include "tiled_kernel.cuh"

namespace dnm {
namespace jit {

struct Kernel {
  void *func = nullptr;  // CUfunction
  size_t smem = 0;
  int threads = 0;
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
  std::vector<int> W;  // window positions, W[j] = index bit of window coordinate j
};

// CUDA source of one kernel per pass, named dnm_jit_p<k>
std::string generate(const std::vector<PassDesc> &passes, int sm_tma);

// nullptr (and `log` filled) when NVRTC or the driver entry points are not available or the
// compilation fails; the caller then keeps the generic kernel
Module *compile(const std::string &src, const std::vector<PassDesc> &passes, std::string &log, bool load);

// compile only (no GPU needed): cubin bytes, empty on failure
std::vector<char> compile_cubin(const std::string &src, std::string &log);

void launch(const Kernel &k, unsigned long long ntiles, cudaStream_t stream, const cplx *x, cplx *y, const double *diag,
            long long rank_bits);

}  // namespace jit
}  // namespace dnm
