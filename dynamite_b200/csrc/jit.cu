// See jit.h.  Generator + NVRTC / driver-API plumbing (both resolved at run time: the library must
// load on a machine without libcuda / libnvrtc, where only the host entry points are used).
#include "jit.h"

#include <cuda.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstring>
#include <mutex>

namespace dnm {
namespace jit {

namespace {

using tiled::PassParams;
using tiled::SmallTables;
typedef unsigned int u32;
typedef unsigned long long u64;

struct Out {
  std::string s;
  void operator()(const char *fmt, ...)
  {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    s += buf;
  }
};

std::string hexd(double v)
{
  char buf[64];
  snprintf(buf, sizeof(buf), "%a", v);
  return buf;
}

int par32(u32 v) { return __builtin_parity(v); }

// "(tid & m) << s | ..." for the runs of consecutive window positions W[lo..hi)
std::string deposit_expr(const char *var, const std::vector<int> &W, int lo, int hi)
{
  std::string e;
  int b = lo;
  while (b < hi) {
    int end = b + 1;
    while (end < hi && W[end] == W[end - 1] + 1) ++end;
    const u64 mask = ((1ull << (end - b)) - 1ull) << b;
    const int shift = W[b] - b;
    char buf[160];
    snprintf(buf, sizeof(buf), "((i64)(%s & 0x%llxull) << %d)", var, mask, shift);
    if (!e.empty()) e += " | ";
    e += buf;
    b = end;
  }
  return e.empty() ? "(i64)0" : e;
}

struct Group {
  u32 lam, s1w, sBw;
  i64 so1, soB;
  double c1, c2;
  bool imag, far;
  u64 farmask;
};

void gen_pass(Out &o, const PassDesc &pd, int index)
{
  const PassParams &P = *pd.p;
  const SmallTables &S = *pd.st;
  const int T = pd.T, R = 8, LOG_NT = T - 3, NT = 1 << LOG_NT;
  const int minb = std::max(1, std::min(8, 65536 / (NT * 64)));
  const std::vector<int> &W = pd.W;

  i64 roff[8];
  i64 rowbits = 0;
  for (int r = 0; r < R; ++r) {
    roff[r] = 0;
    for (int k = 0; k < 3; ++k)
      if ((r >> k) & 1) roff[r] |= (i64)1 << W[LOG_NT + k];
  }
  for (int k = 0; k < 3; ++k) rowbits |= (i64)1 << W[LOG_NT + k];

  o("// ---- pass %d: T=%d, %d groups, %s, far_bits=%d\n", index, T, P.ngroups,
    P.accumulate ? "accumulate" : "write", P.far_bits);
  o("extern \"C\" __global__ void __launch_bounds__(%d, %d)\n", NT, minb);
  o("dnm_jit_p%d(const double2 *__restrict__ x, double2 *__restrict__ y, const double *__restrict__ diag, i64 rank_bits)\n{\n",
    index);
  o("  extern __shared__ double2 tile[];\n");
  o("  const u32 tid = threadIdx.x;\n");
  o("  const u64 tb = blockIdx.x;\n");
  // tile number -> index bits outside the window
  {
    std::string e;
    if (P.n_seg >= 0) {
      for (int j = 0; j < P.n_seg; ++j) {
        char buf[160];
        const int sh = P.seg_shift[j];
        snprintf(buf, sizeof(buf), "((tb & 0x%llxull) %s %d)", (u64)P.seg_mask[j], sh >= 0 ? "<<" : ">>", sh >= 0 ? sh : -sh);
        if (!e.empty()) e += " | ";
        e += buf;
      }
    } else {
      for (int k = 0; k < P.n_outer; ++k) {
        char buf[160];
        snprintf(buf, sizeof(buf), "(((tb >> %d) & 1ull) << %d)", k, (int)P.outer_pos[k]);
        if (!e.empty()) e += " | ";
        e += buf;
      }
    }
    o("  const i64 outer = (i64)(%s);\n", e.empty() ? "0ull" : e.c_str());
  }
  o("  const i64 og = outer | rank_bits;  // index bits shared by the tile (signs)\n");
  o("  const i64 base = outer | %s;\n", deposit_expr("tid", W, 0, LOG_NT).c_str());
  // stage the tile
  for (int r = 0; r < R; ++r) o("  cpa16(&tile[tid + %d], x + (base | 0x%llxll));\n", r * NT, (u64)roff[r]);
  o("  cpa_wait();\n  __syncthreads();\n");
  o("  double ar0, ai0, ar1, ai1, ar2, ai2, ar3, ai3, ar4, ai4, ar5, ai5, ar6, ai6, ar7, ai7;\n");
  o("  if (diag != nullptr) {\n");
  for (int r = 0; r < R; ++r)
    o("    { const double d = __ldg(diag + (base | 0x%llxll)); const double2 v = tile[tid + %d]; ar%d = d * v.x; ai%d = d * v.y; }\n",
      (u64)roff[r], r * NT, r, r);
  o("  } else {\n    ar0 = ai0 = ar1 = ai1 = ar2 = ai2 = ar3 = ai3 = ar4 = ai4 = ar5 = ai5 = ar6 = ai6 = ar7 = ai7 = 0.0;\n  }\n");

  for (int g = 0; g < P.ngroups; ++g) {
    Group G;
    G.lam = S.lam[g];
    G.s1w = S.sw[2 * g];
    G.sBw = S.sw[2 * g + 1];
    G.so1 = S.so[2 * g];
    G.soB = S.so[2 * g] ^ S.so[2 * g + 1];
    G.c1 = S.cf[2 * g];
    G.c2 = S.cf[2 * g + 1];
    G.imag = (S.kp[g] & 1) != 0;
    G.farmask = S.far[g];
    G.far = G.farmask != 0;
    if (G.c2 == 0.0) {  // one sign mask only
      G.sBw = 0;
      G.soB = 0;
    }
    const u32 lamlo = G.lam & (u32)(NT - 1);
    const int HI = (int)(G.lam >> LOG_NT);
    const u32 s1t = G.s1w & (u32)(NT - 1), sBt = G.sBw & (u32)(NT - 1);
    const u32 s1r = G.s1w >> LOG_NT, sBr = G.sBw >> LOG_NT;
    const double A = G.c1 + G.c2, Bc = G.c1 - G.c2;
    if (A == 0.0 && Bc == 0.0) continue;

    o("  {  // group %d: mask window 0x%x%s%s, c1=%s c2=%s\n", g, G.lam, G.imag ? " imag" : "", G.far ? " FAR" : "",
      hexd(G.c1).c_str(), hexd(G.c2).c_str());
    // dynamic parts of the two sign exponents
    auto expo = [&](i64 so, u32 st) -> std::string {
      std::string e;
      char buf[128];
      if (so != 0) {
        snprintf(buf, sizeof(buf), "(__popcll((u64)(og & 0x%llxll)) & 1)", (u64)so);
        e = buf;
      }
      if (st != 0) {
        snprintf(buf, sizeof(buf), "(__popc(tid & 0x%xu) & 1)", st);
        if (!e.empty()) e += " ^ ";
        e += buf;
      }
      return e;
    };
    const std::string e1 = expo(G.so1, s1t), eB = expo(G.soB, sBt);
    if (!e1.empty()) o("    const int e1 = %s;\n", e1.c_str());
    if (!eB.empty()) o("    const int eB = %s;\n", eB.c_str());
    if (G.far) {
      o("    const double2 *src = x + (base ^ 0x%llxll);\n", (u64)((i64)G.farmask & ~rowbits));
    } else {
      o("    const double2 *src = tile + (tid ^ 0x%xu);\n", lamlo);
    }
    // operand of row r
    auto operand = [&](int r) -> std::string {
      char buf[128];
      if (G.far) snprintf(buf, sizeof(buf), "__ldcg(src + 0x%llxll)", (u64)roff[r ^ HI]);
      else snprintf(buf, sizeof(buf), "src[%d]", (r ^ HI) * NT);
      return buf;
    };
    // acc[r] += (+-) c * operand for the rows in `rows`, coefficient variable `cv`
    auto emit_rows = [&](const std::vector<int> &rows, const char *cv, const char *indent) {
      for (size_t h = 0; h < rows.size(); h += 4) {
        const size_t e = std::min(rows.size(), h + 4);
        o("%s{\n", indent);
        for (size_t k = h; k < e; ++k) o("%s  const double2 v%d = %s;\n", indent, rows[k], operand(rows[k]).c_str());
        for (size_t k = h; k < e; ++k) {
          const int r = rows[k];
          const bool neg = par32(s1r & (u32)r) != 0;
          if (G.imag) {
            o("%s  ar%d = fma(%s%s, v%d.y, ar%d); ai%d = fma(%s%s, v%d.x, ai%d);\n", indent, r, neg ? "" : "-", cv, r, r, r,
              neg ? "-" : "", cv, r, r);
          } else {
            o("%s  ar%d = fma(%s%s, v%d.x, ar%d); ai%d = fma(%s%s, v%d.y, ai%d);\n", indent, r, neg ? "-" : "", cv, r, r, r,
              neg ? "-" : "", cv, r, r);
          }
        }
        o("%s}\n", indent);
      }
    };
    std::vector<int> rows0, rows1, all;
    for (int r = 0; r < R; ++r) {
      all.push_back(r);
      (par32(sBr & (u32)r) ? rows1 : rows0).push_back(r);
    }
    if (G.c2 == 0.0) {
      if (e1.empty()) o("    const double c = %s;\n", hexd(G.c1).c_str());
      else o("    const double c = e1 ? %s : %s;\n", hexd(-G.c1).c_str(), hexd(G.c1).c_str());
      emit_rows(all, "c", "    ");
    } else if (A == 0.0 || Bc == 0.0) {
      // rows with (eB ^ kB(r)) == act carry +-V, the others vanish (XX+YY: half of the rows)
      const int act = (A == 0.0) ? 1 : 0;
      const double V = (A == 0.0) ? Bc : A;
      if (e1.empty()) o("    const double c = %s;\n", hexd(V).c_str());
      else o("    const double c = e1 ? %s : %s;\n", hexd(-V).c_str(), hexd(V).c_str());
      if (eB.empty()) {
        emit_rows(act == 0 ? rows0 : rows1, "c", "    ");
      } else {
        o("    if (eB == %d) {\n", act);
        emit_rows(rows0, "c", "      ");
        if (!rows1.empty()) {
          o("    } else {\n");
          emit_rows(rows1, "c", "      ");
        }
        o("    }\n");
      }
    } else {
      // general pair: rows of class kB = 0 use (eB ? c1-c2 : c1+c2), class 1 the other one
      const char *sgn = e1.empty() ? "" : "e1 ? -1.0 : 1.0";
      if (eB.empty()) {
        o("    double c0 = %s, c1v = %s;\n", hexd(A).c_str(), hexd(Bc).c_str());
      } else {
        o("    double c0 = eB ? %s : %s, c1v = eB ? %s : %s;\n", hexd(Bc).c_str(), hexd(A).c_str(), hexd(A).c_str(),
          hexd(Bc).c_str());
      }
      if (!e1.empty()) o("    { const double sg = %s; c0 *= sg; c1v *= sg; }\n", sgn);
      emit_rows(rows0, "c0", "    ");
      if (!rows1.empty()) emit_rows(rows1, "c1v", "    ");
    }
    o("  }\n");
  }

  // epilogue
  if (P.accumulate == 1) {
    o("  __syncthreads();\n");
    for (int r = 0; r < R; ++r) o("  cpa16(&tile[tid + %d], y + (base | 0x%llxll));\n", r * NT, (u64)roff[r]);
    o("  cpa_wait();\n");
    for (int r = 0; r < R; ++r)
      o("  { const double2 old = tile[tid + %d]; y[base | 0x%llxll] = make_double2(ar%d + old.x, ai%d + old.y); }\n", r * NT,
        (u64)roff[r], r, r);
  } else {
    for (int r = 0; r < R; ++r) o("  y[base | 0x%llxll] = make_double2(ar%d, ai%d);\n", (u64)roff[r], r, r);
  }
  o("}\n\n");
}

// ---- run-time resolved NVRTC and driver entry points ------------------------------------------
struct Nvrtc {
  void *lib = nullptr;
  int (*create)(void **, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  int (*compile)(void *, int, const char *const *) = nullptr;
  int (*cubin_size)(void *, size_t *) = nullptr;
  int (*cubin)(void *, char *) = nullptr;
  int (*log_size)(void *, size_t *) = nullptr;
  int (*log)(void *, char *) = nullptr;
  int (*destroy)(void **) = nullptr;
  bool ok = false;
};

Nvrtc &nvrtc()
{
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
    for (const char *nm : names) {
      n.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (n.lib) break;
    }
    if (!n.lib) return;
    n.create = (decltype(n.create))dlsym(n.lib, "nvrtcCreateProgram");
    n.compile = (decltype(n.compile))dlsym(n.lib, "nvrtcCompileProgram");
    n.cubin_size = (decltype(n.cubin_size))dlsym(n.lib, "nvrtcGetCUBINSize");
    n.cubin = (decltype(n.cubin))dlsym(n.lib, "nvrtcGetCUBIN");
    n.log_size = (decltype(n.log_size))dlsym(n.lib, "nvrtcGetProgramLogSize");
    n.log = (decltype(n.log))dlsym(n.lib, "nvrtcGetProgramLog");
    n.destroy = (decltype(n.destroy))dlsym(n.lib, "nvrtcDestroyProgram");
    n.ok = n.create && n.compile && n.cubin_size && n.cubin && n.log_size && n.log && n.destroy;
  });
  return n;
}

struct Driver {
  CUresult (*moduleLoadData)(CUmodule *, const void *) = nullptr;
  CUresult (*moduleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
  CUresult (*moduleUnload)(CUmodule) = nullptr;
  CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **,
                           void **) = nullptr;
  bool ok = false;
};

Driver &driver()
{
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [](const char *sym) -> void * {
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult st;
      if (cudaGetDriverEntryPoint(sym, &fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
      }
      return fn;
    };
    d.moduleLoadData = (decltype(d.moduleLoadData))get("cuModuleLoadData");
    d.moduleGetFunction = (decltype(d.moduleGetFunction))get("cuModuleGetFunction");
    d.moduleUnload = (decltype(d.moduleUnload))get("cuModuleUnload");
    d.funcSetAttribute = (decltype(d.funcSetAttribute))get("cuFuncSetAttribute");
    d.launchKernel = (decltype(d.launchKernel))get("cuLaunchKernel");
    d.ok = d.moduleLoadData && d.moduleGetFunction && d.moduleUnload && d.funcSetAttribute && d.launchKernel;
  });
  return d;
}

}  // namespace

std::string generate(const std::vector<PassDesc> &passes, int)
{
  Out o;
  o("// generated by dynamite_b200 (csrc/jit.cu): operator-specialised window-tiled MatMult passes\n");
  o("typedef long long i64;\ntypedef unsigned long long u64;\ntypedef unsigned int u32;\n");
  o("__device__ __forceinline__ void cpa16(void *s, const void *g)\n{\n"
    "  asm volatile(\"cp.async.cg.shared.global [%%0], [%%1], 16;\\n\" ::\"r\"((u32)__cvta_generic_to_shared(s)), \"l\"(g) : \"memory\");\n}\n");
  o("__device__ __forceinline__ void cpa_wait() { asm volatile(\"cp.async.commit_group;\\ncp.async.wait_group 0;\\n\" ::: \"memory\"); }\n\n");
  for (size_t k = 0; k < passes.size(); ++k) gen_pass(o, passes[k], (int)k);
  return o.s;
}

std::vector<char> compile_cubin(const std::string &src, std::string &log)
{
  std::vector<char> cubin;
  Nvrtc &n = nvrtc();
  if (!n.ok) {
    log = "libnvrtc.so.12 could not be loaded";
    return cubin;
  }
  void *prog = nullptr;
  if (n.create(&prog, src.c_str(), "dnm_jit.cu", 0, nullptr, nullptr) != 0) {
    log = "nvrtcCreateProgram failed";
    return cubin;
  }
  const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--extra-device-vectorization"};
  const int rc = n.compile(prog, 4, opts);
  size_t ls = 0;
  n.log_size(prog, &ls);
  if (ls > 1) {
    log.resize(ls);
    n.log(prog, &log[0]);
  }
  if (rc == 0) {
    size_t cs = 0;
    if (n.cubin_size(prog, &cs) == 0 && cs > 0) {
      cubin.resize(cs);
      n.cubin(prog, cubin.data());
    }
  }
  n.destroy(&prog);
  return cubin;
}

Module::~Module()
{
  if (mod && driver().ok) driver().moduleUnload((CUmodule)mod);
}

Module *compile(const std::string &src, const std::vector<PassDesc> &passes, std::string &log, bool load)
{
  const std::vector<char> cubin = compile_cubin(src, log);
  if (cubin.empty() || !load) return nullptr;
  Driver &d = driver();
  if (!d.ok) {
    log += " [driver entry points unavailable]";
    return nullptr;
  }
  CUmodule mod = nullptr;
  CUresult rc = d.moduleLoadData(&mod, cubin.data());
  if (rc != CUDA_SUCCESS) {
    log += " [cuModuleLoadData failed: " + std::to_string((int)rc) + "]";
    return nullptr;
  }
  Module *m = new Module();
  m->mod = mod;
  for (size_t k = 0; k < passes.size(); ++k) {
    Kernel kn;
    const std::string name = "dnm_jit_p" + std::to_string(k);
    CUfunction f = nullptr;
    rc = d.moduleGetFunction(&f, mod, name.c_str());
    if (rc != CUDA_SUCCESS) {
      log += " [cuModuleGetFunction " + name + " failed]";
      delete m;
      return nullptr;
    }
    kn.func = f;
    kn.smem = sizeof(double2) << passes[k].T;
    kn.threads = 1 << (passes[k].T - 3);
    d.funcSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)kn.smem);
    d.funcSetAttribute(f, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100);
    m->kernels.push_back(kn);
  }
  return m;
}

void launch(const Kernel &k, unsigned long long ntiles, cudaStream_t stream, const cplx *x, cplx *y, const double *diag,
            long long rank_bits)
{
  void *args[] = {(void *)&x, (void *)&y, (void *)&diag, (void *)&rank_bits};
  const CUresult rc = driver().launchKernel((CUfunction)k.func, (unsigned)ntiles, 1, 1, (unsigned)k.threads, 1, 1,
                                            (unsigned)k.smem, (CUstream)stream, args, nullptr);
  DNM_REQUIRE(rc == CUDA_SUCCESS, DNM_ERR_CUDA, "cuLaunchKernel of a generated MatMult pass failed (%d)", (int)rc);
}

}  // namespace jit
}  // namespace dnm
