// See jit.h.  Generator + NVRTC / driver-API plumbing (both resolved at run time: the library must
// load on a machine without libcuda / libnvrtc, where only the host entry points are used).
#include "jit.h"

#include <cuda.h>
#include <dlfcn.h>

#include <algorithm>
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
    char buf[2048];
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

// "(var & m) << s | ..." for the runs of consecutive window positions W[lo..hi)
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

// tile number -> index bits outside the window
std::string outer_expr(const PassParams &P, const char *var)
{
  std::string e;
  char buf[200];
  if (P.n_seg >= 0) {
    for (int j = 0; j < P.n_seg; ++j) {
      const int sh = P.seg_shift[j];
      snprintf(buf, sizeof(buf), "((%s & 0x%llxull) %s %d)", var, (u64)P.seg_mask[j], sh >= 0 ? "<<" : ">>", sh >= 0 ? sh : -sh);
      if (!e.empty()) e += " | ";
      e += buf;
    }
  } else {
    for (int k = 0; k < P.n_outer; ++k) {
      snprintf(buf, sizeof(buf), "(((%s >> %d) & 1ull) << %d)", var, k, (int)P.outer_pos[k]);
      if (!e.empty()) e += " | ";
      e += buf;
    }
  }
  return e.empty() ? "0ull" : e;
}

// geometry of the thread <-> row mapping: thread `tid` owns the rows l = tid + r * NT
struct Geo {
  int T, R, LOGR, LOG_NT, NT;
  i64 roff[8];   // index offset of row group r
  i64 rowbits;   // index positions of the row coordinates
  Geo(const PassDesc &pd)
  {
    T = pd.T;
    R = pd.rows;
    LOGR = (R == 8) ? 3 : 2;
    LOG_NT = T - LOGR;
    NT = 1 << LOG_NT;
    rowbits = 0;
    for (int r = 0; r < R; ++r) {
      roff[r] = 0;
      for (int k = 0; k < LOGR; ++k)
        if ((r >> k) & 1) roff[r] |= (i64)1 << pd.W[LOG_NT + k];
    }
    for (int k = 0; k < LOGR; ++k) rowbits |= (i64)1 << pd.W[LOG_NT + k];
  }
};

// The tile as a box of a tensor over the local index space (element = double).
struct Box {
  int rank = -1;  // -1: not expressible, 0: the tile is contiguous
  u64 dims[MAX_TMA_RANK], strides[MAX_TMA_RANK];
  u32 box[MAX_TMA_RANK];
  int lo[MAX_TMA_RANK], len[MAX_TMA_RANK];  // bit range of the dimension
  bool window[MAX_TMA_RANK];
};

Box tile_box(const PassDesc &pd)
{
  Box b;
  const int n = pd.nloc, T = pd.T;
  std::vector<char> inw(n, 0);
  for (int w : pd.W) inw[w] = 1;
  bool contiguous = true;
  for (int j = 0; j < T; ++j) contiguous = contiguous && pd.W[j] == j;
  if (contiguous) {
    b.rank = 0;
    return b;
  }
  if (!inw[0]) return b;
  int rank = 0, pos = 0;
  while (pos < n) {
    int e = pos + 1;
    while (e < n && inw[e] == inw[pos]) ++e;
    if (inw[pos]) {
      int q = pos;
      while (q < e) {
        const int cap = (q == 0) ? 7 : 8;  // box extents are at most 256 elements (the innermost counts doubles)
        const int l = std::min(cap, e - q);
        if (rank == MAX_TMA_RANK) return b;
        b.lo[rank] = q;
        b.len[rank] = l;
        b.window[rank] = true;
        b.dims[rank] = (q == 0) ? (2ull << l) : (1ull << l);
        b.box[rank] = (u32)b.dims[rank];
        b.strides[rank] = 16ull << q;
        ++rank;
        q += l;
      }
    } else {
      if (rank == MAX_TMA_RANK) return b;
      b.lo[rank] = pos;
      b.len[rank] = e - pos;
      b.window[rank] = false;
      b.dims[rank] = 1ull << (e - pos);
      b.box[rank] = 1;
      b.strides[rank] = 16ull << pos;
      ++rank;
    }
    pos = e;
  }
  if (rank < 2) return b;
  b.rank = rank;
  return b;
}

struct Group {
  u32 lam, s1w, sBw;
  i64 so1, soB;
  double c1, c2;
  bool imag, far;
  u64 farmask;
  int peer;
};

// the real group g and the imaginary group g+1 of one mask, each with a single sign mask (X and Y
// fields on one site): one operand fetch serves both
bool merge_at(const SmallTables &S, int g, int ngroups)
{
  if (g + 1 >= ngroups) return false;
  const bool far0 = ((S.gd[g].w >> 17) & 1u) != 0, far1 = ((S.gd[g + 1].w >> 17) & 1u) != 0;
  return S.lam[g] == S.lam[g + 1] && far0 == far1 && S.far[g] == S.far[g + 1] && S.peer[g] == S.peer[g + 1] &&
         !(S.kp[g] & 1) && (S.kp[g + 1] & 1) && S.cf[2 * g + 1] == 0.0 && S.cf[2 * g + 3] == 0.0 && S.cf[2 * g] != 0.0 &&
         S.cf[2 * g + 2] != 0.0;
}

// operand buffers a classic kernel needs for its folded remote masks
int remote_groups(const PassDesc &pd)
{
  int n = 0;
  const SmallTables &S = *pd.st;
  for (int g = 0; g < pd.p->ngroups; ++g) {
    const bool far = ((S.gd[g].w >> 17) & 1u) != 0;
    const double A = S.cf[2 * g] + S.cf[2 * g + 1], Bc = S.cf[2 * g] - S.cf[2 * g + 1];
    const bool merged = merge_at(S, g, pd.p->ngroups);
    if (far && S.peer[g] != 0 && (merged || !(A == 0.0 && Bc == 0.0))) ++n;
    if (merged) ++g;
  }
  return n;
}

// acc[r] += D_g(row) * x[row ^ mask_g] for every group of the pass.  `tile` = the staged tile,
// `x` = the global operand (far groups), `tid`, `og` (tile-uniform index bits) and `base` (this
// thread's index without the row offsets) are in scope.
// remote_mode 0: groups that read another rank's shard load it straight into registers;
// 1: emit ONLY the asynchronous copies (cp.async, peer memory -> shared buffer k of the pass) of the
//    remote groups, under the same activity conditions as the arithmetic; 2: every group, the remote
//    ones reading their shared buffer (NVLink latency is then paid once per tile, behind the local work)
// select 0: every group; 1: only the FAR groups (operands from global memory: they do not need the
// staged tile); 2: only the groups served from the tile
void gen_groups(Out &o, const PassDesc &pd, const Geo &g_, int remote_mode = 0, int select = 0)
{
  int remote_index = -1;
  const PassParams &P = *pd.p;
  const SmallTables &S = *pd.st;
  const int R = g_.R, LOG_NT = g_.LOG_NT, NT = g_.NT;
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
    G.far = ((S.gd[g].w >> 17) & 1u) != 0;
    G.peer = G.far ? (int)S.peer[g] : 0;
    if (G.c2 == 0.0) {  // one sign mask only
      G.sBw = 0;
      G.soB = 0;
    }
    const u32 lamlo = G.lam & (u32)(NT - 1);
    const int HI = (int)(G.lam >> LOG_NT);
    const u32 s1t = G.s1w & (u32)(NT - 1), sBt = G.sBw & (u32)(NT - 1);
    const u32 s1r = G.s1w >> LOG_NT, sBr = G.sBw >> LOG_NT;
    if (merge_at(S, g, P.ngroups)) {
      // X and Y field on one site: D = +-cr + i (+-ci), one fetch
      const bool remote_m = G.far && G.peer != 0;
      if (remote_m) ++remote_index;
      const int gi = g + 1;
      ++g;
      if (remote_mode == 1 && !remote_m) continue;
      if ((select == 1 && !G.far) || (select == 2 && G.far)) continue;
      const bool issue_m = remote_mode == 1, staged_m = remote_mode == 2 && remote_m;
      if (staged_m && remote_index == 0) o("    cpa_wait_all();  // the staged remote operands (each thread reads back only what it copied)\n");
      const double cr = S.cf[2 * (gi - 1)], ci = S.cf[2 * gi];
      const u32 sit = S.sw[2 * gi] & (u32)(NT - 1), sir = S.sw[2 * gi] >> LOG_NT;
      if (remote_m && pd.filter_bit >= 0)
        o("    if (((og >> %d) & 1) == %d)  // this pass serves the folded remote masks on half of the tiles\n", pd.filter_bit, pd.filter_val);
      o("    {  // groups %d+%d: mask window 0x%x%s, real %s + imaginary %s, one fetch\n", gi - 1, gi, G.lam, G.far ? " FAR" : "",
        hexd(cr).c_str(), hexd(ci).c_str());
      auto expo_m = [&](i64 so, u32 st) -> std::string {
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
      const std::string er = expo_m(G.so1, s1t), ei = expo_m(S.so[2 * gi], sit);
      if (!issue_m) {
        if (er.empty()) o("      const double cr = %s;\n", hexd(cr).c_str());
        else o("      const double cr = (%s) ? %s : %s;\n", er.c_str(), hexd(-cr).c_str(), hexd(cr).c_str());
        if (ei.empty()) o("      const double ci = %s;\n", hexd(ci).c_str());
        else o("      const double ci = (%s) ? %s : %s;\n", ei.c_str(), hexd(-ci).c_str(), hexd(ci).c_str());
      }
      if (staged_m) o("      const double2 *src = tile + %d + tid;  // staged copy of the operand from rank ^ %d\n", (1 + remote_index) << g_.T, G.peer);
      else if (G.far && G.peer) o("      const double2 *src = xs.p[%d] + (base ^ 0x%llxll);  // rank ^ %d over NVLink\n", G.peer, (u64)((i64)G.farmask & ~g_.rowbits), G.peer);
      else if (G.far) o("      const double2 *src = x + (base ^ 0x%llxll);\n", (u64)((i64)G.farmask & ~g_.rowbits));
      else o("      const double2 *src = tile + (tid ^ 0x%xu);\n", lamlo);
      for (int h = 0; h < R; h += 4) {
        o("      {\n");
        for (int r = h; r < std::min(R, h + 4); ++r) {
          if (issue_m) {
            o("        cpa16(&tile[%d + tid + %d], src + 0x%llxll);\n", (1 + remote_index) << g_.T, r * NT, (u64)g_.roff[r ^ HI]);
            continue;
          }
          char opnd[128];
          if (staged_m) snprintf(opnd, sizeof(opnd), "src[%d]", r * NT);
          else if (G.far) snprintf(opnd, sizeof(opnd), "__ldcg(src + 0x%llxll)", (u64)g_.roff[r ^ HI]);
          else snprintf(opnd, sizeof(opnd), "src[%d]", (r ^ HI) * NT);
          o("        const double2 v%d = %s;\n", r, opnd);
        }
        if (!issue_m)
          for (int r = h; r < std::min(R, h + 4); ++r) {
            const bool nr = par32(s1r & (u32)r) != 0, ni = par32(sir & (u32)r) != 0;
            o("        ar%d = fma(%scr, v%d.x, ar%d); ai%d = fma(%scr, v%d.y, ai%d); ar%d = fma(%sci, v%d.y, ar%d); ai%d = fma(%sci, v%d.x, ai%d);\n",
              r, nr ? "-" : "", r, r, r, nr ? "-" : "", r, r, r, ni ? "" : "-", r, r, r, ni ? "-" : "", r, r);
          }
        o("      }\n");
      }
      o("    }\n");
      continue;
    }
    // D(row) = (-1)^(e1 ^ k1(r)) * ((eB ^ kB(r)) ? c1 - c2 : c1 + c2): e = tile and thread part of the
    // sign exponents (run time), k = row part (known here)
    const double A = G.c1 + G.c2, Bc = G.c1 - G.c2;
    if (A == 0.0 && Bc == 0.0) continue;
    const bool remote = G.far && G.peer != 0;
    if (remote) ++remote_index;
    if (remote_mode == 1 && !remote) continue;
    if ((select == 1 && !G.far) || (select == 2 && G.far)) continue;
    const bool issue = remote_mode == 1, staged = remote_mode == 2 && remote;
    if (staged && remote_index == 0) o("    cpa_wait_all();  // the staged remote operands (each thread reads back only what it copied)\n");

    if (remote && pd.filter_bit >= 0)
      o("    if (((og >> %d) & 1) == %d)  // this pass serves the folded remote masks on half of the tiles\n", pd.filter_bit, pd.filter_val);
    o("    {  // group %d: mask window 0x%x%s%s, c1=%s c2=%s\n", g, G.lam, G.imag ? " imag" : "", G.far ? " FAR" : "",
      hexd(G.c1).c_str(), hexd(G.c2).c_str());
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
    if (!e1.empty()) o("      const int e1 = %s;\n", e1.c_str());
    if (!eB.empty()) o("      const int eB = %s;\n", eB.c_str());
    if (staged) o("      const double2 *src = tile + %d + tid;  // staged copy of the operand from rank ^ %d\n", (1 + remote_index) << g_.T, G.peer);
    else if (G.far && G.peer) o("      const double2 *src = xs.p[%d] + (base ^ 0x%llxll);  // rank ^ %d over NVLink\n", G.peer, (u64)((i64)G.farmask & ~g_.rowbits), G.peer);
    else if (G.far) o("      const double2 *src = x + (base ^ 0x%llxll);\n", (u64)((i64)G.farmask & ~g_.rowbits));
    else o("      const double2 *src = tile + (tid ^ 0x%xu);\n", lamlo);
    auto operand = [&](int r) -> std::string {
      char buf[128];
      if (staged) snprintf(buf, sizeof(buf), "src[%d]", r * NT);
      else if (G.far) snprintf(buf, sizeof(buf), "__ldcg(src + 0x%llxll)", (u64)g_.roff[r ^ HI]);
      else snprintf(buf, sizeof(buf), "src[%d]", (r ^ HI) * NT);
      return buf;
    };
    auto emit_rows = [&](const std::vector<int> &rows, const char *cv, const char *indent) {
      const size_t chunk = (G.far && getenv("DNM_JIT_FAR_CHUNK")) ? (size_t)std::max(1, atoi(getenv("DNM_JIT_FAR_CHUNK"))) : 4;
      if (issue) {
        for (int r : rows)
          o("%scpa16(&tile[%d + tid + %d], src + 0x%llxll);\n", indent, (1 + remote_index) << g_.T, r * NT, (u64)g_.roff[r ^ HI]);
        return;
      }
      for (size_t h = 0; h < rows.size(); h += chunk) {
        const size_t e = std::min(rows.size(), h + chunk);
        o("%s{\n", indent);
        for (size_t k = h; k < e; ++k) o("%s  const double2 v%d = %s;\n", indent, rows[k], operand(rows[k]).c_str());
        for (size_t k = h; k < e; ++k) {
          const int r = rows[k];
          const bool neg = par32(s1r & (u32)r) != 0;
          if (G.imag)
            o("%s  ar%d = fma(%s%s, v%d.y, ar%d); ai%d = fma(%s%s, v%d.x, ai%d);\n", indent, r, neg ? "" : "-", cv, r, r, r,
              neg ? "-" : "", cv, r, r);
          else
            o("%s  ar%d = fma(%s%s, v%d.x, ar%d); ai%d = fma(%s%s, v%d.y, ai%d);\n", indent, r, neg ? "-" : "", cv, r, r, r,
              neg ? "-" : "", cv, r, r);
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
      if (e1.empty()) o("      const double c = %s;\n", hexd(G.c1).c_str());
      else o("      const double c = e1 ? %s : %s;\n", hexd(-G.c1).c_str(), hexd(G.c1).c_str());
      emit_rows(all, "c", "      ");
    } else if (A == 0.0 || Bc == 0.0) {
      // rows with (eB ^ kB(r)) == act carry +-V, the others vanish (XX+YY: half of the rows)
      const int act = (A == 0.0) ? 1 : 0;
      const double V = (A == 0.0) ? Bc : A;
      if (e1.empty()) o("      const double c = %s;\n", hexd(V).c_str());
      else o("      const double c = e1 ? %s : %s;\n", hexd(-V).c_str(), hexd(V).c_str());
      if (eB.empty()) {
        emit_rows(act == 0 ? rows0 : rows1, "c", "      ");
      } else {
        o("      if (eB == %d) {\n", act);
        emit_rows(rows0, "c", "        ");
        if (!rows1.empty()) {
          o("      } else {\n");
          emit_rows(rows1, "c", "        ");
        }
        o("      }\n");
      }
    } else {
      if (eB.empty()) o("      double c0 = %s, c1v = %s;\n", hexd(A).c_str(), hexd(Bc).c_str());
      else
        o("      double c0 = eB ? %s : %s, c1v = eB ? %s : %s;\n", hexd(Bc).c_str(), hexd(A).c_str(), hexd(A).c_str(),
          hexd(Bc).c_str());
      if (!e1.empty()) o("      if (e1) { c0 = -c0; c1v = -c1v; }\n");
      emit_rows(rows0, "c0", "      ");
      if (!rows1.empty()) emit_rows(rows1, "c1v", "      ");
    }
    o("    }\n");
  }
}

// streaming (evict-first) accesses for data that is touched once per pass (experiment knob DNM_JIT_HINTS=0)
bool hints_on() { return !(getenv("DNM_JIT_HINTS") && atoi(getenv("DNM_JIT_HINTS")) == 0); }
const char *hint_ld() { return hints_on() ? "__ldcs" : "__ldg"; }
const char *hint_st() { return hints_on() ? "__stcs" : "st_plain"; }

std::string acc_decl(int R)
{
  std::string s = "    double ";
  for (int r = 0; r < R; ++r) {
    char buf[32];
    snprintf(buf, sizeof(buf), "%sar%d, ai%d", r ? ", " : "", r, r);
    s += buf;
  }
  return s + ";\n";
}

void emit_tma_helpers(Out &o, const PassDesc &pd, int index, const Box &bx, bool reduce);

// one tile per CTA; the tile is staged by cp.async (any window) or by TMA (window = a tensor box)
void gen_classic(Out &o, const PassDesc &pd, int index)
{
  const PassParams &P = *pd.p;
  const Geo g(pd);
  const int R = g.R, NT = g.NT;
  int minb = std::max(1, std::min(8, 65536 / (NT * 64)));
  if (getenv("DNM_JIT_MINB")) minb = std::max(1, std::min(minb, atoi(getenv("DNM_JIT_MINB"))));
  const int nrem = pd.stage_remote ? remote_groups(pd) : 0;
  const bool tma = pd.tma_stage;
  const Box bx = tile_box(pd);
  // (an accumulating pass whose window is the contiguous tile has no tensor box to reduce into: it keeps the
  // read-modify-write epilogue -- found by the emulation fuzz, such a pass used to be emitted with a 0-d tensor
  // reduce that NVRTC rejects, which cost the whole plan its generated kernels)
  const bool tma_reduce = tma && pd.tma_reduce && P.accumulate == 1 && bx.rank > 0;
  o("// ---- pass %d (classic%s): T=%d R=%d, %d groups, %s, far_bits=%d\n", index, tma ? ", TMA staging" : "", g.T, R, P.ngroups,
    P.accumulate ? "accumulate" : "write", P.far_bits);
  if (tma) emit_tma_helpers(o, pd, index, bx, tma_reduce);
  o("extern \"C\" __global__ void __launch_bounds__(%d, %d)\n", NT, minb);
  if (tma)
    o("dnm_jit_p%d(const __grid_constant__ TMap tmx, const __grid_constant__ TMap tmy, const double2 *__restrict__ x,\n"
      "           double2 *__restrict__ y, const double *__restrict__ diag, i64 rank_bits, u64 ntiles,\n"
      "           const __grid_constant__ Peers xs)\n{\n",
      index);
  else
    o("dnm_jit_p%d(const double2 *__restrict__ x, double2 *__restrict__ y, const double *__restrict__ diag, i64 rank_bits,\n"
      "           const __grid_constant__ Peers xs)\n{\n",
      index);
  if (tma) {
    o("  extern __shared__ __align__(1024) unsigned char smem[];\n");
    o("  double2 *tile = reinterpret_cast<double2 *>(smem);\n");
    o("  u64 *bar = reinterpret_cast<u64 *>(smem + %zu);\n", ((size_t)16 << g.T) * (size_t)(1 + nrem));
  } else {
    o("  extern __shared__ double2 tile[];\n");
  }
  o("  const u32 tid = threadIdx.x;\n");
  o("  const u64 tb = blockIdx.x;\n");
  if (tma) {
    // one elected thread stages the whole tile: cp.async.bulk[.tensor] completes on the mbarrier
    o("  if (tid == 0) {\n    mbar_init(bar, 1);\n    asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n  }\n");
    o("  __syncthreads();\n");
    o("  if (tid == 0) load_p%d(&tmx, x, tb, tile, bar);\n", index);
  }
  o("  {\n");
  o("    const i64 outer = (i64)(%s);\n", outer_expr(P, "tb").c_str());
  o("    const i64 og = outer | rank_bits;  // index bits shared by the tile (signs)\n");
  o("    const i64 base = outer | %s;\n", deposit_expr("tid", pd.W, 0, g.LOG_NT).c_str());
  if (!tma)
    for (int r = 0; r < R; ++r) o("    cpa16(&tile[tid + %d], x + (base | 0x%llxll));\n", r * NT, (u64)g.roff[r]);
  if (nrem) {
    // operands of the folded remote masks: asynchronous copies from the partners' shards over NVLink into
    // their own buffers, in flight while the local masks are evaluated
    o("    cpa_commit();\n");
    gen_groups(o, pd, g, 1);
    o("    cpa_commit();\n");
  }
  // the cached diagonal is in flight together with the tile
  for (int r = 0; r < R; ++r) o("    double dg%d = 0.0;\n", r);
  o("    if (diag != nullptr) {\n");
  for (int r = 0; r < R; ++r) o("      dg%d = %s(diag + (base | 0x%llxll));\n", r, hint_ld(), (u64)g.roff[r]);
  o("    }\n");
  // experiment (DNM_JIT_FAR_FIRST=1): measured slower, 20.6 against 18.1 ms at L=30 MBL -- the far loads then
  // miss the L2 more often because they run ahead of the partner tiles' own staging
  const bool far_first = !nrem && getenv("DNM_JIT_FAR_FIRST") && atoi(getenv("DNM_JIT_FAR_FIRST")) != 0;
  if (far_first) {
    // the FAR groups read global memory only: they run while the tile is still in flight
    o("%s", acc_decl(R).c_str());
    for (int r = 0; r < R; ++r) o("    ar%d = 0.0; ai%d = 0.0;\n", r, r);
    o("    cpa_commit();\n");
    gen_groups(o, pd, g, 0, 1);
    if (tma) o("    mbar_wait(bar, 0);\n");
    else o("    cpa_wait_all();\n    __syncthreads();\n");
    for (int r = 0; r < R; ++r)
      o("    { const double2 v = tile[tid + %d]; ar%d = fma(dg%d, v.x, ar%d); ai%d = fma(dg%d, v.y, ai%d); }\n", r * NT, r, r, r, r, r, r);
    gen_groups(o, pd, g, 0, 2);
  } else {
    if (tma) o("    mbar_wait(bar, 0);\n");
    else if (nrem) o("    cpa_wait_first();\n    __syncthreads();\n");
    else o("    cpa_wait();\n    __syncthreads();\n");
    o("%s", acc_decl(R).c_str());
    for (int r = 0; r < R; ++r)
      o("    { const double2 v = tile[tid + %d]; ar%d = dg%d * v.x; ai%d = dg%d * v.y; }\n", r * NT, r, r, r, r);
    gen_groups(o, pd, g, nrem ? 2 : 0);
  }
  if (tma_reduce) {
    // the result tile goes through shared memory and is ADDED into y by the TMA unit: the old y never
    // enters the SM
    o("    __syncthreads();  // nobody reads the operand tile any more\n");
    for (int r = 0; r < R; ++r) o("    tile[tid + %d] = make_double2(ar%d, ai%d);\n", r * NT, r, r);
    o("    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n");
    o("    __syncthreads();\n");
    o("    if (tid == 0) {\n      reduce_p%d(&tmy, tb, tile);\n", index);
    o("      asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");  // shared memory must outlive the read\n    }\n");
  } else if (P.accumulate == 1) {
    o("    __syncthreads();\n");
    for (int r = 0; r < R; ++r) o("    cpa16(&tile[tid + %d], y + (base | 0x%llxll));\n", r * NT, (u64)g.roff[r]);
    o("    cpa_wait();\n");
    for (int r = 0; r < R; ++r)
      o("    { const double2 old = tile[tid + %d]; y[base | 0x%llxll] = make_double2(ar%d + old.x, ai%d + old.y); }\n", r * NT,
        (u64)g.roff[r], r, r);
  } else {
    for (int r = 0; r < R; ++r) o("    %s(y + (base | 0x%llxll), make_double2(ar%d, ai%d));\n", hint_st(), (u64)g.roff[r], r, r);
  }
  o("  }\n}\n\n");
}

int pipelined_ctas(const PassDesc &pd)
{
  const Geo g(pd);
  const size_t bytes = (size_t)(pd.nbuf + (pd.p->accumulate == 1 ? 1 : 0)) * ((size_t)16 << pd.T) + 1024 + 1024;
  int by_smem = (int)(232448 / bytes);
  int by_threads = 2048 / g.NT;
  return std::max(1, std::min(std::min(by_smem, by_threads), 4));
}

// Host emulation (tests only, dnm_jit_set_host_emulation): the SAME kernel bodies are emitted, but the
// prelude defines the CUDA vocabulary for a C++ compiler (one OS thread per CUDA thread, a pthread
// barrier for __syncthreads) and the TMA helpers become plain box copies / box additions.  The CUDA
// output is not touched by this mode.
bool g_host_emulation = false;

// nested loops over the box of `bx` (innermost dimension fastest): the statement sees `gi` = element
// offset in the global tensor (doubles) and `si` = running offset in shared memory (doubles)
void emit_host_box_loops(Out &o, const Box &bx, const char *stmt)
{
  o("  long long si = 0;\n");
  for (int d = bx.rank - 1; d >= 0; --d)
    o("  %*sfor (long long i%d = 0; i%d < %u; ++i%d)\n", 2 * (bx.rank - 1 - d), "", d, d, bx.box[d], d);
  std::string gi = "0";
  for (int d = 0; d < bx.rank; ++d) {
    char buf[96];
    snprintf(buf, sizeof(buf), " + (c[%d] + i%d) * %lldll", d, d, (long long)(d == 0 ? 1 : bx.strides[d] / 8));
    gi += buf;
  }
  o("  %*s{ const long long gi = %s; %s ++si; }\n", 2 * bx.rank, "", gi.c_str(), stmt);
}

// outer_p<k>(tile) = index bits outside the window; load_p<k> = TMA load of a tile into shared memory
// (signals `bar`); reduce_p<k> = cp.reduce.async.bulk.tensor add of a result tile into y
void emit_tma_helpers(Out &o, const PassDesc &pd, int index, const Box &bx, bool reduce)
{
  const PassParams &P = *pd.p;
  const u32 tile_bytes = 16u << pd.T;
  o("__device__ __forceinline__ i64 outer_p%d(u64 tb) { return (i64)(%s); }\n", index, outer_expr(P, "tb").c_str());
  auto coords = [&]() -> std::string {
    std::string c;
    for (int d = 0; d < bx.rank; ++d) {
      char buf[96];
      if (bx.window[d]) snprintf(buf, sizeof(buf), "0");
      else snprintf(buf, sizeof(buf), "(int)(((u64)outer >> %d) & 0x%llxull)", bx.lo[d], (1ull << bx.len[d]) - 1ull);
      if (d) c += ", ";
      c += buf;
    }
    return c;
  };
  // issue the TMA load of tile `tb` into ring buffer `dst`
  o("__device__ __forceinline__ void load_p%d(const TMap *tmx, const double2 *x, u64 tb, void *dst, u64 *bar)\n{\n", index);
  o("  const i64 outer = outer_p%d(tb);\n", index);
  o("  mbar_expect_tx(bar, %uu);\n", tile_bytes);
  if (g_host_emulation) {
    if (bx.rank == 0) {
      o("  memcpy(dst, x + outer, %uu);\n", tile_bytes);
    } else {
      o("  const int c[%d] = {%s};\n", bx.rank, coords().c_str());
      o("  const double *g = tmx->base;\n  double *sm = (double *)dst;\n");
      emit_host_box_loops(o, bx, "sm[si] = g[gi];");
    }
    o("  emu_bar_complete(bar);\n}\n");
  } else if (bx.rank == 0) {
    o("  asm volatile(\"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%%0], [%%1], %%2, [%%3];\" "
      "::\"r\"(smem_u32(dst)), \"l\"(x + outer), \"r\"(%uu), \"r\"(smem_u32(bar)) : \"memory\");\n",
      tile_bytes);
  } else {
    std::string ops, cl;
    for (int d = 0; d < bx.rank; ++d) {
      char buf[32];
      snprintf(buf, sizeof(buf), "%s%%%d", d ? ", " : "", 3 + d);
      ops += buf;
      snprintf(buf, sizeof(buf), ", \"r\"(c[%d])", d);
      cl += buf;
    }
    o("  const int c[%d] = {%s};\n", bx.rank, coords().c_str());
    o("  asm volatile(\"cp.async.bulk.tensor.%dd.shared::cluster.global.mbarrier::complete_tx::bytes [%%0], [%%1, {%s}], [%%2];\" "
      "::\"r\"(smem_u32(dst)), \"l\"(tmx), \"r\"(smem_u32(bar))%s : \"memory\");\n",
      bx.rank, ops.c_str(), cl.c_str());
  }
  if (!g_host_emulation) o("}\n");
  if (reduce && g_host_emulation) {
    o("__device__ __forceinline__ void reduce_p%d(const TMap *tmy, u64 tb, const void *src)\n{\n", index);
    o("  const i64 outer = outer_p%d(tb);\n", index);
    o("  const int c[%d] = {%s};\n", bx.rank, coords().c_str());
    o("  double *g = const_cast<double *>(tmy->base);\n  const double *sm = (const double *)src;\n");
    emit_host_box_loops(o, bx, "g[gi] += sm[si];");
    o("}\n");
  } else if (reduce) {
    o("__device__ __forceinline__ void reduce_p%d(const TMap *tmy, u64 tb, const void *src)\n{\n", index);
    o("  const i64 outer = outer_p%d(tb);\n", index);
    std::string ops, cl;
    for (int d = 0; d < bx.rank; ++d) {
      char buf[32];
      snprintf(buf, sizeof(buf), "%s%%%d", d ? ", " : "", 2 + d);
      ops += buf;
      snprintf(buf, sizeof(buf), ", \"r\"(c[%d])", d);
      cl += buf;
    }
    o("  const int c[%d] = {%s};\n", bx.rank, coords().c_str());
    o("  asm volatile(\"cp.reduce.async.bulk.tensor.%dd.global.shared::cta.add.bulk_group [%%0, {%s}], [%%1];\" "
      "::\"l\"(tmy), \"r\"(smem_u32(src))%s : \"memory\");\n",
      bx.rank, ops.c_str(), cl.c_str());
    o("  asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");\n}\n");
  }

}

// persistent CTAs, TMA ring, reduce-add epilogue
void gen_pipelined(Out &o, const PassDesc &pd, int index)
{
  const PassParams &P = *pd.p;
  const Geo g(pd);
  const Box bx = tile_box(pd);
  const int R = g.R, NT = g.NT, T = g.T;
  const bool reduce = P.accumulate == 1;
  const int ctas = pipelined_ctas(pd);
  const u32 tile_bytes = 16u << T;
  o("// ---- pass %d (pipelined): T=%d R=%d nbuf=%d, %d groups, %s, far_bits=%d, TMA rank %d\n", index, T, R, pd.nbuf, P.ngroups,
    reduce ? "accumulate (reduce-add)" : "write", P.far_bits, bx.rank);
  emit_tma_helpers(o, pd, index, bx, reduce);
  o("extern \"C\" __global__ void __launch_bounds__(%d, %d)\n", NT, ctas);
  o("dnm_jit_p%d(const __grid_constant__ TMap tmx, const __grid_constant__ TMap tmy, const double2 *__restrict__ x,\n"
    "           double2 *__restrict__ y, const double *__restrict__ diag, i64 rank_bits, u64 ntiles,\n"
    "           const __grid_constant__ Peers xs)\n{\n",
    index);
  o("  extern __shared__ __align__(1024) unsigned char smem[];\n");
  o("  double2 *ring = reinterpret_cast<double2 *>(smem);\n");
  if (reduce) o("  double2 *outb = ring + %d;\n", pd.nbuf << T);
  o("  u64 *full = reinterpret_cast<u64 *>(smem + %zu);\n", (size_t)(pd.nbuf + (reduce ? 1 : 0)) * tile_bytes);
  o("  const u32 tid = threadIdx.x;\n");
  o("  const u64 stride = gridDim.x;\n");
  o("  if (tid == 0) {\n");
  for (int b = 0; b < pd.nbuf; ++b) o("    mbar_init(&full[%d], 1);\n", b);
  o("    asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n");
  o("  }\n  __syncthreads();\n");
  o("  if (tid == 0) {\n    u64 tt = blockIdx.x;\n");
  o("    for (int b = 0; b < %d && tt < ntiles; ++b, tt += stride) load_p%d(&tmx, x, tt, ring + ((size_t)b << %d), &full[b]);\n",
    pd.nbuf, index, T);
  o("  }\n");
  o("  const i64 toff = %s;\n", deposit_expr("tid", pd.W, 0, g.LOG_NT).c_str());
  // software-pipelined diagonal: the values of the next tile are fetched while this one is evaluated
  for (int r = 0; r < R; ++r) o("  double dn%d = 0.0;\n", r);
  o("  if (diag != nullptr && blockIdx.x < ntiles) {\n    const i64 nb = outer_p%d(blockIdx.x) | toff;\n", index);
  for (int r = 0; r < R; ++r) o("    dn%d = %s(diag + (nb | 0x%llxll));\n", r, hint_ld(), (u64)g.roff[r]);
  o("  }\n");
  o("  int b = 0;\n  u32 phase = 0;\n");
  o("  for (u64 tb = blockIdx.x; tb < ntiles; tb += stride) {\n");
  o("    const i64 outer = outer_p%d(tb);\n", index);
  o("    const i64 og = outer | rank_bits;  // index bits shared by the tile (signs)\n");
  o("    const i64 base = outer | toff;\n");
  for (int r = 0; r < R; ++r) o("    const double dg%d = dn%d;\n", r, r);
  o("    if (diag != nullptr && tb + stride < ntiles) {\n      const i64 nb = outer_p%d(tb + stride) | toff;\n", index);
  for (int r = 0; r < R; ++r) o("      dn%d = %s(diag + (nb | 0x%llxll));\n", r, hint_ld(), (u64)g.roff[r]);
  o("    }\n");
  o("    mbar_wait(&full[b], phase);\n");
  o("    const double2 *tile = ring + ((size_t)b << %d);\n", T);
  o("%s", acc_decl(R).c_str());
  for (int r = 0; r < R; ++r)
    o("    { const double2 v = tile[tid + %d]; ar%d = dg%d * v.x; ai%d = dg%d * v.y; }\n", r * NT, r, r, r, r);
  gen_groups(o, pd, g);
  if (reduce) {
    o("    if (tid == 0) asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");  // the staging buffer is free again\n");
    o("    __syncthreads();\n");
    for (int r = 0; r < R; ++r) o("    outb[tid + %d] = make_double2(ar%d, ai%d);\n", r * NT, r, r);
    o("    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n");
    o("    __syncthreads();\n");
    o("    if (tid == 0) {\n      reduce_p%d(&tmy, tb, outb);\n", index);
  } else {
    for (int r = 0; r < R; ++r) o("    %s(y + (base | 0x%llxll), make_double2(ar%d, ai%d));\n", hint_st(), (u64)g.roff[r], r, r);
    o("    __syncthreads();  // nobody reads ring[b] any more\n");
    o("    if (tid == 0) {\n");
  }
  o("      const u64 tn = tb + %dull * stride;\n", pd.nbuf);
  o("      if (tn < ntiles) load_p%d(&tmx, x, tn, ring + ((size_t)b << %d), &full[b]);\n    }\n", index, T);
  o("    if (++b == %d) { b = 0; phase ^= 1u; }\n", pd.nbuf);
  o("  }\n");
  if (reduce) o("  if (tid == 0) asm volatile(\"cp.async.bulk.wait_group 0;\" ::: \"memory\");\n");
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

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Driver {
  CUresult (*moduleLoadData)(CUmodule *, const void *) = nullptr;
  CUresult (*moduleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
  CUresult (*moduleUnload)(CUmodule) = nullptr;
  CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **,
                           void **) = nullptr;
  EncodeFn encode = nullptr;
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
    d.encode = (EncodeFn)get("cuTensorMapEncodeTiled");
    d.ok = d.moduleLoadData && d.moduleGetFunction && d.moduleUnload && d.funcSetAttribute && d.launchKernel && d.encode;
  });
  return d;
}

const char *PRELUDE =
    "// generated by dynamite_b200 (csrc/jit.cu): operator-specialised window-tiled MatMult passes\n"
    "typedef long long i64;\ntypedef unsigned long long u64;\ntypedef unsigned int u32;\n"
    "struct __align__(64) TMap { u64 opaque[16]; };\n"
    "struct Peers { const double2 *p[16]; };  // the input vector on rank (this ^ h)\n"
    "__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }\n"
    "__device__ __forceinline__ void cpa16(void *s, const void *g)\n{\n"
    "  asm volatile(\"cp.async.cg.shared.global [%0], [%1], 16;\\n\" ::\"r\"(smem_u32(s)), \"l\"(g) : \"memory\");\n}\n"
    "__device__ __forceinline__ void cpa_wait() { asm volatile(\"cp.async.commit_group;\\ncp.async.wait_group 0;\\n\" ::: \"memory\"); }\n"
    "__device__ __forceinline__ void cpa_commit() { asm volatile(\"cp.async.commit_group;\\n\" ::: \"memory\"); }\n"
    "__device__ __forceinline__ void cpa_wait_all() { asm volatile(\"cp.async.wait_group 0;\\n\" ::: \"memory\"); }\n"
    "__device__ __forceinline__ void cpa_wait_first() { asm volatile(\"cp.async.wait_group 1;\\n\" ::: \"memory\"); }\n"
    "__device__ __forceinline__ void st_plain(double2 *p, double2 v) { *p = v; }\n"
    "__device__ __forceinline__ void mbar_init(u64 *bar, int count)\n{\n"
    "  asm volatile(\"mbarrier.init.shared::cta.b64 [%0], %1;\" ::\"r\"(smem_u32(bar)), \"r\"(count));\n}\n"
    "__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)\n{\n"
    "  asm volatile(\"mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\" ::\"r\"(smem_u32(bar)), \"r\"(bytes) : \"memory\");\n}\n"
    "__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)\n{\n"
    "  asm volatile(\"{\\n.reg .pred p;\\nWAIT_%=:\\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\\n@p bra DONE_%=;\\nbra "
    "WAIT_%=;\\nDONE_%=:\\n}\\n\" ::\"r\"(smem_u32(bar)), \"r\"(parity) : \"memory\");\n}\n\n";

}  // namespace

bool tma_eligible(const PassDesc &pd) { return tile_box(pd).rank >= 0; }
bool tma_reducible(const PassDesc &pd) { return tile_box(pd).rank > 0; }

void set_host_emulation(bool on) { g_host_emulation = on; }
bool host_emulation() { return g_host_emulation; }

namespace {
const char *HOST_PRELUDE =
    "// generated by dynamite_b200 (csrc/jit.cu) in HOST EMULATION mode: the kernel bodies below are the ones\n"
    "// that run on the GPU; this prelude maps the CUDA vocabulary to C++ (tests/test_jit_emulation.py)\n"
    "#include <cmath>\n#include <cstdlib>\n#include <cstring>\n#include <pthread.h>\n#include <sched.h>\n"
    "typedef long long i64;\ntypedef unsigned long long u64;\ntypedef unsigned int u32;\n"
    "struct double2 { double x, y; };\n"
    "static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }\n"
    "struct TMap { const double *base; };\n"
    "struct Peers { const double2 *p[16]; };\n"
    "struct EmuIdx { unsigned x; };\n"
    "static thread_local EmuIdx threadIdx, blockIdx, gridDim;\n"
    "static unsigned char *emu_smem_base;\nstatic pthread_barrier_t emu_barrier;\n"
    "#define __global__\n#define __device__\n#define __forceinline__ inline\n#define __launch_bounds__(a, b)\n"
    "#define __grid_constant__\n#define __align__(n)\n"
    "static inline void __syncthreads() { pthread_barrier_wait(&emu_barrier); }\n"
    "static inline int __popc(u32 v) { return __builtin_popcount(v); }\n"
    "static inline int __popcll(u64 v) { return __builtin_popcountll(v); }\n"
    "template <class T> static inline T __ldg(const T *p) { return *p; }\n"
    "template <class T> static inline T __ldcs(const T *p) { return *p; }\n"
    "template <class T> static inline T __ldcg(const T *p) { return *p; }\n"
    "static inline void __stcs(double2 *p, double2 v) { *p = v; }\n"
    "static inline void st_plain(double2 *p, double2 v) { *p = v; }\n"
    "static inline void cpa16(void *s, const void *g) { memcpy(s, g, 16); }\n"
    "static inline void cpa_wait() {}\nstatic inline void cpa_commit() {}\nstatic inline void cpa_wait_all() {}\n"
    "static inline void cpa_wait_first() {}\n"
    "static inline void mbar_init(u64 *bar, int) { __atomic_store_n(bar, 0ull, __ATOMIC_RELEASE); }\n"
    "static inline void mbar_expect_tx(u64 *, u32) {}\n"
    "// an mbarrier as the count of completed phases: wait(parity) returns once the phase of that parity is over\n"
    "static inline void emu_bar_complete(u64 *bar) { __atomic_add_fetch(bar, 1ull, __ATOMIC_RELEASE); }\n"
    "static inline void mbar_wait(u64 *bar, u32 parity)\n{\n"
    "  while ((__atomic_load_n(bar, __ATOMIC_ACQUIRE) & 1ull) == (u64)parity) sched_yield();\n}\n\n";

void replace_all(std::string &s, const std::string &from, const std::string &to)
{
  for (size_t pos = 0; (pos = s.find(from, pos)) != std::string::npos; pos += to.size()) s.replace(pos, from.size(), to);
}
}  // namespace

std::string generate(const std::vector<PassDesc> &passes)
{
  Out o;
  o.s = g_host_emulation ? HOST_PRELUDE : PRELUDE;
  for (size_t k = 0; k < passes.size(); ++k) {
    if (passes[k].pipelined) gen_pipelined(o, passes[k], (int)k);
    else gen_classic(o, passes[k], (int)k);
  }
  if (g_host_emulation) {
    // the few CUDA-only statements inside the kernel bodies
    replace_all(o.s, "asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");", ";");
    replace_all(o.s, "asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");", ";");
    replace_all(o.s, "asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");", ";");
    replace_all(o.s, "asm volatile(\"cp.async.bulk.wait_group 0;\" ::: \"memory\");", ";");
    replace_all(o.s, "extern __shared__ __align__(1024) unsigned char smem[];", "unsigned char *smem = emu_smem_base;");
    replace_all(o.s, "extern __shared__ double2 tile[];", "double2 *tile = reinterpret_cast<double2 *>(emu_smem_base);");
    // a uniform entry point per pass, the table of passes and the harness that walks it: the threads of a
    // block are OS threads, the blocks of a grid run one after the other
    Out t;
    for (size_t k = 0; k < passes.size(); ++k) {
      const PassDesc &pd = passes[k];
      const bool tma = !pd.pipelined && pd.tma_stage;
      t("static void emu_run_p%zu(const double2 *x, double2 *y, const double *diag, i64 rank_bits, u64 ntiles, Peers xs)\n{\n", k);
      if (pd.pipelined || tma) {
        t("  TMap tx; tx.base = (const double *)x;\n  TMap ty; ty.base = (const double *)y;\n");
        t("  dnm_jit_p%zu(tx, ty, x, y, diag, rank_bits, ntiles, xs);\n}\n", k);
      } else {
        t("  (void)ntiles;\n  dnm_jit_p%zu(x, y, diag, rank_bits, xs);\n}\n", k);
      }
    }
    t("struct EmuPass {\n  void (*run)(const double2 *, double2 *, const double *, i64, u64, Peers);\n"
      "  int threads, T, pipelined;\n  i64 rank_bits;\n  long long smem;\n};\n");
    t("static const EmuPass emu_passes[] = {\n");
    for (size_t k = 0; k < passes.size(); ++k) {
      const PassDesc &pd = passes[k];
      const Geo g(pd);
      const int nrem = pd.stage_remote ? remote_groups(pd) : 0;
      t("  {emu_run_p%zu, %d, %d, %d, 0x%llxll, %lldll},\n", k, g.NT, pd.T, pd.pipelined ? 1 : 0, (u64)pd.p->rank_bits,
        (long long)(((size_t)16 << pd.T) * (size_t)(8 + nrem) + 4096));
    }
    t("};\n");
    t("struct EmuJob { const EmuPass *ps; const double2 *x; double2 *y; const double *diag; u64 ntiles; unsigned grid; Peers xs; };\n"
      "struct EmuThread { const EmuJob *job; unsigned tid; };\n"
      "static void *emu_thread(void *arg)\n{\n"
      "  const EmuThread *t = (const EmuThread *)arg;\n  const EmuJob *j = t->job;\n"
      "  threadIdx.x = t->tid;\n  gridDim.x = j->grid;\n"
      "  for (unsigned b = 0; b < j->grid; ++b) {\n    blockIdx.x = b;\n"
      "    j->ps->run(j->x, j->y, j->diag, j->ps->rank_bits, j->ntiles, j->xs);\n"
      "    pthread_barrier_wait(&emu_barrier);  // the next block reuses the shared memory\n  }\n  return 0;\n}\n");
    // peers[h] = the input shard of rank (this ^ h); peers[0] = the local shard.  Pipelined grids are kept to
    // `pipelined_grid` CTAs so that every CTA walks its ring more than once.
    t("extern \"C\" int dnm_emu_npasses() { return %zu; }\n", passes.size());
    t("extern \"C\" int dnm_emu_mult(const double2 *const *peers, int npeers, double2 *y, const double *diag, long long nloc_rows,\n"
      "                            int pipelined_grid)\n{\n"
      "  Peers xs;\n  for (int h = 0; h < 16; ++h) xs.p[h] = h < npeers ? peers[h] : 0;\n"
      "  for (int k = 0; k < %zu; ++k) {\n    const EmuPass *ps = &emu_passes[k];\n"
      "    EmuJob job;\n    job.ps = ps;\n    job.x = peers[0];\n    job.y = y;\n    job.diag = k == 0 ? diag : 0;\n"
      "    job.ntiles = (u64)nloc_rows >> ps->T;\n    job.xs = xs;\n"
      "    job.grid = (unsigned)job.ntiles;\n"
      "    if (ps->pipelined && job.grid > (unsigned)pipelined_grid) job.grid = (unsigned)pipelined_grid;\n"
      "    emu_smem_base = (unsigned char *)aligned_alloc(1024, (size_t)ps->smem);\n"
      "    if (!emu_smem_base) return 1;\n"
      "    pthread_barrier_init(&emu_barrier, 0, (unsigned)ps->threads);\n"
      "    pthread_t *th = new pthread_t[ps->threads];\n    EmuThread *ta = new EmuThread[ps->threads];\n"
      "    for (int i = 0; i < ps->threads; ++i) {\n      ta[i].job = &job;\n      ta[i].tid = (unsigned)i;\n"
      "      if (pthread_create(&th[i], 0, emu_thread, &ta[i]) != 0) return 2;\n    }\n"
      "    for (int i = 0; i < ps->threads; ++i) pthread_join(th[i], 0);\n"
      "    delete[] th;\n    delete[] ta;\n    pthread_barrier_destroy(&emu_barrier);\n    free(emu_smem_base);\n  }\n  return 0;\n}\n",
      passes.size());
    o.s += t.s;
  }
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
  const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  const int rc = n.compile(prog, 3, opts);
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

Module *compile(const std::string &src, const std::vector<PassDesc> &passes, std::string &log)
{
  const std::vector<char> cubin = compile_cubin(src, log);
  if (cubin.empty()) return nullptr;
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
    const PassDesc &pd = passes[k];
    Kernel kn;
    const std::string name = "dnm_jit_p" + std::to_string(k);
    CUfunction f = nullptr;
    rc = d.moduleGetFunction(&f, mod, name.c_str());
    if (rc != CUDA_SUCCESS) {
      log += " [cuModuleGetFunction " + name + " failed]";
      delete m;
      return nullptr;
    }
    const Geo g(pd);
    static unsigned long long next_uid = 1;
    kn.uid = next_uid++;
    kn.func = f;
    kn.threads = g.NT;
    kn.pipelined = pd.pipelined;
    if (pd.pipelined) {
      const Box bx = tile_box(pd);
      kn.reduce = pd.p->accumulate == 1;
      kn.smem = (size_t)(pd.nbuf + (kn.reduce ? 1 : 0)) * ((size_t)16 << pd.T) + 128;
      kn.ctas_per_sm = pipelined_ctas(pd);
      kn.rank = bx.rank;
      for (int i = 0; i < bx.rank; ++i) {
        kn.dims[i] = bx.dims[i];
        kn.strides[i] = bx.strides[i];
        kn.box[i] = bx.box[i];
      }
    } else {
      kn.smem = ((size_t)16 << pd.T) * (size_t)(1 + (pd.stage_remote ? remote_groups(pd) : 0));
      if (pd.tma_stage) {
        const Box bx = tile_box(pd);
        kn.tma_args = true;
        kn.smem += 64;
        kn.reduce = pd.tma_reduce && pd.p->accumulate == 1;
        kn.rank = bx.rank;
        for (int i = 0; i < bx.rank; ++i) {
          kn.dims[i] = bx.dims[i];
          kn.strides[i] = bx.strides[i];
          kn.box[i] = bx.box[i];
        }
      }
    }
    d.funcSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)kn.smem);
    d.funcSetAttribute(f, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100);
    m->kernels.push_back(kn);
  }
  return m;
}

namespace {

// tensor maps are pure functions of (kernel geometry, base pointer): keep the last few
struct MapCache {
  const void *ptr[8];
  unsigned long long uid[8];
  CUtensorMap map[8];
  int next = 0, used = 0;
};

const CUtensorMap *tensor_map(const Kernel &k, const void *ptr)
{
  static MapCache cache;
  for (int i = 0; i < cache.used; ++i)
    if (cache.ptr[i] == ptr && cache.uid[i] == k.uid) return &cache.map[i];
  const int slot = cache.next;
  cache.next = (cache.next + 1) % 8;
  cache.used = std::min(cache.used + 1, 8);
  cuuint64_t dims[MAX_TMA_RANK], strides[MAX_TMA_RANK];
  cuuint32_t box[MAX_TMA_RANK], estr[MAX_TMA_RANK];
  for (int i = 0; i < k.rank; ++i) {
    dims[i] = k.dims[i];
    box[i] = k.box[i];
    estr[i] = 1;
    if (i > 0) strides[i - 1] = k.strides[i];
  }
  cache.ptr[slot] = nullptr;
  const CUresult rc = driver().encode(&cache.map[slot], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)k.rank,
                                      const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DNM_REQUIRE(rc == CUDA_SUCCESS, DNM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
  cache.ptr[slot] = ptr;
  cache.uid[slot] = k.uid;
  return &cache.map[slot];
}

}  // namespace

void launch(const Kernel &k, unsigned long long ntiles, int sm_count, cudaStream_t stream, const cplx *x, cplx *y,
            const double *diag, long long rank_bits, const cplx *const *peers)
{
  CUresult rc;
  struct {
    const cplx *p[MAX_RANKS];
  } xs;
  for (int h = 0; h < MAX_RANKS; ++h) xs.p[h] = peers ? peers[h] : nullptr;
  if (k.pipelined || k.tma_args) {
    static const CUtensorMap zero = {};
    const CUtensorMap *tmx = k.rank > 0 ? tensor_map(k, x) : &zero;
    const CUtensorMap *tmy = (k.rank > 0 && k.reduce) ? tensor_map(k, y) : &zero;
    const unsigned grid = k.pipelined ? (unsigned)std::min<unsigned long long>(ntiles, (unsigned long long)sm_count * k.ctas_per_sm)
                                      : (unsigned)ntiles;
    void *args[] = {(void *)tmx, (void *)tmy, (void *)&x, (void *)&y, (void *)&diag, (void *)&rank_bits, (void *)&ntiles,
                    (void *)&xs};
    rc = driver().launchKernel((CUfunction)k.func, grid, 1, 1, (unsigned)k.threads, 1, 1, (unsigned)k.smem, (CUstream)stream, args,
                               nullptr);
  } else {
    void *args[] = {(void *)&x, (void *)&y, (void *)&diag, (void *)&rank_bits, (void *)&xs};
    rc = driver().launchKernel((CUfunction)k.func, (unsigned)ntiles, 1, 1, (unsigned)k.threads, 1, 1, (unsigned)k.smem,
                               (CUstream)stream, args, nullptr);
  }
  DNM_REQUIRE(rc == CUDA_SUCCESS, DNM_ERR_CUDA,
              "cuLaunchKernel of a generated MatMult pass failed (%d): tiles=%llu threads=%d smem=%zu pipelined=%d", (int)rc,
              ntiles, k.threads, k.smem, (int)k.pipelined);
}

}  // namespace jit
}  // namespace dnm
