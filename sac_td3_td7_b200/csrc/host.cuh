// host.cuh -- host-side plumbing: error string, device buffers, launch tables, launchers.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#ifndef B200_EMU
#include <cuda.h>  // CUtensorMap + the cuTensorMapEncodeTiled prototype (header only: the symbol is fetched at run time)
#endif

#include "common.cuh"
#include "gemm.cuh"
#include "gemm_tc.cuh"
#include "narrow.cuh"
#include "ops.cuh"

namespace b200rl {

inline std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
inline void set_last_error(const std::string& msg) { last_error_ref() = msg; }

// ----------------------------------------------------------------------------- device memory
// All persistent state of a handle lives in a few arenas; nothing is allocated on the hot path.
struct DeviceArena {
  std::vector<void*> blocks;
  size_t total_bytes = 0;
  ~DeviceArena() { release(); }
  void release() {
    for (void* b : blocks) cudaFree(b);
    blocks.clear();
    total_bytes = 0;
  }
  // 256-byte aligned (cudaMalloc alignment); zero-filled
  template <class T>
  int alloc(T** out, size_t count) {
    void* p = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    B200_CHECK_CUDA(cudaMalloc(&p, bytes));
    B200_CHECK_CUDA(cudaMemset(p, 0, bytes));
    // cudaMemset on device memory is ordered on the legacy default stream only and may return before it has run; the
    // engines work on non-blocking streams, so without this a later fill / upload could be overwritten by the zeroing
    B200_CHECK_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    blocks.push_back(p);
    total_bytes += bytes;
    *out = reinterpret_cast<T*>(p);
    return kOk;
  }
};

// ----------------------------------------------------------------------------- gemm tables
inline GemmProblem make_problem(int I, int J, float* C, int c_ld, long long c_gs = 0) {
  GemmProblem p;
  memset(&p, 0, sizeof(p));
  p.I = I;
  p.J = J;
  p.C = C;
  p.c_ld = c_ld;
  p.c_gs = c_gs;
  return p;
}
inline void add_seg(GemmProblem& p, const float* A, int a_ld, long long a_gs, const float* B, int b_ld, long long b_gs, int len) {
  GemmSeg& s = p.seg[p.nseg++];
  s.A = A; s.a_ld = a_ld; s.a_gs = a_gs;
  s.B = B; s.b_ld = b_ld; s.b_gs = b_gs;
  s.len = len;
  s.b_ones = 0;
}

// How many CTAs (one thread-block cluster) share one output tile.  A single agent's layers are 64 tiles of
// 4-12 serial 64-deep chunks each -- less than half of the 148 SMs busy, on a chain that is pure latency --
// so the reduction is cut across a cluster while the launch still fits one wave of 2 CTAs per SM and every
// CTA keeps at least two chunks.  Measured on B200 with profiles/sweep_gemm.py (graph-timed).
inline int chunks_of(const GemmProblem& p) {
  int c = 0;
  for (int s = 0; s < p.nseg; ++s) c += ceil_div(p.seg[s].len, kBK);
  return c;
}
inline int choose_splitk(const std::vector<GemmProblem>& probs, int tiles, int groups) {
#ifdef B200_EMU
  (void)probs; (void)tiles; (void)groups;
  return 1;  // clusters / DSMEM are not emulated
#else
  int max_chunks = 1;
  for (const auto& p : probs) max_chunks = std::max(max_chunks, chunks_of(p));
  const long long ctas = (long long)tiles * groups;
  // sweep (us per launch, 64 tiles): K=256 -> 5.3 / 4.8 / 4.9 for 1 / 2 / 4 CTAs per tile; K=768 -> 11.5 / 7.8 / 7.3;
  // at 128 tiles and above splitting no longer pays (the SMs are already issue-bound)
  int s = 1;
  if (ctas * 2 <= 148 && max_chunks >= 4) s = 2;
  if (ctas * 2 <= 148 && max_chunks >= 8) s = 4;
  return s;
#endif
}

inline FfmaProblem make_ffma_problem(const GemmProblem& p, int splitk) {
  FfmaProblem f;
  memset(&f, 0, sizeof(f));
  for (int s = 0; s < 3; ++s) f.seg[s] = p.seg[s];
  f.nseg = p.nseg; f.I = p.I; f.J = p.J;
  f.total_chunks = chunks_of(p);
  f.per_split = ceil_div(f.total_chunks, splitk);
  f.tile_begin = p.tile_begin;
  f.tiles_j = p.tiles_j;
  f.tiles_j_magic = p.tiles_j > 1 ? (unsigned)((1ull << 32) / (unsigned)p.tiles_j) + 1u : 0u;
  f.C = p.C; f.c_gs = p.c_gs; f.c_ld = p.c_ld; f.act = p.act;
  f.bias = p.bias; f.bias_gs = p.bias_gs;
  f.aux = p.aux; f.aux_gs = p.aux_gs; f.aux_ld = p.aux_ld; f.accumulate = p.accumulate;
  f.P = p.P; f.M = p.M; f.V = p.V; f.p_gs = p.p_gs; f.p_ld = p.p_ld;
  f.noise_sigma = p.noise_sigma; f.noise_clip = p.noise_clip;
  f.adam = p.adam;
  return f;
}

// Stacked seeds (groups >= kTcMinGroups) run their eligible problems on the tensor cores (TF32);
// B200RL_TC=0/1 overrides (tests, A/B measurements).
constexpr int kTcMinGroups = 8;
inline bool use_tensor_cores(int groups) {
#ifdef B200_EMU
  (void)groups;
  return false;
#else
  if (const char* e = getenv("B200RL_TC")) return atoi(e) != 0;
  return groups >= kTcMinGroups;
#endif
}

// ----------------------------------------------------------------------------- host side
// A launch table: problems are appended on the host once, then replayed (the struct is passed by value
// as the kernel parameter; nothing lives in device memory).
struct GemmTable {
  FfmaLaunch launch;      // problems run by the FFMA kernel
  BiasLaunch launch_bias; // bias gradients of stacked seeds (tensor-core path only)
  GemmLaunch launch_tc;   // problems run by the tcgen05 kernel (stacked seeds; gemm_tc.cuh)
  GemmLaunch launch_ws;   // ... of which the TMA-addressable ones, run by the persistent warp-specialised variant
  NarrowWgradLaunch launch_nw;  // narrow.cuh: weight gradients a few input columns wide (+ their layer's bias gradient)
  NarrowRowsLaunch launch_nr;   //             forward / dgrad problems with at most 8 output columns
  Rank1Launch launch_r1;        //             dgrad through a one-output layer
  int nprob = 0;
  int total_tiles = 0;    // FFMA tiles (32 x 32)
  int tc_tiles = 0;       // tcgen05 tiles (128 rows x N)
  int mode = kFwd;
  int groups = 1;
  bool debug_print = false;
};

// Narrow problems of a stacked-seed update (narrow.cuh): on wherever the tensor cores are (default), B200RL_NARROW=0
// keeps them on the FFMA / tcgen05 kernels (A/B), B200RL_NARROW=1 forces them for any seed count (emulation tests).
inline bool use_narrow(int groups) {
  const char* e = getenv("B200RL_NARROW");  // (read at every table build, like B200RL_TC: tests flip it inside one process)
  return e ? atoi(e) != 0 : use_tensor_cores(groups);
}

// The small-tile kernel's product loop: fp32 FFMA (default), or error-compensated 3xTF32 on the warp-level tensor-core
// path (B200RL_MMA=1 / b200rl_set_mma).  Measured on B200: HMMA.1688.F32.TF32 issues at about the FFMA MAC rate per SM
// (the legacy mma.sync path is not the fast tensor path on Blackwell), so with three MMAs per product the chunk loop is
// only 19 % shorter (3 959 vs 4 894 cycles for K = 256) and a TD7 update 2-4 % faster, while ReLU / AvgL1Norm branch
// decisions near zero move: it holds the 2e-6 op bar and the 1e-5 agent bar, but is not the default.
inline int& mma_override() {
  static int v = -1;  // b200rl_set_mma(): -1 = environment / default
  return v;
}
inline bool use_mma() {
#ifdef B200_EMU
  return false;
#else
  static const bool on = [] { const char* e = getenv("B200RL_MMA"); return e ? atoi(e) != 0 : false; }();
  return mma_override() >= 0 ? mma_override() != 0 : on;
#endif
}

#ifndef B200_EMU
// ---- TMA descriptors for the tcgen05 forward problems --------------------------------------------------------------
// Operand of one reduction segment as a 3-D fp32 tensor {reduction index, output index, stacked seed}; box = one
// pipeline stage tile {32, rows, 1}, SWIZZLE_128B -- which is byte for byte the UMMA K-major canonical layout the
// kernel's shared-memory descriptors expect.  Out-of-range reduction indices (segment tails such as the 8 action
// columns) are zero-filled by the TMA unit.
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tma_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
inline bool tma_operand_ok(const float* base, int ld, long long gs, int groups) {
  return base && (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld & 3) == 0 && ld > 0 && (groups == 1 || ((gs & 3) == 0 && gs > 0));
}
inline bool tma_make_map(CUtensorMap* m, const float* base, int ld, long long gs, int len, int rows, int box_rows, int groups) {
  EncodeTiledFn fn = tma_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)len, (cuuint64_t)rows, (cuuint64_t)groups};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(groups > 1 ? gs : (long long)rows * ld) * 4};
  const cuuint32_t box[3] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// MN-major operand (contiguous along the output index c in global memory): 4-D tensor {32 c, r, c / 32, seed}, box
// {32, 32, blocks, 1}, SWIZZLE_128B_ATOM_32B = the UMMA SWIZZLE_128B_BASE32B MN-major layout (4 KB per 32-index block)
inline bool tma_make_map_mn(CUtensorMap* m, const float* base, int ld, long long gs, int len, int cols, int box_blocks, int groups) {
  EncodeTiledFn fn = tma_encode_fn();
  if (!fn || (cols & 31) != 0) return false;
  const cuuint64_t dims[4] = {32u, (cuuint64_t)len, (cuuint64_t)(cols / 32), (cuuint64_t)groups};
  const cuuint64_t strides[3] = {(cuuint64_t)ld * 4, 128u, (cuuint64_t)(groups > 1 ? gs : (long long)len * ld) * 4};
  const cuuint32_t box[4] = {32u, (cuuint32_t)kTcBK, (cuuint32_t)box_blocks, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
inline bool use_tma() {
  static const bool on = [] { const char* e = getenv("B200RL_TC_TMA"); return e ? atoi(e) != 0 : true; }();  // B200RL_TC_TMA=0: cp.async staging
  return on;
}
#endif

inline int gemm_table_upload(GemmTable* t, int mode, std::vector<GemmProblem>& probs, DeviceArena* arena, int groups = 1, int force_splitk = 0) {
  B200_REQUIRE((int)probs.size() <= kMaxProblems, "too many problems in one gemm launch");
  std::vector<GemmProblem> ff, tc, bias, nw, nr, r1;
  const bool want_tc = use_tensor_cores(groups), want_narrow = use_narrow(groups);
  for (auto& p : probs) {
    B200_REQUIRE(p.nseg >= 1 && p.nseg <= 3, "gemm problem needs 1..3 segments");
    B200_REQUIRE(p.I > 0 && p.J > 0, "gemm problem with empty output");
    if (p.act == kActNorm) {
      bool ok = (mode == kFwd) && want_tc && p.aux != nullptr;
#ifndef B200_EMU
      ok = ok && tc_eligible(p);
#else
      ok = false;
#endif
      B200_REQUIRE(ok, "a fused AvgL1Norm epilogue exists on the tensor-core forward path only (128-row tiles, 64..256 columns)");
    }
    if (want_narrow && rank1_eligible(p, mode)) { r1.push_back(p); continue; }
#ifndef B200_EMU
    if (want_tc && tc_eligible(p)) { tc.push_back(p); continue; }
    // A weight gradient whose input width is not a tensor-core tile width (Humanoid's 376 observation columns) is cut along
    // its columns into tcgen05 problems of <= 256 columns, multiples of 32 (TMA-addressable), and a narrow remainder:
    // 376 = 256 + 96 + 24.  As 32 x 32 FFMA tiles these gradients were 20 % of a 256-seed Humanoid update.
    if (want_tc && want_narrow && mode == kWgrad && p.nseg == 1 && !p.seg[0].b_ones && p.J > kTcNMax / 4 && (p.I % kTcM) == 0 && (p.seg[0].b_ld & 3) == 0 &&
        (p.P || p.C) && (!p.P || (p.p_ld & 3) == 0)) {
      int j0 = 0;
      auto sub = [&](int jn) {
        GemmProblem q = p;
        q.J = jn;
        q.seg[0].B = p.seg[0].B + j0;
        if (q.C) q.C += j0;
        if (q.P) { q.P += j0; q.M += j0; q.V += j0; }
        return q;
      };
      while (p.J - j0 >= 64) {
        const int jn = std::min(kTcNMax, (p.J - j0) / 32 * 32);
        tc.push_back(sub(jn));
        j0 += jn;
      }
      if (p.J - j0 > 0) nw.push_back(sub(p.J - j0));
      continue;
    }
#endif
    if (want_narrow && mode == kWgrad && (narrow_wgrad_eligible(p) || narrow_wgrad_t_eligible(p))) { nw.push_back(p); continue; }
    if (want_narrow && narrow_rows_eligible(p, mode)) { nr.push_back(p); continue; }
    if ((want_tc || want_narrow) && mode == kWgrad && p.nseg == 1 && p.seg[0].b_ones && p.J == 1) { bias.push_back(p); continue; }
    ff.push_back(p);
  }
  (void)want_tc;
  // ---- narrow problems (narrow.cuh)
  memset(&t->launch_nw, 0, sizeof(t->launch_nw));
  memset(&t->launch_nr, 0, sizeof(t->launch_nr));
  memset(&t->launch_r1, 0, sizeof(t->launch_r1));
  {
    int units = 0;
    for (size_t i = 0; i < nw.size(); ++i) {
      const GemmProblem& p = nw[i];
      NarrowWgrad& q = t->launch_nw.p[i];
      q.transposed = narrow_wgrad_eligible(p) ? 0 : 1;
      q.rows = p.seg[0].len;
      if (!q.transposed) {
        q.dZ = p.seg[0].A; q.dz_gs = p.seg[0].a_gs; q.dz_ld = p.seg[0].a_ld; q.out = p.I;
        q.X = p.seg[0].B; q.x_gs = p.seg[0].b_gs; q.x_ld = p.seg[0].b_ld; q.J = p.J;
      } else {  // stream the wide operand (the layer's input X), broadcast the few columns of dZ
        q.dZ = p.seg[0].B; q.dz_gs = p.seg[0].b_gs; q.dz_ld = p.seg[0].b_ld; q.out = p.J;
        q.X = p.seg[0].A; q.x_gs = p.seg[0].a_gs; q.x_ld = p.seg[0].a_ld; q.J = p.I;
      }
      q.C = p.C; q.P = p.P; q.M = p.M; q.V = p.V;
      q.p_gs = p.P ? p.p_gs : p.c_gs; q.p_ld = p.P ? p.p_ld : p.c_ld;
      q.adam = p.adam;
      q.jchunks = ceil_div(q.J, kNwCols);
      q.unit_begin = units;
      units += q.jchunks * ceil_div(q.out, kNwThreads);
    }
    t->launch_nw.nprob = (int)nw.size();
    t->launch_nw.total_units = units;
    if (const char* e = getenv("B200RL_NW_DBG")) t->launch_nw.dbg = atoi(e);
    // the bias gradient of a layer (column sums of the same dZ) rides with the first narrow problem of that layer
    std::vector<GemmProblem> bias_left;
    for (auto& b : bias) {
      bool folded = false;
      for (size_t i = 0; i < nw.size() && !folded; ++i) {
        NarrowWgrad& q = t->launch_nw.p[i];
        if (!q.transposed && q.dZ == b.seg[0].A && q.dz_gs == b.seg[0].a_gs && q.out == b.I && !q.bP && !q.bC && (b.P ? b.p_gs : b.c_gs) == q.p_gs && (!b.P || b.adam == q.adam || !q.P)) {
          q.bC = b.C; q.bP = b.P; q.bM = b.M; q.bV = b.V;
          if (!q.adam) q.adam = b.adam;
          folded = true;
        }
      }
      if (!folded) bias_left.push_back(b);
    }
    bias.swap(bias_left);
    units = 0;
    for (size_t i = 0; i < nr.size(); ++i) {
      const GemmProblem& p = nr[i];
      NarrowRows& q = t->launch_nr.p[i];
      for (int s = 0; s < 3; ++s) q.seg[s] = p.seg[s];
      q.nseg = p.nseg; q.I = p.I; q.J = p.J;
      q.C = p.C; q.c_gs = p.c_gs; q.c_ld = p.c_ld; q.act = p.act;
      q.bias = p.bias; q.bias_gs = p.bias_gs;
      q.aux = p.aux; q.aux_gs = p.aux_gs; q.aux_ld = p.aux_ld; q.accumulate = p.accumulate;
      q.noise_sigma = p.noise_sigma; q.noise_clip = p.noise_clip;
      q.unit_begin = units;
      units += ceil_div(p.I, 32);
    }
    t->launch_nr.nprob = (int)nr.size();
    t->launch_nr.total_units = units;
    long long quads = 0;
    for (size_t i = 0; i < r1.size(); ++i) {
      const GemmProblem& p = r1[i];
      Rank1Problem& q = t->launch_r1.p[i];
      q.dz = p.seg[0].A; q.dz_gs = p.seg[0].a_gs; q.dz_ld = p.seg[0].a_ld;
      q.w = p.seg[0].B; q.w_gs = p.seg[0].b_gs;
      q.aux = p.aux; q.aux_gs = p.aux_gs; q.aux_ld = p.aux_ld; q.act = p.act;
      q.C = p.C; q.c_gs = p.c_gs; q.c_ld = p.c_ld; q.accumulate = p.accumulate;
      q.I = p.I; q.J = p.J;
      q.quad_begin = quads;
      quads += (long long)p.I * (p.J / 4);
    }
    t->launch_r1.nprob = (int)r1.size();
    t->launch_r1.total_quads = quads;
  }
  // longest reductions first: their CTAs are scheduled in the first wave and the short problems fill in behind
  std::stable_sort(ff.begin(), ff.end(), [&](const GemmProblem& a, const GemmProblem& b) { return chunks_of(a) > chunks_of(b); });
  std::stable_sort(tc.begin(), tc.end(), [&](const GemmProblem& a, const GemmProblem& b) { return chunks_of(a) > chunks_of(b); });
  int tiles = 0;
  for (auto& p : ff) {
    p.tile_begin = tiles;
    p.tiles_j = ceil_div(p.J, kTile);
    tiles += ceil_div(p.I, kTile) * p.tiles_j;
  }
  int tct = 0;
  for (auto& p : tc) {
    p.tile_begin = tct;
    p.tiles_j = 1;
    tct += p.I / 128;
  }
  t->mode = mode;
  t->nprob = (int)probs.size();
  t->total_tiles = tiles;
  t->tc_tiles = tct;
  t->groups = groups;
  memset(&t->launch, 0, sizeof(t->launch));
  memset(&t->launch_tc, 0, sizeof(t->launch_tc));
  memset(&t->launch_ws, 0, sizeof(t->launch_ws));
  if (const char* e = getenv("B200RL_SPLITK")) force_splitk = atoi(e);  // tuning knob (profiles/sweep_gemm.py)
  int splitk = ff.empty() ? 1 : (force_splitk > 0 ? force_splitk : choose_splitk(ff, tiles, groups));
#ifdef B200_EMU
  splitk = 1;
#endif
  B200_REQUIRE(splitk >= 1 && splitk <= kMaxSplitK, "split-K factor out of range");
  for (size_t i = 0; i < ff.size(); ++i) t->launch.p[i] = make_ffma_problem(ff[i], splitk);
  for (size_t i = 0; i < tc.size(); ++i) t->launch_tc.p[i] = tc[i];
  t->launch.nprob = (int)ff.size();
  t->launch.splitk = splitk;
  memset(&t->launch_bias, 0, sizeof(t->launch_bias));
  for (size_t i = 0; i < bias.size(); ++i) {  // wgrad_problems(): A = dZ [rows][out], C/P/M/V = the bias vector
    const GemmProblem& p = bias[i];
    BiasProblem& b = t->launch_bias.p[i];
    b.dZ = p.seg[0].A; b.dz_gs = p.seg[0].a_gs; b.dz_ld = p.seg[0].a_ld; b.rows = p.seg[0].len; b.out = p.I;
    b.C = p.C; b.P = p.P; b.M = p.M; b.V = p.V; b.p_gs = p.P ? p.p_gs : p.c_gs; b.adam = p.adam;
  }
  t->launch_bias.nprob = (int)bias.size();
  static int table_counter = 0;
  t->launch.tag = ++table_counter;
  if (getenv("B200RL_DEBUG_TABLES")) {
    int mc = 0;
    for (auto& p : ff) mc = std::max(mc, chunks_of(p));
    fprintf(stderr, "[b200rl] gemm table %d mode=%d problems=%zu (+%zu tc, %zu narrow wgrad, %zu narrow rows, %zu rank-1, %zu bias) tiles=%d groups=%d max_chunks=%d splitk=%d\n",
            t->launch.tag, mode, ff.size(), tc.size(), nw.size(), nr.size(), r1.size(), bias.size(), tiles, groups, mc, splitk);
    t->debug_print = true;
  }
  t->launch_tc.nprob = (int)tc.size();
#ifndef B200_EMU
  if (!tc.empty() && use_tma()) {
    const bool a_km = (mode != kWgrad), b_km = (mode == kFwd);
    std::vector<CUtensorMap> maps(tc.size() * 6);
    unsigned mask = 0;
    for (size_t i = 0; i < tc.size(); ++i) {
      const GemmProblem& p = tc[i];
      bool ok = (p.J % 8) == 0 && p.J <= 256 && (b_km || (p.J % 32) == 0);
      for (int sgi = 0; sgi < p.nseg && ok; ++sgi) {
        const GemmSeg& sg = p.seg[sgi];
        ok = tma_operand_ok(sg.A, sg.a_ld, sg.a_gs, groups) && tma_operand_ok(sg.B, sg.b_ld, sg.b_gs, groups);
        if (ok) ok = a_km ? tma_make_map(&maps[i * 6 + sgi * 2 + 0], sg.A, sg.a_ld, sg.a_gs, sg.len, p.I, kTcM, groups)
                          : tma_make_map_mn(&maps[i * 6 + sgi * 2 + 0], sg.A, sg.a_ld, sg.a_gs, sg.len, p.I, kTcM / 32, groups);
        if (ok) ok = b_km ? tma_make_map(&maps[i * 6 + sgi * 2 + 1], sg.B, sg.b_ld, sg.b_gs, sg.len, p.J, p.J, groups)
                          : tma_make_map_mn(&maps[i * 6 + sgi * 2 + 1], sg.B, sg.b_ld, sg.b_gs, sg.len, p.J, p.J / 32, groups);
      }
      if (ok) mask |= 1u << i;
    }
    static const bool use_ws = [] { const char* e = getenv("B200RL_TC_WS"); return e ? atoi(e) != 0 : true; }();  // B200RL_TC_WS=0: one tile per CTA
    if (mask && use_ws) {
      // the TMA-addressable problems move to the persistent warp-specialised kernel; the rest stay where they are
      std::vector<GemmProblem> ws, rest;
      std::vector<CUtensorMap> ws_maps;
      for (size_t i = 0; i < tc.size(); ++i) {
        if ((mask >> i) & 1u) {
          ws.push_back(tc[i]);
          for (int q = 0; q < 6; ++q) ws_maps.push_back(maps[i * 6 + q]);
        } else {
          rest.push_back(tc[i]);
        }
      }
      int wt = 0;
      for (auto& p : ws) { p.tile_begin = wt; wt += p.I / kTcM; }
      int rt = 0;
      for (auto& p : rest) { p.tile_begin = rt; rt += p.I / kTcM; }
      CUtensorMap* d_maps = nullptr;
      B200_TRY(arena->alloc(&d_maps, ws_maps.size()));
      B200_CHECK_CUDA(cudaMemcpy(d_maps, ws_maps.data(), ws_maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
      memset(&t->launch_ws, 0, sizeof(t->launch_ws));
      for (size_t i = 0; i < ws.size(); ++i) t->launch_ws.p[i] = ws[i];
      t->launch_ws.nprob = (int)ws.size();
      t->launch_ws.total_tiles = wt;
      t->launch_ws.tmaps = d_maps;
      t->launch_ws.tma_mask = (1u << ws.size()) - 1u;
      memset(&t->launch_tc, 0, sizeof(t->launch_tc));
      for (size_t i = 0; i < rest.size(); ++i) t->launch_tc.p[i] = rest[i];
      t->launch_tc.nprob = (int)rest.size();
      tct = rt;
      t->tc_tiles = rt;
    } else if (mask) {
      CUtensorMap* d_maps = nullptr;
      B200_TRY(arena->alloc(&d_maps, maps.size()));
      B200_CHECK_CUDA(cudaMemcpy(d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
      t->launch_tc.tmaps = d_maps;
      t->launch_tc.tma_mask = mask;
    }
    if (t->debug_print) fprintf(stderr, "[b200rl]   tcgen05 problems fed by TMA: mask 0x%x of %zu\n", mask, tc.size());
  }
#endif
  t->launch_tc.total_tiles = tct;
  return kOk;
}

inline int gemm_table_launch(const GemmTable& t, int groups, cudaStream_t stream) {
  if (t.nprob == 0) return kOk;
  if (groups != t.groups) {
    set_last_error("gemm table launched with a different number of groups than it was built for");
    return kErrState;
  }
#ifndef B200_EMU
  static bool attr_set = false;  // > 48 KB of dynamic shared memory needs the opt-in once per kernel
  if (!attr_set) {
    void (*ks[6])(FfmaLaunch) = {gemm_kernel<kFwd, false>, gemm_kernel<kDgrad, false>, gemm_kernel<kWgrad, false>,
                                 gemm_kernel<kFwd, true>,  gemm_kernel<kDgrad, true>,  gemm_kernel<kWgrad, true>};
    for (auto k : ks) B200_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes(kMaxSplitK)));
    B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<kFwd>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<kDgrad>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<kWgrad>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    attr_set = true;
  }
  if (t.launch_ws.nprob > 0) {
    static const int n_sm = [] { int v = 148, d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v; }();
    static bool ws_attr = false;
    if (!ws_attr) {
      B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_ws_kernel<kFwd>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmemBytes));
      B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_ws_kernel<kDgrad>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmemBytes));
      B200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_ws_kernel<kWgrad>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmemBytes));
      ws_attr = true;
    }
    const int items = t.launch_ws.total_tiles * groups;
    dim3 grid((unsigned)std::min(items, n_sm), 1, 1), block(kWsThreads, 1, 1);
    switch (t.mode) {
      case kFwd: { auto k = gemm_tc_ws_kernel<kFwd>; B200_LAUNCH(k, grid, block, kWsSmemBytes, stream, t.launch_ws, groups); break; }
      case kDgrad: { auto k = gemm_tc_ws_kernel<kDgrad>; B200_LAUNCH(k, grid, block, kWsSmemBytes, stream, t.launch_ws, groups); break; }
      default: { auto k = gemm_tc_ws_kernel<kWgrad>; B200_LAUNCH(k, grid, block, kWsSmemBytes, stream, t.launch_ws, groups); break; }
    }
    B200_CHECK_LAUNCH();
  }
  if (t.launch_tc.nprob > 0) {
    dim3 grid((unsigned)t.tc_tiles, (unsigned)groups, 1), block(kTcThreads, 1, 1);
    switch (t.mode) {
      case kFwd: { auto k = gemm_tc_kernel<kFwd>; B200_LAUNCH(k, grid, block, kTcSmemBytes, stream, t.launch_tc); break; }
      case kDgrad: { auto k = gemm_tc_kernel<kDgrad>; B200_LAUNCH(k, grid, block, kTcSmemBytes, stream, t.launch_tc); break; }
      default: { auto k = gemm_tc_kernel<kWgrad>; B200_LAUNCH(k, grid, block, kTcSmemBytes, stream, t.launch_tc); break; }
    }
    B200_CHECK_LAUNCH();
  }
#endif
  if (t.launch_bias.nprob > 0) {
    B200_LAUNCH(bias_grad_kernel, dim3((unsigned)t.launch_bias.nprob, (unsigned)groups), dim3(256), 0, stream, t.launch_bias);
    B200_CHECK_LAUNCH();
  }
  if (t.launch_nw.nprob > 0) {
#ifndef B200_EMU
    static bool nw_attr = false;
    if (!nw_attr) {
      B200_CHECK_CUDA(cudaFuncSetAttribute(narrow_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNwSmemBytes));
      nw_attr = true;
    }
#endif
    B200_LAUNCH(narrow_wgrad_kernel, dim3((unsigned)t.launch_nw.total_units, (unsigned)groups), dim3(kNwThreads), (size_t)kNwSmemBytes, stream, t.launch_nw);
    B200_CHECK_LAUNCH();
  }
  if (t.launch_nr.nprob > 0) {
    dim3 grid((unsigned)t.launch_nr.total_units, (unsigned)groups), block(kNrThreads);
    if (t.mode == kFwd) { auto k = narrow_rows_kernel<kFwd>; B200_LAUNCH(k, grid, block, 0, stream, t.launch_nr); }
    else { auto k = narrow_rows_kernel<kDgrad>; B200_LAUNCH(k, grid, block, 0, stream, t.launch_nr); }
    B200_CHECK_LAUNCH();
  }
  if (t.launch_r1.nprob > 0) {
    const long long blocks = (t.launch_r1.total_quads + 255) / 256;
    B200_LAUNCH(rank1_dgrad_kernel, dim3((unsigned)std::min<long long>(blocks, 148 * 8), (unsigned)groups), dim3(256), 0, stream, t.launch_r1);
    B200_CHECK_LAUNCH();
  }
  if (t.launch.nprob > 0) {
    const int sk = t.launch.splitk;
    dim3 grid((unsigned)(t.total_tiles * sk), (unsigned)groups, 1), block(kGemmThreads, 1, 1);
    const size_t smem = (size_t)gemm_smem_bytes(sk);
    void (*k)(FfmaLaunch) = t.mode == kFwd ? gemm_kernel<kFwd, false> : t.mode == kDgrad ? gemm_kernel<kDgrad, false> : gemm_kernel<kWgrad, false>;
#ifndef B200_EMU
    if (use_mma()) k = t.mode == kFwd ? gemm_kernel<kFwd, true> : t.mode == kDgrad ? gemm_kernel<kDgrad, true> : gemm_kernel<kWgrad, true>;
#endif
#ifndef B200_EMU
    if (sk > 1) {
      ++launch_counter();
      launch_cluster(k, grid, block, smem, stream, (unsigned)sk, t.launch);
    } else
#endif
    {
      B200_LAUNCH(k, grid, block, smem, stream, t.launch);
    }
    B200_CHECK_LAUNCH();
  }
  return kOk;
}

// ----------------------------------------------------------------------------- row-op tables (AvgL1Norm)
struct RowTable {
  std::vector<RowOp> h_ops;  // host copy (the persistent dataflow kernel records them, mega.cuh)
  RowOp* d_ops = nullptr;
  int nops = 0;
  int total_rows = 0;
  bool bwd = false;
};
inline RowOp make_rowop(const float* x, int x_ld, long long x_gs, float* y, int y_ld, long long y_gs, int rows, int n,
                        const float* dy = nullptr, int dy_ld = 0, long long dy_gs = 0) {
  RowOp r;
  memset(&r, 0, sizeof(r));
  r.x = x; r.x_ld = x_ld; r.x_gs = x_gs;
  r.y = y; r.y_ld = y_ld; r.y_gs = y_gs;
  r.dy = dy; r.dy_ld = dy_ld; r.dy_gs = dy_gs;
  r.rows = rows; r.n = n;
  return r;
}
inline int row_table_upload(RowTable* t, bool bwd, std::vector<RowOp>& ops, DeviceArena* arena) {
  int rows = 0;
  for (auto& o : ops) { o.row_begin = rows; rows += o.rows; }
  t->bwd = bwd;
  t->h_ops = ops;
  t->nops = (int)ops.size();
  t->total_rows = rows;
  B200_TRY(arena->alloc(&t->d_ops, ops.size()));
  B200_CHECK_CUDA(cudaMemcpy(t->d_ops, ops.data(), ops.size() * sizeof(RowOp), cudaMemcpyHostToDevice));
  return kOk;
}
inline int row_table_launch(const RowTable& t, int groups, cudaStream_t stream) {
  if (t.nops == 0) return kOk;
  dim3 grid((unsigned)ceil_div(t.total_rows, 8), (unsigned)groups, 1), block(256, 1, 1);
  if (t.bwd) { auto k = avgl1norm_kernel<true>; B200_LAUNCH(k, grid, block, 0, stream, t.d_ops, t.nops, t.total_rows); }
  else { auto k = avgl1norm_kernel<false>; B200_LAUNCH(k, grid, block, 0, stream, t.d_ops, t.nops, t.total_rows); }
  B200_CHECK_LAUNCH();
  return kOk;
}

inline int launch_elementwise_grid(long long n) {
  long long b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

// ----------------------------------------------------------------------------- streams for fork/join
// Inside a captured graph, parallel branches are expressed with event record/wait between a few
// side streams; under emulation they degenerate to sequential execution in issue order.
struct Fork {
  cudaStream_t main = nullptr;
  std::vector<cudaStream_t> side;
  std::vector<cudaEvent_t> events;
  size_t next_event = 0;
  // side streams listed in `low_priority_mask` carry work with slack (encoder step, hoisted actor forward):
  // captured kernel nodes inherit the stream priority, so the critical chain gets the SMs first
  int init(cudaStream_t main_stream, int nside, int nevents, unsigned low_priority_mask = 0) {
    main = main_stream;
    side.resize(nside);
    int least = 0, greatest = 0;
    B200_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    for (int i = 0; i < nside; ++i)
      B200_CHECK_CUDA(cudaStreamCreateWithPriority(&side[i], cudaStreamNonBlocking, ((low_priority_mask >> i) & 1u) ? least : greatest));
    events.resize(nevents);
    for (auto& e : events) B200_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return kOk;
  }
  void destroy() {
    for (auto& s : side) cudaStreamDestroy(s);
    for (auto& e : events) cudaEventDestroy(e);
    side.clear();
    events.clear();
  }
  void begin_step() { next_event = 0; }
  // make `to` wait for everything issued so far on `from`
  int link(cudaStream_t from, cudaStream_t to) {
    if (next_event >= events.size()) {
      set_last_error("Fork: out of events");
      return kErrState;
    }
    cudaEvent_t e = events[next_event++];
    B200_CHECK_CUDA(cudaEventRecord(e, from));
    B200_CHECK_CUDA(cudaStreamWaitEvent(to, e, 0));
    return kOk;
  }
};

}  // namespace b200rl
