// mega.cuh -- a whole phase of an update (dozens of dependent layers) as ONE persistent dataflow kernel.
//
// Why: a single agent's update is a chain of 19-35 DEPENDENT 256-row layers (rl/agent/td7.py:287-332).  As one launch per
// layer the chain pays, per link, ~1.2 us between dependent graph nodes, the ramp of 64-256 CTAs being rasterised, the
// parameter parsing, and the drain of the slowest CTA -- about half of the 5-8 us a link takes -- and nothing can start
// before the whole previous layer has drained, although a 32-row block of layer l+1 only needs the same 32 rows of
// layer l.  Here the launch tables of a phase are flattened into one ordered list of TILES (the 32 x 32 output tiles of
// gemm.cuh, 32-row blocks of AvgL1Norm, and a few scalar / noise nodes); a grid of persistent CTAs pulls tiles from the
// list in order and every tile waits on, and then bumps, small counters in global memory:
//     tile (layer l+1, row block r)  waits for  counter(layer l, row block r) == tiles per row block
// so dependent layers overlap at row-block granularity, independent chains (the encoder step, both critics, the target
// side) fill each other's gaps on all 148 SMs, and the cost of a link is one L2 round trip for the counter.
//
// The dependencies are NOT written by hand: the host records the phase as a sequential program (the order the classic
// path issues its launches), computes for every tile the rectangles of memory it reads and writes, and derives RAW / WAR
// / WAW edges between tiles from rectangle overlap (exact when both views share a leading dimension, conservative
// otherwise).  Tiles are then list-scheduled: sorted by depth in the DAG, longest remaining path first.  Since a tile
// only ever waits on tiles that precede it in the list and CTAs take tiles in list order, the kernel cannot deadlock as
// long as every CTA of the grid is resident (grid <= occupancy x SMs); a spin that lasts seconds sets an error word
// instead of hanging the GPU.
//
// Memory model: a producer's threads store, __threadfence(), barrier, then one thread bumps the counters; the consumer's
// waiting lanes ld.acquire.gpu the counters, barrier, and its operands arrive through L2 (cp.async.cg for the tiles,
// ld.global.cg for epilogue operands) -- L1 is never trusted for data produced inside the kernel.
//
// Counters are never reset: launch number e (kept in the control block, bumped by the last CTA to leave) waits for
// (e + 1) x count, compared with wrap-around arithmetic.
#pragma once
#ifndef B200_EMU
#include <map>
#include <memory>
#include <set>

#include "gemm.cuh"
#include "host.cuh"
#include "ops.cuh"

namespace b200rl {

enum MegaKind : int { kMegaFwd = 0, kMegaDgrad = 1, kMegaWgrad = 2, kMegaNormFwd = 3, kMegaNormBwd = 4, kMegaPrep = 5, kMegaGauss = 6 };

struct PrepArgs {
  StepState* st; AdamScalars* adam; const OptimHyper* hyper;
  int nopt; unsigned step_mask; int ring; int pad;
};
struct GaussArgs {
  float* out_base; long long gs; long long n;
  unsigned long long seed0;
  const int* step_counter; const int* inject;
  int stream_id; int pad;
};
union MegaDesc {
  FfmaProblem gemm;
  RowOp rows;
  PrepArgs prep;
  GaussArgs gauss;
  int words[(sizeof(FfmaProblem) + 3) / 4];
};
struct MegaWait { int flag; unsigned need; };
constexpr int kMegaInlineWaits = 4;
struct MegaTile {  // 64 bytes
  int kind, desc, t_local;
  int signal, signal2;  // counters bumped when the tile is done: its row-block group and its whole node (-1: none)
  int nwait, wait_first;  // waits beyond the inline ones live in MegaArgs::waits[wait_first + (i - kMegaInlineWaits)]
  int pad;
  MegaWait w[kMegaInlineWaits];  // the first waits travel with the tile record: no dependent load before polling
};
// every node's descriptor travels in kernel-parameter (constant) memory: the tile bodies address their problem through
// the constant bank like the classic kernels do, instead of spending registers on a shared-memory copy
constexpr int kMegaMaxGemm = 88, kMegaMaxRows = 36;
struct MegaDescTable {  // < 32 KB of kernel parameters; a tile's `desc` indexes the array of its kind
  FfmaProblem gemm[kMegaMaxGemm];
  RowOp rows[kMegaMaxRows];
  PrepArgs prep;
  GaussArgs gauss[2];
};
static_assert(sizeof(MegaDescTable) + 64 <= 32000, "descriptor table must fit the kernel-parameter space");
struct MegaArgs {
  const MegaTile* tiles;
  const MegaWait* waits;
  unsigned* flags;  // [groups][nflags]
  unsigned* ctrl;   // [0] next tile, [1] CTAs that have left, [2] launches completed, [3] error word
  int ntiles, nflags, groups, pad;
};

constexpr int kMegaRowsPerTile = 32;  // rows of an AvgL1Norm tile (8 warps x 4)
constexpr int kMegaGaussPairs = 256;  // Box-Muller pairs per noise tile

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void gauss_fill_pair(const GaussArgs& a, int g, long long pair) {
  if (__ldcg(a.inject)) return;  // the buffer holds caller-provided noise (parity tests)
  if (pair * 2 >= a.n) return;
  float* out = a.out_base + (long long)g * a.gs;
  Philox ph(a.seed0 + 0x9E3779B97F4A7C15ULL * (unsigned long long)g);
  uint32_t r[4];
  ph(((unsigned long long)(unsigned)__ldcg(a.step_counter) << 8) | (unsigned)a.stream_id, (unsigned long long)pair, r);
  float z0, z1;
  gauss2_from_bits(r[0], r[1], &z0, &z1);
  out[pair * 2] = z0;
  if (pair * 2 + 1 < a.n) out[pair * 2 + 1] = z1;
}

#ifdef B200RL_PROBE
// profiling builds: per tile {ticket taken, dependencies met, computed, signalled} (global timer, ns) and {kind, SM}
constexpr int kMegaTraceTiles = 4096, kMegaTracePrograms = 8;
__device__ unsigned long long g_mega_trace[kMegaTracePrograms][kMegaTraceTiles][5];
__device__ __forceinline__ unsigned long long mega_now() {
  unsigned long long n;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n));
  return n;
}
__device__ __forceinline__ unsigned mega_smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}
#define MEGA_TRACE(slot) if (tid == 0 && a.pad < kMegaTracePrograms && t < kMegaTraceTiles) g_mega_trace[a.pad][t][slot] = mega_now();
#else
#define MEGA_TRACE(slot)
#endif

// (each mode gets its own register allocation: inlined into one body the three tile bodies spill)
template <int MODE, bool MMA>
__device__ __forceinline__ void mega_gemm_tile(const FfmaProblem& sp, int t_local, int g, float* smem) {
  gemm_tile<MODE, MMA, true>(sp, t_local, g, 0, 1, smem);
}

template <bool MMA>
__global__ void __launch_bounds__(kGemmThreads, 2) mega_kernel(const __grid_constant__ MegaArgs a, const __grid_constant__ MegaDescTable D) {
  B200_DYN_SMEM(float, smem);
  __shared__ MegaTile cur[2];  // double-buffered: the record of the NEXT tile is fetched while the current one computes
  __shared__ int next_v;
  const int tid = threadIdx.x;
  const int total = a.ntiles * a.groups;
  const unsigned epoch1 = __ldcg(a.ctrl + 2) + 1u;  // stable during the launch: only the last CTA to leave bumps it
  if (tid < 16) {
    int v0 = 0;
    if (tid == 0) v0 = (int)atomicAdd(a.ctrl, 1u);
    v0 = __shfl_sync(0xffffu, v0, 0);
    if (tid == 0) next_v = v0;
    if (v0 < total) reinterpret_cast<int*>(&cur[0])[tid] = __ldg(reinterpret_cast<const int*>(a.tiles + v0 / a.groups) + tid);
  }
  __syncthreads();
  int v = next_v, buf = 0;
  while (v < total) {
    const int t = v / a.groups, g = v - t * a.groups;
    (void)t;
    MEGA_TRACE(0)
    const MegaTile& tl = cur[buf];
    if (tid >= 32 && tid < 48) {  // the next ticket and its tile record, off the critical path (warp 1)
      const int l = tid - 32;
      int nv = 0;
      if (l == 0) nv = (int)atomicAdd(a.ctrl, 1u);
      nv = __shfl_sync(0xffffu, nv, 0);
      if (l == 0) next_v = nv;
      if (nv < total) reinterpret_cast<int*>(&cur[buf ^ 1])[l] = __ldg(reinterpret_cast<const int*>(a.tiles + nv / a.groups) + l);
    }
    if (tid < tl.nwait) {
      const MegaWait w = tid < kMegaInlineWaits ? tl.w[tid] : a.waits[tl.wait_first + tid - kMegaInlineWaits];
      const unsigned* f = a.flags + (size_t)g * a.nflags + w.flag;
      const unsigned target = epoch1 * w.need;
      long long spins = 0;
      while ((int)(ld_acquire_u32(f) - target) < 0) {
        ++spins;
        // about a second of polling is a scheduling bug, not a slow producer: report it in the control block and let
        // every later wait fall through instead of hanging the GPU
        if ((spins & 1023) == 0 && (spins > (1ll << 21) || __ldcg(a.ctrl + 3) != 0u)) {
          atomicCAS(a.ctrl + 3, 0u, 0x80000000u | (unsigned)t);
          break;
        }
      }
    }
    __syncthreads();
    MEGA_TRACE(1)
    const int kind = tl.kind, t_local = tl.t_local, di = tl.desc;
    switch (kind) {
      case kMegaFwd: mega_gemm_tile<kFwd, MMA>(D.gemm[di], t_local, g, smem); break;
#ifndef MEGA_ONLY_FWD
      case kMegaDgrad: mega_gemm_tile<kDgrad, MMA>(D.gemm[di], t_local, g, smem); break;
      case kMegaWgrad: mega_gemm_tile<kWgrad, MMA>(D.gemm[di], t_local, g, smem); break;
#endif
      case kMegaNormFwd:
      case kMegaNormBwd: {
        const int warp = tid >> 5, lane = tid & 31;
        const RowOp& op = D.rows[di];
#pragma unroll 1
        for (int k = 0; k < kMegaRowsPerTile / 8; ++k) {
          const int row = t_local * kMegaRowsPerTile + k * 8 + warp;
          if (row < op.rows) {
            if (kind == kMegaNormFwd) avgl1norm_row<false, true>(op, row, g, lane);
            else avgl1norm_row<true, true>(op, row, g, lane);
          }
        }
        break;
      }
      case kMegaPrep:
        if (tid == 0 && g == 0) prep_step_body(D.prep.st, D.prep.adam, D.prep.hyper, D.prep.nopt, D.prep.step_mask, D.prep.ring);
        break;
      case kMegaGauss: gauss_fill_pair(D.gauss[di], g, (long long)t_local * kMegaGaussPairs + tid); break;
      default: break;
    }
    MEGA_TRACE(2)
    // publish: the barrier orders every thread's stores before thread 0's device-scope fence, which is cumulative
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      unsigned* f = a.flags + (size_t)g * a.nflags;
      if (tl.signal >= 0) atomicAdd(f + tl.signal, 1u);
      if (tl.signal2 >= 0) atomicAdd(f + tl.signal2, 1u);
    }
    MEGA_TRACE(3)
#ifdef B200RL_PROBE
    if (tid == 0 && a.pad < kMegaTracePrograms && t < kMegaTraceTiles) g_mega_trace[a.pad][t][4] = ((unsigned long long)kind << 32) | ((unsigned long long)tl.desc << 12) | mega_smid();
#endif
    __syncthreads();  // (cur[buf] and the pipeline memory are free; next_v / cur[buf ^ 1] were written before the barrier above)
    v = next_v;
    buf ^= 1;
  }
  // the last CTA to leave re-arms the control block for the next launch
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(a.ctrl + 1, 1u) == gridDim.x - 1) {
      a.ctrl[0] = 0u;
      a.ctrl[1] = 0u;
      __threadfence();
      atomicAdd(a.ctrl + 2, 1u);
    }
  }
}

// ======================================================================================================== host side
struct MegaRect {  // a 2-D view of memory in bytes: rows [r0, r1) x byte columns [c0, c1) of a matrix at `base` with row pitch `ld`
  uintptr_t base;
  long long ld;
  long long r0, r1, c0, c1;
  uintptr_t lo() const { return base + (uintptr_t)(r0 * ld + c0); }
  uintptr_t hi() const { return base + (uintptr_t)((r1 - 1) * ld + c1); }
};
inline bool mega_overlap(const MegaRect& a, const MegaRect& b) {
  if (a.hi() <= b.lo() || b.hi() <= a.lo()) return false;
  if (a.ld == b.ld && a.ld > 0) {
    const long long d = (long long)a.base - (long long)b.base;
    long long rs = d / a.ld, cs = d - rs * a.ld;
    if (cs < 0) { cs += a.ld; rs -= 1; }
    if (a.c1 + cs <= a.ld && b.c1 <= b.ld)  // no wrap: exact rectangle test in b's coordinates
      return a.r0 + rs < b.r1 && b.r0 < a.r1 + rs && a.c0 + cs < b.c1 && b.c0 < a.c1 + cs;
  }
  return true;
}
inline MegaRect mega_rect(const void* p, long long ld_elems, long long r0, long long r1, long long c0, long long c1, int esize = 4) {
  MegaRect r;
  r.base = reinterpret_cast<uintptr_t>(p);
  r.ld = ld_elems * esize;
  r.r0 = r0; r.r1 = r1; r.c0 = c0 * esize; r.c1 = c1 * esize;
  return r;
}
inline MegaRect mega_span(const void* p, long long bytes) {
  MegaRect r;
  r.base = reinterpret_cast<uintptr_t>(p);
  r.ld = 0; r.r0 = 0; r.r1 = 1; r.c0 = 0; r.c1 = bytes;
  return r;
}

struct MegaHostTile {
  int kind, node, t_local, group;
  std::vector<MegaRect> reads, writes;
  uintptr_t rlo = ~(uintptr_t)0, rhi = 0, wlo = ~(uintptr_t)0, whi = 0;
  std::vector<int> deps;  // tile indices (program order)
  int depth = 0, height = 0;
  void bounds() {
    for (auto& r : reads) { rlo = std::min(rlo, r.lo()); rhi = std::max(rhi, r.hi()); }
    for (auto& r : writes) { wlo = std::min(wlo, r.lo()); whi = std::max(whi, r.hi()); }
  }
};
struct MegaHostNode {
  int first_tile = 0, ntiles = 0, ngroups = 0;
  std::vector<int> group_size;
  int flag0 = 0;  // counters: flag0 + group, and flag0 + ngroups for the whole node
  std::vector<MegaRect> reads, writes;  // per view (base, pitch): bounding rectangle over the node's tiles
};

struct MegaProgram {
  std::vector<MegaDesc> descs;      // one per node, in recording order
  std::vector<int> desc_kind_index;  // index of that node inside the table array of its kind
  std::vector<MegaHostTile> tiles;
  std::vector<MegaHostNode> nodes;
  // device copies
  MegaArgs args;
  std::unique_ptr<MegaDescTable> table;
  int grid = 0;
  bool built = false;
  int max_depth = 0;

  int open_node() {
    nodes.emplace_back();
    nodes.back().first_tile = (int)tiles.size();
    return (int)nodes.size() - 1;
  }
  MegaHostTile& add_tile(int node, int kind, int desc, int t_local, int group) {
    tiles.emplace_back();
    MegaHostTile& t = tiles.back();
    t.kind = kind; t.node = node; t.t_local = t_local; t.group = group;
    (void)desc;
    return t;
  }
  void close_node(int node) {
    MegaHostNode& n = nodes[node];
    n.ntiles = (int)tiles.size() - n.first_tile;
    n.ngroups = 0;
    for (int i = n.first_tile; i < (int)tiles.size(); ++i) n.ngroups = std::max(n.ngroups, tiles[i].group + 1);
    n.group_size.assign(n.ngroups, 0);
    auto merge = [](std::vector<MegaRect>& into, const MegaRect& r) {
      for (auto& m : into)
        if (m.base == r.base && m.ld == r.ld) {
          m.r0 = std::min(m.r0, r.r0); m.r1 = std::max(m.r1, r.r1); m.c0 = std::min(m.c0, r.c0); m.c1 = std::max(m.c1, r.c1);
          return;
        }
      into.push_back(r);
    };
    for (int i = n.first_tile; i < (int)tiles.size(); ++i) {
      tiles[i].bounds();
      n.group_size[tiles[i].group]++;
      for (auto& r : tiles[i].reads) merge(n.reads, r);
      for (auto& w : tiles[i].writes) merge(n.writes, w);
    }
  }
  static bool node_conflict(const MegaHostNode& later, const MegaHostNode& earlier) {
    for (auto& r : later.reads) for (auto& w : earlier.writes) if (mega_overlap(r, w)) return true;
    for (auto& w : later.writes) for (auto& r : earlier.reads) if (mega_overlap(w, r)) return true;
    for (auto& w : later.writes) for (auto& x : earlier.writes) if (mega_overlap(w, x)) return true;
    return false;
  }

  // ---- recording ------------------------------------------------------------------------------------------------
  std::vector<int> tile_desc;  // desc index per tile (parallel to `tiles`)

  void add_gemm_table(const GemmTable& t) {
    for (int pi = 0; pi < t.launch.nprob; ++pi) {
      FfmaProblem p = t.launch.p[pi];
      p.per_split = p.total_chunks;  // no split-K inside the persistent kernel
      MegaDesc d;
      memset(&d, 0, sizeof(d));
      d.gemm = p;
      const int di = (int)descs.size();
      descs.push_back(d);
      const int node = open_node();
      const int tiles_i = ceil_div(p.I, kTile), tiles_j = ceil_div(p.J, kTile);
      for (int ti = 0; ti < tiles_i; ++ti)
        for (int tj = 0; tj < tiles_j; ++tj) {
          MegaHostTile& T = add_tile(node, t.mode, di, ti * tiles_j + tj, ti);
          tile_desc.push_back(di);
          const long long i0 = ti * kTile, i1 = std::min<long long>(p.I, i0 + kTile), j0 = tj * kTile, j1 = std::min<long long>(p.J, j0 + kTile);
          for (int s = 0; s < p.nseg; ++s) {
            const GemmSeg& sg = p.seg[s];
            if (t.mode == kFwd) {
              T.reads.push_back(mega_rect(sg.A, sg.a_ld, i0, i1, 0, sg.len));
              T.reads.push_back(mega_rect(sg.B, sg.b_ld, j0, j1, 0, sg.len));
            } else if (t.mode == kDgrad) {
              T.reads.push_back(mega_rect(sg.A, sg.a_ld, i0, i1, 0, sg.len));
              T.reads.push_back(mega_rect(sg.B, sg.b_ld, 0, sg.len, j0, j1));
            } else {
              T.reads.push_back(mega_rect(sg.A, sg.a_ld, 0, sg.len, i0, i1));
              if (!sg.b_ones) T.reads.push_back(mega_rect(sg.B, sg.b_ld, 0, sg.len, j0, j1));
            }
          }
          if (t.mode == kFwd) {
            if (p.bias) T.reads.push_back(mega_rect(p.bias, 0, 0, 1, j0, j1));
            if (p.act == kActTanhNoise && p.aux) T.reads.push_back(mega_rect(p.aux, p.aux_ld, i0, i1, j0, j1));
          } else if (t.mode == kDgrad) {
            if (p.aux) T.reads.push_back(mega_rect(p.aux, p.aux_ld, i0, i1, j0, j1));
            if (p.accumulate) T.reads.push_back(mega_rect(p.C, p.c_ld, i0, i1, j0, j1));
          } else if (p.P) {
            for (const float* q : {p.P, p.M, p.V}) {
              T.reads.push_back(mega_rect(q, p.p_ld, i0, i1, j0, j1));
              T.writes.push_back(mega_rect(q, p.p_ld, i0, i1, j0, j1));
            }
            T.reads.push_back(mega_span(p.adam, sizeof(AdamScalars)));
          }
          if (p.C) T.writes.push_back(mega_rect(p.C, p.c_ld, i0, i1, j0, j1));
        }
      close_node(node);
    }
  }
  void add_row_ops(const std::vector<RowOp>& ops, bool bwd) {
    for (const RowOp& op : ops) {
      MegaDesc d;
      memset(&d, 0, sizeof(d));
      d.rows = op;
      const int di = (int)descs.size();
      descs.push_back(d);
      const int node = open_node();
      const int nt = ceil_div(op.rows, kMegaRowsPerTile);
      for (int t = 0; t < nt; ++t) {
        MegaHostTile& T = add_tile(node, bwd ? kMegaNormBwd : kMegaNormFwd, di, t, t);
        tile_desc.push_back(di);
        const long long r0 = (long long)t * kMegaRowsPerTile, r1 = std::min<long long>(op.rows, r0 + kMegaRowsPerTile);
        T.reads.push_back(mega_rect(op.x, op.x_ld, r0, r1, 0, op.n));
        if (bwd) T.reads.push_back(mega_rect(op.dy, op.dy_ld, r0, r1, 0, op.n));
        T.writes.push_back(mega_rect(op.y, op.y_ld, r0, r1, 0, op.n));
      }
      close_node(node);
    }
  }
  void add_prep(const PrepArgs& pa) {
    MegaDesc d;
    memset(&d, 0, sizeof(d));
    d.prep = pa;
    const int di = (int)descs.size();
    descs.push_back(d);
    const int node = open_node();
    MegaHostTile& T = add_tile(node, kMegaPrep, di, 0, 0);
    tile_desc.push_back(di);
    T.reads.push_back(mega_span(pa.st, sizeof(StepState)));
    T.writes.push_back(mega_span(pa.st, sizeof(StepState)));
    T.writes.push_back(mega_span(pa.adam, (long long)pa.nopt * sizeof(AdamScalars)));
    close_node(node);
  }
  void add_gauss(const GaussArgs& ga) {
    MegaDesc d;
    memset(&d, 0, sizeof(d));
    d.gauss = ga;
    const int di = (int)descs.size();
    descs.push_back(d);
    const int node = open_node();
    const long long pairs = (ga.n + 1) / 2;
    const int nt = (int)((pairs + kMegaGaussPairs - 1) / kMegaGaussPairs);
    for (int t = 0; t < nt; ++t) {
      MegaHostTile& T = add_tile(node, kMegaGauss, di, t, 0);
      tile_desc.push_back(di);
      T.reads.push_back(mega_span(ga.step_counter, sizeof(int)));
      const long long e0 = (long long)t * kMegaGaussPairs * 2, e1 = std::min<long long>(ga.n, e0 + kMegaGaussPairs * 2);
      T.writes.push_back(mega_span(ga.out_base + e0, (e1 - e0) * 4));
    }
    close_node(node);
  }
  bool empty() const { return tiles.empty(); }

  // ---- dependency analysis + list scheduling + upload --------------------------------------------------------------
  static bool conflict(const MegaHostTile& later, const MegaHostTile& earlier) {
    // RAW: later reads what earlier wrote; WAR: later writes what earlier read; WAW
    if (later.rlo < earlier.whi && earlier.wlo < later.rhi)
      for (auto& r : later.reads) for (auto& w : earlier.writes) if (mega_overlap(r, w)) return true;
    if (later.wlo < earlier.rhi && earlier.rlo < later.whi)
      for (auto& w : later.writes) for (auto& r : earlier.reads) if (mega_overlap(w, r)) return true;
    if (later.wlo < earlier.whi && earlier.wlo < later.whi)
      for (auto& w : later.writes) for (auto& x : earlier.writes) if (mega_overlap(w, x)) return true;
    return false;
  }

  int build(DeviceArena* arena, int groups, int n_sm) {
    const int nt = (int)tiles.size();
    B200_REQUIRE(nt > 0, "mega: empty program");
    // flags: per node one counter per row-block group + one for the whole node
    int nflags = 0;
    for (auto& n : nodes) { n.flag0 = nflags; nflags += n.ngroups + 1; }
    // waits per tile as (node, group) sets; group == -1: the whole node
    std::vector<std::vector<std::pair<int, int>>> waits(nt);
    for (int nb = 0; nb < (int)nodes.size(); ++nb) {
      const MegaHostNode& B = nodes[nb];
      for (int na = 0; na < nb; ++na) {
        const MegaHostNode& A = nodes[na];
        if (!node_conflict(B, A)) continue;
        for (int tb = B.first_tile; tb < B.first_tile + B.ntiles; ++tb) {
          std::set<int> groups_hit;
          for (int ta = A.first_tile; ta < A.first_tile + A.ntiles; ++ta) {
            if (groups_hit.count(tiles[ta].group)) continue;
            if (conflict(tiles[tb], tiles[ta])) groups_hit.insert(tiles[ta].group);
          }
          if (groups_hit.empty()) continue;
          if ((int)groups_hit.size() == A.ngroups && A.ngroups > 1) waits[tb].push_back({na, -1});
          else for (int gq : groups_hit) waits[tb].push_back({na, gq});
          for (int ta = A.first_tile; ta < A.first_tile + A.ntiles; ++ta)
            if (groups_hit.count(tiles[ta].group)) tiles[tb].deps.push_back(ta);
        }
      }
    }
    // prune waits implied by others: if tile T waits on node A entirely and also on node C, and EVERY tile of A depends
    // (transitively, at node level) on all of C, the wait on C is redundant.  Cheap node-level closure.
    const int nn = (int)nodes.size();
    std::vector<std::vector<char>> full(nn, std::vector<char>(nn, 0));  // full[b][a]: every tile of b waits (transitively) for all of a
    for (int b = 0; b < nn; ++b) {
      for (int a2 = 0; a2 < b; ++a2) {
        bool all = nodes[b].ntiles > 0;
        for (int tb = nodes[b].first_tile; tb < nodes[b].first_tile + nodes[b].ntiles && all; ++tb) {
          bool ok = false;
          for (auto& w : waits[tb]) {
            if (w.first == a2 && (w.second == -1 || nodes[a2].ngroups == 1)) { ok = true; break; }
            if (w.first != a2 && (w.second == -1 || nodes[w.first].ngroups == 1) && full[w.first][a2]) { ok = true; break; }
          }
          all = ok;
        }
        full[b][a2] = all;
      }
    }
    for (int tb = 0; tb < nt; ++tb) {
      auto& W = waits[tb];
      std::vector<std::pair<int, int>> keep;
      for (auto& w : W) {
        bool implied = false;
        for (auto& o : W) {
          if (&o == &w || o.first == w.first) continue;
          if ((o.second == -1 || nodes[o.first].ngroups == 1) && full[o.first][w.first]) { implied = true; break; }
        }
        if (!implied) keep.push_back(w);
      }
      W.swap(keep);
      B200_REQUIRE(W.size() <= 32, "mega: a tile waits on more than 32 counters");
    }
    // depth / height over the (unpruned) tile DAG
    for (int t = 0; t < nt; ++t) {
      int d = 0;
      for (int p : tiles[t].deps) d = std::max(d, tiles[p].depth + 1);
      tiles[t].depth = d;
      max_depth = std::max(max_depth, d);
    }
    for (int t = nt - 1; t >= 0; --t)
      for (int p : tiles[t].deps) tiles[p].height = std::max(tiles[p].height, tiles[t].height + 1);
    std::vector<int> order(nt);
    for (int i = 0; i < nt; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
      if (tiles[x].depth != tiles[y].depth) return tiles[x].depth < tiles[y].depth;
      return tiles[x].height > tiles[y].height;
    });
    // device arrays
    std::vector<MegaTile> dt(nt);
    std::vector<MegaWait> dw;
    for (int pos = 0; pos < nt; ++pos) {
      const int t = order[pos];
      const MegaHostTile& T = tiles[t];
      const MegaHostNode& N = nodes[T.node];
      MegaTile& o = dt[pos];
      memset(&o, 0, sizeof(o));
      o.kind = T.kind; o.desc = tile_desc[t]; o.t_local = T.t_local;
      o.signal = N.flag0 + T.group;
      o.signal2 = N.ngroups > 1 ? N.flag0 + N.ngroups : -1;
      o.wait_first = (int)dw.size();
      int nw = 0;
      for (auto& w : waits[t]) {
        const MegaHostNode& Q = nodes[w.first];
        MegaWait mw;
        if (w.second == -1) { mw.flag = Q.flag0 + Q.ngroups; mw.need = (unsigned)Q.ntiles; }
        else { mw.flag = Q.flag0 + w.second; mw.need = (unsigned)Q.group_size[w.second]; }
        if (nw < kMegaInlineWaits) o.w[nw] = mw;
        else dw.push_back(mw);
        ++nw;
      }
      o.nwait = nw;
    }
    if (dw.empty()) dw.push_back(MegaWait{0, 0});
    table = std::make_unique<MegaDescTable>();
    memset(table.get(), 0, sizeof(MegaDescTable));
    {
      // node -> kind from its first tile; place every descriptor in the array of its kind
      std::vector<int> node_kind(descs.size(), -1);
      for (int t = 0; t < nt; ++t) node_kind[tile_desc[t]] = tiles[t].kind;
      desc_kind_index.assign(descs.size(), 0);
      int ng = 0, nr = 0, ngs = 0, np = 0;
      for (size_t i = 0; i < descs.size(); ++i) {
        const int k = node_kind[i];
        if (k <= kMegaWgrad) {
          B200_REQUIRE(ng < kMegaMaxGemm, "mega: too many GEMM nodes for one launch (the descriptor table travels in kernel parameters)");
          table->gemm[ng] = descs[i].gemm; desc_kind_index[i] = ng++;
        } else if (k == kMegaNormFwd || k == kMegaNormBwd) {
          B200_REQUIRE(nr < kMegaMaxRows, "mega: too many row-op nodes for one launch");
          table->rows[nr] = descs[i].rows; desc_kind_index[i] = nr++;
        } else if (k == kMegaPrep) {
          B200_REQUIRE(np < 1, "mega: one counter-preparation node per launch");
          table->prep = descs[i].prep; desc_kind_index[i] = np++;
        } else {
          B200_REQUIRE(ngs < 2, "mega: at most two noise nodes per launch");
          table->gauss[ngs] = descs[i].gauss; desc_kind_index[i] = ngs++;
        }
      }
    }
    for (int pos = 0; pos < nt; ++pos) dt[pos].desc = desc_kind_index[dt[pos].desc];
    MegaTile* d_tiles; MegaWait* d_waits; unsigned* d_flags; unsigned* d_ctrl;
    B200_TRY(arena->alloc(&d_tiles, dt.size()));
    B200_TRY(arena->alloc(&d_waits, dw.size()));
    B200_TRY(arena->alloc(&d_flags, (size_t)groups * nflags));
    B200_TRY(arena->alloc(&d_ctrl, 8));
    B200_CHECK_CUDA(cudaMemcpy(d_tiles, dt.data(), dt.size() * sizeof(MegaTile), cudaMemcpyHostToDevice));
    B200_CHECK_CUDA(cudaMemcpy(d_waits, dw.data(), dw.size() * sizeof(MegaWait), cudaMemcpyHostToDevice));
    args.tiles = d_tiles; args.waits = d_waits; args.flags = d_flags; args.ctrl = d_ctrl;
    static int program_counter = 0;
    args.ntiles = nt; args.nflags = nflags; args.groups = groups; args.pad = program_counter++;  // (profiling builds: trace slot)
    grid = std::min(nt * groups, 2 * n_sm);
    built = true;
    if (getenv("B200RL_DEBUG_TABLES")) {
      size_t nw = 0;
      for (auto& w : waits) nw += w.size();
      fprintf(stderr, "[b200rl] mega program: %zu nodes, %d tiles, %d counters, %zu waits, depth %d, grid %d\n", nodes.size(), nt, nflags, nw, max_depth + 1, grid);
    }
    return kOk;
  }

  int launch(cudaStream_t stream) const {
    static bool attr = false;
    if (!attr) {
      B200_CHECK_CUDA(cudaFuncSetAttribute(mega_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPipeBytes));
      attr = true;
    }
    B200_LAUNCH(mega_kernel<false>, dim3((unsigned)grid), dim3(kGemmThreads), (size_t)kPipeBytes, stream, args, *table);
    B200_CHECK_LAUNCH();
    return kOk;
  }
};

}  // namespace b200rl
#endif  // !B200_EMU
