// replay.cuh -- GPU-resident ring replay buffer with LAP (prioritized) and uniform sampling.
//
// Restates rl/replay_memory/lap.py and rl/replay_memory/simple.py of the reference:
//   sample  (lap.py:45-64):  P = cumsum(priority[:size]) in float64, each prefix rounded to
//           float32; v = u * P[-1] (float32); ind = first i with P[i] >= v; gather 5 arrays.
//   update_priority (lap.py:66-69), reset_max_priority (lap.py:71-73), append (lap.py:31-43).
//
// HBM layout: one packed fp32 row per transition; both observation vectors start on a 16-byte boundary and the row is
// a multiple of 4 floats, so a warp reads them with coalesced 128-bit loads:
//     [ state(S) pad4 | next_state(S) pad4 | action(A) | reward | notdone | pad4 ]      (RowLayout below)
// (fp32 storage is value-identical to the reference's fp64-store + fp32-cast).
// priority: fp32[capacity] (+pad) per stacked seed; max_priority: one fp32 per seed, on device.
//
// Sampling never materialises the 1M-entry prefix array.  A three-level sum tree is kept in HBM:
//     priority[cap]  ->  tile_sums[cap/256] (fp64)  ->  super_sums[cap/16384] (fp64)
// and every kernel that writes priorities (append, the LAP write-back inside the critic-loss
// kernel, update_priority) repairs the <= B touched tiles and their super-tiles right away
// (lap_fix_sums_kernel / the tail of lap_append_priority_kernel), so a draw costs no pass over
// the priorities at all:
//   lap_sample_kernel: one warp per draw walks the tree top-down -- warp scan of the super sums,
//   of the 64 tile sums of the chosen super-tile, of the 256 priorities of the chosen tile; every
//   comparison is made on the fp64 prefix rounded to fp32, exactly like the reference's fp32
//   prefix array -- then copies the chosen row out with 128-bit loads.
//   lap_tile_sums_kernel + lap_super_sums_kernel rebuild the tree from scratch (bulk loads,
//   restored checkpoints).
// Priorities are fp32 values >= 1 (or 0), so fp64 partial sums are EXACT for any summation
// order while max_priority * size < 2^30; the sampled indices are then bit-identical to
// torch.cumsum + torch.searchsorted on the CPU, and an incrementally repaired tree is
// bit-identical to a rebuilt one.
#pragma once
#include "host.cuh"

namespace b200rl {

constexpr int kPrioTile = 256;       // priorities per tile (one warp: 32 lanes x 8)
constexpr int kPrioTilesPerCta = 8;  // 2048 priorities per CTA in the rebuild kernel
constexpr int kSuperTiles = 64;      // tiles per super-tile (one warp: 32 lanes x 2)
constexpr int kSmallRowFloats = 128; // packed rows up to this many floats are gathered through one 128-bit load per lane

__device__ __forceinline__ float4 ld_cg4(const float* p) {
#ifdef B200_EMU
  return *reinterpret_cast<const float4*>(p);
#else
  return __ldcg(reinterpret_cast<const float4*>(p));
#endif
}
__device__ __forceinline__ double ld_cgd(const double* p) {
#ifdef B200_EMU
  return *p;
#else
  return __ldcg(p);
#endif
}
__device__ __forceinline__ double shfl_double(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// exact fp64 sum of one whole tile (entries past `size` are zero by construction), by one warp
__device__ __forceinline__ double tile_sum_warp(const float* tile_base, int lane) {
  const float4 a = ld_cg4(tile_base + lane * 4);
  const float4 b = ld_cg4(tile_base + 128 + lane * 4);
  const double s = ((double)a.x + (double)a.y) + ((double)a.z + (double)a.w) + (((double)b.x + (double)b.y) + ((double)b.z + (double)b.w));
  return warp_sum(s);
}
__device__ __forceinline__ double super_sum_warp(const double* first_tile, int lane) {
  return warp_sum(ld_cgd(first_tile + lane) + ld_cgd(first_tile + 32 + lane));
}

__global__ void __launch_bounds__(256) lap_tile_sums_kernel(const float* __restrict__ priority, long long prio_gs, int size,
                                                            double* __restrict__ tile_sums, int ntiles, long long ts_gs) {
  B200_PDL_PROLOGUE();
  const int g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * kPrioTilesPerCta + warp;
  if (tile >= ntiles) return;
  const float* p = priority + (long long)g * prio_gs + (long long)tile * kPrioTile;
  const int base = tile * kPrioTile;
  // two coalesced float4 per lane (the array is padded to a multiple of 2048 entries)
  const float4 a = *reinterpret_cast<const float4*>(p + lane * 4);
  const float4 b = *reinterpret_cast<const float4*>(p + 128 + lane * 4);
  const int ia = base + lane * 4, ib = base + 128 + lane * 4;
  double s = 0.0;
  if (ia + 0 < size) s += (double)a.x;
  if (ia + 1 < size) s += (double)a.y;
  if (ia + 2 < size) s += (double)a.z;
  if (ia + 3 < size) s += (double)a.w;
  if (ib + 0 < size) s += (double)b.x;
  if (ib + 1 < size) s += (double)b.y;
  if (ib + 2 < size) s += (double)b.z;
  if (ib + 3 < size) s += (double)b.w;
  s = warp_sum(s);
  if (lane == 0) tile_sums[(long long)g * ts_gs + tile] = s;
}

__global__ void __launch_bounds__(256) lap_super_sums_kernel(const double* __restrict__ tile_sums, long long ts_gs, double* __restrict__ super_sums, long long ss_gs,
                                                             int nsuper) {
  B200_PDL_PROLOGUE();
  const int g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sup = blockIdx.x * 8 + warp;
  if (sup >= nsuper) return;
  const double s = super_sum_warp(tile_sums + (long long)g * ts_gs + (long long)sup * kSuperTiles, lane);
  if (lane == 0) super_sums[(long long)g * ss_gs + sup] = s;
}

// Repair the sum tree after up to `n` priorities changed.  One CTA per stacked seed: phase 1 recomputes the
// tile of every touched entry (duplicates rewrite the same value), phase 2 the super-tile of every touched
// entry from the repaired tile sums.  Entries: idx[i] when idx != null, else (pos + i) % cap.
struct FixArgs {
  const float* priority; long long prio_gs;
  double* tile_sums; long long ts_gs;
  double* super_sums; long long ss_gs;
  const long long* idx; long long idx_gs;
  long long pos, cap;
  int n;
};
__device__ __forceinline__ void lap_fix_block(const FixArgs& a, int g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* prio = a.priority + (long long)g * a.prio_gs;
  double* ts = a.tile_sums + (long long)g * a.ts_gs;
  double* ss = a.super_sums + (long long)g * a.ss_gs;
  const long long* idx = a.idx ? a.idx + (long long)g * a.idx_gs : nullptr;
  if (idx) {
    for (int i = warp; i < a.n; i += nw) {
      const long long tile = idx[i] / kPrioTile;
      const double s = tile_sum_warp(prio + tile * kPrioTile, lane);
      if (lane == 0) ts[tile] = s;
    }
  } else {  // a contiguous (possibly wrapping) range: visit every covered tile once
    for (int part = 0; part < 2; ++part) {
      const long long lo = part == 0 ? a.pos : 0, hi = part == 0 ? (a.pos + a.n < a.cap ? a.pos + a.n : a.cap) : a.pos + a.n - a.cap;  // [lo, hi)
      if (hi <= lo) continue;
      for (long long tile = lo / kPrioTile + warp; tile <= (hi - 1) / kPrioTile; tile += nw) {
        const double s = tile_sum_warp(prio + tile * kPrioTile, lane);
        if (lane == 0) ts[tile] = s;
      }
    }
  }
  __threadfence_block();
  __syncthreads();
  if (idx) {
    for (int i = warp; i < a.n; i += nw) {
      const long long sup = idx[i] / (kPrioTile * kSuperTiles);
      const double s = super_sum_warp(ts + sup * kSuperTiles, lane);
      if (lane == 0) ss[sup] = s;
    }
  } else {
    for (int part = 0; part < 2; ++part) {
      const long long lo = part == 0 ? a.pos : 0, hi = part == 0 ? (a.pos + a.n < a.cap ? a.pos + a.n : a.cap) : a.pos + a.n - a.cap;
      if (hi <= lo) continue;
      for (long long sup = lo / (kPrioTile * kSuperTiles) + warp; sup <= (hi - 1) / (kPrioTile * kSuperTiles); sup += nw) {
        const double s = super_sum_warp(ts + sup * kSuperTiles, lane);
        if (lane == 0) ss[sup] = s;
      }
    }
  }
}
__global__ void __launch_bounds__(1024) lap_fix_sums_kernel(FixArgs a) {
  B200_PDL_PROLOGUE();
  lap_fix_block(a, blockIdx.x);
}

// offsets (floats) of the five fields inside a packed row
struct RowLayout {
  int state, next_state, action, reward, notdone, stride;
  static RowLayout make(int S, int A) {
    RowLayout L;
    L.state = 0;
    L.next_state = (int)((S + 3) / 4 * 4);
    L.action = 2 * L.next_state;
    L.reward = L.action + A;
    L.notdone = L.reward + 1;
    L.stride = (L.notdone + 1 + 3) / 4 * 4;
    return L;
  }
};

struct SampleArgs {
  const float* priority; long long prio_gs;
  const double* tile_sums; long long ts_gs; int ntiles;
  const double* super_sums; long long ss_gs; int nsuper;
  int size;
  const float* u; long long u_gs;          // injected uniforms [B] (null -> Philox)
  unsigned long long seed, counter;        // Philox stream when u == null
  const int* counter_ptr;                  // optional device-side counter added to `counter` (graph replays)
  const float* rows; long long rows_gs; int row_stride; int S, A;
  RowLayout lay;
  long long* idx_out; long long idx_gs;
  float* state; int ld_state; long long state_gs;
  float* action; int ld_action; long long action_gs;
  float* reward; long long reward_gs;
  float* next_state; int ld_next; long long next_gs;
  float* done; long long done_gs;
  int B;
  int prioritized;
};

// K consecutive values per lane (lane l holds elements l*K .. l*K+K-1 of a 32*K run whose exclusive prefix
// is `base`): position of the first element whose fp32-rounded inclusive prefix reaches v (-1: none), and
// the exact exclusive prefix of that element.
template <int K>
__device__ __forceinline__ int warp_first_reaching(const double (&x)[K], double base, float v, int lane, double* excl_out) {
  double c[K];
  double run = 0.0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    run += x[k];
    c[k] = run;
  }
  const double incl = warp_inclusive_scan(run, lane);
  const double excl = base + (incl - run);
  int kfirst = K;
#pragma unroll
  for (int k = K - 1; k >= 0; --k)
    if ((float)(excl + c[k]) >= v) kfirst = k;
  const unsigned hit = __ballot_sync(0xffffffffu, kfirst < K);
  double before = excl;
#pragma unroll
  for (int k = 1; k < K; ++k)
    if (kfirst == k) before = excl + c[k - 1];
  if (hit == 0u) {
    *excl_out = base;
    return -1;
  }
  const int src = __ffs((int)hit) - 1;
  const int kk = __shfl_sync(0xffffffffu, kfirst, src);
  *excl_out = shfl_double(before, src);
  return src * K + kk;
}

// One warp copies `len` floats of a replay row (read-only for the whole launch: ld.global.nc) to a batch tensor row.
// Every load of a batch is issued before the first store so a lane keeps up to four 16-byte requests in flight;
// a 16-byte aligned destination gets 128-bit stores (the source side then reads 128-bit when it is aligned too, else
// four 32-bit loads of the same sectors), anything else goes lane-per-float, coalesced on both sides.
__device__ __forceinline__ void warp_copy_row_segment(float* __restrict__ dst, const float* __restrict__ src, int len, int lane) {
  if ((reinterpret_cast<unsigned long long>(dst) & 15ull) == 0ull && len >= 64) {
    const int n4 = len >> 2;
    const bool src_al = (reinterpret_cast<unsigned long long>(src) & 15ull) == 0ull;
    for (int q0 = 0; q0 < n4; q0 += 128) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = q0 + j * 32 + lane;
        if (q < n4) {
          if (src_al) {
            v[j] = __ldg(reinterpret_cast<const float4*>(src) + q);
          } else {
            v[j].x = __ldg(src + 4 * q);
            v[j].y = __ldg(src + 4 * q + 1);
            v[j].z = __ldg(src + 4 * q + 2);
            v[j].w = __ldg(src + 4 * q + 3);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = q0 + j * 32 + lane;
        if (q < n4) reinterpret_cast<float4*>(dst)[q] = v[j];
      }
    }
    const int i = n4 * 4 + lane;
    if (i < len) dst[i] = __ldg(src + i);
    return;
  }
  for (int i0 = 0; i0 < len; i0 += 128) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * 32 + lane;
      if (i < len) v[j] = __ldg(src + i);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * 32 + lane;
      if (i < len) dst[i] = v[j];
    }
  }
}

template <int kMinCtas>
__global__ void __launch_bounds__(256, kMinCtas) lap_sample_kernel(SampleArgs a) {
  B200_PDL_PROLOGUE();
  const int g = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int size = a.size;
  const int b = blockIdx.x * 8 + warp;

  // one Philox block yields four uniforms: the CTA's eight draws need two blocks, computed by two lanes and shared (every
  // warp recomputing its own block was ~10 % of the instructions of a draw on narrow rows)
  __shared__ float u_sh[8];
  __shared__ __align__(16) float row_sh[8][kSmallRowFloats];
  float u;
  if (a.u) {
    u = b < a.B ? a.u[(long long)g * a.u_gs + b] : 0.f;
  } else {
    if (tid < 2) {
      Philox ph(a.seed);
      uint32_t r[4];
      const unsigned long long ctr = a.counter + (a.counter_ptr ? (unsigned long long)(unsigned)(*a.counter_ptr) : 0ULL);
      ph(ctr, ((unsigned long long)g << 32) | (unsigned)(blockIdx.x * 2 + tid), r);
#pragma unroll
      for (int k = 0; k < 4; ++k) u_sh[tid * 4 + k] = u01_from_bits(r[k]);
    }
    __syncthreads();
    u = u_sh[warp];
  }
  if (b >= a.B) return;

  long long idx;
  if (a.prioritized) {
    // ---- level 0: super-tiles.  Lane l owns the contiguous run [l*per, l*per+per) of super sums.
    const double* ss = a.super_sums + (long long)g * a.ss_gs;
    const int nsup = a.nsuper;
    const int per = (nsup + 31) >> 5;
    const int s_lo = min(lane * per, nsup), s_hi = min(s_lo + per, nsup);
    // one pass over the super sums: up to 4 per lane stay in registers (capacities up to 2 M rows); beyond that the
    // second loop reads them again
    double sv[4] = {0.0, 0.0, 0.0, 0.0};
    double run = 0.0;
    if (per <= 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (s_lo + k < s_hi) sv[k] = ld_cgd(ss + s_lo + k);
      run = ((sv[0] + sv[1]) + sv[2]) + sv[3];
    } else {
      for (int t = s_lo; t < s_hi; ++t) run += ld_cgd(ss + t);
    }
    const double incl = warp_inclusive_scan(run, lane);
    const double total = shfl_double(incl, 31);
    const float v = __fmul_rn(u, (float)total);
    double acc = incl - run;  // exclusive prefix of this lane's run
    int sfirst = -1;
    double sbase = 0.0;
    if (per <= 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double nx = acc + sv[k];
        if (sfirst < 0 && s_lo + k < s_hi && (float)nx >= v) { sfirst = s_lo + k; sbase = acc; }
        acc = nx;
      }
    } else {
      for (int t = s_lo; t < s_hi; ++t) {
        const double nx = acc + ld_cgd(ss + t);
        if (sfirst < 0 && (float)nx >= v) { sfirst = t; sbase = acc; }
        acc = nx;
      }
    }
    const unsigned shit = __ballot_sync(0xffffffffu, sfirst >= 0);
    int sup = nsup - 1;
    double base = 0.0;
    if (shit != 0u) {
      const int src = __ffs((int)shit) - 1;
      sup = __shfl_sync(0xffffffffu, sfirst, src);
      base = shfl_double(sbase, src);
    }
    // ---- level 1: the 64 tiles of that super-tile
    const double* ts = a.tile_sums + (long long)g * a.ts_gs + (long long)sup * kSuperTiles;
    const double t2[2] = {ld_cgd(ts + 2 * lane), ld_cgd(ts + 2 * lane + 1)};
    double tbase;
    int tin = warp_first_reaching<2>(t2, base, v, lane, &tbase);
    if (tin < 0) { tin = min(kSuperTiles - 1, a.ntiles - 1 - sup * kSuperTiles); tbase = base; }
    const int tile = sup * kSuperTiles + tin;
    // ---- level 2: the 256 priorities of that tile
    const float* p = a.priority + (long long)g * a.prio_gs + (long long)tile * kPrioTile + lane * 8;
    const float4 x0 = ld_cg4(p);
    const float4 x1 = ld_cg4(p + 4);
    const double pv[8] = {(double)x0.x, (double)x0.y, (double)x0.z, (double)x0.w, (double)x1.x, (double)x1.y, (double)x1.z, (double)x1.w};
    double ebase;
    const int ein = warp_first_reaching<8>(pv, tbase, v, lane, &ebase);
    idx = ein >= 0 ? (long long)tile * kPrioTile + ein : (long long)size - 1;
    if (idx > size - 1) idx = size - 1;  // unreachable for u < 1; keeps the gather in range
  } else {
    // uniform: prefix is (float)(i+1); first i with (float)(i+1) >= u * (float)size
    const float v = __fmul_rn(u, (float)size);
    int lo = 0, hi = size - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((float)(mid + 1) >= v) hi = mid; else lo = mid + 1;
    }
    idx = lo;
  }
  if (lane == 0 && a.idx_out) a.idx_out[(long long)g * a.idx_gs + b] = idx;

  // ---- gather the packed row into the five batch tensors ----
  const float* row = a.rows + (long long)g * a.rows_gs + idx * (long long)a.row_stride;
  const int S = a.S, A = a.A;
  if (a.row_stride <= kSmallRowFloats) {
    // narrow rows (<= 512 B: every MuJoCo task but Humanoid): ONE 128-bit load per lane brings the whole packed row into
    // the warp's shared-memory slot, then each batch tensor row is written lane-per-float, coalesced
    float* rs = row_sh[warp];
    if (lane * 4 < a.row_stride) *reinterpret_cast<float4*>(rs + lane * 4) = __ldg(reinterpret_cast<const float4*>(row) + lane);
    __syncwarp();
    float* ds = a.state + (long long)g * a.state_gs + (long long)b * a.ld_state;
    float* dn = a.next_state + (long long)g * a.next_gs + (long long)b * a.ld_next;
    float* da = a.action + (long long)g * a.action_gs + (long long)b * a.ld_action;
    for (int i = lane; i < S; i += 32) {
      ds[i] = rs[a.lay.state + i];
      dn[i] = rs[a.lay.next_state + i];
    }
    if (lane < A) da[lane] = rs[a.lay.action + lane];
    for (int i = 32 + lane; i < A; i += 32) da[i] = rs[a.lay.action + i];
    if (lane == 0) {
      a.reward[(long long)g * a.reward_gs + b] = rs[a.lay.reward];
      a.done[(long long)g * a.done_gs + b] = rs[a.lay.notdone];
    }
    return;
  }
  if (lane == 0) {
    a.reward[(long long)g * a.reward_gs + b] = __ldg(row + a.lay.reward);
    a.done[(long long)g * a.done_gs + b] = __ldg(row + a.lay.notdone);
  }
  warp_copy_row_segment(a.state + (long long)g * a.state_gs + (long long)b * a.ld_state, row + a.lay.state, S, lane);
  warp_copy_row_segment(a.next_state + (long long)g * a.next_gs + (long long)b * a.ld_next, row + a.lay.next_state, S, lane);
  warp_copy_row_segment(a.action + (long long)g * a.action_gs + (long long)b * a.ld_action, row + a.lay.action, A, lane);
}

// (An eight-lanes-per-draw variant of the walk -- every lane summing a contiguous eighth of a node, 3-step scans, four
// draws per warp instruction -- was built for the narrow rows, where this kernel is bound by instruction issue, and
// measured on 256 stacked Ant seeds: 80.9 us against 45.3 us.  The per-element fp64 work is the same per draw either
// way, only the scans shrink, and the serial per-lane chains halve the occupancy: dropped.  profiles/README.md.)

__global__ void __launch_bounds__(1024) lap_update_priority_kernel(float* __restrict__ priority, long long prio_gs, float* __restrict__ max_priority,
                                                                   const long long* __restrict__ idx, long long idx_gs, const float* __restrict__ prio,
                                                                   long long p_gs, int B, FixArgs fix) {
  B200_PDL_PROLOGUE();
  B200_DYN_SMEM(ScatterTable, tab);
  const int g = blockIdx.x, b = threadIdx.x;
  const float p = (b < B) ? prio[(long long)g * p_gs + b] : 0.f;
  lap_scatter_block(priority + (long long)g * prio_gs, max_priority + g, idx + (long long)g * idx_gs, p, b, B, tab);
  __threadfence_block();
  __syncthreads();
  lap_fix_block(fix, g);
}

__global__ void __launch_bounds__(256) lap_max_kernel(const float* __restrict__ priority, long long prio_gs, int size, float* __restrict__ max_priority) {
  B200_PDL_PROLOGUE();
  __shared__ float red[32];
  const int g = blockIdx.y;
  const float* p = priority + (long long)g * prio_gs;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, p[i]);
  m = block_max_256(m, red);
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(max_priority + g), __float_as_int(m));
}

// new rows take the current running max as priority (lap.py:41)
__global__ void lap_append_priority_kernel(float* __restrict__ priority, const float* __restrict__ max_priority, long long pos, int count, long long cap) {
  B200_PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) priority[(pos + i) % cap] = *max_priority;
}
// the per-environment-step form (a handful of rows per stacked seed): one CTA per seed copies that seed's rows -- `src`
// may be device memory or pinned host memory read over the bus, so the step needs no separate H2D copy --, applies the
// reference's action normalisation a/scale - bias (lap.py:35; IEEE fp32 division and subtraction, i.e. exactly numpy's
// result for the fp32 actions the agents emit) when asked to, writes the priorities and repairs the sum tree.
// src == null: rows were copied by the caller; priority == null: uniform buffer; dst_gs == 0: one shared row store
// (only the first seed's CTA copies).
struct AppendArgs {
  const float* src; long long src_gs;
  float* dst; long long dst_gs;
  int row_stride, group0;
  float* priority; long long prio_gs;
  const float* max_priority;
  const float* act_scale; const float* act_bias;
  int act_off, A;  // offset of the action columns inside a row
  FixArgs fix;  // base pointers of group 0, per-group strides inside
};
__global__ void __launch_bounds__(1024) lap_append_fix_kernel(AppendArgs a) {
  B200_PDL_PROLOGUE();
  const int g = a.group0 + (int)blockIdx.x;
  const FixArgs& fix = a.fix;
  if (a.src && (a.dst_gs != 0 || blockIdx.x == 0)) {
    const float* src = a.src + (long long)blockIdx.x * a.src_gs;
    float* dst_base = a.dst + (long long)g * a.dst_gs;
    const int q4 = a.row_stride >> 2;  // float4 per row
    for (int i = threadIdx.x; i < fix.n * q4; i += blockDim.x) {
      const int r = i / q4, c = i - r * q4;
      const float4 v = *reinterpret_cast<const float4*>(src + (long long)r * a.row_stride + c * 4);
      *reinterpret_cast<float4*>(dst_base + ((fix.pos + r) % fix.cap) * a.row_stride + c * 4) = v;
    }
    if (a.act_scale) {
      __syncthreads();  // the action columns were just written by other threads of this block
      for (int i = threadIdx.x; i < fix.n * a.A; i += blockDim.x) {
        const int r = i / a.A, j = i - r * a.A;
        float* p = dst_base + ((fix.pos + r) % fix.cap) * a.row_stride + a.act_off + j;
        *p = __fsub_rn(__fdiv_rn(*p, a.act_scale[j]), a.act_bias[j]);
      }
    }
  }
  if (!a.priority) return;
  float* priority = a.priority + (long long)g * a.prio_gs;
  const float m = a.max_priority[g];
  for (int i = threadIdx.x; i < fix.n; i += blockDim.x) priority[(fix.pos + i) % fix.cap] = m;
  __threadfence_block();
  __syncthreads();
  lap_fix_block(fix, g);
}

// ----------------------------------------------------------------------------- host object
struct Replay {
  long long capacity = 0, cap_pad = 0;
  int S = 0, A = 0, row_stride = 0;
  RowLayout lay;
  int groups = 1;
  bool prioritized = true;
  bool share_rows = false;
  float* d_rows = nullptr;
  float* d_priority = nullptr;
  float* d_max_priority = nullptr;
  double* d_tile_sums = nullptr;
  double* d_super_sums = nullptr;
  int max_supers = 0;
  bool sums_valid = false;  // the sum tree matches the priorities (false after bulk writes: rebuilt by the next sample)
  long long* d_idx = nullptr;  // last sampled indices [groups][max_batch]
  float* d_prio_stage = nullptr;
  int max_batch = 4096;
  int max_tiles = 0;
  std::vector<long long> ptr, size;
  int last_batch = 0;
  unsigned long long rng_seed = 0x5eed5eedULL, rng_counter = 0;
  DeviceArena arena;

  long long rows_gs() const { return share_rows ? 0 : capacity * (long long)row_stride; }

  // pinned staging ring for per-step appends (rl/rollout.py:45): the append kernel reads a slot over the bus when it runs, an
  // event per slot tells when it may be rewritten
  static constexpr int kStageSlots = 16;
  float* h_stage = nullptr;
  cudaEvent_t stage_ev[kStageSlots] = {};
  bool stage_busy[kStageSlots] = {};
  int stage_next = 0;
  long long stage_rows() const { return share_rows ? 1 : groups; }
  ~Replay() {
#ifndef B200_EMU
    if (h_stage) {
      cudaFreeHost(h_stage);
      for (auto& e : stage_ev) cudaEventDestroy(e);
    }
#else
    if (h_stage) free(h_stage);
#endif
  }
  int stage_slot(float** out) {
    if (!h_stage) {
      const size_t bytes = (size_t)kStageSlots * stage_rows() * row_stride * sizeof(float);
#ifndef B200_EMU
      B200_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_stage), bytes, cudaHostAllocMapped));
      for (auto& e : stage_ev) B200_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
#else
      h_stage = static_cast<float*>(aligned_alloc(64, (bytes + 63) / 64 * 64));
#endif
      memset(h_stage, 0, bytes);
    }
    const int sl = stage_next;
#ifndef B200_EMU
    if (stage_busy[sl]) B200_CHECK_CUDA(cudaEventSynchronize(stage_ev[sl]));  // (16 appends ago: practically never waits)
#endif
    *out = h_stage + (long long)sl * stage_rows() * row_stride;
    return kOk;
  }
  int stage_commit(cudaStream_t stream) {
#ifndef B200_EMU
    B200_CHECK_CUDA(cudaEventRecord(stage_ev[stage_next], stream));
#else
    (void)stream;
#endif
    stage_busy[stage_next] = true;
    stage_next = (stage_next + 1) % kStageSlots;
    return kOk;
  }
  // rl/rollout.py:45 + lap.py:31-43 for one environment step: `n` = 1 (one agent / one stacked seed `g`) or `groups`
  // (every stacked seed; a shared row store takes the transition of seed 0) transitions given as separate host arrays
  int append_transitions(int g, int n, const float* obs, const float* action, const float* reward, const float* next_obs, const float* notdone, bool raw_action,
                         cudaStream_t stream) {
    B200_REQUIRE(n == 1 || n == groups, "append_transitions: one transition, or one per stacked seed");
    float* slot = nullptr;
    B200_TRY(stage_slot(&slot));
    const int nrows = (n == 1 || share_rows) ? 1 : groups;
    for (int r = 0; r < nrows; ++r) {
      float* row = slot + (long long)r * row_stride;
      memcpy(row + lay.state, obs + (long long)r * S, S * sizeof(float));
      memcpy(row + lay.action, action + (long long)r * A, A * sizeof(float));
      row[lay.reward] = reward[r];
      row[lay.notdone] = notdone[r];
      memcpy(row + lay.next_state, next_obs + (long long)r * S, S * sizeof(float));
    }
    int st;
    if (n == 1 && groups > 1) st = append(g, slot, 1, cudaMemcpyHostToDevice, stream, raw_action);
    else if (groups == 1) st = append(0, slot, 1, cudaMemcpyHostToDevice, stream, raw_action);
    else st = append_all(slot, 1, cudaMemcpyHostToDevice, stream, raw_action);
    if (st != kOk) return st;
    return stage_commit(stream);
  }

  float* d_act_scale = nullptr;  // action normalisation applied by append when requested (lap.py:35)
  float* d_act_bias = nullptr;
  int set_action_norm(const float* scale_host, const float* bias_host) {
    if (!d_act_scale) { B200_TRY(arena.alloc(&d_act_scale, (size_t)A)); B200_TRY(arena.alloc(&d_act_bias, (size_t)A)); }
    B200_CHECK_CUDA(cudaMemcpy(d_act_scale, scale_host, A * sizeof(float), cudaMemcpyHostToDevice));
    B200_CHECK_CUDA(cudaMemcpy(d_act_bias, bias_host, A * sizeof(float), cudaMemcpyHostToDevice));
    return kOk;
  }

  int init(long long cap, int s, int a, int g, bool prio, bool share) {
    B200_REQUIRE(cap > 0 && s > 0 && a > 0 && g > 0, "replay: capacity, dims and groups must be positive");
    B200_REQUIRE(cap < (1LL << 31) - 4096, "replay: capacity must fit in int32");
    capacity = cap; S = s; A = a; groups = g; prioritized = prio; share_rows = share;
    lay = RowLayout::make(S, A);
    row_stride = lay.stride;
    cap_pad = round_up(capacity, kPrioTile * kPrioTilesPerCta);
    max_tiles = (int)round_up(cap_pad / kPrioTile, kSuperTiles);
    max_supers = max_tiles / kSuperTiles;
    const long long row_sets = share_rows ? 1 : groups;
    B200_TRY(arena.alloc(&d_rows, (size_t)row_sets * capacity * row_stride));
    B200_TRY(arena.alloc(&d_idx, (size_t)groups * max_batch));
    B200_TRY(arena.alloc(&d_prio_stage, (size_t)groups * max_batch));
    B200_TRY(arena.alloc(&d_max_priority, (size_t)groups));
    if (prioritized) {
      B200_TRY(arena.alloc(&d_priority, (size_t)groups * cap_pad));
      B200_TRY(arena.alloc(&d_tile_sums, (size_t)groups * max_tiles));
      B200_TRY(arena.alloc(&d_super_sums, (size_t)groups * max_supers));
      sums_valid = true;  // all zero
    }
    std::vector<float> ones(groups, 1.0f);  // lap.py:29 max_priority = 1
    B200_CHECK_CUDA(cudaMemcpy(d_max_priority, ones.data(), groups * sizeof(float), cudaMemcpyHostToDevice));
    ptr.assign(groups, 0);
    size.assign(groups, 0);
    return kOk;
  }

  // if the GPU can read `rows` directly (device memory, or pinned / registered host memory): the pointer to use in a kernel
  static const float* device_readable(const float* rows, cudaMemcpyKind kind) {
#ifndef B200_EMU
    if ((reinterpret_cast<uintptr_t>(rows) & 15) != 0) return nullptr;
    if (kind == cudaMemcpyDeviceToDevice) return rows;
    cudaPointerAttributes at;
    const float* direct = nullptr;
    if (cudaPointerGetAttributes(&at, rows) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) direct = static_cast<const float*>(at.devicePointer);
    (void)cudaGetLastError();
    return direct;
#else
    (void)kind;
    return rows;
#endif
  }

  // rows: host or device pointer to `count` packed rows (row_stride floats each); normalise: apply a/scale - bias to the
  // action columns on the way in (needs set_action_norm)
  int append(int g, const float* rows, long long count, cudaMemcpyKind kind, cudaStream_t stream, bool normalise = false) {
    B200_REQUIRE(g >= 0 && g < groups, "replay.append: bad group");
    B200_REQUIRE(count >= 0 && count <= capacity, "replay.append: count must be within capacity");
    B200_REQUIRE(!normalise || d_act_scale, "replay.append: action normalisation requested before set_action_norm");
    if (count == 0) return kOk;
    const long long pos = ptr[g];
    const bool want_rows = !(share_rows && g > 0);
    float* dst_base = d_rows + (share_rows ? 0 : (long long)g * capacity * row_stride);
    // small appends (the per-step case) whose source the GPU can read directly -- device memory, or pinned / registered
    // host memory -- are ONE kernel: row copy (+ action normalisation) + priorities + tree repair
    const float* direct = nullptr;
    if (want_rows && count <= 64 && (!prioritized || sums_valid)) direct = device_readable(rows, kind);
#ifdef B200_EMU
    if (!normalise) direct = nullptr;  // (the emulation keeps the memcpy path it always had, except where the kernel must transform)
#endif
    B200_REQUIRE(!normalise || direct || !want_rows, "replay.append: device-side normalisation needs <= 64 rows the GPU can read (pinned host or device memory)");
    if (want_rows && !direct) {
      const long long first = std::min(count, capacity - pos);
      B200_CHECK_CUDA(cudaMemcpyAsync(dst_base + pos * row_stride, rows, (size_t)first * row_stride * sizeof(float), kind, stream));
      if (first < count)
        B200_CHECK_CUDA(cudaMemcpyAsync(dst_base, rows + first * row_stride, (size_t)(count - first) * row_stride * sizeof(float), kind, stream));
    }
    if (prioritized || direct) {
      if (count <= 4096 && (sums_valid || !prioritized)) {
        AppendArgs a;
        memset(&a, 0, sizeof(a));
        if (prioritized) a.fix = fix_args(nullptr, 0, (int)count);
        a.fix.n = (int)count; a.fix.cap = capacity; a.fix.pos = pos;
        a.src = direct; a.src_gs = 0;
        a.dst = d_rows; a.dst_gs = share_rows ? 0 : capacity * (long long)row_stride;
        a.row_stride = row_stride; a.group0 = g;
        a.priority = prioritized ? d_priority : nullptr; a.prio_gs = cap_pad;
        a.max_priority = d_max_priority;
        if (normalise) { a.act_scale = d_act_scale; a.act_bias = d_act_bias; }
        a.act_off = lay.action; a.A = A;
        if (share_rows && g > 0) a.src = nullptr;
        B200_LAUNCH(lap_append_fix_kernel, dim3(1), dim3(count <= 32 ? 64 : 1024), 0, stream, a);
      } else {
        dim3 grid((unsigned)ceil_div((int)count, 256)), block(256);
        B200_LAUNCH(lap_append_priority_kernel, grid, block, 0, stream, d_priority + (long long)g * cap_pad, d_max_priority + g, pos, (int)count, capacity);
        sums_valid = false;
      }
      B200_CHECK_LAUNCH();
    }
    ptr[g] = (pos + count) % capacity;
    size[g] = std::min(size[g] + count, capacity);
    return kOk;
  }

  // one environment step of a whole population: `count` (<= 64) rows for EVERY stacked seed in one launch.  rows:
  // [groups][count][row_stride] (private row stores) or [count][row_stride] (shared row store), readable by the GPU.
  int append_all(const float* rows, long long count, cudaMemcpyKind kind, cudaStream_t stream, bool normalise) {
    B200_REQUIRE(count > 0 && count <= 64, "replay.append_all: 1..64 rows per seed");
    B200_REQUIRE(!normalise || d_act_scale, "replay.append_all: action normalisation requested before set_action_norm");
    for (int g = 1; g < groups; ++g) B200_REQUIRE(ptr[g] == ptr[0] && size[g] == size[0], "replay.append_all: stacked seeds must share one cursor");
    if (prioritized && !sums_valid) {
      long long n;
      B200_TRY(common_size(&n));
      B200_TRY(rebuild_sums(n, stream));
    }
    const float* direct = device_readable(rows, kind);
    B200_REQUIRE(direct != nullptr, "replay.append_all: rows must be device memory or pinned host memory, 16-byte aligned");
    AppendArgs a;
    memset(&a, 0, sizeof(a));
    if (prioritized) a.fix = fix_args(nullptr, 0, (int)count);
    a.fix.n = (int)count; a.fix.cap = capacity; a.fix.pos = ptr[0];
    a.src = direct; a.src_gs = share_rows ? 0 : count * (long long)row_stride;
    a.dst = d_rows; a.dst_gs = share_rows ? 0 : capacity * (long long)row_stride;
    a.row_stride = row_stride; a.group0 = 0;
    a.priority = prioritized ? d_priority : nullptr; a.prio_gs = cap_pad;
    a.max_priority = d_max_priority;
    if (normalise) { a.act_scale = d_act_scale; a.act_bias = d_act_bias; }
    a.act_off = lay.action; a.A = A;
    B200_LAUNCH(lap_append_fix_kernel, dim3((unsigned)groups), dim3(count <= 32 ? 64 : 1024), 0, stream, a);
    B200_CHECK_LAUNCH();
    for (int g = 0; g < groups; ++g) {
      ptr[g] = (ptr[g] + count) % capacity;
      size[g] = std::min(size[g] + count, capacity);
    }
    return kOk;
  }

  FixArgs fix_args(const long long* idx, long long idx_gs, int n) const {
    FixArgs f;
    memset(&f, 0, sizeof(f));
    f.priority = d_priority; f.prio_gs = cap_pad;
    f.tile_sums = d_tile_sums; f.ts_gs = max_tiles;
    f.super_sums = d_super_sums; f.ss_gs = max_supers;
    f.idx = idx; f.idx_gs = idx_gs;
    f.pos = 0; f.cap = capacity;
    f.n = n;
    return f;
  }
  // priority write-back for the indices of the last sample + repair of the sum tree, as one kernel (graph node of
  // the agents' update: runs beside the critic's backward pass, off the critical chain)
  int scatter_and_fix(const float* prio_dev, long long p_gs, int B, cudaStream_t stream) {
    B200_LAUNCH(lap_update_priority_kernel, dim3((unsigned)groups), dim3(1024), sizeof(ScatterTable), stream, d_priority, cap_pad, d_max_priority, d_idx,
                (long long)max_batch, prio_dev, p_gs, B, fix_args(d_idx, (long long)max_batch, B));
    B200_CHECK_LAUNCH();
    return kOk;
  }
  // rebuild the whole sum tree from the priorities
  int rebuild_sums(long long n, cudaStream_t stream) {
    const int ntiles = ceil_div((int)n, kPrioTile);
    // nothing past `size` may contribute (a restored cursor can leave stale priorities behind it)
    B200_CHECK_CUDA(cudaMemsetAsync(d_tile_sums, 0, (size_t)groups * max_tiles * sizeof(double), stream));
    B200_CHECK_CUDA(cudaMemsetAsync(d_super_sums, 0, (size_t)groups * max_supers * sizeof(double), stream));
    for (int g = 0; g < groups && n < cap_pad; ++g)
      B200_CHECK_CUDA(cudaMemsetAsync(d_priority + (long long)g * cap_pad + n, 0, (size_t)(cap_pad - n) * sizeof(float), stream));
    if (ntiles > 0) {
      dim3 grid((unsigned)ceil_div(ntiles, kPrioTilesPerCta), (unsigned)groups), block(256);
      B200_LAUNCH(lap_tile_sums_kernel, grid, block, 0, stream, d_priority, cap_pad, (int)n, d_tile_sums, ntiles, (long long)max_tiles);
      B200_CHECK_LAUNCH();
      const int nsup = ceil_div(ntiles, kSuperTiles);
      dim3 grid2((unsigned)ceil_div(nsup, 8), (unsigned)groups);
      B200_LAUNCH(lap_super_sums_kernel, grid2, block, 0, stream, d_tile_sums, (long long)max_tiles, d_super_sums, (long long)max_supers, nsup);
      B200_CHECK_LAUNCH();
    }
    sums_valid = true;
    return kOk;
  }

  int common_size(long long* out) const {
    for (int g = 1; g < groups; ++g)
      if (size[g] != size[0]) {
        set_last_error("replay.sample: stacked seeds must hold the same number of rows");
        return kErrState;
      }
    *out = size[0];
    return kOk;
  }

  struct Outputs {
    float* state; int ld_state; long long state_gs;
    float* action; int ld_action; long long action_gs;
    float* reward; long long reward_gs;
    float* next_state; int ld_next; long long next_gs;
    float* done; long long done_gs;
  };

  int sample(int B, const float* u_dev, const Outputs& o, cudaStream_t stream, const int* counter_ptr = nullptr) {
    B200_REQUIRE(B > 0 && B <= max_batch, "replay.sample: batch size must be in 1..4096");
    long long n;
    B200_TRY(common_size(&n));
    if (n <= 0) { set_last_error("replay.sample: buffer is empty"); return kErrState; }
    const int ntiles = ceil_div((int)n, kPrioTile);
    if (prioritized && !sums_valid) B200_TRY(rebuild_sums(n, stream));
    SampleArgs a;
    memset(&a, 0, sizeof(a));
    a.priority = d_priority; a.prio_gs = cap_pad;
    a.tile_sums = d_tile_sums; a.ts_gs = max_tiles; a.ntiles = ntiles;
    a.super_sums = d_super_sums; a.ss_gs = max_supers; a.nsuper = ceil_div(ntiles, kSuperTiles);
    a.size = (int)n;
    a.u = u_dev; a.u_gs = B;
    a.seed = rng_seed; a.counter = counter_ptr ? 0ULL : rng_counter++; a.counter_ptr = counter_ptr;
    a.rows = d_rows; a.rows_gs = rows_gs(); a.row_stride = row_stride; a.S = S; a.A = A; a.lay = lay;
    a.idx_out = d_idx; a.idx_gs = max_batch;
    a.state = o.state; a.ld_state = o.ld_state; a.state_gs = o.state_gs;
    a.action = o.action; a.ld_action = o.ld_action; a.action_gs = o.action_gs;
    a.reward = o.reward; a.reward_gs = o.reward_gs;
    a.next_state = o.next_state; a.ld_next = o.ld_next; a.next_gs = o.next_gs;
    a.done = o.done; a.done_gs = o.done_gs;
    a.B = B;
    a.prioritized = prioritized ? 1 : 0;
    dim3 grid((unsigned)ceil_div(B, 8), (unsigned)groups), block(256);
    // stacked seeds: thousands of latency-bound warps, so 32 registers (8 CTAs = 64 warps per SM) beat 40 (6 CTAs);
    // a single agent's 32 CTAs keep the spill-free build.  B200RL_SAMPLE_OCC8=0|1 forces either.
    static const int occ8_env = [] { const char* e = getenv("B200RL_SAMPLE_OCC8"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    const bool occ8 = occ8_env >= 0 ? occ8_env == 1 : groups > 1;
    if (occ8) { auto k = lap_sample_kernel<8>; B200_LAUNCH(k, grid, block, 0, stream, a); }
    else { auto k = lap_sample_kernel<6>; B200_LAUNCH(k, grid, block, 0, stream, a); }
    B200_CHECK_LAUNCH();
    last_batch = B;
    return kOk;
  }

  // priority write-back for the indices of the most recent sample (lap.py:66-69)
  int update_priority(const float* prio_dev, long long p_gs, int B, cudaStream_t stream) {
    B200_REQUIRE(prioritized, "update_priority on a uniform buffer");
    B200_REQUIRE(B == last_batch, "update_priority: batch size differs from the last sample()");
    dim3 grid((unsigned)groups), block((unsigned)round_up(B, 32));
    B200_LAUNCH(lap_update_priority_kernel, grid, block, sizeof(ScatterTable), stream, d_priority, cap_pad, d_max_priority, d_idx,
                (long long)max_batch, prio_dev, p_gs, B, fix_args(d_idx, (long long)max_batch, sums_valid ? B : 0));
    B200_CHECK_LAUNCH();
    return kOk;
  }

  int reset_max_priority(cudaStream_t stream) {
    B200_REQUIRE(prioritized, "reset_max_priority on a uniform buffer");
    long long n;
    B200_TRY(common_size(&n));
    B200_CHECK_CUDA(cudaMemsetAsync(d_max_priority, 0, groups * sizeof(float), stream));
    if (n > 0) {
      dim3 grid((unsigned)std::min<long long>(296, (n + 255) / 256), (unsigned)groups), block(256);
      B200_LAUNCH(lap_max_kernel, grid, block, 0, stream, d_priority, cap_pad, (int)n, d_max_priority);
      B200_CHECK_LAUNCH();
    }
    return kOk;
  }
};

}  // namespace b200rl
