// replay.cuh -- GPU-resident ring replay buffer with LAP (prioritized) and uniform sampling.
//
// Restates rl/replay_memory/lap.py and rl/replay_memory/simple.py of the reference:
//   sample  (lap.py:45-64):  P = cumsum(priority[:size]) in float64, each prefix rounded to
//           float32; v = u * P[-1] (float32); ind = first i with P[i] >= v; gather 5 arrays.
//   update_priority (lap.py:66-69), reset_max_priority (lap.py:71-73), append (lap.py:31-43).
//
// HBM layout: one packed fp32 row per transition, padded to a multiple of 4 floats so a
// warp reads it with coalesced 128-bit loads:
//     [ state(S) | action(A) | reward | notdone | next_state(S) | pad ]
// (fp32 storage is value-identical to the reference's fp64-store + fp32-cast).
// priority: fp32[capacity] (+pad) per stacked seed; max_priority: one fp32 per seed, on device.
//
// Sampling is reduce-then-search, never materialising the 1M-entry prefix array:
//   1. lap_tile_sums_kernel: fp64 sum of every 256-priority tile  (reads 4*size bytes)
//   2. lap_sample_kernel:    each CTA scans the tile sums in shared memory (fp64), each warp
//      takes one draw: binary search over tile prefixes (compared after rounding to fp32,
//      exactly like the reference's fp32 prefix array), warp-shuffle scan inside the tile,
//      then the warp copies the chosen row out with 128-bit loads.
// Priorities are fp32 values >= 1 (or 0), so fp64 partial sums are EXACT for any summation
// order while max_priority * size < 2^30; the sampled indices are then bit-identical to
// torch.cumsum + torch.searchsorted on the CPU.
#pragma once
#include "host.cuh"

namespace b200rl {

constexpr int kPrioTile = 256;       // priorities per tile (one warp: 32 lanes x 8)
constexpr int kPrioTilesPerCta = 8;  // 2048 priorities per CTA in the reduce kernel
// draws handled by one warp of the search/gather kernel: 1 for a single agent (latency), 4 when many seeds
// are stacked (fewer CTAs re-scan the tile sums)
inline int samples_per_warp(int groups) { return groups >= 8 ? 4 : 1; }

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

struct SampleArgs {
  const float* priority; long long prio_gs;
  const double* tile_sums; long long ts_gs; int ntiles;
  int size;
  const float* u; long long u_gs;          // injected uniforms [B] (null -> Philox)
  unsigned long long seed, counter;        // Philox stream when u == null
  const int* counter_ptr;                  // optional device-side counter added to `counter` (graph replays)
  const float* rows; long long rows_gs; int row_stride; int S, A;
  long long* idx_out; long long idx_gs;
  float* state; int ld_state; long long state_gs;
  float* action; int ld_action; long long action_gs;
  float* reward; long long reward_gs;
  float* next_state; int ld_next; long long next_gs;
  float* done; long long done_gs;
  int B;
  int prioritized;
  int spw;  // samples per warp
};

#define PFX(i) ((i) + ((i) >> 4))
__global__ void __launch_bounds__(256) lap_sample_kernel(SampleArgs a) {
  B200_PDL_PROLOGUE();
  B200_DYN_SMEM(double, prefix);  // [ntiles] inclusive tile prefix sums (prioritized mode)
  __shared__ double warp_tot[8];
  const int g = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int size = a.size;

  if (a.prioritized) {
    // inclusive scan of the tile sums in shared memory: coalesced copy in, per-thread serial scan of
    // a contiguous run, warp/block scan of the run totals, add the offsets back
    const int nt = a.ntiles;
    const double* ts = a.tile_sums + (long long)g * a.ts_gs;
    // (element i lives at PFX(i) = i + i/16: one pad slot per 16 keeps the per-thread runs off a single bank)
    for (int t = tid; t < nt; t += 256) prefix[PFX(t)] = ts[t];
    __syncthreads();
    const int per = (nt + 255) / 256;
    const int lo = min(tid * per, nt), hi = min(lo + per, nt);
    double local = 0.0;
    for (int t = lo; t < hi; ++t) { local += prefix[PFX(t)]; prefix[PFX(t)] = local; }
    const double incl = warp_inclusive_scan(local, lane);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    double off = incl - local;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    for (int t = lo; t < hi; ++t) prefix[PFX(t)] += off;
    __syncthreads();
  }

  // each warp takes kSamplesPerWarp draws in turn (fewer CTAs re-scan the tile sums)
  for (int sw = 0; sw < a.spw; ++sw) {
  const int b = (blockIdx.x * 8 + warp) * a.spw + sw;
  if (b >= a.B) return;

  float u;
  if (a.u) {
    u = a.u[(long long)g * a.u_gs + b];
  } else {
    Philox ph(a.seed);
    uint32_t r[4];
    const unsigned long long ctr = a.counter + (a.counter_ptr ? (unsigned long long)(unsigned)(*a.counter_ptr) : 0ULL);
    ph(ctr, ((unsigned long long)g << 32) | (unsigned)(b >> 2), r);
    u = u01_from_bits(r[b & 3]);
  }

  long long idx;
  if (a.prioritized) {
    const int nt = a.ntiles;
    const float total32 = (float)prefix[PFX(nt - 1)];
    const float v = __fmul_rn(u, total32);
    // first tile whose fp32-rounded inclusive prefix reaches v
    int lo = 0, hi = nt - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((float)prefix[PFX(mid)] >= v) hi = mid; else lo = mid + 1;
    }
    const int tile = lo;
    const double base = tile > 0 ? prefix[PFX(tile - 1)] : 0.0;
    const float* p = a.priority + (long long)g * a.prio_gs + (long long)tile * kPrioTile + lane * 8;
    const float4 x0 = *reinterpret_cast<const float4*>(p);
    const float4 x1 = *reinterpret_cast<const float4*>(p + 4);
    const int e0 = tile * kPrioTile + lane * 8;
    float pv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    double c[8];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (e0 + k < size) run += (double)pv[k];
      c[k] = run;
    }
    const double incl = warp_inclusive_scan(run, lane);
    const double excl = base + (incl - run);
    int kfirst = 8;
#pragma unroll
    for (int k = 7; k >= 0; --k)
      if (e0 + k < size && (float)(excl + c[k]) >= v) kfirst = k;
    const unsigned hit = __ballot_sync(0xffffffffu, kfirst < 8);
    if (hit == 0u) {
      idx = size - 1;  // unreachable for u < 1; keeps the gather in range
    } else {
      const int src = __ffs((int)hit) - 1;
      const int kk = __shfl_sync(0xffffffffu, kfirst, src);
      idx = (long long)tile * kPrioTile + src * 8 + kk;
    }
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

  // ---- gather the packed row with 128-bit loads, scatter into the five batch tensors ----
  const float* row = a.rows + (long long)g * a.rows_gs + idx * (long long)a.row_stride;
  const int S = a.S, A = a.A;
  const int o_act = S, o_rew = S + A, o_done = S + A + 1, o_next = S + A + 2, o_end = 2 * S + A + 2;
  float* st = a.state + (long long)g * a.state_gs + (long long)b * a.ld_state;
  float* ac = a.action + (long long)g * a.action_gs + (long long)b * a.ld_action;
  float* ns = a.next_state + (long long)g * a.next_gs + (long long)b * a.ld_next;
  for (int q = lane; q * 4 < o_end; q += 32) {
    const float4 v4 = *reinterpret_cast<const float4*>(row + q * 4);
    const float vals[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int pos = q * 4 + k;
      const float val = vals[k];
      if (pos < o_act) st[pos] = val;
      else if (pos < o_rew) ac[pos - o_act] = val;
      else if (pos == o_rew) a.reward[(long long)g * a.reward_gs + b] = val;
      else if (pos == o_done) a.done[(long long)g * a.done_gs + b] = val;
      else if (pos < o_end) ns[pos - o_next] = val;
    }
  }
  }  // samples of this warp
}

__global__ void __launch_bounds__(1024) lap_update_priority_kernel(float* __restrict__ priority, long long prio_gs, float* __restrict__ max_priority,
                                                                   const long long* __restrict__ idx, long long idx_gs, const float* __restrict__ prio,
                                                                   long long p_gs, int B) {
  B200_PDL_PROLOGUE();
  B200_DYN_SMEM(ScatterTable, tab);
  const int g = blockIdx.x, b = threadIdx.x;
  const float p = (b < B) ? prio[(long long)g * p_gs + b] : 0.f;
  lap_scatter_block(priority + (long long)g * prio_gs, max_priority + g, idx + (long long)g * idx_gs, p, b, B, tab);
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

// ----------------------------------------------------------------------------- host object
struct Replay {
  long long capacity = 0, cap_pad = 0;
  int S = 0, A = 0, row_stride = 0;
  int groups = 1;
  bool prioritized = true;
  bool share_rows = false;
  float* d_rows = nullptr;
  float* d_priority = nullptr;
  float* d_max_priority = nullptr;
  double* d_tile_sums = nullptr;
  long long* d_idx = nullptr;  // last sampled indices [groups][max_batch]
  float* d_prio_stage = nullptr;
  int max_batch = 4096;
  int max_tiles = 0;
  std::vector<long long> ptr, size;
  int last_batch = 0;
  unsigned long long rng_seed = 0x5eed5eedULL, rng_counter = 0;
  DeviceArena arena;

  long long rows_gs() const { return share_rows ? 0 : capacity * (long long)row_stride; }

  int init(long long cap, int s, int a, int g, bool prio, bool share) {
    B200_REQUIRE(cap > 0 && s > 0 && a > 0 && g > 0, "replay: capacity, dims and groups must be positive");
    B200_REQUIRE(cap < (1LL << 31) - 4096, "replay: capacity must fit in int32");
    capacity = cap; S = s; A = a; groups = g; prioritized = prio; share_rows = share;
    row_stride = (int)round_up(2 * S + A + 2, 4);
    cap_pad = round_up(capacity, kPrioTile * kPrioTilesPerCta);
    max_tiles = (int)(cap_pad / kPrioTile);
    B200_REQUIRE((size_t)(max_tiles + max_tiles / 16 + 1) * sizeof(double) <= 200 * 1024, "replay: capacity too large for the shared-memory tile scan (max ~26M rows)");
    const long long row_sets = share_rows ? 1 : groups;
    B200_TRY(arena.alloc(&d_rows, (size_t)row_sets * capacity * row_stride));
    B200_TRY(arena.alloc(&d_idx, (size_t)groups * max_batch));
    B200_TRY(arena.alloc(&d_prio_stage, (size_t)groups * max_batch));
    B200_TRY(arena.alloc(&d_max_priority, (size_t)groups));
    if (prioritized) {
      B200_TRY(arena.alloc(&d_priority, (size_t)groups * cap_pad));
      B200_TRY(arena.alloc(&d_tile_sums, (size_t)groups * max_tiles));
    }
    std::vector<float> ones(groups, 1.0f);  // lap.py:29 max_priority = 1
    B200_CHECK_CUDA(cudaMemcpy(d_max_priority, ones.data(), groups * sizeof(float), cudaMemcpyHostToDevice));
    ptr.assign(groups, 0);
    size.assign(groups, 0);
#ifndef B200_EMU
    B200_CHECK_CUDA(cudaFuncSetAttribute(lap_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
#endif
    return kOk;
  }

  // rows: host or device pointer to `count` packed rows (row_stride floats each, already normalised)
  int append(int g, const float* rows, long long count, cudaMemcpyKind kind, cudaStream_t stream) {
    B200_REQUIRE(g >= 0 && g < groups, "replay.append: bad group");
    B200_REQUIRE(count >= 0 && count <= capacity, "replay.append: count must be within capacity");
    if (count == 0) return kOk;
    const long long pos = ptr[g];
    if (!(share_rows && g > 0)) {
      float* dst_base = d_rows + (share_rows ? 0 : (long long)g * capacity * row_stride);
      const long long first = std::min(count, capacity - pos);
      B200_CHECK_CUDA(cudaMemcpyAsync(dst_base + pos * row_stride, rows, (size_t)first * row_stride * sizeof(float), kind, stream));
      if (first < count)
        B200_CHECK_CUDA(cudaMemcpyAsync(dst_base, rows + first * row_stride, (size_t)(count - first) * row_stride * sizeof(float), kind, stream));
    }
    if (prioritized) {
      dim3 grid((unsigned)ceil_div((int)count, 256)), block(256);
      B200_LAUNCH(lap_append_priority_kernel, grid, block, 0, stream, d_priority + (long long)g * cap_pad, d_max_priority + g, pos, (int)count, capacity);
      B200_CHECK_LAUNCH();
    }
    ptr[g] = (pos + count) % capacity;
    size[g] = std::min(size[g] + count, capacity);
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
    if (prioritized) {
      dim3 grid((unsigned)ceil_div(ntiles, kPrioTilesPerCta), (unsigned)groups), block(256);
      B200_LAUNCH(lap_tile_sums_kernel, grid, block, 0, stream, d_priority, cap_pad, (int)n, d_tile_sums, ntiles, (long long)max_tiles);
      B200_CHECK_LAUNCH();
    }
    SampleArgs a;
    memset(&a, 0, sizeof(a));
    a.priority = d_priority; a.prio_gs = cap_pad;
    a.tile_sums = d_tile_sums; a.ts_gs = max_tiles; a.ntiles = ntiles;
    a.size = (int)n;
    a.u = u_dev; a.u_gs = B;
    a.seed = rng_seed; a.counter = counter_ptr ? 0ULL : rng_counter++; a.counter_ptr = counter_ptr;
    a.rows = d_rows; a.rows_gs = rows_gs(); a.row_stride = row_stride; a.S = S; a.A = A;
    a.idx_out = d_idx; a.idx_gs = max_batch;
    a.state = o.state; a.ld_state = o.ld_state; a.state_gs = o.state_gs;
    a.action = o.action; a.ld_action = o.ld_action; a.action_gs = o.action_gs;
    a.reward = o.reward; a.reward_gs = o.reward_gs;
    a.next_state = o.next_state; a.ld_next = o.ld_next; a.next_gs = o.next_gs;
    a.done = o.done; a.done_gs = o.done_gs;
    a.B = B;
    a.prioritized = prioritized ? 1 : 0;
    a.spw = samples_per_warp(groups);
    dim3 grid((unsigned)ceil_div(B, 8 * a.spw), (unsigned)groups), block(256);
    const size_t smem = prioritized ? (size_t)(ntiles + ntiles / 16 + 1) * sizeof(double) : 0;
    B200_LAUNCH(lap_sample_kernel, grid, block, smem, stream, a);
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
                (long long)max_batch, prio_dev, p_gs, B);
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
