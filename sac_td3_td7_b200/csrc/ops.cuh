// ops.cuh -- the non-GEMM kernels of the update chain: AvgL1Norm fwd/bwd, the TD7 / TD3 / SAC
// target + loss + priority kernels, SAC tanh-Gaussian sampling fwd/bwd, Adam / Polyak over
// packed parameter blobs, and the per-step scalar preparation.
//
// Reference arithmetic each kernel restates (file:line in /root/reference):
//   AvgL1Norm                    rl/nn/sale.py:11-13
//   TD7 encoder loss             rl/agent/td7.py:246-257
//   TD7 target / LAP Huber / priority / value range   rl/agent/td7.py:165-244
//   TD7 actor loss               rl/agent/td7.py:259-276
//   TD3 critic / actor loss      rl/agent/td3.py:137-192
//   SAC sample / log-prob        rl/agent/sac.py:154-172
//   SAC critic / actor / temperature loss   rl/agent/sac.py:174-241
//   Polyak                       rl/agent/td3.py:195-204, rl/agent/sac.py:244-249
//   priority scatter + running max          rl/replay_memory/lap.py:66-69
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace b200rl {

// ----------------------------------------------------------------------------- AvgL1Norm
struct RowOp {
  const float* x;   // pre-norm input  [rows][n]
  const float* dy;  // bwd: gradient wrt the normalised output
  float* y;         // fwd: normalised output; bwd: gradient wrt x
  long long x_gs, dy_gs, y_gs;
  int x_ld, dy_ld, y_ld;
  int rows, n;
  int row_begin;  // first linear row of this tensor inside the launch
};

constexpr float kNormEps = 1e-8f;

// loads of data another SM may have produced inside the same (persistent) kernel go through L2
template <bool CG>
__device__ __forceinline__ float ldf(const float* p) {
#ifndef B200_EMU
  if (CG) return __ldcg(p);
#endif
  return *p;
}
template <bool CG>
__device__ __forceinline__ float4 ldf4(const float* p) {
#ifndef B200_EMU
  if (CG) return __ldcg(reinterpret_cast<const float4*>(p));
#endif
  return *reinterpret_cast<const float4*>(p);
}

// one row of one tensor, by one warp
template <bool BWD, bool CG>
__device__ __forceinline__ void avgl1norm_row(const RowOp& op, int row, int g, int lane) {
  const int n = op.n;
  const float* x = op.x + (long long)g * op.x_gs + (long long)row * op.x_ld;
  float* y = op.y + (long long)g * op.y_gs + (long long)row * op.y_ld;
  float s_abs = 0.f, s_dot = 0.f;
  const float* dy = nullptr;
  if (BWD) dy = op.dy + (long long)g * op.dy_gs + (long long)row * op.dy_ld;
  // 128-bit path: rows of up to 512 floats whose base addresses are 16-byte aligned (the 256-wide SALE layers) are
  // held in registers -- every load of the row is in flight at once and the row is read from memory only once
  const bool vec = (n % 128 == 0) && n <= 512 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | (BWD ? reinterpret_cast<uintptr_t>(dy) : 0)) & 15) == 0;
  if (vec) {
    const int nv = n >> 7;  // float4 per lane
    float4 xv[4], dv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nv) {
        xv[k] = ldf4<CG>(x + k * 128 + lane * 4);
        if (BWD) dv[k] = ldf4<CG>(dy + k * 128 + lane * 4);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nv) {
        s_abs += (fabsf(xv[k].x) + fabsf(xv[k].y)) + (fabsf(xv[k].z) + fabsf(xv[k].w));
        if (BWD) s_dot = fmaf(dv[k].x, xv[k].x, fmaf(dv[k].y, xv[k].y, fmaf(dv[k].z, xv[k].z, fmaf(dv[k].w, xv[k].w, s_dot))));
      }
    }
    s_abs = warp_sum(s_abs);
    const float mean = s_abs / (float)n;
    const float c = fmaxf(mean, kNormEps);
    if (!BWD) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < nv) *reinterpret_cast<float4*>(y + k * 128 + lane * 4) = make_float4(xv[k].x / c, xv[k].y / c, xv[k].z / c, xv[k].w / c);
    } else {
      s_dot = warp_sum(s_dot);
      const float gm = (mean >= kNormEps) ? (-s_dot / (c * c)) / (float)n : 0.f;
      auto sgn = [](float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); };
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < nv)
          *reinterpret_cast<float4*>(y + k * 128 + lane * 4) = make_float4(dv[k].x / c + sgn(xv[k].x) * gm, dv[k].y / c + sgn(xv[k].y) * gm,
                                                                            dv[k].z / c + sgn(xv[k].z) * gm, dv[k].w / c + sgn(xv[k].w) * gm);
    }
    return;
  }
  for (int j = lane; j < n; j += 32) {
    float v = ldf<CG>(x + j);
    s_abs += fabsf(v);
    if (BWD) s_dot = fmaf(ldf<CG>(dy + j), v, s_dot);
  }
  s_abs = warp_sum(s_abs);
  const float mean = s_abs / (float)n;
  const float c = fmaxf(mean, kNormEps);
  if (!BWD) {
    for (int j = lane; j < n; j += 32) y[j] = ldf<CG>(x + j) / c;
  } else {
    s_dot = warp_sum(s_dot);
    // y = x / c ; c = clamp(mean|x|, eps):  dx = dy/c + sign(x) * [mean>=eps] * (-sum(dy*x)/c^2) / n
    const float gm = (mean >= kNormEps) ? (-s_dot / (c * c)) / (float)n : 0.f;
    for (int j = lane; j < n; j += 32) {
      float v = ldf<CG>(x + j);
      float sg = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
      y[j] = ldf<CG>(dy + j) / c + sg * gm;
    }
  }
}

template <bool BWD>
__global__ void __launch_bounds__(256) avgl1norm_kernel(const RowOp* __restrict__ ops, int nops, int total_rows) {
  B200_PDL_PROLOGUE();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_lin = blockIdx.x * 8 + warp;
  const int g = blockIdx.y;
  if (row_lin >= total_rows) return;
  int oi = 0;
  while (oi + 1 < nops && row_lin >= ops[oi + 1].row_begin) ++oi;
  const RowOp op = ops[oi];
  avgl1norm_row<BWD, false>(op, row_lin - op.row_begin, g, lane);
}

// ----------------------------------------------------------------------------- per-step scalars
struct OptimHyper {
  double lr, beta1, beta2, eps;
};
// Host-visible mailbox (pinned, mapped memory): the kernel that produces an info value of seed 0 also posts it here as
// ONE 8-byte store {value bits, update number} -- a single posted write, so no system-scope fence is needed to order
// value and tag (a fence.sys costs ~1.5 us on the chain) -- and the host can read a loss the moment it exists, without
// waiting for the rest of the update graph, by polling the tag (b200rl_agent_wait_info).
struct InfoMail { unsigned long long bits; };  // low 32: float bits, high 32: update number
constexpr int kMailCols = 8;
// device-resident counters shared by the update graphs
struct StepState {
  int n_runs;         // number of completed train_ops calls (TD7 semantics: incremented first)
  int adam_t[4];      // step count per optimizer
  int info_slot;      // ring slot the loss kernels of the current step write to
  int pad[2];
  double b1pow[4];    // beta1^t, beta2^t per optimizer, carried as running products (1.0 at t = 0)
  double b2pow[4];
  InfoMail* mail;     // device address of the host mailbox (null: none)
};
__device__ __forceinline__ void publish_info(const StepState* st, long long g, int col, float v) {
  if (g != 0 || st->mail == nullptr || col < 0 || col >= kMailCols) return;
  const unsigned long long packed = ((unsigned long long)(unsigned)st->n_runs << 32) | (unsigned long long)__float_as_uint(v);
  *reinterpret_cast<volatile unsigned long long*>(&st->mail[col].bits) = packed;
}

// one thread: advance counters and refresh the Adam scalars of the optimizers that step now
__device__ __forceinline__ void prep_step_body(StepState* st, AdamScalars* adam, const OptimHyper* hyper, int nopt, unsigned step_mask, int ring);
__global__ void prep_step_kernel(StepState* st, AdamScalars* adam, const OptimHyper* hyper, int nopt, unsigned step_mask, int ring) {
  B200_PDL_PROLOGUE();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  prep_step_body(st, adam, hyper, nopt, step_mask, ring);
}
// (one thread)
__device__ __forceinline__ void prep_step_body(StepState* st, AdamScalars* adam, const OptimHyper* hyper, int nopt, unsigned step_mask, int ring) {
  st->n_runs += 1;
  st->info_slot = (st->n_runs - 1) % ring;
  for (int k = 0; k < nopt; ++k) {
    if (!((step_mask >> k) & 1u)) continue;
    const int t = ++st->adam_t[k];
    const OptimHyper h = hyper[k];
    st->b1pow[k] *= h.beta1;
    st->b2pow[k] *= h.beta2;
    const double bc1 = 1.0 - st->b1pow[k];
    const double bc2 = 1.0 - st->b2pow[k];
    AdamScalars s;
    s.w1 = (float)(1.0 - h.beta1);
    s.beta2 = (float)h.beta2;
    s.w2 = (float)(1.0 - h.beta2);
    s.bc2_sqrt = (float)sqrt(bc2);
    s.neg_step = (float)(-(h.lr / bc1));
    s.eps = (float)h.eps;
    s.step = t;
    s.pad = 0;
    adam[k] = s;
  }
}

// ----------------------------------------------------------------------------- elementwise over blobs
__global__ void __launch_bounds__(256) adam_blob_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                        const float* __restrict__ g, long long n, const AdamScalars* __restrict__ sc) {
  B200_PDL_PROLOGUE();
  const AdamScalars s = *sc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    adam_update(g[i], pi, mi, vi, s);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

// target <- tau*src + target*(1-tau), in the reference's operation order (two roundings + add)
__global__ void __launch_bounds__(256) polyak_kernel(float* __restrict__ target, const float* __restrict__ src, long long n, float tau, float one_minus_tau) {
  B200_PDL_PROLOGUE();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    target[i] = __fadd_rn(__fmul_rn(tau, src[i]), __fmul_rn(target[i], one_minus_tau));
  }
}

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ p, long long n, float v) {
  B200_PDL_PROLOGUE();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// sum over tensors of ||grad||_2 (rl/nn/utils.py:13-19).  The tensors are cut into chunks of kNormChunk elements; one
// CTA reduces one chunk (fp64 sum of squares, fixed order), and the CTA that finishes last adds the chunk partials of
// every tensor in chunk order, takes the square roots and writes the total -- deterministic, and a few microseconds
// instead of one CTA walking 70k+ elements at one L2 round trip per pass.
struct TensorSlice { long long offset; long long numel; };
struct NormChunk { long long offset; int count; int slice; };
constexpr int kNormChunk = 4096;
__global__ void __launch_bounds__(256) grad_norm_sum_kernel(const float* __restrict__ grad_base, long long grad_gs, const NormChunk* __restrict__ chunks, int nchunks,
                                                            int nslices, double* __restrict__ partial, unsigned* __restrict__ counter, float* __restrict__ info,
                                                            int info_stride, int info_col, long long info_gs, const StepState* st) {
  B200_PDL_PROLOGUE();
  __shared__ double red[8];
  __shared__ bool is_last;
  const int g = blockIdx.y, c = blockIdx.x;
  const NormChunk ch = chunks[c];
  const float* gp = grad_base + (long long)g * grad_gs + ch.offset;
  double s = 0.0;
  for (int i = threadIdx.x; i < ch.count; i += blockDim.x) {
    const double x = gp[i];
    s += x * x;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < 8; ++w) a += red[w];
    partial[(long long)g * nchunks + c] = a;
    __threadfence();
    is_last = atomicAdd(&counter[g], 1u) == (unsigned)(nchunks - 1);
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  const volatile double* pp = partial + (long long)g * nchunks;
  double total = 0.0, acc = 0.0;
  int cur = 0;
  for (int k = 0; k < nchunks; ++k) {  // chunks are stored slice by slice
    if (chunks[k].slice != cur) { total += (double)sqrtf((float)acc); acc = 0.0; cur = chunks[k].slice; }
    acc += pp[k];
  }
  total += (double)sqrtf((float)acc);
  (void)nslices;
  info[g * info_gs + (long long)st->info_slot * info_stride + info_col] = (float)total;
  publish_info(st, g, info_col, (float)total);
  counter[g] = 0u;  // ready for the next replay of the graph
}

// ----------------------------------------------------------------------------- block reductions (fixed order)
__device__ __forceinline__ float block_sum_256(float v, float* red /*[32]*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}
__device__ __forceinline__ float block_max_256(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float s = red[0];
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 1; i < nw; ++i) s = fmaxf(s, red[i]);
  return s;
}
__device__ __forceinline__ float block_min_256(float v, float* red) { return -block_max_256(-v, red); }

// ----------------------------------------------------------------------------- LAP priority write-back
// priority[idx[b]] = prio[b]; duplicates: the LAST batch position wins (numpy/torch index
// assignment); running max via atomicMax on the (positive) float bits.
// Duplicates are resolved with a small open-addressing hash table in shared memory
// (key = row index, value = highest batch position that sampled it): O(1) per thread.
constexpr int kScatterSlots = 2048;  // >= 2 * max batch (1024)
struct ScatterTable {
  unsigned long long keys[kScatterSlots];
  int last[kScatterSlots];
};
__device__ __forceinline__ size_t scatter_smem_bytes() { return sizeof(ScatterTable); }

__device__ __forceinline__ void lap_scatter_block(float* __restrict__ priority, float* __restrict__ max_priority,
                                                  const long long* __restrict__ idx, float p, int b, int B, ScatterTable* tab) {
  const unsigned long long kEmpty = ~0ULL;
  for (int i = threadIdx.x; i < kScatterSlots; i += blockDim.x) { tab->keys[i] = kEmpty; tab->last[i] = -1; }
  __syncthreads();
  unsigned long long my = 0;
  int slot = -1;
  if (b < B) {
    my = (unsigned long long)idx[b];
    unsigned h = (unsigned)((my * 0x9E3779B97F4A7C15ULL) >> 40) & (kScatterSlots - 1);
    for (;;) {
      const unsigned long long old = atomicCAS(&tab->keys[h], kEmpty, my);
      if (old == kEmpty || old == my) break;
      h = (h + 1) & (kScatterSlots - 1);
    }
    slot = (int)h;
    atomicMax(&tab->last[slot], b);
  }
  __syncthreads();
  if (b < B) {
    if (tab->last[slot] == b) priority[my] = p;
    atomicMax(reinterpret_cast<int*>(max_priority), __float_as_int(p));
  }
}

// ----------------------------------------------------------------------------- TD7
struct TD7Range { float value_max, value_min, value_target_max, value_target_min; };

// encoder loss: L = mean((zsa - zs_next)^2); dzsa = 2 (zsa - zs_next) / (B*Z).  One warp per row;
// block partials are combined by the last block to finish, in block order.
__global__ void __launch_bounds__(256) td7_encoder_loss_kernel(const float* __restrict__ zsa, int zsa_ld, const float* __restrict__ zs_next, int zn_ld,
                                                               float* __restrict__ dzsa, int d_ld, int B, int Z, long long gs_act,
                                                               float* __restrict__ partial, unsigned* __restrict__ counter,
                                                               float* __restrict__ info, int info_stride, int info_col, const StepState* st, long long info_gs) {
  B200_PDL_PROLOGUE();
  __shared__ float red[32];
  __shared__ bool is_last;
  const int g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  float s = 0.f;
  if (row < B) {
    const float* a = zsa + g * gs_act + (long long)row * zsa_ld;
    const float* t = zs_next + g * gs_act + (long long)row * zn_ld;
    float* d = dzsa + g * gs_act + (long long)row * d_ld;
    const float scale = 2.0f / ((float)B * (float)Z);
    for (int j = lane; j < Z; j += 32) {
      float diff = a[j] - t[j];
      s = fmaf(diff, diff, s);
      d[j] = diff * scale;
    }
  }
  s = block_sum_256(s, red);
  if (threadIdx.x == 0) {
    partial[g * gridDim.x + blockIdx.x] = s;
    __threadfence();
    unsigned done = atomicAdd(&counter[g], 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double tot = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) tot += (double)partial[g * gridDim.x + k];
    info[g * info_gs + (long long)st->info_slot * info_stride + info_col] = (float)(tot / ((double)B * (double)Z));
    publish_info(st, g, info_col, (float)(tot / ((double)B * (double)Z)));
    counter[g] = 0;
  }
}

// critic target, loss, priorities, output gradients.  One block of >= B threads per seed.
//   y   = r + gamma * clamp(min(qt1,qt2), vt_min, vt_max) * notdone
//   LAP : L = mean_B(h(|q1-y|) + h(|q2-y|)), h(d) = d<1 ? d^2/2 : d ; prio = max(|.|,1)^0.4
//   MSE : L = mean((y-q1)^2)/2 + mean((y-q2)^2)/2
struct CriticLossArgs {
  const float* q1; const float* q2; const float* qt1; const float* qt2;
  const float* reward; const float* notdone;
  float* dq1; float* dq2;          // gradient wrt the critic outputs
  float* y_out;                    // optional: q target, for parity tests
  float* prio_out;                 // optional: priorities, for parity tests / API
  TD7Range* range;                 // null for TD3/SAC
  const float* logp_next;          // SAC: log pi(a'|s') ; else null
  const float* log_alpha;          // SAC: pointer to log-temperature (device) or null
  float alpha_const;               // SAC fixed temperature when log_alpha == null
  float* priority; float* max_priority; const long long* idx;  // replay write-back (null = none)
  long long prio_gs;               // stride of the priority array between stacked seeds
  long long idx_gs;                // stride of the sampled-index array between stacked seeds
  float* info; int info_stride; int info_col; long long info_gs;
  const StepState* st;
  float gamma;
  int B;
  int use_lap;
  int huber_sum;                   // TD7 sums the two critics then means (same value; kept for clarity)
  long long gs;                    // stride between stacked seeds of the [B] vectors
};

__global__ void __launch_bounds__(1024) critic_loss_kernel(CriticLossArgs a) {
  B200_PDL_PROLOGUE();
  __shared__ float red[32];
  B200_DYN_SMEM(ScatterTable, scatter_tab);  // hash table for the last-writer-wins scatter
  const int g = blockIdx.x;
  const int b = threadIdx.x;
  const int B = a.B;
  const long long o = (long long)g * a.gs;
  float y = 0.f, q1 = 0.f, q2 = 0.f;
  float vt_min = 0.f, vt_max = 0.f;
  const bool clip = (a.range != nullptr);
  if (clip) { vt_min = a.range[g].value_target_min; vt_max = a.range[g].value_target_max; }
  if (b < B) {
    float nv = fminf(a.qt1[o + b], a.qt2[o + b]);
    if (clip) nv = fminf(fmaxf(nv, vt_min), vt_max);
    if (a.logp_next) {
      const float alpha = a.log_alpha ? expf(a.log_alpha[g]) : a.alpha_const;
      nv = __fsub_rn(nv, __fmul_rn(alpha, a.logp_next[o + b]));
    }
    y = __fadd_rn(a.reward[o + b], __fmul_rn(__fmul_rn(a.gamma, nv), a.notdone[o + b]));
    q1 = a.q1[o + b];
    q2 = a.q2[o + b];
    if (a.y_out) a.y_out[o + b] = y;
  }
  if (clip) {
    float ymax = block_max_256(b < B ? y : -3.0e38f, red);
    float ymin = block_min_256(b < B ? y : 3.0e38f, red);
    if (b == 0) {
      a.range[g].value_max = fmaxf(a.range[g].value_max, ymax);
      a.range[g].value_min = fminf(a.range[g].value_min, ymin);
    }
  }
  const float e1 = q1 - y, e2 = q2 - y;
  const float td1 = fabsf(e1), td2 = fabsf(e2);
  float l = 0.f, g1 = 0.f, g2 = 0.f;
  const float invB = 1.0f / (float)B;
  if (b < B) {
    if (a.use_lap) {
      l = (td1 < 1.f ? 0.5f * td1 * td1 : td1) + (td2 < 1.f ? 0.5f * td2 * td2 : td2);
      g1 = (td1 < 1.f ? e1 : (e1 > 0.f ? 1.f : -1.f)) * invB;
      g2 = (td2 < 1.f ? e2 : (e2 > 0.f ? 1.f : -1.f)) * invB;
    } else {
      l = 0.5f * (e1 * e1 + e2 * e2);
      g1 = e1 * invB;
      g2 = e2 * invB;
    }
    a.dq1[o + b] = g1;
    a.dq2[o + b] = g2;
  }
  const float tot = block_sum_256(l, red);
  if (b == 0) {
    a.info[g * a.info_gs + (long long)a.st->info_slot * a.info_stride + a.info_col] = tot * invB;
    publish_info(a.st, g, a.info_col, tot * invB);
  }
  if (a.use_lap) {
    const float p = (b < B) ? powf(fmaxf(fmaxf(td1, td2), 1.0f), 0.4f) : 0.f;
    if (a.prio_out && b < B) a.prio_out[o + b] = p;
    if (a.priority) {
      lap_scatter_block(a.priority + g * a.prio_gs, a.max_priority + g, a.idx + g * a.idx_gs, p, b, B, scatter_tab);
    }
  }
}

// actor loss TD7: L = -mean(cat[q1,q2]); the output gradients are the constant -1/(2B)
__global__ void __launch_bounds__(1024) td7_actor_loss_kernel(const float* __restrict__ q1, const float* __restrict__ q2, int B, long long gs,
                                                              float* __restrict__ info, int info_stride, int info_col, long long info_gs, const StepState* st) {
  B200_PDL_PROLOGUE();
  __shared__ float red[32];
  const int g = blockIdx.x, b = threadIdx.x;
  float v = (b < B) ? (q1[g * gs + b] + q2[g * gs + b]) : 0.f;
  const float tot = block_sum_256(v, red);
  if (b == 0) {
    info[g * info_gs + (long long)st->info_slot * info_stride + info_col] = -tot / (2.0f * (float)B);
    publish_info(st, g, info_col, -tot / (2.0f * (float)B));
  }
}

// actor loss TD3 / SAC: L = mean(-min(q1,q2) [+ alpha*logp]); gradient of min splits evenly on ties.
// SAC additionally: L_alpha = mean(exp(t) * (-logp - target_entropy)), dL_alpha/dt = L_alpha,
// entropy = -mean(logp), dlogp = alpha/B.
struct ActorLossArgs {
  const float* q1; const float* q2;
  float* dq1; float* dq2;
  const float* logp;          // SAC or null
  const float* log_alpha;     // SAC auto mode or null
  float alpha_const;
  float target_entropy;
  float* dlogp;               // SAC: [B]
  float* grad_log_alpha;      // SAC auto: scalar gradient
  float* info; int info_stride; long long info_gs; const StepState* st;
  int col_policy, col_tmp_loss, col_entropy, col_alpha, col_tmp_grad;  // info columns (-1 = skip)
  int B; long long gs;
};
__global__ void __launch_bounds__(1024) minq_actor_loss_kernel(ActorLossArgs a) {
  B200_PDL_PROLOGUE();
  __shared__ float red[32];
  const int g = blockIdx.x, b = threadIdx.x, B = a.B;
  const long long o = (long long)g * a.gs;
  const float invB = 1.0f / (float)B;
  float q1 = 0.f, q2 = 0.f, lp = 0.f;
  float alpha = 0.f;
  const bool sac = (a.logp != nullptr);
  if (sac) alpha = a.log_alpha ? expf(a.log_alpha[g]) : a.alpha_const;
  if (b < B) {
    q1 = a.q1[o + b]; q2 = a.q2[o + b];
    float w1 = q1 < q2 ? 1.f : (q1 == q2 ? 0.5f : 0.f);
    a.dq1[o + b] = -w1 * invB;
    a.dq2[o + b] = -(1.f - w1) * invB;
    if (sac) { lp = a.logp[o + b]; a.dlogp[o + b] = alpha * invB; }
  }
  float term = (b < B) ? (-fminf(q1, q2) + (sac ? lp * alpha : 0.f)) : 0.f;
  const float pol = block_sum_256(term, red) * invB;
  float* info = a.info + g * a.info_gs + (long long)a.st->info_slot * a.info_stride;
  if (!sac) {
    if (b == 0 && a.col_policy >= 0) {
      info[a.col_policy] = pol;
      publish_info(a.st, g, a.col_policy, pol);
    }
    return;
  }
  const float slp = block_sum_256(b < B ? lp : 0.f, red);
  float tl = 0.f;
  if (a.log_alpha) {
    const float t = block_sum_256(b < B ? alpha * (-lp - a.target_entropy) : 0.f, red);
    tl = t * invB;
  }
  if (b == 0) {
    // the reference logs policy_obj AFTER `obj += tmp_loss` mutated it in place
    if (a.col_policy >= 0) { info[a.col_policy] = a.log_alpha ? pol + tl : pol; publish_info(a.st, g, a.col_policy, a.log_alpha ? pol + tl : pol); }
    if (a.col_entropy >= 0) { info[a.col_entropy] = -slp * invB; publish_info(a.st, g, a.col_entropy, -slp * invB); }
    if (a.log_alpha) {
      if (a.col_tmp_loss >= 0) { info[a.col_tmp_loss] = tl; publish_info(a.st, g, a.col_tmp_loss, tl); }
      if (a.col_alpha >= 0) { info[a.col_alpha] = alpha; publish_info(a.st, g, a.col_alpha, alpha); }
      if (a.col_tmp_grad >= 0) { info[a.col_tmp_grad] = tl; publish_info(a.st, g, a.col_tmp_grad, tl); }
      a.grad_log_alpha[g] = tl;
    }
  }
}

// ----------------------------------------------------------------------------- SAC tanh-Gaussian
constexpr float kSacEps = 1e-6f;            // rl/utils/annotation.py:12
constexpr float kHalfLog2Pi = 0.9189385332046727f;

// net_out [B][2A] = (mean | log_std); eps [B][A]  ->  action [B][A], logp [B].  One warp per row.
__global__ void __launch_bounds__(256) sac_sample_fwd_kernel(const float* __restrict__ net_out, int ld_out, const float* __restrict__ eps, int ld_eps,
                                                            float* __restrict__ action, int ld_a, float* __restrict__ logp, int B, int A,
                                                            float min_ls, float max_ls, long long gs_out, long long gs_eps, long long gs_a, long long gs_lp) {
  B200_PDL_PROLOGUE();
  const int g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= B) return;
  const float* o = net_out + g * gs_out + (long long)row * ld_out;
  const float* e = eps + g * gs_eps + (long long)row * ld_eps;
  float* act = action + g * gs_a + (long long)row * ld_a;
  float s1 = 0.f, s2 = 0.f;
  for (int j = lane; j < A; j += 32) {
    const float mean = o[j];
    const float ls = fminf(fmaxf(o[A + j], min_ls), max_ls);
    const float sd = expf(ls);
    const float x = __fadd_rn(mean, __fmul_rn(e[j], sd));
    const float av = tanhf(x);
    act[j] = av;
    const float d = x - mean;
    s1 += -(d * d) / (2.f * (sd * sd)) - logf(sd) - kHalfLog2Pi;
    s2 += logf(1.f - av * av + kSacEps);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) logp[g * gs_lp + row] = s1 - s2;
}

// backward of the sample for the actor step: given dL/da [B][A] and dL/dlogp [B] produce the
// gradient wrt the network output [B][2A].  (The mean / x terms of Normal.log_prob cancel
// analytically: d logN/d mean = 0, d logN/d std = -1/std.)
__global__ void __launch_bounds__(256) sac_sample_bwd_kernel(const float* __restrict__ net_out, int ld_out, const float* __restrict__ eps, int ld_eps,
                                                            const float* __restrict__ action, int ld_a, const float* __restrict__ da, int ld_da,
                                                            const float* __restrict__ dlogp, float* __restrict__ dout, int ld_dout, int B, int A,
                                                            float min_ls, float max_ls, long long gs_out, long long gs_eps, long long gs_a, long long gs_lp) {
  B200_PDL_PROLOGUE();
  const int g = blockIdx.y;
  const int row = blockIdx.x * blockDim.y + threadIdx.y;
  if (row >= B) return;
  const float* o = net_out + g * gs_out + (long long)row * ld_out;
  const float* e = eps + g * gs_eps + (long long)row * ld_eps;
  const float* act = action + g * gs_a + (long long)row * ld_a;
  const float* dA = da + g * gs_a + (long long)row * ld_da;
  float* d_o = dout + g * gs_out + (long long)row * ld_dout;
  const float dlp = dlogp[g * gs_lp + row];
  for (int j = threadIdx.x; j < A; j += blockDim.x) {
    const float raw = o[A + j];
    const float ls = fminf(fmaxf(raw, min_ls), max_ls);
    const float sd = expf(ls);
    const float av = act[j];
    const float one_m = 1.f - av * av;
    // d/da of -log(1 - a^2 + eps) = 2a / (1 - a^2 + eps)
    const float dx = (dA[j] + dlp * (2.f * av / (one_m + kSacEps))) * one_m;
    d_o[j] = dx;                                      // d mean
    const float dsd = dx * e[j] - dlp / sd;           // d std
    const bool pass = (raw >= min_ls) && (raw <= max_ls);
    d_o[A + j] = pass ? dsd * sd : 0.f;               // d log_std (exp, clamp)
  }
}

}  // namespace b200rl
