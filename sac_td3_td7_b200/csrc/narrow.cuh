// narrow.cuh -- the NARROW problems of a stacked-seed update, as bandwidth-shaped kernels.
//
// With hundreds of seeds stacked, the 256-wide layers run on the tensor cores (gemm_tc.cuh).  What is left are
// problems whose output is a handful of columns wide -- the observation / action columns of the first layers'
// weight gradients (27 and 8 wide for Ant), the policy's action head and the critics' scalar head, the gradient
// with respect to the action -- and the rank-1 "dgrad" through the critics' scalar head.  Arithmetically they are
// nothing; as 32 x 32 FFMA tiles they were 12 % of a 256-seed TD7 update (8 tiles per seed and problem, each paying the
// full tile overhead for a few useful columns), and the rank-1 product ran as a padded K = 32 tensor-core problem.
// Here each one is a single pass over its one large operand:
//   narrow_wgrad_kernel  dW[:, j0:j0+32] = dZ^T X[:, j0:j0+32] (+ the bias gradient, + Adam in place): thread = output
//                        row, the <= 32 input columns broadcast from shared memory, dZ read once, coalesced;
//   narrow_rows_kernel   Y[i, 0:J] = epilogue(sum_seg X_seg[i, :] . B_seg), J <= 32 (forward and dgrad layouts): eight
//                        lanes per row with 128-bit loads, the small operand staged in shared memory;
//   rank1_dgrad_kernel   dX[i, j] = dz[i] * w[j] (* act'(H[i, j])): elementwise.
// All three are plain CUDA C (they also run under the host emulation of tests/cuemu).
#pragma once
#include "gemm.cuh"

namespace b200rl {

// ----------------------------------------------------------------------------- narrow weight gradients
constexpr int kNwCols = 32;      // input columns per CTA
constexpr int kNwThreads = 256;  // output rows per CTA
constexpr int kNwRows = 256;     // batch rows per pass (X staged whole: 32 KB)
constexpr int kNwStageRows = 16; // rows of dZ per pipeline stage (16 KB)
constexpr int kNwStages = 4;
constexpr int kNwSmemBytes = (kNwRows * kNwCols + kNwStages * kNwStageRows * kNwThreads) * (int)sizeof(float);  // 96 KB: two CTAs per SM (8-row stages at three CTAs per SM were measured: 35 -> 43 us per 8-column problem)
struct NarrowWgrad {
  const float* dZ; long long dz_gs; int dz_ld, rows, out;  // dZ [rows][out]
  const float* X; long long x_gs; int x_ld, J;             // X  [rows][J]
  float* C; float* P; float* M; float* V;                  // gradient out (may be null) / Adam state, [out][p_ld]
  long long p_gs; int p_ld;
  int unit_begin, jchunks;
  int transposed, pad;  // 1: the result block is stored transposed, W[j][i] (a layer with a few OUTPUTS: the roles of dZ and X swap)
  // bias of the same layer (column sums of dZ), folded into the pass over dZ of column chunk 0 (null: none)
  float* bC; float* bP; float* bM; float* bV;
  const AdamScalars* adam;
};
struct NarrowWgradLaunch {
  NarrowWgrad p[kMaxProblems];
  int nprob, total_units;
  int dbg, pad;  // measurement only (B200RL_NW_DBG): bit 0 skips the X staging, bit 1 the product loop, bit 2 the epilogue
};

// product loop of one pipeline stage: thread (ti = t & 63, tj = t >> 6) owns output rows 4 ti .. 4 ti + 3 and input
// columns 8 tj .. 8 tj + 7 -- per batch row one 128-bit load of dZ (conflict-free) and two broadcast 128-bit loads of X for
// sixteen FFMA2.  (One output row per thread with the whole column block broadcast was bound by the shared-memory
// return path: 7 LDS.128 per 28 FMA, 30 us of a 64 us launch.)  Warps whose column group lies beyond J skip the loop.
__device__ __forceinline__ void nw_stage(const float* zs, const float* xs, int rn, float2 (&acc)[4][4], float4& bsum, bool sum_bias) {
#pragma unroll 4
  for (int r = 0; r < rn; ++r) {
    const float4 a = *reinterpret_cast<const float4*>(zs + r * kNwThreads);
    const float4 x0 = *reinterpret_cast<const float4*>(xs + r * kNwCols);
    const float4 x1 = *reinterpret_cast<const float4*>(xs + r * kNwCols + 4);
    if (sum_bias) { bsum.x += a.x; bsum.y += a.y; bsum.z += a.z; bsum.w += a.w; }
    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 aa = make_float2(av[k], av[k]);
      ffma2(acc[k][0], aa, make_float2(x0.x, x0.y));
      ffma2(acc[k][1], aa, make_float2(x0.z, x0.w));
      ffma2(acc[k][2], aa, make_float2(x1.x, x1.y));
      ffma2(acc[k][3], aa, make_float2(x1.z, x1.w));
    }
  }
}

// dZ streams through a 4-stage cp.async ring (16 rows x 256 output columns per stage, two CTAs per SM: ~96 KB in flight
// per SM, enough for the HBM rate); thread i owns output row i and reads dZ[b][i] from the stage (conflict-free) and
// the <= 32 input columns of batch row b as broadcast 128-bit loads.  The epilogue goes back through shared memory so
// that gradient / Adam state accesses run along the rows of W (coalesced) instead of one row per thread.
__global__ void __launch_bounds__(kNwThreads, 2) narrow_wgrad_kernel(const __grid_constant__ NarrowWgradLaunch L) {
  B200_PDL_PROLOGUE();
  B200_DYN_SMEM(float, sm);
  float* Xs = sm;                          // [kNwRows][kNwCols]
  float* Zs = sm + kNwRows * kNwCols;      // [kNwStages][kNwStageRows][kNwThreads]
  int pi = 0;
  while (pi + 1 < L.nprob && (int)blockIdx.x >= L.p[pi + 1].unit_begin) ++pi;
  const NarrowWgrad& q = L.p[pi];
  const int u = (int)blockIdx.x - q.unit_begin;
  const int ic = u / q.jchunks, jc = u - ic * q.jchunks;
  const int g = blockIdx.y;
  const int j0 = jc * kNwCols, jn = min(kNwCols, q.J - j0);
  const int nq = (jn + 3) >> 2;
  const int i0 = ic * kNwThreads;
  const int i = i0 + (int)threadIdx.x;
  const bool valid = i < q.out;
  const float* dzg = q.dZ + (long long)g * q.dz_gs;
  const float* x = q.X + (long long)g * q.x_gs + j0;
  const bool with_bias = (jc == 0) && (q.bP != nullptr || q.bC != nullptr);
  // a stage row is 256 consecutive output columns: 64 16-byte copies, thread t copies quad (t & 63) of rows (t >> 6) + 4k
  const bool vec = (q.dz_ld & 3) == 0 && (q.dz_gs & 3) == 0 && aligned16(q.dZ) && i0 + kNwThreads <= q.out;
  const int ti = threadIdx.x & 63, tj = threadIdx.x >> 6;  // (tj is uniform over a warp)
  const bool active = 8 * tj < jn;
  float2 acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = make_float2(0.f, 0.f);
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  AdamScalars ad = {};  // (requested now: its round trip hides behind the stream instead of opening the epilogue)
  if (q.P || (with_bias && q.bP)) ad = *q.adam;
  for (int b0 = 0; b0 < q.rows; b0 += kNwRows) {
    const int nb = min(kNwRows, q.rows - b0);
    const int nst = (nb + kNwStageRows - 1) / kNwStageRows;
    __syncthreads();  // (previous pass has finished reading Xs / Zs)
    auto issue = [&](int st) {
      float* dst = Zs + (st % kNwStages) * kNwStageRows * kNwThreads;
      const int r0 = st * kNwStageRows;
      if (vec) {
        const int cq = (threadIdx.x & 63) * 4;
#pragma unroll
        for (int k = 0; k < kNwStageRows / 4; ++k) {
          const int r = (threadIdx.x >> 6) + 4 * k;
          if (r0 + r < nb) cp_async16(dst + r * kNwThreads + cq, dzg + (long long)(b0 + r0 + r) * q.dz_ld + i0 + cq);
        }
      } else {
        for (int r = 0; r < kNwStageRows; ++r)
          dst[r * kNwThreads + threadIdx.x] = (valid && r0 + r < nb) ? dzg[(long long)(b0 + r0 + r) * q.dz_ld + i] : 0.f;
      }
    };
#pragma unroll
    for (int st = 0; st < kNwStages - 1; ++st) {
      if (st < nst) issue(st);
      cp_async_commit();
    }
    // X: only the jn live columns are read (read-only path, sixteen independent loads in flight per thread); the rest of
    // a 32-wide row is zero
    if (!(L.dbg & 1)) {
      if (jn < kNwCols)
        for (int idx = threadIdx.x; idx < kNwRows * kNwCols; idx += kNwThreads)
          if ((idx & (kNwCols - 1)) >= jn || idx / kNwCols >= nb) Xs[idx] = 0.f;
      const int live = nb * jn;
      for (int e0 = threadIdx.x; e0 < live; e0 += 16 * kNwThreads) {
        float xv[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int e = e0 + k * kNwThreads;
          const int b = e / jn, j = e - b * jn;
          xv[k] = e < live ? __ldg(x + (long long)(b0 + b) * q.x_ld + j) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int e = e0 + k * kNwThreads;
          const int b = e / jn, j = e - b * jn;
          if (e < live) Xs[b * kNwCols + j] = xv[k];
        }
      }
    }
    for (int st = 0; st < nst; ++st) {
      cp_async_wait<kNwStages - 2>();
      __syncthreads();  // stage st has landed for everybody (and Xs on the first turn); the stage refilled below is free
      if (st + kNwStages - 1 < nst) issue(st + kNwStages - 1);
      cp_async_commit();
      const float* zs = Zs + (st % kNwStages) * kNwStageRows * kNwThreads + 4 * ti;
      const float* xs = Xs + st * kNwStageRows * kNwCols + 8 * tj;
      const int rn = (L.dbg & 2) ? 0 : min(kNwStageRows, nb - st * kNwStageRows);
      if (active) nw_stage(zs, xs, rn, acc, bsum, with_bias && tj == 0);
    }
    cp_async_wait<0>();
  }
  if (L.dbg & 4) return;
  // ---- epilogue: the [256 rows][jn] gradient block goes through shared memory (row stride 33: conflict-free both ways)
  __syncthreads();  // everybody is done with the stages
  float* G = sm;  // (X and the stages are dead)
  constexpr int kGs = kNwCols + 1;
  static_assert(kNwThreads * (kNwCols + 1) * (int)sizeof(float) <= kNwSmemBytes, "the gradient block must fit the kernel's shared memory");
  if (active) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        G[(4 * ti + a) * kGs + 8 * tj + 2 * b] = acc[a][b].x;
        G[(4 * ti + a) * kGs + 8 * tj + 2 * b + 1] = acc[a][b].y;
      }
  }
  __syncthreads();
  const int rows_here = min(kNwThreads, q.out - i0);
  const bool tr = q.transposed != 0;
  const long long base = (long long)g * q.p_gs + (tr ? (long long)j0 * q.p_ld + i0 : (long long)i0 * q.p_ld + j0);
  const int total = rows_here * jn;
  // element e -> (row e / jn, column e % jn): runs of jn consecutive floats of W's rows; eight elements in flight
  for (int e0 = threadIdx.x; e0 < total; e0 += 8 * kNwThreads) {
    float gq[8], pp[8], mm[8], vv[8];
    long long off[8];
    bool ok[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = e0 + k * kNwThreads;
      ok[k] = e < total;
      int r, c;  // (consecutive e run along the contiguous direction of W)
      if (tr) { c = ok[k] ? e / rows_here : 0; r = ok[k] ? e - c * rows_here : 0; }
      else { r = ok[k] ? e / jn : 0; c = ok[k] ? e - r * jn : 0; }
      off[k] = base + (tr ? (long long)c * q.p_ld + r : (long long)r * q.p_ld + c);
      gq[k] = G[r * kGs + c];
      if (ok[k] && q.P) { pp[k] = q.P[off[k]]; mm[k] = q.M[off[k]]; vv[k] = q.V[off[k]]; }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (!ok[k]) continue;
      if (q.P) {
        adam_update(gq[k], pp[k], mm[k], vv[k], ad);
        q.P[off[k]] = pp[k]; q.M[off[k]] = mm[k]; q.V[off[k]] = vv[k];
      }
      if (q.C) q.C[off[k]] = gq[k];
    }
  }
  if (with_bias && tj == 0) {
    const float bs[4] = {bsum.x, bsum.y, bsum.z, bsum.w};
    const long long o = (long long)g * q.p_gs + i0 + 4 * ti;
    const int nv = min(4, q.out - (i0 + 4 * ti));
    if (nv > 0) {
      float p[4], m[4], v[4];
      if (q.bP) {
        load_quad(p, q.bP + o, nv); load_quad(m, q.bM + o, nv); load_quad(v, q.bV + o, nv);
        float gq[4] = {bs[0], bs[1], bs[2], bs[3]};
#pragma unroll
        for (int a = 0; a < 4; ++a) adam_update(gq[a], p[a], m[a], v[a], ad);
        store_quad(q.bP + o, p, nv); store_quad(q.bM + o, m, nv); store_quad(q.bV + o, v, nv);
      }
      if (q.bC) store_quad(q.bC + o, bs, nv);
    }
  }
}

// ----------------------------------------------------------------------------- narrow outputs (forward / dgrad, J <= 8)
constexpr int kNrMaxJ = 8;      // output columns per pass
constexpr int kNrMaxCols = 32;  // widest output handled (in passes of kNrMaxJ)
constexpr int kNrMaxK = 1024;   // total reduction length over the segments
constexpr int kNrThreads = 256; // 32 rows x 8 lanes
struct NarrowRows {
  GemmSeg seg[3];
  int nseg, I, J;
  float* C; long long c_gs; int c_ld, act;
  const float* bias; long long bias_gs;
  const float* aux; long long aux_gs; int aux_ld, accumulate;
  float noise_sigma, noise_clip;
  int unit_begin, pad;
};
struct NarrowRowsLaunch {
  NarrowRows p[kMaxProblems];
  int nprob, total_units;
};

// MODE kFwd: B_seg(r, j) = B[j * b_ld + r];  kDgrad: B_seg(r, j) = B[r * b_ld + j]
template <int MODE>
__global__ void __launch_bounds__(kNrThreads) narrow_rows_kernel(const __grid_constant__ NarrowRowsLaunch L) {
  B200_PDL_PROLOGUE();
  __shared__ __align__(16) float Ws[kNrMaxJ][kNrMaxK];
  int pi = 0;
  while (pi + 1 < L.nprob && (int)blockIdx.x >= L.p[pi + 1].unit_begin) ++pi;
  const NarrowRows& q = L.p[pi];
  const int g = blockIdx.y;
  const int row = ((int)blockIdx.x - q.unit_begin) * 32 + (threadIdx.x >> 3), lq = threadIdx.x & 7;
  const bool valid = row < q.I;
  // output columns in passes of eight (Humanoid's 17-wide action head: three passes over rows that stay in L1 / L2)
  for (int jb = 0; jb < q.J; jb += kNrMaxJ) {
    const int J = min(kNrMaxJ, q.J - jb);
    if (jb) __syncthreads();  // the previous pass has finished reading Ws
    // stage the small operand: Ws[j][k] over the concatenated reduction index k
    int koff = 0;
    for (int s = 0; s < q.nseg; ++s) {
      const GemmSeg& sg = q.seg[s];
      const float* Bp = sg.B + (long long)g * sg.b_gs;
      // (read-only loads, four in flight per thread: a plain load / store loop serialises on possible aliasing)
      for (int base = 0; base < sg.len * J; base += 4 * kNrThreads) {
        float wv[4];
        int wj[4], wr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int idx = base + k * kNrThreads + (int)threadIdx.x;
          const bool ok = idx < sg.len * J;
          int r, j;
          if (MODE == kFwd) { j = idx / sg.len; r = idx - j * sg.len; }
          else { r = idx / J; j = idx - r * J; }
          wj[k] = ok ? j : -1; wr[k] = r;
          wv[k] = ok ? __ldg(Bp + (MODE == kFwd ? (long long)(jb + j) * sg.b_ld + r : (long long)r * sg.b_ld + jb + j)) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (wj[k] >= 0) Ws[wj[k]][koff + wr[k]] = wv[k];
      }
      koff += sg.len;
    }
    __syncthreads();
    float acc[kNrMaxJ];
#pragma unroll
    for (int j = 0; j < kNrMaxJ; ++j) acc[j] = 0.f;
    koff = 0;
    for (int s = 0; s < q.nseg; ++s) {
      const GemmSeg& sg = q.seg[s];
      const float* a = sg.A + (long long)g * sg.a_gs + (long long)(valid ? row : 0) * sg.a_ld;
      // (eligibility: len, a_ld multiples of 4 and 16-byte aligned bases) four 128-bit loads in flight per lane
      for (int r0 = lq * 4; r0 < sg.len; r0 += 128) {
        float4 xv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int r = r0 + 32 * t;
          xv[t] = (valid && r < sg.len) ? *reinterpret_cast<const float4*>(a + r) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int r = r0 + 32 * t;
          if (r < sg.len) {
#pragma unroll
            for (int j = 0; j < kNrMaxJ; ++j) {
              if (j < J) {
                const float4 wv = *reinterpret_cast<const float4*>(&Ws[j][koff + r]);
                acc[j] = fmaf(xv[t].x, wv.x, fmaf(xv[t].y, wv.y, fmaf(xv[t].z, wv.z, fmaf(xv[t].w, wv.w, acc[j]))));
              }
            }
          }
        }
      }
      koff += sg.len;
    }
    // the eight lanes of a row combine; lane j finishes output column jb + j
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < kNrMaxJ; ++j) {
      float v = acc[j];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (j == lq) mine = v;
    }
    if (valid && lq < J) {
      const int col = jb + lq;
      float* c = q.C + (long long)g * q.c_gs + (long long)row * q.c_ld + col;
      float v = mine;
      if (MODE == kFwd) {
        if (q.bias) v += q.bias[(long long)g * q.bias_gs + col];
        if (q.act == kActTanhNoise) {
          const float n = fminf(fmaxf(q.aux[(long long)g * q.aux_gs + (long long)row * q.aux_ld + col] * q.noise_sigma, -q.noise_clip), q.noise_clip);
          v = fminf(fmaxf(tanhf(v) + n, -1.f), 1.f);
        } else {
          v = act_forward(v, q.act);
        }
      } else {
        if (q.accumulate) v += *c;
        if (q.aux) v *= act_backward(q.aux[(long long)g * q.aux_gs + (long long)row * q.aux_ld + col], q.act);
      }
      *c = v;
    }
  }
}

// ----------------------------------------------------------------------------- rank-1 dgrad (through a one-output layer)
struct Rank1Problem {
  const float* dz; long long dz_gs; int dz_ld;  // dZ [I][1]
  const float* w; long long w_gs;               // W[0][k_lo : k_lo + J]
  const float* aux; long long aux_gs; int aux_ld, act;
  float* C; long long c_gs; int c_ld, accumulate;
  int I, J;
  long long quad_begin;  // first (row, column quad) of this problem within the launch
};
struct Rank1Launch {
  Rank1Problem p[kMaxProblems];
  int nprob, pad;
  long long total_quads;
};
__global__ void __launch_bounds__(256) rank1_dgrad_kernel(const __grid_constant__ Rank1Launch L) {
  B200_PDL_PROLOGUE();
  const int g = blockIdx.y;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < L.total_quads; t += (long long)gridDim.x * blockDim.x) {
    int pi = 0;
    while (pi + 1 < L.nprob && t >= L.p[pi + 1].quad_begin) ++pi;
    const Rank1Problem& q = L.p[pi];
    const long long lt = t - q.quad_begin;
    const int nq = q.J >> 2;
    const int i = (int)(lt / nq), j = (int)(lt - (long long)i * nq) * 4;
    const float d = q.dz[(long long)g * q.dz_gs + (long long)i * q.dz_ld];
    const float4 w = *reinterpret_cast<const float4*>(q.w + (long long)g * q.w_gs + j);
    float o[4] = {d * w.x, d * w.y, d * w.z, d * w.w};
    float* c = q.C + (long long)g * q.c_gs + (long long)i * q.c_ld + j;
    if (q.accumulate) {
      const float4 cv = *reinterpret_cast<const float4*>(c);
      o[0] += cv.x; o[1] += cv.y; o[2] += cv.z; o[3] += cv.w;
    }
    if (q.aux) {
      const float4 h = *reinterpret_cast<const float4*>(q.aux + (long long)g * q.aux_gs + (long long)i * q.aux_ld + j);
      o[0] *= act_backward(h.x, q.act); o[1] *= act_backward(h.y, q.act); o[2] *= act_backward(h.z, q.act); o[3] *= act_backward(h.w, q.act);
    }
    *reinterpret_cast<float4*>(c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ----------------------------------------------------------------------------- eligibility (host)
inline bool ptr16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool narrow_wgrad_eligible(const GemmProblem& p) {
  return p.nseg == 1 && !p.seg[0].b_ones && p.J < 64 && (p.P != nullptr || p.C != nullptr);
}
// dW[i, :] for a layer with at most 32 outputs (action head, scalar head): computed as its transpose
inline bool narrow_wgrad_t_eligible(const GemmProblem& p) {
  return p.nseg == 1 && !p.seg[0].b_ones && p.I <= kNwCols && p.J >= 64 && (p.P != nullptr || p.C != nullptr);
}
inline bool narrow_rows_eligible(const GemmProblem& p, int mode) {
  if (mode == kWgrad || p.J > kNrMaxCols) return false;
  int k = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const GemmSeg& sg = p.seg[s];
    if (sg.b_ones || (sg.len & 3) || (sg.a_ld & 3) || (sg.a_gs & 3) || !ptr16(sg.A)) return false;
    k += sg.len;
  }
  return k <= kNrMaxK;
}
inline bool rank1_eligible(const GemmProblem& p, int mode) {
  if (mode != kDgrad || p.nseg != 1 || p.seg[0].len != 1 || p.seg[0].b_ones) return false;
  const GemmSeg& sg = p.seg[0];
  return (p.J & 3) == 0 && (p.c_ld & 3) == 0 && (p.c_gs & 3) == 0 && ptr16(p.C) && ptr16(sg.B) && (sg.b_gs & 3) == 0 &&
         (!p.aux || ((p.aux_ld & 3) == 0 && (p.aux_gs & 3) == 0 && ptr16(p.aux)));
}

}  // namespace b200rl
