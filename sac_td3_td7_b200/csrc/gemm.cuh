// gemm.cuh -- grouped, fused fp32 FFMA GEMM for the 256-row MLP layers of the update path.
//
// One launch executes a TABLE of independent problems (the layers of a "phase" that have
// no data dependence on each other: e.g. both critics, or the target and the current
// encoder pass), each optionally replicated over `groups` stacked seeds (blockIdx.y).
// Every problem computes
//     C[i,j] = epilogue( sum_seg sum_r A_seg(i,r) * B_seg(r,j) )
// with up to three reduction segments, which is how torch.cat([...],1) inputs
// (rl/nn/sale.py:52,80,114-117 of the reference) are consumed without materialising the
// concatenation.  Three operand layouts cover forward / dgrad / wgrad:
//     kFwd   (Y = X W^T)   A = X[i][r]  (r contiguous)   B = W[j][r]  (r contiguous)
//     kDgrad (dX = dZ W)   A = dZ[i][r] (r contiguous)   B = W[r][j]  (j contiguous)
//     kWgrad (dW = dZ^T X) A = dZ[r][i] (i contiguous)   B = X[r][j]  (j contiguous)
// Epilogues: forward adds bias and applies ReLU / ELU / tanh (+ TD3/TD7 target-policy
// smoothing noise); dgrad multiplies by the activation derivative taken from the saved
// post-activation output (and can accumulate); wgrad optionally stores the gradient and
// applies Adam in place (torch.optim.Adam single-tensor ordering), so gradients never
// round-trip through HBM in the stacked-seed case.
//
// Tiling: CTA = 128 threads, 32x32 output tile, 32-deep reduction chunks, double-buffered
// through registers.  The four warps split each chunk's reduction (8 r's each) and every
// lane owns a 4x8 register tile (32 FFMA per 3 LDS.128); partial tiles are summed in a
// fixed order through shared memory, so results are bit-reproducible run to run.
// Shared tiles are stored reduction-major with an XOR swizzle at float4 granularity so
// that both the transposing stores (r-contiguous global operands) and the float4 reads
// are bank-conflict free.
//
// Why FFMA and not tcgen05 here: a single agent's layer is 256x256x(<=768) -- 64 tiles --
// and parity is fp32 at 1e-5 (BASELINE.json north_star), which single-pass TF32/BF16
// tensor-core math cannot give.  The stacked-seed path is where tensor cores apply.
#pragma once
#include "common.cuh"

namespace b200rl {

enum GemmMode : int { kFwd = 0, kDgrad = 1, kWgrad = 2 };
enum Act : int { kActNone = 0, kActRelu = 1, kActElu = 2, kActTanh = 3, kActTanhNoise = 4 };

// Adam scalars for one optimizer, refreshed on device once per update by `prep_step_kernel`
// (float casts of the Python doubles torch computes in `_single_tensor_adam`).
struct AdamScalars {
  float w1;          // 1 - beta1                      (lerp weight)
  float beta2;       // beta2
  float w2;          // 1 - beta2                      (addcmul value)
  float bc2_sqrt;    // sqrt(1 - beta2^t)
  float neg_step;    // -(lr / (1 - beta1^t))          (addcdiv value)
  float eps;
  int step;          // t (1-based), for inspection
  int pad;
};

__device__ __forceinline__ void adam_update(float g, float& p, float& m, float& v, const AdamScalars& s) {
  // exp_avg.lerp_(g, 1-b1)  -> fma(w, g-m, m)                    (ATen lerp_vec)
  m = __fmaf_rn(s.w1, __fsub_rn(g, m), m);
  // exp_avg_sq.mul_(b2).addcmul_(g, g, value=1-b2) -> fma(w2*g, g, v*b2)
  v = __fmaf_rn(__fmul_rn(s.w2, g), g, __fmul_rn(v, s.beta2));
  // denom = sqrt(v)/sqrt(bc2) + eps ; p += (neg_step*m)/denom
  float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt), s.eps);
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(s.neg_step, m), denom));
}

struct GemmSeg {
  const float* A;
  const float* B;
  long long a_gs, b_gs;  // per-group (stacked seed) strides, in elements
  int a_ld, b_ld;
  int len;     // reduction length of this segment
  int b_ones;  // B(r,j) := (j == 0) ? 1 : 0   -- bias-gradient problems (column sums of dZ)
};

struct GemmProblem {
  GemmSeg seg[3];
  int nseg;
  int I, J;  // output rows, cols
  float* C;  // fwd: Y; dgrad: dX; wgrad: gradient out (may be null)
  long long c_gs;
  int c_ld;
  int act;         // fwd: activation; dgrad: activation whose derivative multiplies (0 = none)
  const float* bias;  // fwd
  long long bias_gs;
  const float* aux;  // fwd+kActTanhNoise: noise[i][j]; dgrad: H (saved post-activation output)
  long long aux_gs;
  int aux_ld;
  int accumulate;  // dgrad: C += result (before the derivative multiply)
  float* P;        // wgrad: Adam parameter / exp_avg / exp_avg_sq (null = no Adam)
  float* M;
  float* V;
  long long p_gs;
  int p_ld;
  float noise_sigma, noise_clip;
  const AdamScalars* adam;
  int tile_begin;  // first linear tile of this problem within the launch
  int tiles_j;
};

constexpr int kTile = 32;
constexpr int kGemmThreads = 128;

__device__ __forceinline__ int swz(int r, int c) { return r * kTile + (c ^ (((r >> 2) & 7) << 2)); }

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// load a 4-wide strip along the CONTIGUOUS direction of a global operand, zero filled
__device__ __forceinline__ float4 load_strip(const float* row_ptr, int pos, int limit, bool row_ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok || pos >= limit) return v;
  const float* p = row_ptr + pos;
  if (pos + 3 < limit && aligned16(p)) {
    v = *reinterpret_cast<const float4*>(p);
  } else {
    v.x = p[0];
    if (pos + 1 < limit) v.y = p[1];
    if (pos + 2 < limit) v.z = p[2];
    if (pos + 3 < limit) v.w = p[3];
  }
  return v;
}

// operand element (c, r): RC layout = r contiguous, CC layout = c contiguous.
// Each thread moves two float4 strips per operand per chunk.
template <bool RC>
__device__ __forceinline__ void operand_fetch(const float* base, int ld, int c0, int cmax, int r0, int rlen, int tid, float4 (&reg)[2]) {
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    int q = tid + kGemmThreads * t;
    if (RC) {
      int c = q >> 3, r4 = (q & 7) << 2;
      bool ok = (c0 + c) < cmax;
      reg[t] = load_strip(base + (long long)(c0 + c) * ld, r0 + r4, rlen, ok);
    } else {
      int r = q >> 3, c4 = (q & 7) << 2;
      bool ok = (r0 + r) < rlen;
      reg[t] = load_strip(base + (long long)(r0 + r) * ld, c0 + c4, cmax, ok);
    }
  }
}

template <bool RC>
__device__ __forceinline__ void operand_stash(float* tile, int tid, const float4 (&reg)[2]) {
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    int q = tid + kGemmThreads * t;
    if (RC) {
      int c = q >> 3, r4 = (q & 7) << 2;
      tile[swz(r4 + 0, c)] = reg[t].x;
      tile[swz(r4 + 1, c)] = reg[t].y;
      tile[swz(r4 + 2, c)] = reg[t].z;
      tile[swz(r4 + 3, c)] = reg[t].w;
    } else {
      int r = q >> 3, c4 = (q & 7) << 2;
      *reinterpret_cast<float4*>(&tile[swz(r, c4)]) = reg[t];
    }
  }
}

__device__ __forceinline__ void ones_fetch(int c0, int r0, int rlen, int tid, float4 (&reg)[2]) {
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    int q = tid + kGemmThreads * t;
    int r = q >> 3, c4 = (q & 7) << 2;
    float one = ((r0 + r) < rlen && (c0 + c4) == 0) ? 1.f : 0.f;
    reg[t] = make_float4(one, 0.f, 0.f, 0.f);
  }
}

__device__ __forceinline__ float act_forward(float v, int act) {
  switch (act) {
    case kActRelu: return v > 0.f ? v : 0.f;
    case kActElu: return v > 0.f ? v : expm1f(v);
    case kActTanh:
    case kActTanhNoise: return tanhf(v);
    default: return v;
  }
}
// derivative expressed through the saved post-activation output h
__device__ __forceinline__ float act_backward(float h, int act) {
  switch (act) {
    case kActRelu: return h > 0.f ? 1.f : 0.f;
    case kActElu: return h > 0.f ? 1.f : h + 1.f;
    case kActTanh: return 1.f - h * h;
    default: return 1.f;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kGemmThreads) gemm_kernel(const GemmProblem* __restrict__ probs, int nprob) {
  constexpr bool A_RC = (MODE != kWgrad);
  constexpr bool B_RC = (MODE == kFwd);
  __shared__ __align__(16) float smem[4 * kTile * kTile];  // 2 stages x (A,B) tiles; reused for the 4 partial tiles
  __shared__ GemmProblem sp;

  const int tid = threadIdx.x;
  const int g = blockIdx.y;
  {
    int pi = 0;
    const int tile_id = blockIdx.x;
    while (pi + 1 < nprob && tile_id >= probs[pi + 1].tile_begin) ++pi;
    const int* src = reinterpret_cast<const int*>(&probs[pi]);
    int* dst = reinterpret_cast<int*>(&sp);
    for (int w = tid; w < (int)(sizeof(GemmProblem) / sizeof(int)); w += kGemmThreads) dst[w] = src[w];
  }
  __syncthreads();
  const int t_local = (int)blockIdx.x - sp.tile_begin;
  const int i0 = (t_local / sp.tiles_j) * kTile;
  const int j0 = (t_local % sp.tiles_j) * kTile;
  const int I = sp.I, J = sp.J;

  const int warp = tid >> 5, lane = tid & 31;
  const int li = lane >> 2, lj = lane & 3;

  float acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

  // flattened (segment, chunk) iteration
  int seg = 0, r0 = 0;
  int nchunks = 0;
  for (int s = 0; s < sp.nseg; ++s) nchunks += (sp.seg[s].len + kTile - 1) / kTile;

  float4 ra[2], rb[2];
  auto fetch = [&](int s, int r) {
    const GemmSeg& sg = sp.seg[s];
    const float* A = sg.A + (long long)g * sg.a_gs;
    operand_fetch<A_RC>(A, sg.a_ld, i0, I, r, sg.len, tid, ra);
    if (sg.b_ones) {
      ones_fetch(j0, r, sg.len, tid, rb);
    } else {
      const float* B = sg.B + (long long)g * sg.b_gs;
      operand_fetch<B_RC>(B, sg.b_ld, j0, J, r, sg.len, tid, rb);
    }
  };
  auto advance = [&]() {
    r0 += kTile;
    if (r0 >= sp.seg[seg].len) { r0 = 0; ++seg; }
  };

  if (nchunks > 0) {
    fetch(seg, r0);
    advance();
    operand_stash<A_RC>(smem, tid, ra);
    // a b_ones tile is produced in CC form
    if (B_RC) operand_stash<true>(smem + kTile * kTile, tid, rb); else operand_stash<false>(smem + kTile * kTile, tid, rb);
  }
  __syncthreads();
  int stage = 0;
  for (int c = 0; c < nchunks; ++c) {
    const bool more = (c + 1 < nchunks);
    if (more) { fetch(seg, r0); advance(); }
    const float* As = smem + stage * 2 * kTile * kTile;
    const float* Bs = As + kTile * kTile;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int r = warp * 8 + rr;
      const float4 a4 = *reinterpret_cast<const float4*>(&As[swz(r, li * 4)]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[swz(r, lj * 8)]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[swz(r, lj * 8 + 4)]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    if (more) {
      float* An = smem + (stage ^ 1) * 2 * kTile * kTile;
      operand_stash<A_RC>(An, tid, ra);
      if (B_RC) operand_stash<true>(An + kTile * kTile, tid, rb); else operand_stash<false>(An + kTile * kTile, tid, rb);
    }
    __syncthreads();
    stage ^= 1;
  }

  // ---- cross-warp reduction of the four K-slices (fixed order) ----
  float* part = smem;  // [4][32][32]
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float* row = part + warp * kTile * kTile + (li * 4 + a) * kTile + lj * 8;
    *reinterpret_cast<float4*>(row) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    *reinterpret_cast<float4*>(row + 4) = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
  }
  __syncthreads();
  const int oi = tid >> 2, ojb = (tid & 3) * 8;
  float out[8];
  {
    const float* p0 = part + oi * kTile + ojb;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s = p0[e];
      s += p0[kTile * kTile + e];
      s += p0[2 * kTile * kTile + e];
      s += p0[3 * kTile * kTile + e];
      out[e] = s;
    }
  }
  const int gi = i0 + oi;
  if (gi >= I) return;
  const int gj0 = j0 + ojb;
  if (gj0 >= J) return;
  const int nvalid = min(8, J - gj0);

  if (MODE == kFwd) {
    const float* bias = sp.bias ? sp.bias + (long long)g * sp.bias_gs : nullptr;
    const float* noise = (sp.act == kActTanhNoise) ? sp.aux + (long long)g * sp.aux_gs + (long long)gi * sp.aux_ld : nullptr;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (e < nvalid) {
        float v = out[e] + (bias ? bias[gj0 + e] : 0.f);
        v = act_forward(v, sp.act);
        if (noise) {
          float n = fminf(fmaxf(noise[gj0 + e] * sp.noise_sigma, -sp.noise_clip), sp.noise_clip);
          v = fminf(fmaxf(v + n, -1.f), 1.f);
        }
        out[e] = v;
      }
    }
  } else if (MODE == kDgrad) {
    float* Cg = sp.C + (long long)g * sp.c_gs + (long long)gi * sp.c_ld;
    const float* H = sp.aux ? sp.aux + (long long)g * sp.aux_gs + (long long)gi * sp.aux_ld : nullptr;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (e < nvalid) {
        float v = out[e];
        if (sp.accumulate) v += Cg[gj0 + e];
        if (H) v *= act_backward(H[gj0 + e], sp.act);
        out[e] = v;
      }
    }
  } else {  // kWgrad
    if (sp.P) {
      const AdamScalars s = *sp.adam;
      float* Pp = sp.P + (long long)g * sp.p_gs + (long long)gi * sp.p_ld;
      float* Mp = sp.M + (long long)g * sp.p_gs + (long long)gi * sp.p_ld;
      float* Vp = sp.V + (long long)g * sp.p_gs + (long long)gi * sp.p_ld;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (e < nvalid) {
          float p = Pp[gj0 + e], m = Mp[gj0 + e], v = Vp[gj0 + e];
          adam_update(out[e], p, m, v, s);
          Pp[gj0 + e] = p;
          Mp[gj0 + e] = m;
          Vp[gj0 + e] = v;
        }
      }
    }
    if (!sp.C) return;
  }
  float* Crow = sp.C + (long long)g * sp.c_gs + (long long)gi * sp.c_ld + gj0;
  if (nvalid == 8 && aligned16(Crow)) {
    *reinterpret_cast<float4*>(Crow) = make_float4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<float4*>(Crow + 4) = make_float4(out[4], out[5], out[6], out[7]);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (e < nvalid) Crow[e] = out[e];
  }
}

// ----------------------------------------------------------------------------- host side
// A launch table: problems are appended on the host, uploaded once, then replayed.
struct GemmTable {
  GemmProblem* d_probs = nullptr;
  int nprob = 0;
  int total_tiles = 0;
  int mode = kFwd;
};

}  // namespace b200rl
