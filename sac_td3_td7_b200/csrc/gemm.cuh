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
// Tiling: CTA = 256 threads, 32x32 output tile, 32-deep reduction chunks in a 3-stage
// cp.async (LDGSTS) pipeline: every thread issues one 16-byte copy per operand per chunk
// straight from global/L2 into shared memory (no register staging, no transposition), two
// chunks ahead of the math.  The eight warps split each chunk's reduction (4 r's each) and
// every lane owns a 4x8 register tile (128 FFMA per 12 LDS.128); the eight partial tiles
// are summed in a fixed order through shared memory, so results are bit-reproducible run
// to run.  Operands that are contiguous along the reduction in global memory keep that
// orientation in shared memory, XOR-swizzled at float4 granularity so the strided float4
// reads are bank-conflict free; ragged / unaligned strips (obs and action widths such as
// 17, 6, 23) take a zero-filling register path.
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
constexpr int kGemmThreads = 256;
constexpr int kStages = 3;
constexpr int kTileElems = kTile * kTile;

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- cp.async (LDGSTS) helpers; under host emulation they degrade to plain copies ----
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
#ifdef B200_EMU
  memcpy(smem_dst, gmem_src, 16);
#else
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src));
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef B200_EMU
  asm volatile("cp.async.commit_group;\n" ::);
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef B200_EMU
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
#endif
}

// A 32x32 operand tile in shared memory, 8 float4 "slots" per 32-float row.
//   K-major (operand contiguous along the reduction r in global): row = c (output index), slots along r,
//     XOR-swizzled by (c >> 2) so the per-lane float4 reads along r of 8 different rows hit 8 bank groups.
//   C-major (operand contiguous along the output index c in global): row = r, slots along c, no swizzle.
// Either way a thread copies one 16-byte strip per operand per chunk, straight from global with cp.async
// when the strip is complete and 16-byte aligned, else through registers with zero fill.
template <bool KMAJOR>
__device__ __forceinline__ void stage_operand(float* tile, const float* base, int ld, int c0, int cmax, int r0, int rlen, int tid) {
  const int row = tid >> 3, s = tid & 7;
  float* dst;
  const float* src;
  int pos, limit;
  bool row_ok;
  if (KMAJOR) {
    dst = tile + row * kTile + ((s ^ ((row >> 2) & 7)) << 2);
    row_ok = (c0 + row) < cmax;
    pos = r0 + (s << 2);
    limit = rlen;
    src = base + (long long)(c0 + row) * ld + pos;
  } else {
    dst = tile + row * kTile + (s << 2);
    row_ok = (r0 + row) < rlen;
    pos = c0 + (s << 2);
    limit = cmax;
    src = base + (long long)(r0 + row) * ld + pos;
  }
  if (row_ok && pos + 3 < limit && aligned16(src)) {
    cp_async16(dst, src);
  } else {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_ok && pos < limit) {
      v.x = src[0];
      if (pos + 1 < limit) v.y = src[1];
      if (pos + 2 < limit) v.z = src[2];
      if (pos + 3 < limit) v.w = src[3];
    }
    *reinterpret_cast<float4*>(dst) = v;
  }
}

// bias-gradient problems: B(r, j) = (j == 0)
__device__ __forceinline__ void stage_ones(float* tile, int c0, int r0, int rlen, int tid) {
  const int row = tid >> 3, s = tid & 7;
  const float one = ((r0 + row) < rlen && s == 0 && c0 == 0) ? 1.f : 0.f;
  *reinterpret_cast<float4*>(tile + row * kTile + (s << 2)) = make_float4(one, 0.f, 0.f, 0.f);
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
  constexpr bool A_KM = (MODE != kWgrad);  // A contiguous along the reduction in global memory
  constexpr bool B_KM = (MODE == kFwd);
  // 3 pipeline stages x (A tile, B tile) = 24 KB; the same 32 KB later holds the 8 partial output tiles
  __shared__ __align__(16) float smem[8 * kTileElems];
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

  int nchunks = 0;
  for (int s = 0; s < sp.nseg; ++s) nchunks += (sp.seg[s].len + kTile - 1) / kTile;

  // producer cursor over the flattened (segment, chunk) sequence
  int pseg = 0, pr0 = 0;
  auto issue = [&](int stage) {
    const GemmSeg& sg = sp.seg[pseg];
    float* At = smem + stage * 2 * kTileElems;
    float* Bt = At + kTileElems;
    stage_operand<A_KM>(At, sg.A + (long long)g * sg.a_gs, sg.a_ld, i0, I, pr0, sg.len, tid);
    if (sg.b_ones) stage_ones(Bt, j0, pr0, sg.len, tid);
    else stage_operand<B_KM>(Bt, sg.B + (long long)g * sg.b_gs, sg.b_ld, j0, J, pr0, sg.len, tid);
    pr0 += kTile;
    if (pr0 >= sg.len) { pr0 = 0; ++pseg; }
  };

  // prologue: fill kStages-1 stages
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nchunks) issue(s);
    cp_async_commit();
  }

  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<kStages - 2>();  // chunk c has landed (for this thread's copies)
    __syncthreads();               // ... and for everybody's; also: stage (c-1)%3 is no longer being read
    if (c + kStages - 1 < nchunks) issue((c + kStages - 1) % kStages);
    cp_async_commit();

    const float* As = smem + (c % kStages) * 2 * kTileElems;
    const float* Bs = As + kTileElems;
    float av[4][4], bv[8][4];  // [output index][r within this warp's 4-slice]
    if (A_KM) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float4 v = *reinterpret_cast<const float4*>(&As[(li * 4 + a) * kTile + ((warp ^ li) << 2)]);
        av[a][0] = v.x; av[a][1] = v.y; av[a][2] = v.z; av[a][3] = v.w;
      }
    } else {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const float4 v = *reinterpret_cast<const float4*>(&As[(warp * 4 + rr) * kTile + li * 4]);
        av[0][rr] = v.x; av[1][rr] = v.y; av[2][rr] = v.z; av[3][rr] = v.w;
      }
    }
    if (B_KM) {
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int col = lj * 8 + b;
        const float4 v = *reinterpret_cast<const float4*>(&Bs[col * kTile + ((warp ^ ((col >> 2) & 7)) << 2)]);
        bv[b][0] = v.x; bv[b][1] = v.y; bv[b][2] = v.z; bv[b][3] = v.w;
      }
    } else {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const float4 v0 = *reinterpret_cast<const float4*>(&Bs[(warp * 4 + rr) * kTile + lj * 8]);
        const float4 v1 = *reinterpret_cast<const float4*>(&Bs[(warp * 4 + rr) * kTile + lj * 8 + 4]);
        bv[0][rr] = v0.x; bv[1][rr] = v0.y; bv[2][rr] = v0.z; bv[3][rr] = v0.w;
        bv[4][rr] = v1.x; bv[5][rr] = v1.y; bv[6][rr] = v1.z; bv[7][rr] = v1.w;
      }
    }
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a][rr], bv[b][rr], acc[a][b]);
  }
  cp_async_wait<0>();
  __syncthreads();  // everyone is done reading the pipeline stages: reuse the memory for the partial tiles

  // ---- cross-warp reduction of the eight K-slices, fixed order (bit-reproducible) ----
  // partial tile layout: [warp][row][8 slots of float4], slot index XOR (row >> 2) to spread banks
  float* part = smem;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int row = li * 4 + a;
    float* base = part + warp * kTileElems + row * kTile;
    *reinterpret_cast<float4*>(base + (((2 * lj) ^ li) << 2)) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    *reinterpret_cast<float4*>(base + (((2 * lj + 1) ^ li) << 2)) = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
  }
  __syncthreads();
  const int oi = tid >> 3, oslot = tid & 7;
  float out[4];
  {
    const float* p0 = part + oi * kTile + ((oslot ^ ((oi >> 2) & 7)) << 2);
    float4 s = *reinterpret_cast<const float4*>(p0);
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(p0 + w * kTileElems);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[0] = s.x; out[1] = s.y; out[2] = s.z; out[3] = s.w;
  }
  const int gi = i0 + oi;
  if (gi >= I) return;
  const int gj0 = j0 + oslot * 4;
  if (gj0 >= J) return;
  const int nvalid = min(4, J - gj0);

  if (MODE == kFwd) {
    const float* bias = sp.bias ? sp.bias + (long long)g * sp.bias_gs : nullptr;
    const float* noise = (sp.act == kActTanhNoise) ? sp.aux + (long long)g * sp.aux_gs + (long long)gi * sp.aux_ld : nullptr;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
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
    for (int e = 0; e < 4; ++e) {
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
      for (int e = 0; e < 4; ++e) {
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
  if (nvalid == 4 && aligned16(Crow)) {
    *reinterpret_cast<float4*>(Crow) = make_float4(out[0], out[1], out[2], out[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
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
