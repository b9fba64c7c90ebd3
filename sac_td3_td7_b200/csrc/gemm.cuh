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

// One launch: the problem table travels in kernel-parameter (constant) memory, so a CTA finds its
// problem without a dependent global load.  With splitk > 1 every output tile is computed by `splitk`
// CTAs, each over a contiguous slice of the reduction chunks; the partial tiles meet in `scratch` and the
// CTA that arrives last (ticket counter) sums them in split order -- so the result does not depend on
// arrival order -- and runs the epilogue.  Splitting is what gives a single agent's 64-tile layers
// enough CTAs to occupy the machine and cuts the serial chunk chain per CTA.
constexpr int kMaxProblems = 24;
struct GemmLaunch {
  GemmProblem p[kMaxProblems];
  int nprob;
  int splitk;
  int total_tiles;
  int pad;
  float* scratch;      // [groups][total_tiles][splitk][32*32]
  unsigned* tickets;   // [groups][total_tiles], zero between launches
};

constexpr int kTile = 32;   // output tile edge
constexpr int kBK = 64;     // reduction depth of one pipeline stage
constexpr int kGemmThreads = 256;
constexpr int kStages = 3;
constexpr int kTileElems = kTile * kTile;          // one partial output tile
constexpr int kStageElems = 2 * kTile * kBK;       // A tile + B tile of one stage
constexpr int kGemmSmemBytes = kStages * kStageElems * (int)sizeof(float);  // 48 KB (>= 8 partial tiles = 32 KB)

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// L2-coherent 128-bit load (bypasses the non-coherent L1): data written by other CTAs of this launch
__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
#ifdef B200_EMU
  return *reinterpret_cast<const float4*>(p);
#else
  return __ldcg(reinterpret_cast<const float4*>(p));
#endif
}

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

// One operand of one reduction segment, as seen by one thread: the two 16-byte strips it copies per
// chunk.  Shared-memory tile layouts (32 output indices x 64 reduction indices):
//   K-major (operand contiguous along the reduction r in global): row = c, 16 float4 slots along r,
//     slot index XOR (c >> 2) & 7 so the strided per-lane float4 reads hit 8 distinct bank groups;
//   C-major (operand contiguous along the output index c in global): row = r, 8 float4 slots along c.
// Everything that does not change from chunk to chunk (row validity, alignment, base pointers, the
// shared-memory offset) is computed once per segment; per chunk the fast path is one bounds test,
// one cp.async and one pointer bump per strip.
template <bool KMAJOR>
struct OperandCursor {
  const float* src[2];
  int dst[2];        // float offset inside the operand tile
  int pos;           // K-major: r offset of the strip inside the chunk (4*s); C-major: c0 + 4*s
  int row[2];        // C-major: r offset of the strip inside the chunk
  bool row_ok[2];    // K-major: output row inside the problem
  bool fast[2];      // 16-byte aligned source and (C-major) complete strip
  long long bump;    // pointer advance per chunk

  __device__ __forceinline__ void init(const float* base, int ld, int c0, int cmax, int tid, int r_start) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (KMAJOR) {
        const int c = (tid >> 4) + 16 * h, s = tid & 15;
        dst[h] = c * kBK + ((s ^ ((c >> 2) & 7)) << 2);
        pos = s << 2;
        row[h] = 0;
        row_ok[h] = (c0 + c) < cmax;
        src[h] = base + (long long)(c0 + c) * ld + pos + r_start;
        fast[h] = row_ok[h] && aligned16(src[h]);
        bump = kBK;
      } else {
        const int r = (tid >> 3) + 32 * h, s = tid & 7;
        dst[h] = r * kTile + (s << 2);
        pos = c0 + (s << 2);
        row[h] = r;
        row_ok[h] = true;
        src[h] = base + (long long)(r_start + r) * ld + pos;
        fast[h] = (pos + 3 < cmax) && aligned16(src[h]);
        bump = (long long)kBK * ld;
      }
    }
  }
  // copy this thread's strips of the chunk starting at reduction offset r0 (segment length rlen)
  __device__ __forceinline__ void stage(float* tile, int r0, int rlen, int cmax) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* d = tile + dst[h];
      if (KMAJOR) {
        const int p = r0 + pos;
        if (fast[h] && p + 3 < rlen) {
          cp_async16(d, src[h]);
        } else {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok[h] && p < rlen) {
            const float* q = src[h];
            v.x = q[0];
            if (p + 1 < rlen) v.y = q[1];
            if (p + 2 < rlen) v.z = q[2];
            if (p + 3 < rlen) v.w = q[3];
          }
          *reinterpret_cast<float4*>(d) = v;
        }
      } else {
        const bool in = (r0 + row[h]) < rlen;
        if (fast[h] && in) {
          cp_async16(d, src[h]);
        } else {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (in && pos < cmax) {
            const float* q = src[h];
            v.x = q[0];
            if (pos + 1 < cmax) v.y = q[1];
            if (pos + 2 < cmax) v.z = q[2];
            if (pos + 3 < cmax) v.w = q[3];
          }
          *reinterpret_cast<float4*>(d) = v;
        }
      }
      src[h] += bump;
    }
  }
};

// bias-gradient problems: B(r, j) = (j == 0)   (C-major tile)
__device__ __forceinline__ void stage_ones(float* tile, int c0, int r0, int rlen, int tid) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = (tid >> 3) + 32 * h, s = tid & 7;
    const float one = ((r0 + r) < rlen && s == 0 && c0 == 0) ? 1.f : 0.f;
    *reinterpret_cast<float4*>(tile + r * kTile + (s << 2)) = make_float4(one, 0.f, 0.f, 0.f);
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
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_kernel(const __grid_constant__ GemmLaunch L) {
  constexpr bool A_KM = (MODE != kWgrad);  // A contiguous along the reduction in global memory
  constexpr bool B_KM = (MODE == kFwd);
  // 3 pipeline stages x (A tile 32x64, B tile 32x64) = 48 KB; the first 32 KB later hold the 8 partial tiles
  B200_DYN_SMEM(float, smem);
  __shared__ unsigned s_ticket;

  const int tid = threadIdx.x;
  const int g = blockIdx.y;
  const int splitk = L.splitk;
  const int tile_id = (int)blockIdx.x / splitk;
  const int split = (int)blockIdx.x - tile_id * splitk;
  int pi = 0;
  while (pi + 1 < L.nprob && tile_id >= L.p[pi + 1].tile_begin) ++pi;
  const GemmProblem& sp = L.p[pi];
  const int t_local = tile_id - sp.tile_begin;
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

  int total_chunks = 0;
  for (int s = 0; s < sp.nseg; ++s) total_chunks += (sp.seg[s].len + kBK - 1) / kBK;
  // this CTA's slice of the flattened (segment, chunk) sequence
  const int per_split = (total_chunks + splitk - 1) / splitk;
  const int c_begin = min(split * per_split, total_chunks);
  const int nchunks = min(c_begin + per_split, total_chunks) - c_begin;

  // producer cursor
  OperandCursor<A_KM> ca;
  OperandCursor<B_KM> cb;
  int pseg = 0, pr0 = 0, plen = 0;
  bool ones = false;
  auto open_segment = [&](int r_start) {
    const GemmSeg& sg = sp.seg[pseg];
    plen = sg.len;
    ones = sg.b_ones != 0;
    ca.init(sg.A + (long long)g * sg.a_gs, sg.a_ld, i0, I, tid, r_start);
    if (!ones) cb.init(sg.B + (long long)g * sg.b_gs, sg.b_ld, j0, J, tid, r_start);
    pr0 = r_start;
  };
  auto issue = [&](int stage) {
    float* At = smem + stage * kStageElems;
    float* Bt = At + kTile * kBK;
    ca.stage(At, pr0, plen, I);
    if (ones) stage_ones(Bt, j0, pr0, plen, tid);
    else cb.stage(Bt, pr0, plen, J);
    pr0 += kBK;
    if (pr0 >= plen && ++pseg < sp.nseg) open_segment(0);
  };
  if (nchunks > 0) {
    int skip = c_begin;  // walk to the segment that holds chunk c_begin
    for (;;) {
      const int sc = (sp.seg[pseg].len + kBK - 1) / kBK;
      if (skip < sc) break;
      skip -= sc;
      ++pseg;
    }
    open_segment(skip * kBK);
  }

  // prologue: fill kStages-1 stages
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nchunks) issue(s);
    cp_async_commit();
  }

  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<kStages - 2>();  // chunk c has landed (this thread's copies)
    __syncthreads();               // ... and everybody's; also: the stage refilled below is no longer being read
    if (c + kStages - 1 < nchunks) issue((c + kStages - 1) % kStages);
    cp_async_commit();

    const float* As = smem + (c % kStages) * kStageElems;
    const float* Bs = As + kTile * kBK;
    // this warp's 8 reduction indices of the chunk: r = 8*warp .. 8*warp+7, in two halves of 4
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float av[4][4], bv[8][4];  // [output index][r within the half]
      if (A_KM) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float4 v = *reinterpret_cast<const float4*>(&As[(li * 4 + a) * kBK + (((2 * warp + h) ^ li) << 2)]);
          av[a][0] = v.x; av[a][1] = v.y; av[a][2] = v.z; av[a][3] = v.w;
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const float4 v = *reinterpret_cast<const float4*>(&As[(warp * 8 + h * 4 + rr) * kTile + li * 4]);
          av[0][rr] = v.x; av[1][rr] = v.y; av[2][rr] = v.z; av[3][rr] = v.w;
        }
      }
      if (B_KM) {
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int col = lj * 8 + b;
          const float4 v = *reinterpret_cast<const float4*>(&Bs[col * kBK + (((2 * warp + h) ^ ((col >> 2) & 7)) << 2)]);
          bv[b][0] = v.x; bv[b][1] = v.y; bv[b][2] = v.z; bv[b][3] = v.w;
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const float4 v0 = *reinterpret_cast<const float4*>(&Bs[(warp * 8 + h * 4 + rr) * kTile + lj * 8]);
          const float4 v1 = *reinterpret_cast<const float4*>(&Bs[(warp * 8 + h * 4 + rr) * kTile + lj * 8 + 4]);
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
  if (splitk > 1) {
    // publish this split's partial tile, take a ticket; the last arriver sums all splits in split order
    const long long tile_lin = (long long)g * L.total_tiles + tile_id;
    float* mine = L.scratch + (tile_lin * splitk + split) * kTileElems;
    *reinterpret_cast<float4*>(mine + tid * 4) = make_float4(out[0], out[1], out[2], out[3]);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&L.tickets[tile_lin], 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(splitk - 1)) return;
    __threadfence();
    const float* all = L.scratch + tile_lin * splitk * kTileElems + tid * 4;
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < splitk; ++k) {
      const float4 v = ld_cg_f4(all + (long long)k * kTileElems);
      s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
    }
    out[0] = s4.x; out[1] = s4.y; out[2] = s4.z; out[3] = s4.w;
    if (tid == 0) L.tickets[tile_lin] = 0u;  // ready for the next launch / graph replay
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
// A launch table: problems are appended on the host once, then replayed (the struct is passed by value
// as the kernel parameter; scratch / tickets for split-K live in device memory owned by the table).
struct GemmTable {
  GemmLaunch launch;      // problems run by the FFMA kernel
  GemmLaunch launch_tc;   // problems run by the tcgen05 kernel (stacked seeds; gemm_tc.cuh)
  int nprob = 0;
  int total_tiles = 0;    // FFMA tiles (32 x 32)
  int tc_tiles = 0;       // tcgen05 tiles (128 rows x N)
  int mode = kFwd;
  int groups = 1;
};

}  // namespace b200rl
