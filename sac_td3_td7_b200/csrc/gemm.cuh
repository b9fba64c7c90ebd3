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
// Tiling: CTA = 256 threads, 32x32 output tile, 64-deep reduction chunks in a 3-stage
// cp.async (LDGSTS) pipeline: every thread issues two 16-byte copies per operand per chunk
// straight from global/L2 into shared memory (no register staging, no transposition), two
// chunks ahead of the math.  The eight warps split each chunk's reduction (8 r's each) and
// every lane owns a 4x8 register tile (128 FFMA per 12 LDS.128; dgrad / wgrad pair
// neighbouring columns into FFMA2); the eight partial tiles are summed in a fixed order
// through shared memory, so results are bit-reproducible run to run.  Operands that are
// contiguous along the reduction in global memory keep that orientation in shared memory,
// XOR-swizzled at float4 granularity so the strided float4 reads are bank-conflict free;
// ragged / unaligned strips (obs and action widths such as 17, 6, 23) take an out-of-line
// zero-filling path.
//
// A single agent's layer is only 64 tiles of 4-12 serial chunks, so launches that would
// leave more than half of the 148 SMs idle run as thread-block CLUSTERS of 2 or 4 CTAs per
// tile: each CTA reduces a contiguous slice of the chunks, the non-leaders push their partial
// tile into the leader's shared memory (DSMEM) and the leader adds them in rank order and runs
// the epilogue -- deterministic, no global scratch, no atomics.  The epilogue's operands (bias,
// saved activations, Adam moments and scalars) are requested before the cross-warp reduction so
// that their L2 round trips overlap it.  Measured per launch on B200 (profiles/probe_phases.py):
// ~1.2 us between dependent launches, ~0.4 us setup, ~0.6 us per chunk, ~0.35 us reduction.
//
// Why FFMA and not tcgen05 here: a single agent's layer is 256x256x(<=768) -- 64 tiles --
// and parity is fp32 at 1e-5 (BASELINE.json north_star), which single-pass TF32/BF16
// tensor-core math cannot give.  The stacked-seed path is where tensor cores apply.
#pragma once
#include "common.cuh"

namespace b200rl {

enum GemmMode : int { kFwd = 0, kDgrad = 1, kWgrad = 2 };
enum Act : int { kActNone = 0, kActRelu = 1, kActElu = 2, kActTanh = 3, kActTanhNoise = 4,
                 // forward, tensor-core path only: C = x (bias added, no activation) AND aux = AvgL1Norm(x) = x / max(mean|x|, 1e-8)
                 // over the J columns of each row (rl/nn/sale.py:11-13) -- the normalisation fused into the producing layer
                 kActNorm = 5 };

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

// One launch of the tcgen05 kernel (gemm_tc.cuh): the problem table travels in kernel-parameter (constant)
// memory, so a CTA finds its problem without a dependent global load.
#ifndef B200RL_MAX_PROBLEMS
#define B200RL_MAX_PROBLEMS 24
#endif
constexpr int kMaxProblems = B200RL_MAX_PROBLEMS;
struct GemmLaunch {
  GemmProblem p[kMaxProblems];
  int nprob;
  int total_tiles;
  // forward problems whose operands TMA can address (16-byte aligned bases and row strides): bit pi of tma_mask, and
  // tmaps = device array of CUtensorMap [problem][segment][A, B] (gemm_tc.cuh)
  unsigned tma_mask;
  int pad;
  const void* tmaps;
};

// Device-side form of a problem for the FFMA kernel: GemmProblem plus everything the host can work out
// once (chunk counts, the split of the reduction, a reciprocal for the tile -> (row, col) division), so the
// per-CTA setup is a handful of constant loads.
struct FfmaProblem {
  GemmSeg seg[3];
  int nseg, I, J;
  int total_chunks;        // sum over segments of ceil(len / kBK)
  int per_split;           // chunks handled by each of the `splitk` CTAs of a tile
  int tile_begin, tiles_j;
  unsigned tiles_j_magic;  // floor(2^32 / tiles_j) + 1 for tiles_j > 1 (exact quotient for the tile counts used here)
  float* C;
  long long c_gs;
  int c_ld, act;
  const float* bias;
  long long bias_gs;
  const float* aux;
  long long aux_gs;
  int aux_ld, accumulate;
  float* P;
  float* M;
  float* V;
  long long p_gs;
  int p_ld;
  float noise_sigma, noise_clip;
  int pad;
  const AdamScalars* adam;
};
// With splitk > 1 the kernel is launched in thread-block clusters of `splitk` CTAs: the CTAs of a cluster
// compute the same output tile over disjoint, contiguous slices of the reduction chunks, the non-leaders
// push their partial tiles into the leader's shared memory (DSMEM) and the leader sums them in rank order
// -- deterministic, no global scratch, no atomics -- and runs the epilogue.  This is what lets a single
// agent's 64-tile layers occupy the whole machine and shortens the serial chunk chain per CTA.
constexpr int kMaxSplitK = 4;
struct FfmaLaunch {
  FfmaProblem p[kMaxProblems];
  int nprob;
  int splitk;
  int tag;  // table number in creation order (profiling builds: timeline slot)
  int pad;
};

constexpr int kTile = 32;   // output tile edge
constexpr int kBK = 64;     // reduction depth of one pipeline stage
constexpr int kGemmThreads = 256;
constexpr int kStages = 3;
constexpr int kTileElems = kTile * kTile;          // one partial output tile
constexpr int kStageElems = 2 * kTile * kBK;       // A tile + B tile of one stage
constexpr int kPipeBytes = kStages * kStageElems * (int)sizeof(float);  // 48 KB (>= 8 partial tiles = 32 KB)
// + one 4 KB receive slot per non-leader CTA of a split-K cluster
constexpr int gemm_smem_bytes(int splitk) { return kPipeBytes + (splitk - 1) * kTileElems * (int)sizeof(float); }

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

// Shared-memory swizzles.  Both tile layouts are XOR-swizzled at float4 (16-byte) granularity so that the cp.async
// writes stay 16-byte contiguous and the reads of the math loop are bank-conflict free:
//   FFMA loop: K-major slot ^= (row >> 2) & 7 (strided per-lane float4 reads); C-major rows are read as they are.
//   MMA loop (mma.sync fragments, scalar reads of lane (g = lane >> 2, t = lane & 3)): K-major slot ^= row & 7 (eight
//   consecutive rows x four t's hit 32 different banks); C-major slot ^= (r & 3) << 1 (four consecutive reduction
//   rows x eight consecutive output indices hit 32 different banks).
template <bool MMA>
__device__ __forceinline__ int km_swz(int row) { return MMA ? (row & 7) : ((row >> 2) & 7); }
template <bool MMA>
__device__ __forceinline__ int cm_swz(int r) { return MMA ? ((r & 3) << 1) : 0; }
// K-major: float offset of the strip 16 rows below `dst` (same 16-byte slot of the unswizzled row)
template <bool MMA>
__device__ __forceinline__ int km_second_strip(int dst) { return MMA ? dst + 16 * kBK : (dst ^ 16) + 16 * kBK; }

// copy one thread's two strips of the chunk starting at reduction offset r0 (segment length rlen), zero-filling
// whatever lies outside the operand
template <bool KMAJOR, bool MMA>
B200_NOINLINE __device__ void stage_ragged(float* tile, const float* src, int delta, int dst, int pos, int row, bool ok0, bool ok1, bool vec, int r0, int rlen,
                                           int cmax) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float* q = src + (h ? delta : 0);
    if (KMAJOR) {
      float* d = tile + (h ? (km_second_strip<MMA>(dst)) : dst);
      const int p = r0 + pos;
      const bool ok = h ? ok1 : ok0;
      if (vec && ok && p + 3 < rlen) {
        cp_async16(d, q);
      } else {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && p < rlen) {
          v.x = q[0];
          if (p + 1 < rlen) v.y = q[1];
          if (p + 2 < rlen) v.z = q[2];
          if (p + 3 < rlen) v.w = q[3];
        }
        *reinterpret_cast<float4*>(d) = v;
      }
    } else {
      float* d = tile + dst + h * (32 * kTile);
      const bool in = (r0 + row + 32 * h) < rlen;
      if (vec && in) {
        cp_async16(d, q);
      } else {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in && ok0) {
          v.x = q[0];
          if (pos + 1 < cmax) v.y = q[1];
          if (pos + 2 < cmax) v.z = q[2];
          if (pos + 3 < cmax) v.w = q[3];
        }
        *reinterpret_cast<float4*>(d) = v;
      }
    }
  }
}

// One operand of one reduction segment, as seen by one thread: the two 16-byte strips it copies per
// chunk.  Shared-memory tile layouts (32 output indices x 64 reduction indices):
//   K-major (operand contiguous along the reduction r in global): row = c, 16 float4 slots along r,
//     slot index XOR (c >> 2) & 7 so the strided per-lane float4 reads hit 8 distinct bank groups;
//   C-major (operand contiguous along the output index c in global): row = r, 8 float4 slots along c.
// Everything that does not change from chunk to chunk (row validity, alignment, base pointer, the
// shared-memory offset) is computed once per segment; per chunk the fast path is one bounds test,
// one cp.async per strip and one pointer bump.  The second strip sits 16 rows (K-major) / 32 rows
// (C-major) below the first, so one pointer, one stride and one offset describe both.
template <bool KMAJOR, bool MMA>
struct OperandCursor {
  const float* src;  // strip 0 of the next chunk; strip 1 = src + delta
  int delta;         // K-major: 16 * ld; C-major: 32 * ld
  int dst;           // float offset of strip 0 inside the operand tile
  int pos;           // K-major: r offset of the strip inside the chunk (4*s); C-major: c0 + 4*s
  int row;           // C-major: r offset of strip 0 inside the chunk (strip 1: + 32)
  bool ok0, ok1;     // K-major: output row of strip 0 / 1 inside the problem; C-major: strip has >= 1 valid column
  bool vec;          // 16-byte aligned source (and, C-major, a complete strip): cp.async path
  bool all;          // both strips of a chunk that lies fully inside the segment take the cp.async path

  __device__ __forceinline__ void init(const float* base, int ld, int c0, int cmax, int tid, int r_start) {
    if (KMAJOR) {
      const int c = tid >> 4, s = tid & 15;
      dst = c * kBK + ((s ^ km_swz<MMA>(c)) << 2);
      pos = s << 2;
      row = 0;
      ok0 = (c0 + c) < cmax;
      ok1 = (c0 + c + 16) < cmax;
      src = base + (long long)(c0 + c) * ld + (pos + r_start);
      delta = 16 * ld;
      vec = aligned16(src);  // 16*ld floats and the 64-float bump keep the alignment
      all = vec && ok0 && ok1;
    } else {
      const int r = tid >> 3, s = tid & 7;
      dst = r * kTile + ((s ^ cm_swz<MMA>(r)) << 2);
      pos = c0 + (s << 2);
      row = r;
      ok0 = ok1 = pos < cmax;
      src = base + (long long)(r_start + r) * ld + pos;
      delta = 32 * ld;
      vec = (pos + 3 < cmax) && aligned16(src);
      all = vec;
    }
  }
  // fast path: the whole chunk lies inside the segment and both strips are complete and aligned
  __device__ __forceinline__ void stage_full(float* tile) {
    if (KMAJOR) {
      cp_async16(tile + dst, src);
      cp_async16(tile + km_second_strip<MMA>(dst), src + delta);
      src += kBK;
    } else {
      cp_async16(tile + dst, src);
      cp_async16(tile + dst + 32 * kTile, src + delta);
      src += 2ll * delta;
    }
  }
  // general path (ragged widths such as 17 / 6 / 23, segment tails, unaligned rows): out of line, it is rare
  __device__ __forceinline__ void stage(float* tile, int r0, int rlen, int cmax) {
    stage_ragged<KMAJOR, MMA>(tile, src, delta, dst, pos, row, ok0, ok1, vec, r0, rlen, cmax);
    src += KMAJOR ? (long long)kBK : 2ll * delta;
  }
};

// bias-gradient problems: B(r, j) = (j == 0)   (C-major tile)
template <bool MMA>
__device__ __forceinline__ void stage_ones(float* tile, int c0, int r0, int rlen, int tid) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = (tid >> 3) + 32 * h, s = tid & 7;
    const float one = ((r0 + r) < rlen && s == 0 && c0 == 0) ? 1.f : 0.f;
    *reinterpret_cast<float4*>(tile + r * kTile + ((s ^ cm_swz<MMA>(r)) << 2)) = make_float4(one, 0.f, 0.f, 0.f);
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

// Optional in-kernel timeline (build with -DB200RL_PROBE): CTA (0,0) thread 0 records the global
// timer at entry / exit and the SM clock at the phase boundaries of every launch into a ring.
#ifdef B200RL_PROBE
__device__ long long g_probe[64][10];
__device__ unsigned g_probe_n;
#define B200_PROBE_DECL long long pb_[10]; const bool pb_on_ = (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0); if (pb_on_) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pb_[0])); pb_[1] = clock64(); }
#define B200_PROBE(i) if (pb_on_) pb_[i] = clock64();
#define B200_PROBE_END if (pb_on_) { pb_[8] = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pb_[9])); unsigned sl_ = atomicAdd(&g_probe_n, 1u) & 63u; for (int q_ = 0; q_ < 10; ++q_) g_probe[sl_][q_] = pb_[q_]; }
// per-table timeline: earliest CTA entry / latest CTA exit of the most recent launch of every table (tag)
__device__ unsigned long long g_tl[256][2];
struct TimelineGuard {
  int tag;
  __device__ explicit TimelineGuard(int t) : tag(t & 255) {
    if (threadIdx.x == 0) { unsigned long long n; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n)); atomicMin(&g_tl[tag][0], n); }
  }
  __device__ ~TimelineGuard() {
    if (threadIdx.x == 0) { unsigned long long n; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n)); atomicMax(&g_tl[tag][1], n); }
  }
};
#define B200_TIMELINE(tag) TimelineGuard tl_guard_(tag);
#else
#define B200_TIMELINE(tag)
#define B200_PROBE_DECL
#define B200_PROBE(i)
#define B200_PROBE_END
#endif

// ---- thread-block cluster helpers (split-K through distributed shared memory) ----
#ifndef B200_EMU
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_id_x() {
  unsigned r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// store 16 bytes into the shared memory of CTA `rank` of this cluster, at the address `local` has in ours
__device__ __forceinline__ void st_cluster_f4(float* local, unsigned rank, float a, float b, float c, float d) {
  const unsigned la = (unsigned)__cvta_generic_to_shared(local);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
#endif

// Two fp32 FMAs in one instruction (Blackwell FFMA2, PTX fma.rn.f32x2): d.x = a.x*b.x + d.x, d.y = a.y*b.y + d.y,
// each rounded to nearest exactly like fmaf.  Half the issue slots and register-file reads of two FFMAs, which is
// what the small-tile main loop is short of.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
#ifdef B200_EMU
  d.x = fmaf(a.x, b.x, d.x);
  d.y = fmaf(a.y, b.y, d.y);
#else
  unsigned long long dd, aa, bb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d.x), "f"(d.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(dd));
#endif
}

// ---- error-compensated TF32 on the warp-level tensor-core path (3xTF32) -------------------------------------------
// x = hi + lo with hi = tf32(x) and lo = tf32(x - hi) (both round-to-nearest): hi carries 11 significant bits, lo the next
// 11, so a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi is short of the exact fp32 product by ~2^-22 relative (the dropped
// a_lo*b_lo term) -- below the rounding of an fp32 FFMA chain.  The products are accumulated in fp32 by
// mma.sync.m16n8k8 (small terms first).  Not available to the host emulation, which keeps the FFMA loop.
#ifndef B200_EMU
__device__ __forceinline__ void tf32_split(float x, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
// D(16x8) += A(16x8, row) * B(8x8, col); lane (g = lane >> 2, t = lane & 3) holds a0 = A[g][t], a1 = A[g+8][t],
// a2 = A[g][t+4], a3 = A[g+8][t+4]; b0 = B[t][g], b1 = B[t+4][g]; d0 = D[g][2t], d1 = D[g][2t+1], d2 = D[g+8][2t], d3 = D[g+8][2t+1]
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#endif

// four consecutive floats of an epilogue operand, all loads in flight at once
template <bool CG = false>
__device__ __forceinline__ void load_quad(float (&o)[4], const float* p, int nvalid) {
#ifndef B200_EMU
  if (CG) {
    if (nvalid == 4 && aligned16(p)) {
      const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
      o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = (e < nvalid) ? __ldcg(p + e) : 0.f;
    }
    return;
  }
#endif
  if (nvalid == 4 && aligned16(p)) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = (e < nvalid) ? p[e] : 0.f;
  }
}
template <bool CG>
__device__ __forceinline__ AdamScalars load_adam(const AdamScalars* p) {
#ifndef B200_EMU
  if (CG) {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldcg(reinterpret_cast<const float4*>(p) + 1);
    AdamScalars s;
    s.w1 = a.x; s.beta2 = a.y; s.w2 = a.z; s.bc2_sqrt = a.w; s.neg_step = b.x; s.eps = b.y; s.step = __float_as_int(b.z); s.pad = 0;
    return s;
  }
#endif
  return *p;
}
__device__ __forceinline__ void store_quad(float* p, const float (&o)[4], int nvalid) {
  if (nvalid == 4 && aligned16(p)) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < nvalid) p[e] = o[e];
  }
}

// Register budget: two CTAs per SM (<= 128 registers).  A variant without the cap was measured (fewer address
// recomputations in the loop) and lost: a single agent's update keeps several small launches of different streams in
// flight, and they only overlap if two CTAs fit on one SM.
// MMA = true: the chunk products run on the tensor cores as 3xTF32 (mma.sync fragments read straight from the same
// cp.async-staged tiles, with the fragment-friendly swizzles); everything around the product loop -- staging, split-K
// clusters, the fixed-order cross-warp reduction, the fused epilogues -- is shared with the FFMA variant.
// One output tile of one problem.  `sp` may live in kernel-parameter memory (classic launches) or in shared memory (the
// persistent dataflow kernel, mega.cuh).  PERSIST = true: no split-K cluster, no programmatic-launch wait, and every
// operand the epilogue reads goes through L2 (ld.global.cg) -- it may have been written by another SM inside the
// same kernel, where L1 is not invalidated between tiles.
template <int MODE, bool MMA, bool PERSIST>
__device__ __forceinline__ void gemm_tile(const FfmaProblem& sp, int t_local, int g, int split, int splitk_arg, float* smem) {
  B200_PROBE_DECL
  constexpr bool A_KM = (MODE != kWgrad);  // A contiguous along the reduction in global memory
  constexpr bool B_KM = (MODE == kFwd);
  const int tid = threadIdx.x;
#ifndef B200_EMU
  const int splitk = PERSIST ? 1 : splitk_arg;
#else
  constexpr int splitk = 1;
  (void)splitk_arg;
#endif
  const int ti = sp.tiles_j > 1 ? (int)__umulhi((unsigned)t_local, sp.tiles_j_magic) : t_local;
  const int i0 = ti * kTile;
  const int j0 = (t_local - ti * sp.tiles_j) * kTile;
  const int I = sp.I, J = sp.J;

  const int warp = tid >> 5, lane = tid & 31;
  // (row group li, column group lj) of this lane: li varies fastest, so the 8 lanes of a quarter-warp read 8 different A
  // fragments and ONE B fragment (measured against li = lane >> 2: forward equal, dgrad -3 %, wgrad -5 %)
  const int li = lane & 7, lj = lane >> 3;

  // dgrad / wgrad read B contiguous along the output column, so neighbouring outputs (a, 2bp) and (a, 2bp+1) sit in
  // adjacent registers and one FFMA2 updates both (same r order, same rounding as two FFMAs).  Forward reads both
  // operands contiguous along r; pairing there would need split even/odd accumulators (64 more registers): plain FFMA.
  constexpr bool USE_F2 = (MODE != kFwd);
  float2 acc2[4][4];
  float acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = 0; b < 4; ++b) acc2[a][b] = make_float2(0.f, 0.f);
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  }

  // this CTA's slice of the flattened (segment, chunk) sequence
  const int c_begin = split * sp.per_split;
  const int nchunks = max(0, min(sp.per_split, sp.total_chunks - c_begin));

  // producer cursor
  OperandCursor<A_KM, MMA> ca;
  OperandCursor<B_KM, MMA> cb;
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
    const bool full = pr0 + kBK <= plen;
    if (full && ca.all) ca.stage_full(At);
    else ca.stage(At, pr0, plen, I);
    if (ones) stage_ones<MMA>(Bt, j0, pr0, plen, tid);
    else if (full && cb.all) cb.stage_full(Bt);
    else cb.stage(Bt, pr0, plen, J);
    pr0 += kBK;
    if (pr0 >= plen && ++pseg < sp.nseg) open_segment(0);
  };
  if (nchunks > 0) {
    int skip = c_begin;  // walk to the segment that holds chunk c_begin (split 0: no iterations)
    while (skip > 0) {
      const int sc = (sp.seg[pseg].len + kBK - 1) / kBK;
      if (skip < sc) break;
      skip -= sc;
      ++pseg;
    }
    open_segment(skip * kBK);
  }
  B200_PROBE(2)
  if (!PERSIST) B200_PDL_WAIT();  // operands written by earlier kernels are read from here on

  // prologue: fill kStages-1 stages
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nchunks) issue(s);
    cp_async_commit();
  }
  B200_PROBE(3)

  // loop-invariant shared-memory read offsets of this thread (floats, relative to a stage's A / B tile)
  int a_off[4], b_off[8];
  if (A_KM) {
#pragma unroll
    for (int a = 0; a < 4; ++a) a_off[a] = (li * 4 + a) * kBK + (((2 * warp) ^ li) << 2);  // h = 1: ^ 4
  } else {
#pragma unroll
    for (int a = 0; a < 4; ++a) a_off[a] = (warp * 8 + a) * kTile + li * 4;  // rr = a; h = 1: + 4 * kTile
  }
  if (B_KM) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int col = lj * 8 + b;
      b_off[b] = col * kBK + (((2 * warp) ^ ((col >> 2) & 7)) << 2);  // h = 1: ^ 4
    }
  } else {
#pragma unroll
    for (int b = 0; b < 8; ++b) b_off[b] = (warp * 8 + (b >> 1)) * kTile + lj * 8 + 4 * (b & 1);  // rr = b >> 1, half-row b & 1
  }

  // MMA fragments: lane (fg = lane >> 2, ft = lane & 3); the warp's k8 step is r = 8*warp + {ft, ft + 4}
  int ma_off[2][4], mb_off[4][2];
  float macc[2][4][4];
  {
    const int fg = lane >> 2, ft = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = 16 * mi + fg + 8 * (q & 1), r = 8 * warp + ft + 4 * (q >> 1);
        ma_off[mi][q] = A_KM ? i * kBK + ((((r >> 2) ^ km_swz<true>(i)) << 2) | (r & 3)) : r * kTile + ((((i >> 2) ^ cm_swz<true>(r)) << 2) | (i & 3));
#pragma unroll
        for (int e = 0; e < 4; ++e) macc[mi][q][e] = 0.f;  // (q doubles as the n index here: 2 x 4 x 4 accumulators)
      }
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int n = 8 * ni + fg, r = 8 * warp + ft + 4 * q;
        mb_off[ni][q] = B_KM ? n * kBK + ((((r >> 2) ^ km_swz<true>(n)) << 2) | (r & 3)) : r * kTile + ((((n >> 2) ^ cm_swz<true>(r)) << 2) | (n & 3));
      }
  }

  int rs = 0, ws = kStages - 1;  // stage read by chunk c / stage refilled during chunk c
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<kStages - 2>();  // chunk c has landed (this thread's copies)
    __syncthreads();               // ... and everybody's; also: the stage refilled below is no longer being read
    if (c == 0) { B200_PROBE(4) }
    if (c + kStages - 1 < nchunks) issue(ws);
    cp_async_commit();

    const float* As = smem + rs * kStageElems;
    const float* Bs = As + kTile * kBK;
    rs = (rs + 1 == kStages) ? 0 : rs + 1;
    ws = (ws + 1 == kStages) ? 0 : ws + 1;
    // this warp's 8 reduction indices of the chunk: r = 8*warp .. 8*warp+7
#ifndef B200_EMU
    if (MMA) {
      // one k8 step of the whole 32 x 32 tile: 2 (m16) x 4 (n8) fragments, each product as three TF32 MMAs
      unsigned ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int q = 0; q < 4; ++q) tf32_split(As[ma_off[mi][q]], ah[mi][q], al[mi][q]);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int q = 0; q < 2; ++q) tf32_split(Bs[mb_off[ni][q]], bh[ni][q], bl[ni][q]);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          mma_tf32(macc[mi][ni], al[mi], bh[ni]);
          mma_tf32(macc[mi][ni], ah[mi], bl[ni]);
          mma_tf32(macc[mi][ni], ah[mi], bh[ni]);
        }
    } else
#endif
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float av[4][4], bv[8][4];  // [output index][r within the half]
      if (A_KM) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float4 v = *reinterpret_cast<const float4*>(&As[a_off[a] ^ (h << 2)]);
          av[a][0] = v.x; av[a][1] = v.y; av[a][2] = v.z; av[a][3] = v.w;
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const float4 v = *reinterpret_cast<const float4*>(&As[a_off[rr] + h * 4 * kTile]);
          av[0][rr] = v.x; av[1][rr] = v.y; av[2][rr] = v.z; av[3][rr] = v.w;
        }
      }
      if (B_KM) {
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const float4 v = *reinterpret_cast<const float4*>(&Bs[b_off[b] ^ (h << 2)]);
          bv[b][0] = v.x; bv[b][1] = v.y; bv[b][2] = v.z; bv[b][3] = v.w;
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const float4 v0 = *reinterpret_cast<const float4*>(&Bs[b_off[2 * rr] + h * 4 * kTile]);
          const float4 v1 = *reinterpret_cast<const float4*>(&Bs[b_off[2 * rr + 1] + h * 4 * kTile]);
          bv[0][rr] = v0.x; bv[1][rr] = v0.y; bv[2][rr] = v0.z; bv[3][rr] = v0.w;
          bv[4][rr] = v1.x; bv[5][rr] = v1.y; bv[6][rr] = v1.z; bv[7][rr] = v1.w;
        }
      }
      if (!USE_F2) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a][rr], bv[b][rr], acc[a][b]);
      } else {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const float2 aa = make_float2(av[a][rr], av[a][rr]);
#pragma unroll
            for (int bp = 0; bp < 4; ++bp) ffma2(acc2[a][bp], aa, make_float2(bv[2 * bp][rr], bv[2 * bp + 1][rr]));
          }
      }
    }
  }

  // ---- epilogue operands: every load is issued here, before the reduction, so that their L2 round trips
  // overlap the barriers below instead of being paid one after the other at the end ----
  const int oi = tid >> 3, oslot = tid & 7;
  const int gi = i0 + oi;
  const int gj0 = j0 + oslot * 4;
  const bool live = (split == 0) && gi < I && gj0 < J;  // this thread writes outputs
  const int nvalid = live ? min(4, J - gj0) : 0;
  float e0[4] = {0.f, 0.f, 0.f, 0.f}, e1[4] = {0.f, 0.f, 0.f, 0.f}, e2[4] = {0.f, 0.f, 0.f, 0.f};
  const int act = sp.act;
  bool has0 = false, has1 = false;
  AdamScalars adam_s = {};
  if (MODE == kFwd) {
    has0 = sp.bias != nullptr;
    has1 = act == kActTanhNoise;
    if (live && has0) load_quad<PERSIST>(e0, sp.bias + (long long)g * sp.bias_gs + gj0, nvalid);
    if (live && has1) load_quad<PERSIST>(e1, sp.aux + (long long)g * sp.aux_gs + (long long)gi * sp.aux_ld + gj0, nvalid);
  } else if (MODE == kDgrad) {
    has0 = sp.aux != nullptr;
    has1 = sp.accumulate != 0;
    if (live && has0) load_quad<PERSIST>(e0, sp.aux + (long long)g * sp.aux_gs + (long long)gi * sp.aux_ld + gj0, nvalid);
    if (live && has1) load_quad<PERSIST>(e1, sp.C + (long long)g * sp.c_gs + (long long)gi * sp.c_ld + gj0, nvalid);
  } else {
    has0 = sp.P != nullptr;
    if (live && has0) {
      const long long off = (long long)g * sp.p_gs + (long long)gi * sp.p_ld + gj0;
      load_quad<PERSIST>(e0, sp.P + off, nvalid);
      load_quad<PERSIST>(e1, sp.M + off, nvalid);
      load_quad<PERSIST>(e2, sp.V + off, nvalid);
      adam_s = load_adam<PERSIST>(sp.adam);
    }
  }

  cp_async_wait<0>();
  __syncthreads();  // everyone is done reading the pipeline stages: reuse the memory for the partial tiles
  B200_PROBE(5)

  // ---- cross-warp reduction of the eight K-slices, fixed order (bit-reproducible) ----
  // partial tile layout: [warp][row][8 slots of float4], slot index XOR (row >> 2) to spread banks
  float* part = smem;
  if (USE_F2) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = (b & 1) ? acc2[a][b >> 1].y : acc2[a][b >> 1].x;
  }
  if (MMA) {
    const int fg = lane >> 2, ft = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int row = 16 * mi + fg + 8 * hh, col = 8 * ni + 2 * ft;
          float* base = part + warp * kTileElems + row * kTile;
          *reinterpret_cast<float2*>(base + ((((col >> 2) ^ km_swz<true>(row)) << 2) | (col & 3))) = make_float2(macc[mi][ni][2 * hh], macc[mi][ni][2 * hh + 1]);
        }
  } else {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int row = li * 4 + a;
      float* base = part + warp * kTileElems + row * kTile;
      *reinterpret_cast<float4*>(base + (((2 * lj) ^ li) << 2)) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      *reinterpret_cast<float4*>(base + (((2 * lj + 1) ^ li) << 2)) = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
    }
  }
  __syncthreads();
  float out[4];
  {
    const float* p0 = part + oi * kTile + ((oslot ^ km_swz<MMA>(oi)) << 2);
    float4 s = *reinterpret_cast<const float4*>(p0);
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(p0 + w * kTileElems);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[0] = s.x; out[1] = s.y; out[2] = s.z; out[3] = s.w;
  }
#ifndef B200_EMU
  if (splitk > 1) {
    // non-leaders push their partial tile into the leader's receive slot; the leader adds them in rank order
    float* recv = smem + kStages * kStageElems;
    cluster_wait();  // phase 1 complete: every CTA of the cluster is running
    if (split != 0) st_cluster_f4(recv + (split - 1) * kTileElems + tid * 4, 0u, out[0], out[1], out[2], out[3]);
    cluster_arrive();  // phase 2 (release): my push is visible to whoever waits
    if (split != 0) return;
    cluster_wait();
    for (int k = 1; k < splitk; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(recv + (k - 1) * kTileElems + tid * 4);
      out[0] += v.x; out[1] += v.y; out[2] += v.z; out[3] += v.w;
    }
  }
#endif
  B200_PROBE(6)
  if (!live) return;

  if (MODE == kFwd) {
    if (has0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) out[e] += e0[e];
    }
    switch (act) {
      case kActRelu:
#pragma unroll
        for (int e = 0; e < 4; ++e) out[e] = out[e] > 0.f ? out[e] : 0.f;
        break;
      case kActElu:
#pragma unroll
        for (int e = 0; e < 4; ++e) out[e] = out[e] > 0.f ? out[e] : expm1f(out[e]);
        break;
      case kActTanh:
#pragma unroll
        for (int e = 0; e < 4; ++e) out[e] = tanhf(out[e]);
        break;
      case kActTanhNoise:
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float n = fminf(fmaxf(e1[e] * sp.noise_sigma, -sp.noise_clip), sp.noise_clip);
          out[e] = fminf(fmaxf(tanhf(out[e]) + n, -1.f), 1.f);
        }
        break;
      default: break;
    }
  } else if (MODE == kDgrad) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = out[e];
      if (has1) v += e1[e];
      if (has0) v *= act_backward(e0[e], act);
      out[e] = v;
    }
  } else {  // kWgrad
    if (has0) {
      const long long off = (long long)g * sp.p_gs + (long long)gi * sp.p_ld + gj0;
#pragma unroll
      for (int e = 0; e < 4; ++e) adam_update(out[e], e0[e], e1[e], e2[e], adam_s);
      store_quad(sp.P + off, e0, nvalid);
      store_quad(sp.M + off, e1, nvalid);
      store_quad(sp.V + off, e2, nvalid);
    }
    if (!sp.C) return;
  }
  B200_PROBE(7)
  store_quad(sp.C + (long long)g * sp.c_gs + (long long)gi * sp.c_ld + gj0, out, nvalid);
  B200_PROBE_END
}

template <int MODE, bool MMA>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_kernel(const __grid_constant__ FfmaLaunch L) {
  B200_PDL_TRIGGER();  // the setup below reads kernel parameters only: it may overlap the previous kernel
  B200_TIMELINE(L.tag)
  // 3 pipeline stages x (A tile 32x64, B tile 32x64) = 48 KB; the first 32 KB later hold the 8 partial tiles;
  // behind them one 4 KB receive slot per non-leader CTA of the split-K cluster
  B200_DYN_SMEM(float, smem);
  int tile_id = (int)blockIdx.x, split = 0;
#ifndef B200_EMU
  const int splitk = L.splitk;
  if (splitk > 1) {
    cluster_arrive();  // phase 1: "my shared memory exists" (waited on just before the DSMEM pushes)
    tile_id = (int)cluster_id_x();
    split = (int)cluster_ctarank();
  }
#else
  constexpr int splitk = 1;
#endif
  int pi = 0;
  while (pi + 1 < L.nprob && tile_id >= L.p[pi + 1].tile_begin) ++pi;
  gemm_tile<MODE, MMA, false>(L.p[pi], tile_id - L.p[pi].tile_begin, (int)blockIdx.y, split, splitk, smem);
}

// ---- bias gradients of stacked seeds --------------------------------------------------------------
// db[i] = sum_r dZ[r][i] (+ Adam).  As a "B := ones column" problem of the tile kernel a bias gradient costs eight full
// 32x32 tiles of FFMA for one useful column; for one agent that rides along in a launch that exists anyway, but with
// hundreds of stacked seeds it is most of the narrow-problem launch.  There (the tensor-core path) bias gradients are
// column sums: one CTA per (problem, seed), thread i sums column i over the batch with coalesced row reads.
struct BiasProblem {
  const float* dZ; long long dz_gs; int dz_ld; int rows, out;
  float* C; float* P; float* M; float* V; long long p_gs;
  const AdamScalars* adam;
};
struct BiasLaunch {
  BiasProblem p[kMaxProblems];
  int nprob;
  int pad;
};
constexpr int kBiasMaxVecOut = 1024;
__global__ void __launch_bounds__(256) bias_grad_kernel(const __grid_constant__ BiasLaunch L) {
  B200_PDL_PROLOGUE();
  const BiasProblem& bp = L.p[blockIdx.x];
  const int g = blockIdx.y;
  const float* dz = bp.dZ + (long long)g * bp.dz_gs;
  // Column sums are four chains over r mod 4, combined as (s0 + s1) + (s2 + s3).  128-bit path: thread (chain, column quad)
  // walks its rows with eight 16-byte loads in flight and the chains meet in shared memory -- a launch with one problem has
  // only one CTA per seed, so the bytes in flight per thread are what reaches the HBM rate; same sums, same order as the
  // scalar path below.
  __shared__ float4 part[4][kBiasMaxVecOut / 4];
  const bool vec = (bp.out & 3) == 0 && bp.out <= kBiasMaxVecOut && (bp.dz_ld & 3) == 0 && (reinterpret_cast<unsigned long long>(dz) & 15ull) == 0ull;
  if (vec) {
    const int nq = bp.out >> 2;
    const int rows4 = bp.rows & ~3;
    for (int item = threadIdx.x; item < nq * 4; item += blockDim.x) {
      const int cq = item % nq, chain = item / nq;
      const float* col = dz + 4 * cq;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      int r = chain;
      for (; r + 28 < rows4; r += 32) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = *reinterpret_cast<const float4*>(col + (long long)(r + 4 * k) * bp.dz_ld);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
      }
      for (; r < rows4; r += 4) {
        const float4 v = *reinterpret_cast<const float4*>(col + (long long)r * bp.dz_ld);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      if (chain == 0)
        for (r = rows4; r < bp.rows; ++r) {
          const float4 v = *reinterpret_cast<const float4*>(col + (long long)r * bp.dz_ld);
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
      part[chain][cq] = s;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < bp.out; i += blockDim.x) {
    float gsum;
    if (vec) {
      const float* pf = reinterpret_cast<const float*>(&part[0][0]);
      gsum = (pf[i] + pf[kBiasMaxVecOut + i]) + (pf[2 * kBiasMaxVecOut + i] + pf[3 * kBiasMaxVecOut + i]);
    } else {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int r = 0;
      for (; r + 3 < bp.rows; r += 4) {
        s0 += dz[(long long)(r + 0) * bp.dz_ld + i];
        s1 += dz[(long long)(r + 1) * bp.dz_ld + i];
        s2 += dz[(long long)(r + 2) * bp.dz_ld + i];
        s3 += dz[(long long)(r + 3) * bp.dz_ld + i];
      }
      for (; r < bp.rows; ++r) s0 += dz[(long long)r * bp.dz_ld + i];
      gsum = (s0 + s1) + (s2 + s3);
    }
    const long long o = (long long)g * bp.p_gs + i;
    if (bp.P) {
      float p = bp.P[o], m = bp.M[o], v = bp.V[o];
      adam_update(gsum, p, m, v, *bp.adam);
      bp.P[o] = p; bp.M[o] = m; bp.V[o] = v;
    }
    if (bp.C) bp.C[o] = gsum;
  }
}


}  // namespace b200rl
