// gemm_tc.cuh -- tcgen05 (5th-gen tensor core) grouped GEMM for STACKED seeds.
//
// When G independent agents are stacked (gridDim.y = seed), every layer is a real dense batch of
// 256x256xK contractions and the FFMA path of gemm.cuh becomes compute bound (640 GFLOP per stacked
// TD7 update).  This kernel runs the same problem table on the tensor cores:
//   * one CTA per (problem, 128-row half, seed): D[128 x N<=256] accumulates in TMEM (256 columns),
//   * operands stay fp32 in HBM/L2 and are consumed as TF32 (`kind::tf32`, 10-bit mantissa, fp32
//     accumulate) -- the north star's "stated TF32 tolerance" applies to the stacked path only; the
//     single-agent path stays fp32 FFMA,
//   * 32-deep (128-byte) reduction chunks land straight in the canonical UMMA SWIZZLE_128B shared-memory
//     layouts (K-major for operands contiguous along the reduction, MN-major otherwise, so forward /
//     dgrad / wgrad need no transposed copies) -- by TMA (cp.async.bulk.tensor, 3-D / 4-D tensor maps
//     built on the host, one issuing thread) wherever the operand's base and row stride are 16-byte
//     aligned, else by per-thread cp.async,
//   * one elected thread issues 4 x tcgen05.mma (M=128, N, K=8) per chunk and commits to an mbarrier
//     that frees the stage; the epilogue reads TMEM with tcgen05.ld (32 lanes x 32 columns per warp)
//     and applies the same fused epilogues as the FFMA kernel (bias+activation(+noise), act', Adam).
// Two kernels share this machinery: gemm_tc_kernel (one tile per CTA, 2 stages x 48 KB so two CTAs share
// an SM and one CTA's epilogue overlaps the other's main loop; cp.async or TMA), and gemm_tc_ws_kernel
// (persistent, warp-specialised: TMA producer warp, MMA warp, two TMEM accumulators, sixteen epilogue
// warps; TMA-addressable problems only -- the default for them).
// Eligibility (host): rows multiple of 128, 64 <= N <= 256 and N % 16 == 0; everything else in a table
// (action-width layers) stays on the FFMA kernel in the same phase, bias gradients go to bias_grad_kernel.
#pragma once
#ifndef B200_EMU
#include "gemm.cuh"
#ifndef B200RL_TC_SPIN_ALL
#define B200RL_TC_SPIN_ALL 1
#endif

namespace b200rl {

constexpr int kTcM = 128;
constexpr int kTcNMax = 256;
constexpr int kTcBK = 32;  // fp32 elements per 128-byte swizzle row
constexpr int kTcStages = 2;
constexpr int kTcThreads = 256;
constexpr int kTcATileBytes = kTcM * 128;                       // 16 KB
constexpr int kTcBTileBytes = kTcNMax * 128;                    // 32 KB
constexpr int kTcStageBytes = kTcATileBytes + kTcBTileBytes;    // 48 KB
constexpr int kTcSmemBytes = 1024 + kTcStages * kTcStageBytes + 64;  // + alignment slack + 2 x kTcStages mbarriers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  // bounded spin: a descriptor / pipeline bug must surface as a trapped launch, never as a hung GPU
  for (unsigned spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 22)) __trap();
  }
}
// ---- TMA (cp.async.bulk.tensor): one thread moves a whole operand tile into the UMMA SWIZZLE_128B layout ----
constexpr int kTmapBytes = 128;  // sizeof(CUtensorMap)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(smem_dst)),
               "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(smem_u32(smem_dst)),
               "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, fp32 accumulate, M = 128, N from the instruction descriptor
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

// 32 TMEM lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread t <-> lane base + t)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B = 2 [61,64)
// layout: 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (the only layout the tensor core accepts
// for MN-major 32-bit operands: 32-byte swizzle granularity, 4-row atoms)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate
__host__ __device__ inline uint32_t tc_instr_desc(int n, bool a_mn_major, bool b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                        // c_format = F32
  d |= 2u << 7;                        // a_format = TF32
  d |= 2u << 10;                       // b_format = TF32
  d |= (a_mn_major ? 1u : 0u) << 15;   // a_major
  d |= (b_mn_major ? 1u : 0u) << 16;   // b_major
  d |= (uint32_t)(n >> 3) << 17;       // n_dim
  d |= (uint32_t)(kTcM >> 4) << 24;    // m_dim
  return d;
}

// One operand tile (R = 128 or 256 output indices x 32 reduction indices) as seen by one thread:
// NS = R/32 strips of 16 bytes per chunk, laid out for the UMMA SWIZZLE_128B canonical layouts:
//   K-major  (SWIZZLE_128B)        : byte(c, r) = (c/8)*1024 + (c%8)*128 + ((r/4 ^ c%8) * 16)        (SBO = 1024)
//   MN-major (SWIZZLE_128B_BASE32B): byte(c, r) = (c/32)*4096 + (r/4)*512 + (r%4)*128
//                                                 + (((c%32)/8 ^ r%4) * 32) + (c%8)*4            (LBO = 4096, SBO = 512)
template <int R, bool KMAJOR>
struct TcOperand {
  static constexpr int NS = R / 32;
  const float* src[NS];
  int dst[NS];
  unsigned fast;   // bit t: strip t is a complete, 16-byte aligned, in-range strip (modulo the reduction bound)
  unsigned rowok;  // bit t: K-major: output row inside the problem
  int pos;         // K-major: 4*(tid&7); MN-major: c0 + 4*cc
  int rr[NS];      // MN-major: reduction row of strip t inside the chunk
  long long bump;

  __device__ __forceinline__ void init(const float* base, int ld, int c0, int cmax, int tid, int r_start) {
    fast = 0; rowok = 0;
#pragma unroll
    for (int t = 0; t < NS; ++t) {
      if (KMAJOR) {
        const int c = (tid >> 3) + 32 * t, ch = tid & 7;
        dst[t] = (c >> 3) * 1024 + (c & 7) * 128 + ((ch ^ (c & 7)) << 4);
        pos = ch << 2;
        rr[t] = 0;
        const bool ok = (c0 + c) < cmax;
        src[t] = base + (long long)(c0 + c) * ld + pos + r_start;
        if (ok) rowok |= 1u << t;
        if (ok && aligned16(src[t])) fast |= 1u << t;
        bump = kTcBK;
      } else {
        constexpr int CPR = R / 4;  // 16-byte column chunks per reduction row
        const int q = tid + kTcThreads * t;
        const int r = q / CPR, cc = q % CPR;
        const int c = cc * 4;
        dst[t] = (c >> 5) * 4096 + (r >> 2) * 512 + (r & 3) * 128 + ((((c & 31) >> 3) ^ (r & 3)) << 5) + (((c & 7) >> 2) << 4);
        pos = c0 + c;
        rr[t] = r;
        src[t] = base + (long long)(r_start + r) * ld + pos;
        rowok |= 1u << t;
        if (pos + 3 < cmax && aligned16(src[t])) fast |= 1u << t;
        bump = (long long)kTcBK * ld;
      }
    }
  }
  __device__ __forceinline__ void stage(unsigned char* tile, int r0, int rlen, int cmax) {
#pragma unroll
    for (int t = 0; t < NS; ++t) {
      float* d = reinterpret_cast<float*>(tile + dst[t]);
      bool in, full;
      if (KMAJOR) { const int p = r0 + pos; in = ((rowok >> t) & 1u) && p < rlen; full = p + 3 < rlen; }
      else { in = (r0 + rr[t]) < rlen && pos < cmax; full = true; }
      if (((fast >> t) & 1u) && in && full) {
        cp_async16(d, src[t]);
      } else {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in) {
          const float* q = src[t];
          const int lim = KMAJOR ? (rlen - (r0 + pos)) : (cmax - pos);
          v.x = q[0];
          if (lim > 1) v.y = q[1];
          if (lim > 2) v.z = q[2];
          if (lim > 3) v.w = q[3];
        }
        *reinterpret_cast<float4*>(d) = v;
      }
      src[t] += bump;
    }
  }
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// (out of line: its address arithmetic must not add to the register pressure of the epilogue proper)
template <int MODE>
__device__ __noinline__ void tc_prefetch_epilogue_operands(const GemmProblem& sp, int g, int prow, int pcol, int ncols) {
  if (prow >= sp.I || pcol >= sp.J) return;
  if (MODE == kWgrad && sp.P) {
    const long long o = (long long)g * sp.p_gs + (long long)prow * sp.p_ld + pcol;
    for (int c = 0; c < ncols && pcol + c < sp.J; c += 32) { prefetch_l2(sp.P + o + c); prefetch_l2(sp.M + o + c); prefetch_l2(sp.V + o + c); }
  }
  if (MODE == kDgrad && sp.aux) {
    const long long o = (long long)g * sp.aux_gs + (long long)prow * sp.aux_ld + pcol;
    for (int c = 0; c < ncols && pcol + c < sp.J; c += 32) prefetch_l2(sp.aux + o + c);
  }
}

// Epilogue of one warp: its 32 output rows (TMEM lanes [32*rowgrp, +32) of the accumulator at `tmem_acc`) x the 128
// columns starting at `colhalf`.  tcgen05.ld hands every thread one ROW (32 consecutive columns); going to global memory
// like that would make each warp instruction touch 32 different rows.  So each 32x32 block is transposed through a padded
// shared-memory tile `st` (32 x 33 floats, private to the warp) and written row by row with lane = column: every global
// access of the epilogue -- C, bias, aux, and Adam's p/m/v -- is a coalesced 128-byte line.
// Fused AvgL1Norm (act == kActNorm): the warps that share a row group (`nshare` of them, this one is number `qidx`) first
// add up |x| of their columns straight from TMEM (thread = row), meet at a named barrier, and every thread then knows
// its row's mean; the second pass is the ordinary transposed write, of x and of x * (1 / max(mean|x|, eps)).  `nred`:
// [2][4][4][32] floats of shared memory (double-buffered by `parity`: a fast warp may start the next tile while a slow one
// still reads this tile's sums), `ncs`: 32 floats private to the warp.
template <int MODE>
__device__ __forceinline__ void tc_epilogue_warp(const GemmProblem& sp, int g, int I, int J, int n_pad, int i0, int rowgrp, int colhalf, float* st, int lane,
                                                 uint32_t tmem_acc, int total_chunks, int ncb = 4, float* nred = nullptr, float* ncs = nullptr, int nshare = 1,
                                                 int qidx = 0, unsigned parity = 0) {
  const int row_base = i0 + 32 * rowgrp;
  AdamScalars sc = {};
  if (MODE == kWgrad && sp.P) sc = *sp.adam;  // (once per tile, its round trip behind the accumulator wait of the first block)
  const bool fused_norm = (MODE == kFwd) && sp.act == kActNorm;
  if (fused_norm) {
    float part = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < ncb; ++cc) {
      const int col0 = colhalf + cc * 32;
      if (col0 >= n_pad) break;
      st[lane] = (sp.bias && col0 + lane < J) ? sp.bias[(long long)g * sp.bias_gs + col0 + lane] : 0.f;
      __syncwarp();
      float v[32];
      tc_ld32(tmem_acc + ((uint32_t)(32 * rowgrp) << 16) + (uint32_t)col0, v);
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (col0 + e < J) part += fabsf((total_chunks > 0 ? v[e] : 0.f) + st[e]);
      __syncwarp();
    }
    float* slot = nred + ((parity * 4 + rowgrp) * 4) * 32;
    slot[qidx * 32 + lane] = part;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + rowgrp), "r"(32 * nshare) : "memory");
    float tot = 0.f;
    for (int q = 0; q < nshare; ++q) tot += slot[q * 32 + lane];
    ncs[lane] = 1.f / fmaxf(tot / (float)J, 1e-8f);  // the reciprocal, once per row: 64 IEEE divisions per thread in the write pass
    __syncwarp();                                     // cost as much as the standalone normalisation launches had

  }
#pragma unroll 1
  for (int cc = 0; cc < ncb; ++cc) {
    const int col0 = colhalf + cc * 32;
    if (col0 >= n_pad) break;  // uniform across the warp
    // (requested before the TMEM load and the transposition, consumed after them)
    const float bias = (MODE == kFwd && sp.bias && col0 + lane < J) ? sp.bias[(long long)g * sp.bias_gs + col0 + lane] : 0.f;
    {
      float v[32];
      if (total_chunks > 0) {
        tc_ld32(tmem_acc + ((uint32_t)(32 * rowgrp) << 16) + (uint32_t)col0, v);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 32; ++e) st[lane * 33 + e] = v[e];
    }
    __syncwarp();
    const int col = col0 + lane;
    const bool col_ok = col < J;
    const int nrows = min(32, I - row_base);
    constexpr int U = 8;  // rows in flight: all global loads of a batch are issued before the first use
    if (MODE == kFwd) {
      // (the smoothing-noise epilogue never reaches this kernel: its layer has J = action_dim < 64)
      float* cptr = sp.C + (long long)g * sp.c_gs + (long long)row_base * sp.c_ld + col;
      if (col_ok) {
        // The activation is resolved once per block and the 32 rows go in unrolled batches of eight independent
        // load -> activation -> store chains: as one rolled loop with the activation switch inside, a 128 x 256 tile's
        // epilogue was a ~12 us dependent chain per warp (what a K = 27 first-layer tile took in total), longer than
        // the loads of a K = 256 tile.  ELU uses ex2.approx: |error| <= 2 ulp of exp(x), far inside the TF32 tolerance
        // of this path (the fp32 single-agent kernels keep expm1f).
        const long long ld = sp.c_ld;
        auto rows = [&](auto f) {
          if (nrows == 32) {
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 8) {
              float y[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) y[u] = f(st[(r0 + u) * 33 + lane] + bias);
#pragma unroll
              for (int u = 0; u < 8; ++u) cptr[(long long)(r0 + u) * ld] = y[u];
            }
          } else {
            for (int rr = 0; rr < nrows; ++rr) cptr[(long long)rr * ld] = f(st[rr * 33 + lane] + bias);
          }
        };
        if (fused_norm) {
          float* yptr = const_cast<float*>(sp.aux) + (long long)g * sp.aux_gs + (long long)row_base * sp.aux_ld + col;
          const long long ldy = sp.aux_ld;
          for (int r0 = 0; r0 < nrows; r0 += 8) {
            float x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = (r0 + u < nrows) ? st[(r0 + u) * 33 + lane] + bias : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (r0 + u < nrows) {
                cptr[(long long)(r0 + u) * ld] = x[u];
                yptr[(long long)(r0 + u) * ldy] = x[u] * ncs[r0 + u];
              }
            }
          }
        } else
        switch (sp.act) {
          case kActRelu: rows([](float v) { return v > 0.f ? v : 0.f; }); break;
          case kActElu: rows([](float v) { return v > 0.f ? v : __expf(v) - 1.f; }); break;
          case kActTanh: rows([](float v) { return tanhf(v); }); break;
          default: rows([](float v) { return v; }); break;
        }
      }
    } else if (MODE == kDgrad) {
      float* cbase = sp.C + (long long)g * sp.c_gs + (long long)row_base * sp.c_ld + col;
      const float* hbase = sp.aux ? sp.aux + (long long)g * sp.aux_gs + (long long)row_base * sp.aux_ld + col : nullptr;
      const bool acc_in = sp.accumulate != 0;
      // same shape as the forward epilogue: the derivative is resolved once per block, full blocks run in unrolled
      // batches of eight rows with every global load of a batch issued before its first use
      auto rows = [&](auto dact) {
        if (nrows == 32 && col_ok) {
#pragma unroll
          for (int r0 = 0; r0 < 32; r0 += U) {
            float hv[U], cv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              hv[u] = hbase ? hbase[(long long)(r0 + u) * sp.aux_ld] : 0.f;
              cv[u] = acc_in ? cbase[(long long)(r0 + u) * sp.c_ld] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              float x = st[(r0 + u) * 33 + lane] + cv[u];
              if (hbase) x *= dact(hv[u]);
              cbase[(long long)(r0 + u) * sp.c_ld] = x;
            }
          }
        } else {
          for (int rr = 0; rr < nrows; ++rr) {
            if (!col_ok) break;
            float x = st[rr * 33 + lane] + (acc_in ? cbase[(long long)rr * sp.c_ld] : 0.f);
            if (hbase) x *= dact(hbase[(long long)rr * sp.aux_ld]);
            cbase[(long long)rr * sp.c_ld] = x;
          }
        }
      };
      switch (sp.act) {
        case kActRelu: rows([](float h) { return h > 0.f ? 1.f : 0.f; }); break;
        case kActElu: rows([](float h) { return h > 0.f ? 1.f : h + 1.f; }); break;
        case kActTanh: rows([](float h) { return 1.f - h * h; }); break;
        default: rows([](float) { return 1.f; }); break;
      }
    } else {
      const long long poff = (long long)g * sp.p_gs + (long long)row_base * sp.p_ld + col;
      float* cbase = sp.C ? sp.C + (long long)g * sp.c_gs + (long long)row_base * sp.c_ld + col : nullptr;
      if (nrows == 32 && col_ok && sp.P) {  // full block: no per-element guards, 24 loads in flight per batch
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += U) {
          float pv[U], mv[U], vv[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const long long o = poff + (long long)(r0 + u) * sp.p_ld;
            pv[u] = sp.P[o]; mv[u] = sp.M[o]; vv[u] = sp.V[o];
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const float gval = st[(r0 + u) * 33 + lane];
            const long long o = poff + (long long)(r0 + u) * sp.p_ld;
            adam_update(gval, pv[u], mv[u], vv[u], sc);
            sp.P[o] = pv[u]; sp.M[o] = mv[u]; sp.V[o] = vv[u];
            if (cbase) cbase[(long long)(r0 + u) * sp.c_ld] = gval;
          }
        }
      } else if (col_ok) {  // ragged block or gradient-only problem: one row at a time
        for (int rr = 0; rr < nrows; ++rr) {
          const float gval = st[rr * 33 + lane];
          if (sp.P) {
            const long long o = poff + (long long)rr * sp.p_ld;
            float pv = sp.P[o], mv = sp.M[o], vv = sp.V[o];
            adam_update(gval, pv, mv, vv, sc);
            sp.P[o] = pv; sp.M[o] = mv; sp.V[o] = vv;
          }
          if (cbase) cbase[(long long)rr * sp.c_ld] = gval;
        }
      }
    }
    __syncwarp();
  }
}

template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 2) gemm_tc_kernel(const __grid_constant__ GemmLaunch L) {
  B200_PDL_PROLOGUE();
  constexpr bool A_KM = (MODE != kWgrad);
  constexpr bool B_KM = (MODE == kFwd);
  extern __shared__ __align__(16) unsigned char tc_smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes);  // [kTcStages] "stage free"
  uint64_t* full = bars + kTcStages;                                                // [kTcStages] "stage filled" (TMA path)
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.y;
  const int tile_id = blockIdx.x;
  int pi = 0;
  while (pi + 1 < L.nprob && tile_id >= L.p[pi + 1].tile_begin) ++pi;
  const GemmProblem& sp = L.p[pi];
  const int i0 = (tile_id - sp.tile_begin) * kTcM;
  const int I = sp.I, J = sp.J;
  const int n_pad = (J + 15) & ~15;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem_base)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&bars[s], 1); mbar_init(&full[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  int total_chunks = 0;
  for (int s = 0; s < sp.nseg; ++s) total_chunks += (sp.seg[s].len + kTcBK - 1) / kTcBK;

  const uint32_t idesc = tc_instr_desc(n_pad, !A_KM, !B_KM);
  const uint32_t smem_base_u32 = smem_u32(smem);
  if ((L.tma_mask >> pi) & 1u) {
    // ---- TMA main loop: thread 0 is producer and MMA issuer; nobody stages anything, nobody synchronises per chunk.
    // For chunk c it first refills the stage freed by the MMAs of chunk c-1 (tile of chunk c+S-1: one bulk tensor copy
    // per operand, completing on full[stage] with the box byte count), then waits for chunk c's tile and issues its MMAs.
    if (tid == 0) {
      const unsigned char* maps = reinterpret_cast<const unsigned char*>(L.tmaps) + (size_t)pi * 6 * kTmapBytes;
      const uint32_t tx_bytes = (uint32_t)(kTcM * 128 + J * 128);
      int lseg = 0, lr0 = 0;  // (segment, reduction offset) of the next chunk to load
      auto load_chunk = [&](int stage) {
        unsigned char* At = smem + stage * kTcStageBytes;
        mbar_expect_tx(&full[stage], tx_bytes);
        // K-major operand: 3-D tensor {r, output index, seed}; MN-major: 4-D {32 output indices, r, 32-index block, seed}
        if (A_KM) tma_load_3d(At, maps + (lseg * 2 + 0) * kTmapBytes, lr0, i0, g, &full[stage]);
        else tma_load_4d(At, maps + (lseg * 2 + 0) * kTmapBytes, 0, lr0, i0 / 32, g, &full[stage]);
        if (B_KM) tma_load_3d(At + kTcATileBytes, maps + (lseg * 2 + 1) * kTmapBytes, lr0, 0, g, &full[stage]);
        else tma_load_4d(At + kTcATileBytes, maps + (lseg * 2 + 1) * kTmapBytes, 0, lr0, 0, g, &full[stage]);
        lr0 += kTcBK;
        if (lr0 >= sp.seg[lseg].len) { ++lseg; lr0 = 0; }
      };
      for (int s = 0; s < kTcStages - 1 && s < total_chunks; ++s) load_chunk(s);
      for (int c = 0; c < total_chunks; ++c) {
        const int s = c % kTcStages;
        const int cn = c + kTcStages - 1;
        if (cn < total_chunks) {
          if (c >= 1) mbar_wait(&bars[cn % kTcStages], (uint32_t)(((c - 1) / kTcStages) & 1));
          load_chunk(cn % kTcStages);
        }
        mbar_wait(&full[s], (uint32_t)((c / kTcStages) & 1));
        tc_fence_after();
        const uint32_t a_addr = smem_base_u32 + s * kTcStageBytes;
        const uint32_t b_addr = a_addr + kTcATileBytes;
#pragma unroll
        for (int kk = 0; kk < kTcBK / 8; ++kk) {
          const uint64_t da = A_KM ? tc_smem_desc(a_addr + kk * 32, 16, 1024, 2) : tc_smem_desc(a_addr + kk * 1024, 4096, 512, 1);
          const uint64_t db = B_KM ? tc_smem_desc(b_addr + kk * 32, 16, 1024, 2) : tc_smem_desc(b_addr + kk * 1024, 4096, 512, 1);
          tc_mma_tf32(tmem_base, da, db, idesc, (c > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&bars[s]);
      }
      if (total_chunks > 0) {
        const int last = total_chunks - 1;
        mbar_wait(&bars[last % kTcStages], (uint32_t)((last / kTcStages) & 1));
      }
    }
  } else {
  TcOperand<kTcM, A_KM> ca;
  TcOperand<kTcNMax, B_KM> cb;
  int pseg = 0, pr0 = 0, plen = 0;
  auto open_segment = [&]() {
    const GemmSeg& sg = sp.seg[pseg];
    plen = sg.len;
    ca.init(sg.A + (long long)g * sg.a_gs, sg.a_ld, i0, I, tid, 0);
    cb.init(sg.B + (long long)g * sg.b_gs, sg.b_ld, 0, J, tid, 0);
    pr0 = 0;
  };
  auto issue = [&](int stage) {
    unsigned char* At = smem + stage * kTcStageBytes;
    unsigned char* Bt = At + kTcATileBytes;
    ca.stage(At, pr0, plen, I);
    cb.stage(Bt, pr0, plen, J);
    pr0 += kTcBK;
    if (pr0 >= plen && ++pseg < sp.nseg) open_segment();
  };
  open_segment();


  // prologue
#pragma unroll
  for (int s = 0; s < kTcStages - 1; ++s) {
    if (s < total_chunks) issue(s);
    cp_async_commit();
  }
  for (int c = 0; c < total_chunks; ++c) {
    const int s = c % kTcStages;
    const int cn = c + kTcStages - 1;  // chunk to prefetch: reuses the stage read by the MMAs of chunk c-1
    cp_async_wait<kTcStages - 2>();
    fence_proxy_async();  // cp.async / st.shared wrote through the generic proxy; the tensor core reads through the async proxy
    __syncthreads();      // chunk c has landed for everybody
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_base_u32 + s * kTcStageBytes;
      const uint32_t b_addr = a_addr + kTcATileBytes;
#pragma unroll
      for (int kk = 0; kk < kTcBK / 8; ++kk) {
        const uint64_t da = A_KM ? tc_smem_desc(a_addr + kk * 32, 16, 1024, 2) : tc_smem_desc(a_addr + kk * 1024, 4096, 512, 1);
        const uint64_t db = B_KM ? tc_smem_desc(b_addr + kk * 32, 16, 1024, 2) : tc_smem_desc(b_addr + kk * 1024, 4096, 512, 1);
        tc_mma_tf32(tmem_base, da, db, idesc, (c > 0 || kk > 0) ? 1u : 0u);
      }
      tc_commit(&bars[s]);  // arrives when the MMAs above (and all earlier ones) have finished reading smem / writing TMEM
      // the MMAs of chunk c are queued behind those of chunk c-1; wait (one thread only -- a block of spinning
      // threads would eat the issue slots) until chunk c-1 is done so that its stage can be refilled
    }
#if B200RL_TC_SPIN_ALL
    if (c >= 1 && cn < total_chunks) mbar_wait(&bars[cn % kTcStages], (uint32_t)(((c - 1) / kTcStages) & 1));
#else
    if (c >= 1 && cn < total_chunks) {
      if (tid == 0) mbar_wait(&bars[cn % kTcStages], (uint32_t)(((c - 1) / kTcStages) & 1));
      __syncthreads();  // uniform condition: publish "stage cn%S is free"
    }
#endif
    if (cn < total_chunks) issue(cn % kTcStages);
    cp_async_commit();
  }
  cp_async_wait<0>();
  if (total_chunks > 0 && tid == 0) {
    const int last = total_chunks - 1;
    mbar_wait(&bars[last % kTcStages], (uint32_t)((last / kTcStages) & 1));
  }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- epilogue: warp w owns rows [32*(w&3), +32) and columns [128*(w>>2), +128); the pipeline stages are free now
  // and serve as the warps' transposition tiles ----
  __shared__ float s_nred[2 * 4 * 4 * 32];
  __shared__ float s_ncs[(kTcThreads / 32) * 32];
  tc_epilogue_warp<MODE>(sp, g, I, J, n_pad, i0, warp & 3, (warp >> 2) * 128, reinterpret_cast<float*>(smem) + warp * (32 * 33), lane, tmem_base, total_chunks, 4,
                         s_nred, s_ncs + warp * 32, kTcThreads / 128, warp >> 2, 0u);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// ---- persistent, warp-specialised variant (TMA-fed problems only) ----------------------------------------------------
// One CTA per SM walks the (seed, tile) work list.  Warp 0 (one lane) is the TMA producer and runs ahead through a ring of
// kWsStages operand stages; warp 1 (one lane) issues the tcgen05 MMAs into one of TWO TMEM accumulators (2 x 256 columns);
// warps 4..11 are the epilogue: they drain accumulator a while the MMAs of the next tile fill accumulator a^1, so the
// transposing, HBM-bound epilogue (Adam state!) overlaps the operand stream of the next tile instead of leaving the SM's
// load pipeline empty.  mbarriers: full[s] (TMA bytes landed) / empty[s] (tcgen05.commit: MMAs done reading the stage),
// acc_full[a] (tcgen05.commit after the tile's last MMA) / acc_empty[a] (one arrive per epilogue warp).
constexpr int kWsStages = 3;
#ifndef B200RL_WS_EPI
#define B200RL_WS_EPI 16
#endif
constexpr int kWsEpiWarps = B200RL_WS_EPI;                 // 8: two column halves; 16: four column quarters
constexpr int kWsThreads = (4 + kWsEpiWarps) * 32;
constexpr int kWsSmemBytes = 1024 + kWsStages * kTcStageBytes + 128 + kWsEpiWarps * 32 * 33 * 4;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kWsThreads, 1) gemm_tc_ws_kernel(const __grid_constant__ GemmLaunch L, int groups) {
  __shared__ float s_nred[2 * 4 * 4 * 32];  // fused AvgL1Norm: partial row sums of the column groups (tc_epilogue_warp)
  __shared__ float s_ncs[kWsEpiWarps * 32];
  B200_PDL_PROLOGUE();
  constexpr bool A_KM = (MODE != kWgrad);
  constexpr bool B_KM = (MODE == kFwd);
  extern __shared__ __align__(16) unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWsStages * kTcStageBytes);
  uint64_t* empty = full + kWsStages;
  uint64_t* acc_full = empty + kWsStages;
  uint64_t* acc_empty = acc_full + 2;
  float* epi = reinterpret_cast<float*>(smem + kWsStages * kTcStageBytes + 128);
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < kWsStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], kWsEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const int total_items = L.total_tiles * groups;

  // work item -> (seed, problem, first row)
  auto decode = [&](int item, int& g, int& pi, int& i0) {
    g = item / L.total_tiles;
    const int tile = item - g * L.total_tiles;
    pi = 0;
    while (pi + 1 < L.nprob && tile >= L.p[pi + 1].tile_begin) ++pi;
    i0 = (tile - L.p[pi].tile_begin) * kTcM;
  };

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer
      unsigned it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int g, pi, i0;
        decode(item, g, pi, i0);
        const GemmProblem& sp = L.p[pi];
        const unsigned char* maps = reinterpret_cast<const unsigned char*>(L.tmaps) + (size_t)pi * 6 * kTmapBytes;
        const uint32_t tx_bytes = (uint32_t)(kTcM * 128 + sp.J * 128);
        for (int sg = 0; sg < sp.nseg; ++sg) {
          for (int r0 = 0; r0 < sp.seg[sg].len; r0 += kTcBK, ++it) {
            const unsigned s = it % kWsStages, use = it / kWsStages;
            if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
            unsigned char* At = smem + s * kTcStageBytes;
            mbar_expect_tx(&full[s], tx_bytes);
            if (A_KM) tma_load_3d(At, maps + (sg * 2 + 0) * kTmapBytes, r0, i0, g, &full[s]);
            else tma_load_4d(At, maps + (sg * 2 + 0) * kTmapBytes, 0, r0, i0 / 32, g, &full[s]);
            if (B_KM) tma_load_3d(At + kTcATileBytes, maps + (sg * 2 + 1) * kTmapBytes, r0, 0, g, &full[s]);
            else tma_load_4d(At + kTcATileBytes, maps + (sg * 2 + 1) * kTmapBytes, 0, r0, 0, g, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer
      const uint32_t smem_base_u32 = smem_u32(smem);
      unsigned it = 0, n = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
        int g, pi, i0;
        decode(item, g, pi, i0);
        const GemmProblem& sp = L.p[pi];
        const uint32_t idesc = tc_instr_desc((sp.J + 15) & ~15, !A_KM, !B_KM);
        const unsigned a = n & 1, ause = n >> 1;
        if (ause > 0) mbar_wait(&acc_empty[a], (ause - 1) & 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + a * 256u;
        int c = 0;
        for (int sg = 0; sg < sp.nseg; ++sg) {
          for (int r0 = 0; r0 < sp.seg[sg].len; r0 += kTcBK, ++it, ++c) {
            const unsigned s = it % kWsStages, use = it / kWsStages;
            mbar_wait(&full[s], use & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_base_u32 + s * kTcStageBytes;
            const uint32_t b_addr = a_addr + kTcATileBytes;
#pragma unroll
            for (int kk = 0; kk < kTcBK / 8; ++kk) {
              const uint64_t da = A_KM ? tc_smem_desc(a_addr + kk * 32, 16, 1024, 2) : tc_smem_desc(a_addr + kk * 1024, 4096, 512, 1);
              const uint64_t db = B_KM ? tc_smem_desc(b_addr + kk * 32, 16, 1024, 2) : tc_smem_desc(b_addr + kk * 1024, 4096, 512, 1);
              tc_mma_tf32(tmem_acc, da, db, idesc, (c > 0 || kk > 0) ? 1u : 0u);
            }
            tc_commit(&empty[s]);
          }
        }
        tc_commit(&acc_full[a]);
      }
    }
  } else if (warp >= 4) {  // ---- epilogue warps
    const int ew = warp - 4;
    float* st = epi + ew * (32 * 33);
    unsigned n = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
      int g, pi, i0;
      decode(item, g, pi, i0);
      const GemmProblem& sp = L.p[pi];
      const unsigned a = n & 1, ause = n >> 1;
      constexpr int kColsPerWarp = 256 / (kWsEpiWarps / 4);
      // The epilogue's own global operands -- Adam's p / m / v (3 x 128 KB per tile: two thirds of the kernel's bytes) and
      // the saved activations of a dgrad -- do not depend on the accumulator: their lines are requested into L2 here,
      // before the wait for the MMAs, at no cost in registers (a warp's loads in flight are bounded by the registers that
      // receive them: 24 per thread = 48 KB per SM, which kept the weight-gradient kernel latency-bound at 71 % of the HBM
      // rate; from L2 the same loads return in a fraction of the DRAM latency).
      if (MODE != kFwd) tc_prefetch_epilogue_operands<MODE>(sp, g, i0 + 32 * (warp & 3) + lane, (ew >> 2) * kColsPerWarp, kColsPerWarp);
      mbar_wait(&acc_full[a], ause & 1);
      tc_fence_after();
      tc_epilogue_warp<MODE>(sp, g, sp.I, sp.J, (sp.J + 15) & ~15, i0, warp & 3, (ew >> 2) * kColsPerWarp, st, lane, tmem_base + a * 256u, 1, kColsPerWarp / 32,
                             s_nred, s_ncs + ew * 32, kWsEpiWarps / 4, ew >> 2, n & 1u);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

inline bool tc_eligible(const GemmProblem& p) {
  if (p.I % kTcM != 0 || p.J < 64 || p.J > kTcNMax || (p.J % 16) != 0 || p.act == kActTanhNoise) return false;
  for (int s = 0; s < p.nseg; ++s)
    if (p.seg[s].b_ones) return false;
  return true;
}

}  // namespace b200rl
#endif  // !B200_EMU
