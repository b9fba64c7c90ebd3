// act.cuh -- the acting forward of agent.sample() as ONE kernel (rl/agent/td7.py:141-162, td3.py:114-141,
// sac.py:131-159): one observation per stacked seed through
//   TD7:      zs = AvgL1Norm(zs3(elu(zs2(elu(zs1(s))))))  of the FIXED encoder (rl/nn/sale.py:43-45), then
//             a  = tanh(l3(relu(l2(relu(l1([AvgL1Norm(l0(s)), zs]))))))       (rl/nn/sale.py:79-83)
//   TD3/SAC:  mlp.4(relu(mlp.2(relu(mlp.0(s)))))  (+ tanh for TD3)           (rl/nn/mlp.py:55-57)
//
// A batch-1 forward is seven dependent mat-vecs over ~1.6 MB of weights: as seven launches it is all launch latency, as
// one CTA it is bounded by what a single SM can pull out of L2.  So one thread-block CLUSTER of 8 CTAs runs one seed:
// every CTA computes an eighth of each layer's outputs (a warp per output row, lanes striding the reduction with
// 128-bit loads, all loads of a phase in flight at once), pushes them into the activation vector of all eight CTAs
// through distributed shared memory, and one cluster barrier separates the layers.  The observation may live in
// pinned host memory (read over the bus by the kernel itself) and the action may be written to a pinned host mailbox
// followed by a sequence tag, so a host-to-host call is: memcpy into the slot, one launch, poll the tag.
// Option (B200RL_ACT_PREFETCH=1, measured and left off): weights do not depend on the observation, so every CTA can
// request its slice of ALL later layers (cp.async into ~165 KB of shared memory) before the first phase.  It changes
// nothing -- TD7 15.4 us per back-to-back launch against 14.4 us, TD3 8.2 against 6.2 -- the ~2.7 us a phase costs is
// the warp reductions, the DSMEM push and the cluster barrier, not the weight reads (which come from L2 anyway).
#pragma once
#include "gemm.cuh"
#include "ops.cuh"

namespace b200rl {

constexpr int kActCluster = 8;
constexpr int kActThreads = 256;
#define ACT_PHASE_BARRIER(a) do { if ((a).dbg & 1) __syncthreads(); else { cluster_arrive(); cluster_wait(); } } while (0)

struct ActLayer {
  const float* W;  // [out][ld] of stacked seed 0, bias right behind the layer as laid out by nets.cuh
  const float* b;
  long long gs;    // blob stride between stacked seeds
  int ld, in, out;
};
struct ActArgs {
  int kind;  // 0 = TD7 (SALE), 1 = MLP + tanh (TD3), 2 = MLP raw output (SAC: mean | log_std)
  int S, H, Z, Aout;
  int vmax;  // floats per activation vector (>= H, Z; multiple of 4)
  ActLayer L[7];  // TD7: zs1 zs2 zs3 l0 l1 l2 l3;  MLP: mlp.0 mlp.2 mlp.4
  const float* obs; long long obs_gs;
  float* out; long long out_gs;
  int pre_off[7];  // float offset (from the prefetch area) of this CTA's slice of layer l, -1: read from global memory
  int pre_total;   // floats of the prefetch area (0: no prefetch)
  int dbg;         // measurement only (B200RL_ACT_DBG; results are wrong): bit 0 block barriers instead of cluster barriers, bit 1 no DSMEM pushes, bit 2 no weight loads
  unsigned* done;                    // device counter of finished CTAs (tag only)
  volatile unsigned long long* tag;  // written with tag_value once every seed's action is in `out` (null: none)
  unsigned long long tag_value;
  unsigned total_ctas;
};

#ifndef B200_EMU
__device__ __forceinline__ void st_cluster_f32(float* local, unsigned rank, float v) {
  const unsigned la = (unsigned)__cvta_generic_to_shared(local);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

// y[j] = act(b[j] + W[j][0:na] . xa + W[j][na:na+nb] . xb) for this CTA's share of the outputs.  dst != null: the
// result goes into dst[j] of EVERY CTA of the cluster; gout != null: straight to global memory.  A second layer L2r
// (same input, same output count: zs1 and l0 both read the observation) rides along so that all weight loads of the
// phase are in flight together.
// Wsm != null: this CTA's rows [j_lo, j_hi) of W have been prefetched to shared memory (row stride ld).
__device__ __forceinline__ void act_matvec(const ActLayer& Lr, int g, const float* xa, int na, const float* xb, int nb, int act, float* dst, float* gout, unsigned rank,
                                           const ActLayer* L2r = nullptr, int act2 = 0, float* dst2 = nullptr, const float* Wsm = nullptr, int dbg = 0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (Lr.out + kActCluster - 1) / kActCluster;
  const int j_lo = (int)rank * per, j_hi = min(Lr.out, j_lo + per);
  const float* W = Lr.W + (long long)g * Lr.gs;
  const float* bias = Lr.b + (long long)g * Lr.gs;
  const float* W2 = L2r ? L2r->W + (long long)g * L2r->gs : nullptr;
  const float* bias2 = L2r ? L2r->b + (long long)g * L2r->gs : nullptr;
  const bool vec = ((na | nb) & 3) == 0;  // both segments start 16-byte aligned in the weight row
  const int o_fin = lane >> 3, r_fin = lane & 7;  // lane (o, r) finishes output j0 + o and delivers it to CTA r
  for (int j0 = j_lo + warp * 4; j0 < j_hi; j0 += 8 * 4) {  // four output rows per warp and round
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
    const float* wr[4];
    const float* wr2[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      wr[o] = Wsm ? Wsm + (long long)(min(j0 + o, j_hi - 1) - j_lo) * Lr.ld : W + (long long)min(j0 + o, Lr.out - 1) * Lr.ld;
      wr2[o] = W2 ? W2 + (long long)min(j0 + o, Lr.out - 1) * L2r->ld : nullptr;
    }
    const int jf = min(j0 + o_fin, Lr.out - 1);
    const float bv = __ldg(bias + jf);  // requested before the reduction, consumed after it
    const float bv2 = bias2 ? __ldg(bias2 + jf) : 0.f;
#pragma unroll
    for (int sgm = 0; sgm < 2; ++sgm) {
      const float* x = sgm ? xb : xa;
      const int n = sgm ? nb : na, koff = sgm ? na : 0;
      if (n <= 0) continue;
      if (vec) {
#pragma unroll 2
        for (int k = lane * 4; k < n; k += 128) {
          const float4 xv = *reinterpret_cast<const float4*>(x + k);
          float4 wv[4], wv2[4];
#pragma unroll
          for (int o = 0; o < 4; ++o) wv[o] = (dbg & 4) ? make_float4(1.f, 1.f, 1.f, 1.f) : Wsm ? *reinterpret_cast<const float4*>(wr[o] + koff + k) : __ldg(reinterpret_cast<const float4*>(wr[o] + koff + k));
          if (W2) {
#pragma unroll
            for (int o = 0; o < 4; ++o) wv2[o] = __ldg(reinterpret_cast<const float4*>(wr2[o] + koff + k));
          }
#pragma unroll
          for (int o = 0; o < 4; ++o) acc[o] = fmaf(wv[o].x, xv.x, fmaf(wv[o].y, xv.y, fmaf(wv[o].z, xv.z, fmaf(wv[o].w, xv.w, acc[o]))));
          if (W2) {
#pragma unroll
            for (int o = 0; o < 4; ++o) acc2[o] = fmaf(wv2[o].x, xv.x, fmaf(wv2[o].y, xv.y, fmaf(wv2[o].z, xv.z, fmaf(wv2[o].w, xv.w, acc2[o]))));
          }
        }
      } else {
        for (int k = lane; k < n; k += 32) {
          const float xv = x[k];
#pragma unroll
          for (int o = 0; o < 4; ++o) acc[o] = fmaf(Wsm ? wr[o][koff + k] : __ldg(wr[o] + koff + k), xv, acc[o]);
          if (W2) {
#pragma unroll
            for (int o = 0; o < 4; ++o) acc2[o] = fmaf(__ldg(wr2[o] + koff + k), xv, acc2[o]);
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[o] = warp_sum(acc[o]);
    if (W2) {
#pragma unroll
      for (int o = 0; o < 4; ++o) acc2[o] = warp_sum(acc2[o]);
    }
    const int j = j0 + o_fin;
    if (j < j_hi) {
      const float y = act_forward((o_fin == 0 ? acc[0] : o_fin == 1 ? acc[1] : o_fin == 2 ? acc[2] : acc[3]) + bv, act);
      if (dst && !(dbg & 2)) st_cluster_f32(dst + j, (unsigned)r_fin, y);
      if (gout && r_fin == 0) gout[j] = y;
      if (W2) {
        const float y2 = act_forward((o_fin == 0 ? acc2[0] : o_fin == 1 ? acc2[1] : o_fin == 2 ? acc2[2] : acc2[3]) + bv2, act2);
        if (!(dbg & 2)) st_cluster_f32(dst2 + j, (unsigned)r_fin, y2);
      }
    }
  }
}

// in-place AvgL1Norm of a full vector every CTA holds: x / max(mean|x|, 1e-8) (rl/nn/sale.py:11-13)
__device__ __forceinline__ void act_norm(float* x, int n, float* red) {
  float s = 0.f;
  for (int k = threadIdx.x; k < n; k += kActThreads) s += fabsf(x[k]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kActThreads / 32; ++w) t += red[w];
  const float c = fmaxf(t / (float)n, kNormEps);
  for (int k = threadIdx.x; k < n; k += kActThreads) x[k] = x[k] / c;
  __syncthreads();
}

// request this CTA's rows of layer l (one contiguous block of the packed blob) into the prefetch area; one commit group
__device__ __forceinline__ void act_prefetch(const ActArgs& a, int l, int g, unsigned rank, float* wb) {
  if (a.pre_off[l] >= 0) {
    const ActLayer& Lr = a.L[l];
    const int per = (Lr.out + kActCluster - 1) / kActCluster;
    const int j_lo = (int)rank * per, j_hi = min(Lr.out, j_lo + per);
    const int n = max(0, j_hi - j_lo) * Lr.ld;  // (ld is a multiple of 4 floats and the blob is 16-byte aligned)
    const float* src = Lr.W + (long long)g * Lr.gs + (long long)j_lo * Lr.ld;
    float* dst = wb + a.pre_off[l];
    for (int k = threadIdx.x * 4; k < n; k += kActThreads * 4) cp_async16(dst + k, src + k);
  }
  cp_async_commit();
}

// a single agent's observation travels BY VALUE in the kernel parameters (<= 4 KB): no pinned slot, no read over the bus
constexpr int kActInlineObs = 1024;
struct ActObs { float v[kActInlineObs]; };

template <bool INLINE_OBS>
__global__ void __cluster_dims__(kActCluster, 1, 1) __launch_bounds__(kActThreads) act_fused_kernel(const __grid_constant__ ActArgs a, const __grid_constant__ ActObs ob) {
  B200_DYN_SMEM(float, sm);
  __shared__ float red[kActThreads / 32];
  const unsigned rank = cluster_ctarank();
  const int g = (int)cluster_id_x();
  float* xo = sm;                              // observation [S]
  float* v0 = xo + ((a.S + 3) & ~3);           // four activation vectors
  float* v1 = v0 + a.vmax;
  float* v2 = v1 + a.vmax;
  float* v3 = v2 + a.vmax;
  float* wb = v3 + a.vmax;                     // prefetched weight slices
  auto wsm = [&](int l) -> const float* { return a.pre_off[l] >= 0 ? wb + a.pre_off[l] : nullptr; };
  act_prefetch(a, 1, g, rank, wb);  // the second phase's weights first: they are needed soonest
  if (INLINE_OBS) {
    for (int k = threadIdx.x; k < a.S; k += kActThreads) xo[k] = ob.v[k];
  } else {
    const float* obs = a.obs + (long long)g * a.obs_gs;
    for (int k = threadIdx.x; k < a.S; k += kActThreads) xo[k] = obs[k];
  }
  __syncthreads();
  cluster_arrive();  // every CTA of the cluster is running (its shared memory may be written) ...
  cluster_wait();
  float* out = a.out + (long long)g * a.out_gs;
  if (a.kind == 0) {
    if (a.L[0].out == a.L[3].out) {  // h1 = elu(zs1 s) and p0 = l0 s in one pass over the observation
      act_matvec(a.L[0], g, xo, a.S, nullptr, 0, kActElu, v0, nullptr, rank, &a.L[3], kActNone, v1);
    } else {
      act_matvec(a.L[0], g, xo, a.S, nullptr, 0, kActElu, v0, nullptr, rank);
      act_matvec(a.L[3], g, xo, a.S, nullptr, 0, kActNone, v1, nullptr, rank);
    }
    act_prefetch(a, 2, g, rank, wb); act_prefetch(a, 4, g, rank, wb); act_prefetch(a, 5, g, rank, wb); act_prefetch(a, 6, g, rank, wb);
    cp_async_wait<4>();  // (this thread's part of layer 1; the cluster barrier covers everybody's)
    ACT_PHASE_BARRIER(a);
    act_matvec(a.L[1], g, v0, a.H, nullptr, 0, kActElu, v2, nullptr, rank, nullptr, 0, nullptr, wsm(1), a.dbg);   // h2 = elu(zs2 h1)
    act_norm(v1, a.H, red);                                                   // p0n (local copy; nobody writes v1 any more)
    cp_async_wait<3>();
    ACT_PHASE_BARRIER(a);
    act_matvec(a.L[2], g, v2, a.H, nullptr, 0, kActNone, v3, nullptr, rank, nullptr, 0, nullptr, wsm(2), a.dbg);  // u3 = zs3 h2
    cp_async_wait<2>();
    ACT_PHASE_BARRIER(a);
    act_norm(v3, a.Z, red);                                                   // zs
    act_matvec(a.L[4], g, v1, a.H, v3, a.Z, kActRelu, v0, nullptr, rank, nullptr, 0, nullptr, wsm(4), a.dbg);     // a1 = relu(l1 [p0n, zs])
    cp_async_wait<1>();
    ACT_PHASE_BARRIER(a);
    act_matvec(a.L[5], g, v0, a.H, nullptr, 0, kActRelu, v2, nullptr, rank, nullptr, 0, nullptr, wsm(5), a.dbg);  // a2 = relu(l2 a1)
    cp_async_wait<0>();
    ACT_PHASE_BARRIER(a);
    act_matvec(a.L[6], g, v2, a.H, nullptr, 0, kActTanh, nullptr, out, rank, nullptr, 0, nullptr, wsm(6), a.dbg);  // a = tanh(l3 a2)
  } else {
    act_matvec(a.L[0], g, xo, a.S, nullptr, 0, kActRelu, v0, nullptr, rank);
    act_prefetch(a, 2, g, rank, wb);
    cp_async_wait<1>();
    ACT_PHASE_BARRIER(a);
    act_matvec(a.L[1], g, v0, a.H, nullptr, 0, kActRelu, v1, nullptr, rank, nullptr, 0, nullptr, wsm(1), a.dbg);
    cp_async_wait<0>();
    ACT_PHASE_BARRIER(a);
    act_matvec(a.L[2], g, v1, a.H, nullptr, 0, a.kind == 1 ? kActTanh : kActNone, nullptr, out, rank, nullptr, 0, nullptr, wsm(2), a.dbg);
  }
  if (a.tag && a.total_ctas == (unsigned)kActCluster) {  // one seed: the cluster barrier orders the eight CTAs' stores, CTA 0 posts the tag
    __syncthreads();
    __threadfence_system();
    cluster_arrive(); cluster_wait();
    if (rank == 0 && threadIdx.x == 0) {
      __threadfence_system();
      *a.tag = a.tag_value;
    }
  } else if (a.tag) {  // last CTA of the launch posts the sequence tag behind every seed's action
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned n = atomicAdd(a.done, 1u);
      if (n == a.total_ctas - 1) {
        *a.done = 0u;
        __threadfence_system();
        *a.tag = a.tag_value;
      }
    }
  }
}
#endif

}  // namespace b200rl
