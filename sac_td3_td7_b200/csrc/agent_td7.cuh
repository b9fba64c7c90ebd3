// agent_td7.cuh -- the TD7 update (rl/agent/td7.py:287-332) as a replayable launch program.
//
// One update = encoder step | critic step -> (every policy_freq-th call) actor step, with the
// hard target copies / value-range sync / reset_max_priority every target_update_rate calls.
// The encoder step touches only `encoder`, so it runs as a parallel branch of the graph; the
// part of the actor step that does not depend on the freshly updated critics (policy forward,
// fixed-encoder zsa) is hoisted onto a second branch.  Layers with no mutual dependence are
// grouped into one GEMM launch (see gemm.cuh).
#pragma once
#include "agent_base.cuh"

namespace b200rl {

__global__ void td7_range_sync_kernel(TD7Range* r, int G) {
  B200_PDL_PROLOGUE();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) { r[g].value_target_max = r[g].value_max; r[g].value_target_min = r[g].value_min; }
}

struct TD7Config {
  int S, A, H, Z, B, G;
  double discount, policy_lr, critic_lr;
  int target_update_rate, policy_freq;
  double target_policy_noise, noise_clip;
  int use_lap, keep_grads;
};

struct TD7Agent : AgentBase {
  TD7Config cfg;
  // encoder(trained) policy(trained) q1 q2(trained) | target_q1 target_q2 fixed_encoder fixed_encoder_target
  Net enc, pol, q[2], qt[2], fe, fet;
  TD7Range* d_range = nullptr;
  float* d_dqa = nullptr;  // constant -1/(2B) seed gradient of the actor loss
  // activations
  Mat noise;
  Mat e_h1, e_h2, e_u3, e_z, e_g1, e_g2, e_zsa, d_zsa, d_g2, d_g1, d_z, d_u3, d_h2, d_h1;
  Mat t1, t2, tu3, zs_t, f1, f2, fu3, zs_f, pa0, pa0n, p1, p2, a_next, fa1, fa2, zsa_f, ta1, ta2, zsa_t;
  Mat tq0[2], tq0n[2], tqa[2], tqb[2], qtv[2], cq0[2], cq0n[2], cqa[2], cqb[2], qv[2], dq[2], d_cqb[2], d_cqa[2], d_cq0n[2], d_cq0[2];
  Mat pb0, pb0n, pp1, pp2, a_pi, pa1, pa2, zsa_pi, aq0[2], aq0n[2], aqa[2], aqb[2], qpi[2], dqa[2];
  Mat d_aqb[2], d_aqa[2], d_aq0n[2], d_aq0[2], d_zsa_pi, d_pa2, d_pa1, d_api, d_pp2, d_pp1, d_pb0n, d_pb0;
  Mat y_dbg, prio_dbg;
  float* enc_partial = nullptr;
  unsigned* enc_counter = nullptr;

  // launch tables
  GemmTable E[16], C[24], Aa[24];
  GemmTable We[6];         // encoder weight gradients (+Adam), one launch per layer
  GemmTable Cq[3];         // one agent: the current critics' forward q1 q2 q3 as a branch of its own (split_q)
  bool split_q = false;
  // stacked seeds on the tensor cores: every AvgL1Norm rides in the epilogue of the layer that produces its input (kActNorm),
  // the six normalisation launches of an update (each a full pass over 67-134 MB at 256 seeds) are not issued
  bool fuse_norm = false;
  static GemmProblem with_norm(GemmProblem pr, const Mat& y) {
    pr.act = kActNorm; pr.aux = y.p; pr.aux_ld = y.ld; pr.aux_gs = y.gs;
    return pr;
  }
  GemmTable C2v[2];        // second launch of the critic chain, without / with the actor step's l0(s)
  GemmTable Wq[4], Wp[4];  // critic / policy weight gradients (+Adam), one launch per layer: [0] = last layer
  RowTable En[2], Cn[4], An[4];
  RowTable CnQ, CnP[2];

  void layout_acts() {
    act_cursor = 0;
    const int B2 = 2 * B;
    ss = alloc_act(B2, S, true); act_in = alloc_act(B, A, true); rew = alloc_act(B, 1); nd = alloc_act(B, 1); noise = alloc_act(B, A);
    e_h1 = alloc_act(B2, H); e_h2 = alloc_act(B2, H); e_u3 = alloc_act(B2, Z); e_z = alloc_act(B2, Z);
    e_g1 = alloc_act(B, H); e_g2 = alloc_act(B, H); e_zsa = alloc_act(B, Z); d_zsa = alloc_act(B, Z);
    d_g2 = alloc_act(B, H); d_g1 = alloc_act(B, H); d_z = alloc_act(B, Z); d_u3 = alloc_act(B, Z); d_h2 = alloc_act(B, H); d_h1 = alloc_act(B, H);
    t1 = alloc_act(B, H); t2 = alloc_act(B, H); tu3 = alloc_act(B, Z); zs_t = alloc_act(B, Z);
    f1 = alloc_act(B, H); f2 = alloc_act(B, H); fu3 = alloc_act(B, Z); zs_f = alloc_act(B, Z);
    pa0 = alloc_act(B, H); pa0n = alloc_act(B, H); p1 = alloc_act(B, H); p2 = alloc_act(B, H); a_next = alloc_act(B, A, true);
    fa1 = alloc_act(B, H); fa2 = alloc_act(B, H); zsa_f = alloc_act(B, Z); ta1 = alloc_act(B, H); ta2 = alloc_act(B, H); zsa_t = alloc_act(B, Z);
    for (int k = 0; k < 2; ++k) {
      tq0[k] = alloc_act(B, H); tq0n[k] = alloc_act(B, H); tqa[k] = alloc_act(B, H); tqb[k] = alloc_act(B, H); qtv[k] = alloc_act(B, 1);
      cq0[k] = alloc_act(B, H); cq0n[k] = alloc_act(B, H); cqa[k] = alloc_act(B, H); cqb[k] = alloc_act(B, H); qv[k] = alloc_act(B, 1);
      dq[k] = alloc_act(B, 1); d_cqb[k] = alloc_act(B, H); d_cqa[k] = alloc_act(B, H); d_cq0n[k] = alloc_act(B, H); d_cq0[k] = alloc_act(B, H);
      aq0[k] = alloc_act(B, H); aq0n[k] = alloc_act(B, H); aqa[k] = alloc_act(B, H); aqb[k] = alloc_act(B, H); qpi[k] = alloc_act(B, 1);
      dqa[k] = alloc_act(B, 1);
      d_aqb[k] = alloc_act(B, H); d_aqa[k] = alloc_act(B, H); d_aq0n[k] = alloc_act(B, H); d_aq0[k] = alloc_act(B, H);
    }
    pb0 = alloc_act(B, H); pb0n = alloc_act(B, H); pp1 = alloc_act(B, H); pp2 = alloc_act(B, H); a_pi = alloc_act(B, A, true);
    pa1 = alloc_act(B, H); pa2 = alloc_act(B, H); zsa_pi = alloc_act(B, Z);
    d_zsa_pi = alloc_act(B, Z); d_pa2 = alloc_act(B, H); d_pa1 = alloc_act(B, H); d_api = alloc_act(B, A);
    d_pp2 = alloc_act(B, H); d_pp1 = alloc_act(B, H); d_pb0n = alloc_act(B, H); d_pb0 = alloc_act(B, H);
    y_dbg = alloc_act(B, 1); prio_dbg = alloc_act(B, 1);
  }

  int init(const TD7Config& c) {
    cfg = c;
    S = c.S; A = c.A; H = c.H; Z = c.Z; B = c.B; G = c.G;
    use_lap = c.use_lap != 0; keep_grads = c.keep_grads != 0;
    B200_REQUIRE(S > 0 && A > 0 && H > 0 && Z > 0 && G > 0, "td7: dims must be positive");
    B200_REQUIRE(B > 0 && B <= 1024, "td7: batch size must be in 1..1024");
    B200_REQUIRE(H <= 1024 && Z <= 1024, "td7: hidden sizes above 1024 are not supported");
    OptimHyper hy[3] = {{c.policy_lr, 0.9, 0.999, 1e-8}, {c.critic_lr, 0.9, 0.999, 1e-8}, {c.policy_lr, 0.9, 0.999, 1e-8}};  // policy, q, encoder
    low_priority_streams = 0x3u;  // side[0] = encoder branch, side[1] = hoisted actor forward
    B200_TRY(init_common(3, 3, hy));
    B200_TRY(alloc_net(&enc, make_layout(kNetSaleEncoder, S, A, H, Z), true, d_adam + 2));
    B200_TRY(alloc_net(&pol, make_layout(kNetSaleActor, S, A, H, Z), true, d_adam + 0));
    for (int k = 0; k < 2; ++k) {
      B200_TRY(alloc_net(&q[k], make_layout(kNetSaleCritic, S, A, H, Z), true, d_adam + 1));
      B200_TRY(alloc_net(&qt[k], make_layout(kNetSaleCritic, S, A, H, Z), false, nullptr));
    }
    B200_TRY(alloc_net(&fe, make_layout(kNetSaleEncoder, S, A, H, Z), false, nullptr));
    B200_TRY(alloc_net(&fet, make_layout(kNetSaleEncoder, S, A, H, Z), false, nullptr));
    B200_TRY(arena.alloc(&d_range, (size_t)G));
    std::vector<TD7Range> r0(G, TD7Range{-1e8f, 1e8f, 0.f, 0.f});  // td7.py:84-87
    B200_CHECK_CUDA(cudaMemcpy(d_range, r0.data(), G * sizeof(TD7Range), cudaMemcpyHostToDevice));
    B200_TRY(arena.alloc(&enc_partial, (size_t)G * ceil_div(B, 8)));
    B200_TRY(arena.alloc(&enc_counter, (size_t)G));
    layout_acts();
    act_stride = act_cursor;
    B200_TRY(arena.alloc(&act_base, (size_t)G * act_stride));
    layout_acts();
    // constant seed gradient of L = -mean(cat[q1,q2])
    for (int k = 0; k < 2; ++k) {
      B200_LAUNCH(fill_kernel, dim3(1), dim3(256), 0, stream, dqa[k].p, (long long)B, -1.0f / (2.0f * (float)B));
      for (int g = 1; g < G; ++g)
        B200_LAUNCH(fill_kernel, dim3(1), dim3(256), 0, stream, dqa[k].p + (long long)g * act_stride, (long long)B, -1.0f / (2.0f * (float)B));
    }
    B200_CHECK_LAUNCH();
    B200_TRY(build_tables());
    B200_TRY(sync());
    B200_CHECK_CUDA(cudaDeviceSynchronize());  // (uploads above went through the legacy default stream; the engine runs on non-blocking ones)
    return kOk;
  }

  // side = the launch belongs to a branch with slack (encoder step, hoisted actor forward): no split-K there, so that
  // it takes as few SMs as possible away from the critical chain it runs beside
  template <class V>
  int up(GemmTable& t, int mode, V probs, bool side = false) {
    std::vector<GemmProblem> v(probs);
    return gemm_table_upload(&t, mode, v, &arena, G, side ? 1 : 0);
  }
  int upn(RowTable& t, bool bwd, std::vector<RowOp> ops) { return row_table_upload(&t, bwd, ops, &arena); }
  static RowOp nf(const Mat& x, const Mat& y) { return make_rowop(x.p, x.ld, x.gs, y.p, y.ld, y.gs, x.rows, x.cols); }
  static RowOp nb(const Mat& x, const Mat& dy, const Mat& dx) { return make_rowop(x.p, x.ld, x.gs, dx.p, dx.ld, dx.gs, x.rows, x.cols, dy.p, dy.ld, dy.gs); }

  int build_tables() {
    using PV = std::vector<GemmProblem>;
    const Mat s_all = ss, s_cur = ss.rows_from(0, B), s_nxt = ss.rows_from(B, B);
    const Mat z_cur = e_z.rows_from(0, B), z_nxt = e_z.rows_from(B, B);
    const float sigma = (float)cfg.target_policy_noise, clipv = (float)cfg.noise_clip;
    // (only where every producing layer is a tensor-core problem: 128-row tiles, 64..256 columns in steps of 16)
    fuse_norm = use_tensor_cores(G) && (B % 128) == 0 && H >= 64 && H <= 256 && (H % 16) == 0 && Z >= 64 && Z <= 256 && (Z % 16) == 0;
    auto nrm = [&](GemmProblem pr, const Mat& y) { return fuse_norm ? with_norm(pr, y) : pr; };
    // ------------------------------------------------ encoder branch (td7.py:246-257, 299-303)
    B200_TRY(up(E[0], kFwd, PV{fwd_problem(enc, 0, {s_all}, e_h1, kActElu)}, true));
    B200_TRY(up(E[1], kFwd, PV{fwd_problem(enc, 1, {e_h1}, e_h2, kActElu)}, true));
    B200_TRY(up(E[2], kFwd, PV{nrm(fwd_problem(enc, 2, {e_h2}, e_u3, kActNone), e_z)}, true));
    B200_TRY(upn(En[0], false, {nf(e_u3, e_z)}));
    B200_TRY(up(E[3], kFwd, PV{fwd_problem(enc, 3, {z_cur, act_in}, e_g1, kActElu)}, true));
    B200_TRY(up(E[4], kFwd, PV{fwd_problem(enc, 4, {e_g1}, e_g2, kActElu)}, true));
    B200_TRY(up(E[5], kFwd, PV{fwd_problem(enc, 5, {e_g2}, e_zsa, kActNone)}, true));
    B200_TRY(up(E[6], kDgrad, PV{dgrad_problem({{&enc, 5, d_zsa, 0}}, H, d_g2, &e_g2, kActElu)}, true));
    B200_TRY(up(E[7], kDgrad, PV{dgrad_problem({{&enc, 4, d_g2, 0}}, H, d_g1, &e_g1, kActElu)}, true));
    B200_TRY(up(E[8], kDgrad, PV{dgrad_problem({{&enc, 3, d_g1, 0}}, Z, d_z, nullptr, 0)}, true));
    B200_TRY(upn(En[1], true, {nb(e_u3.rows_from(0, B), d_z, d_u3)}));
    {
      Mat h2c = e_h2.rows_from(0, B), h1c = e_h1.rows_from(0, B);
      B200_TRY(up(E[9], kDgrad, PV{dgrad_problem({{&enc, 2, d_u3, 0}}, H, d_h2, &h2c, kActElu)}, true));
      B200_TRY(up(E[10], kDgrad, PV{dgrad_problem({{&enc, 1, d_h2, 0}}, H, d_h1, &h1c, kActElu)}, true));
      // one small launch per layer, issued right after the dgrad that reads that layer's weights: the encoder step runs
      // beside the critic chain, and 64-tile launches slip into its idle SMs where one 384-tile burst would stall it
      const Mat* dz[6] = {&d_h1, &d_h2, &d_u3, &d_g1, &d_g2, &d_zsa};
      for (int l = 0; l < 6; ++l) {
        PV w;
        if (l == 0) wgrad_problems(w, enc, 0, *dz[0], {s_cur}, true);
        else if (l == 1) wgrad_problems(w, enc, 1, *dz[1], {h1c}, true);
        else if (l == 2) wgrad_problems(w, enc, 2, *dz[2], {h2c}, true);
        else if (l == 3) wgrad_problems(w, enc, 3, *dz[3], {z_cur, act_in}, true);
        else if (l == 4) wgrad_problems(w, enc, 4, *dz[4], {e_g1}, true);
        else wgrad_problems(w, enc, 5, *dz[5], {e_g2}, true);
        B200_TRY(up(We[l], kWgrad, w, true));
      }
    }
    // ------------------------------------------------ critic chain (td7.py:175-244, 307-315)
    // first launch: 4 x 64 tiles = one wave at 2 CTAs per SM; the policy's l0 rides with the second launch (128 tiles)
    B200_TRY(up(C[0], kFwd, PV{fwd_problem(fet, 0, {s_nxt}, t1, kActElu), fwd_problem(fe, 0, {s_cur}, f1, kActElu),
                               nrm(fwd_problem(q[0], 0, {s_cur, act_in}, cq0[0], kActNone), cq0n[0]), nrm(fwd_problem(q[1], 0, {s_cur, act_in}, cq0[1], kActNone), cq0n[1])}));
    B200_TRY(upn(CnQ, false, {nf(cq0[0], cq0n[0]), nf(cq0[1], cq0n[1])}));
    {
      PV c2{fwd_problem(fet, 1, {t1}, t2, kActElu), fwd_problem(fe, 1, {f1}, f2, kActElu), nrm(fwd_problem(pol, 0, {s_nxt}, pa0, kActNone), pa0n)};
      B200_TRY(up(C2v[0], kFwd, c2));
      c2.push_back(nrm(fwd_problem(pol, 0, {s_cur}, pb0, kActNone), pb0n));  // actor steps also need l0(s)
      B200_TRY(up(C2v[1], kFwd, c2));
    }
    B200_TRY(upn(CnP[0], false, {nf(pa0, pa0n)}));
    B200_TRY(upn(CnP[1], false, {nf(pa0, pa0n), nf(pb0, pb0n)}));
    B200_TRY(up(C[3], kFwd, PV{nrm(fwd_problem(fet, 2, {t2}, tu3, kActNone), zs_t), nrm(fwd_problem(fe, 2, {f2}, fu3, kActNone), zs_f)}));
    B200_TRY(upn(Cn[2], false, {nf(tu3, zs_t), nf(fu3, zs_f)}));
    B200_TRY(up(C[4], kFwd, PV{fwd_problem(pol, 1, {pa0n, zs_t}, p1, kActRelu), fwd_problem(fe, 3, {zs_f, act_in}, fa1, kActElu)}));
    B200_TRY(up(C[5], kFwd, PV{fwd_problem(pol, 2, {p1}, p2, kActRelu), fwd_problem(fe, 4, {fa1}, fa2, kActElu)}));
    {
      GemmProblem pn = fwd_problem(pol, 3, {p2}, a_next, kActTanhNoise);  // + clipped smoothing noise, clamp to [-1,1]
      pn.aux = noise.p; pn.aux_ld = noise.ld; pn.aux_gs = noise.gs; pn.noise_sigma = sigma; pn.noise_clip = clipv;
      B200_TRY(up(C[6], kFwd, PV{pn, fwd_problem(fe, 5, {fa2}, zsa_f, kActNone)}));
    }
    // The current critics' forward (three launches) has slack -- the target side still has six launches to go when its
    // inputs are ready -- so for one agent it runs as its own branch: riding with the target chain's launches it made
    // the first of them 320 tiles (a second wave at two CTAs per SM for 24 CTAs: 16 us instead of ~7).  Stacked seeds
    // fill the machine either way and keep the grouped launches.
    split_q = (G == 1);
    if (split_q) {
      B200_TRY(up(C[7], kFwd, PV{fwd_problem(fet, 3, {zs_t, a_next}, ta1, kActElu), nrm(fwd_problem(qt[0], 0, {s_nxt, a_next}, tq0[0], kActNone), tq0n[0]),
                                 nrm(fwd_problem(qt[1], 0, {s_nxt, a_next}, tq0[1], kActNone), tq0n[1])}));
      B200_TRY(up(C[8], kFwd, PV{fwd_problem(fet, 4, {ta1}, ta2, kActElu)}));
      B200_TRY(up(C[9], kFwd, PV{fwd_problem(fet, 5, {ta2}, zsa_t, kActNone)}));
      B200_TRY(up(Cq[0], kFwd, PV{fwd_problem(q[0], 1, {cq0n[0], zsa_f, zs_f}, cqa[0], kActElu), fwd_problem(q[1], 1, {cq0n[1], zsa_f, zs_f}, cqa[1], kActElu)}, true));
      B200_TRY(up(Cq[1], kFwd, PV{fwd_problem(q[0], 2, {cqa[0]}, cqb[0], kActElu), fwd_problem(q[1], 2, {cqa[1]}, cqb[1], kActElu)}, true));
      B200_TRY(up(Cq[2], kFwd, PV{fwd_problem(q[0], 3, {cqb[0]}, qv[0], kActNone), fwd_problem(q[1], 3, {cqb[1]}, qv[1], kActNone)}, true));
    } else {
      B200_TRY(up(C[7], kFwd, PV{fwd_problem(fet, 3, {zs_t, a_next}, ta1, kActElu), nrm(fwd_problem(qt[0], 0, {s_nxt, a_next}, tq0[0], kActNone), tq0n[0]),
                                 nrm(fwd_problem(qt[1], 0, {s_nxt, a_next}, tq0[1], kActNone), tq0n[1]), fwd_problem(q[0], 1, {cq0n[0], zsa_f, zs_f}, cqa[0], kActElu),
                                 fwd_problem(q[1], 1, {cq0n[1], zsa_f, zs_f}, cqa[1], kActElu)}));
      B200_TRY(up(C[8], kFwd, PV{fwd_problem(fet, 4, {ta1}, ta2, kActElu), fwd_problem(q[0], 2, {cqa[0]}, cqb[0], kActElu), fwd_problem(q[1], 2, {cqa[1]}, cqb[1], kActElu)}));
      B200_TRY(up(C[9], kFwd, PV{fwd_problem(fet, 5, {ta2}, zsa_t, kActNone), fwd_problem(q[0], 3, {cqb[0]}, qv[0], kActNone), fwd_problem(q[1], 3, {cqb[1]}, qv[1], kActNone)}));
    }
    B200_TRY(upn(Cn[3], false, {nf(tq0[0], tq0n[0]), nf(tq0[1], tq0n[1])}));
    B200_TRY(up(C[10], kFwd, PV{fwd_problem(qt[0], 1, {tq0n[0], zsa_t, zs_t}, tqa[0], kActElu), fwd_problem(qt[1], 1, {tq0n[1], zsa_t, zs_t}, tqa[1], kActElu)}));
    B200_TRY(up(C[11], kFwd, PV{fwd_problem(qt[0], 2, {tqa[0]}, tqb[0], kActElu), fwd_problem(qt[1], 2, {tqa[1]}, tqb[1], kActElu)}));
    B200_TRY(up(C[12], kFwd, PV{fwd_problem(qt[0], 3, {tqb[0]}, qtv[0], kActNone), fwd_problem(qt[1], 3, {tqb[1]}, qtv[1], kActNone)}));
    B200_TRY(up(C[13], kDgrad, PV{dgrad_problem({{&q[0], 3, dq[0], 0}}, H, d_cqb[0], &cqb[0], kActElu), dgrad_problem({{&q[1], 3, dq[1], 0}}, H, d_cqb[1], &cqb[1], kActElu)}));
    B200_TRY(up(C[14], kDgrad, PV{dgrad_problem({{&q[0], 2, d_cqb[0], 0}}, H, d_cqa[0], &cqa[0], kActElu), dgrad_problem({{&q[1], 2, d_cqb[1], 0}}, H, d_cqa[1], &cqa[1], kActElu)}));
    B200_TRY(up(C[15], kDgrad, PV{dgrad_problem({{&q[0], 1, d_cqa[0], 0}}, H, d_cq0n[0], nullptr, 0), dgrad_problem({{&q[1], 1, d_cqa[1], 0}}, H, d_cq0n[1], nullptr, 0)}));
    {
      std::vector<RowOp> nbq{nb(cq0[0], d_cq0n[0], d_cq0[0]), nb(cq0[1], d_cq0n[1], d_cq0[1])};
      B200_TRY(row_table_upload(&critic_norm_bwd, true, nbq, &arena));
      // one launch per layer: layer l's gradient (and its in-place Adam step) can start as soon as the dgrad that READS
      // layer l's weights has finished, instead of waiting for the whole backward chain
      PV w3, w2, w1, w0;
      for (int k = 0; k < 2; ++k) {
        wgrad_problems(w3, q[k], 3, dq[k], {cqb[k]}, true);
        wgrad_problems(w2, q[k], 2, d_cqb[k], {cqa[k]}, true);
        wgrad_problems(w1, q[k], 1, d_cqa[k], {cq0n[k], zsa_f, zs_f}, true);
        wgrad_problems(w0, q[k], 0, d_cq0[k], {s_cur, act_in}, true);
      }
      B200_TRY(up(Wq[0], kWgrad, w3));
      B200_TRY(up(Wq[1], kWgrad, w2));
      B200_TRY(up(Wq[2], kWgrad, w1));
      B200_TRY(up(Wq[3], kWgrad, w0));
    }
    // ------------------------------------------------ actor step (td7.py:259-276, 319-324)
    B200_TRY(up(Aa[0], kFwd, PV{fwd_problem(pol, 1, {pb0n, zs_f}, pp1, kActRelu)}, true));
    B200_TRY(up(Aa[1], kFwd, PV{fwd_problem(pol, 2, {pp1}, pp2, kActRelu)}, true));
    B200_TRY(up(Aa[2], kFwd, PV{fwd_problem(pol, 3, {pp2}, a_pi, kActTanh)}, true));
    B200_TRY(up(Aa[3], kFwd, PV{fwd_problem(fe, 3, {zs_f, a_pi}, pa1, kActElu)}, true));
    B200_TRY(up(Aa[4], kFwd, PV{fwd_problem(fe, 4, {pa1}, pa2, kActElu)}, true));
    B200_TRY(up(Aa[5], kFwd, PV{fwd_problem(fe, 5, {pa2}, zsa_pi, kActNone)}, true));
    B200_TRY(up(Aa[6], kFwd, PV{nrm(fwd_problem(q[0], 0, {s_cur, a_pi}, aq0[0], kActNone), aq0n[0]), nrm(fwd_problem(q[1], 0, {s_cur, a_pi}, aq0[1], kActNone), aq0n[1])}));
    B200_TRY(upn(An[0], false, {nf(aq0[0], aq0n[0]), nf(aq0[1], aq0n[1])}));
    B200_TRY(up(Aa[7], kFwd, PV{fwd_problem(q[0], 1, {aq0n[0], zsa_pi, zs_f}, aqa[0], kActElu), fwd_problem(q[1], 1, {aq0n[1], zsa_pi, zs_f}, aqa[1], kActElu)}));
    B200_TRY(up(Aa[8], kFwd, PV{fwd_problem(q[0], 2, {aqa[0]}, aqb[0], kActElu), fwd_problem(q[1], 2, {aqa[1]}, aqb[1], kActElu)}));
    B200_TRY(up(Aa[9], kFwd, PV{fwd_problem(q[0], 3, {aqb[0]}, qpi[0], kActNone), fwd_problem(q[1], 3, {aqb[1]}, qpi[1], kActNone)}));
    B200_TRY(up(Aa[10], kDgrad, PV{dgrad_problem({{&q[0], 3, dqa[0], 0}}, H, d_aqb[0], &aqb[0], kActElu), dgrad_problem({{&q[1], 3, dqa[1], 0}}, H, d_aqb[1], &aqb[1], kActElu)}));
    B200_TRY(up(Aa[11], kDgrad, PV{dgrad_problem({{&q[0], 2, d_aqb[0], 0}}, H, d_aqa[0], &aqa[0], kActElu), dgrad_problem({{&q[1], 2, d_aqb[1], 0}}, H, d_aqa[1], &aqa[1], kActElu)}));
    B200_TRY(up(Aa[12], kDgrad, PV{dgrad_problem({{&q[0], 1, d_aqa[0], 0}}, H, d_aq0n[0], nullptr, 0), dgrad_problem({{&q[1], 1, d_aqa[1], 0}}, H, d_aq0n[1], nullptr, 0),
                                   dgrad_problem({{&q[0], 1, d_aqa[0], H}, {&q[1], 1, d_aqa[1], H}}, Z, d_zsa_pi, nullptr, 0)}));
    B200_TRY(upn(An[1], true, {nb(aq0[0], d_aq0n[0], d_aq0[0]), nb(aq0[1], d_aq0n[1], d_aq0[1])}));
    B200_TRY(up(Aa[13], kDgrad, PV{dgrad_problem({{&fe, 5, d_zsa_pi, 0}}, H, d_pa2, &pa2, kActElu)}));
    B200_TRY(up(Aa[14], kDgrad, PV{dgrad_problem({{&fe, 4, d_pa2, 0}}, H, d_pa1, &pa1, kActElu)}));
    // d a = dq01_1[:, S:S+A] + dq01_2[:, S:S+A] + dzsa1[:, Z:Z+A]; times tanh'
    B200_TRY(up(Aa[15], kDgrad, PV{dgrad_problem({{&q[0], 0, d_aq0[0], S}, {&q[1], 0, d_aq0[1], S}, {&fe, 3, d_pa1, Z}}, A, d_api, &a_pi, kActTanh)}));
    B200_TRY(up(Aa[16], kDgrad, PV{dgrad_problem({{&pol, 3, d_api, 0}}, H, d_pp2, &pp2, kActRelu)}));
    B200_TRY(up(Aa[17], kDgrad, PV{dgrad_problem({{&pol, 2, d_pp2, 0}}, H, d_pp1, &pp1, kActRelu)}));
    B200_TRY(up(Aa[18], kDgrad, PV{dgrad_problem({{&pol, 1, d_pp1, 0}}, H, d_pb0n, nullptr, 0)}));
    B200_TRY(upn(An[2], true, {nb(pb0, d_pb0n, d_pb0)}));
    {
      PV w3, w2, w1, w0;
      wgrad_problems(w3, pol, 3, d_api, {pp2}, true);
      wgrad_problems(w2, pol, 2, d_pp2, {pp1}, true);
      wgrad_problems(w1, pol, 1, d_pp1, {pb0n, zs_f}, true);
      wgrad_problems(w0, pol, 0, d_pb0, {s_cur}, true);
      B200_TRY(up(Wp[0], kWgrad, w3));
      B200_TRY(up(Wp[1], kWgrad, w2));
      B200_TRY(up(Wp[2], kWgrad, w1));
      B200_TRY(up(Wp[3], kWgrad, w0));
    }
    built = true;
    B200_TRY(rebuild_mega());
    return kOk;
  }
  RowTable critic_norm_bwd;

#ifndef B200_EMU
  // ---- the update as three persistent dataflow launches (mega.cuh): [A] everything forward, [B] encoder / critic
  // backward with the weight gradients (+ Adam) and, on actor steps, the critics' forward at pi(s), [C] the actor's
  // backward.  Only the loss kernels (and the LAP write-back beside [B]) remain separate launches.
  std::unique_ptr<MegaProgram> MA[2], MB[2], MC;
  int rebuild_mega() override {
    if (!use_mega()) return kOk;
    for (int v = 0; v < 2; ++v) {
      const bool actor = v == 1;
      auto a = std::make_unique<MegaProgram>();
      a->add_prep(prep_args(actor ? 0x7u : 0x6u));
      a->add_gauss(gauss_args(noise, 0));
      for (int i : {0, 1, 2}) a->add_gemm_table(E[i]);
      a->add_row_ops(En[0].h_ops, false);
      for (int i : {3, 4, 5}) a->add_gemm_table(E[i]);
      a->add_gemm_table(C[0]);
      a->add_row_ops(CnQ.h_ops, false);
      a->add_gemm_table(C2v[v]);
      a->add_row_ops(CnP[v].h_ops, false);
      a->add_gemm_table(C[3]);
      a->add_row_ops(Cn[2].h_ops, false);
      if (actor) for (int i = 0; i <= 5; ++i) a->add_gemm_table(Aa[i]);
      for (int i : {4, 5, 6, 7}) a->add_gemm_table(C[i]);
      if (split_q) for (int i = 0; i < 3; ++i) a->add_gemm_table(Cq[i]);
      a->add_row_ops(Cn[3].h_ops, false);
      for (int i : {8, 9, 10, 11, 12}) a->add_gemm_table(C[i]);
      B200_TRY(a->build(&arena, G, sm_count()));
      MA[v] = std::move(a);

      auto b = std::make_unique<MegaProgram>();
      b->add_gemm_table(C[13]); b->add_gemm_table(Wq[0]);
      b->add_gemm_table(C[14]); b->add_gemm_table(Wq[1]);
      b->add_gemm_table(C[15]); b->add_gemm_table(Wq[2]);
      b->add_row_ops(critic_norm_bwd.h_ops, true);
      b->add_gemm_table(Wq[3]);
      b->add_gemm_table(E[6]); b->add_gemm_table(We[5]);
      b->add_gemm_table(E[7]); b->add_gemm_table(We[4]);
      b->add_gemm_table(E[8]); b->add_gemm_table(We[3]);
      b->add_row_ops(En[1].h_ops, true);
      b->add_gemm_table(E[9]); b->add_gemm_table(We[2]);
      b->add_gemm_table(E[10]); b->add_gemm_table(We[1]); b->add_gemm_table(We[0]);
      if (actor) {
        b->add_gemm_table(Aa[6]);
        b->add_row_ops(An[0].h_ops, false);
        for (int i : {7, 8, 9}) b->add_gemm_table(Aa[i]);
      }
      B200_TRY(b->build(&arena, G, sm_count()));
      MB[v] = std::move(b);
    }
    auto c = std::make_unique<MegaProgram>();
    for (int i : {10, 11, 12}) c->add_gemm_table(Aa[i]);
    c->add_row_ops(An[1].h_ops, true);
    for (int i : {13, 14, 15, 16}) c->add_gemm_table(Aa[i]);
    c->add_gemm_table(Wp[0]);
    c->add_gemm_table(Aa[17]); c->add_gemm_table(Wp[1]);
    c->add_gemm_table(Aa[18]); c->add_gemm_table(Wp[2]);
    c->add_row_ops(An[2].h_ops, true);
    c->add_gemm_table(Wp[3]);
    B200_TRY(c->build(&arena, G, sm_count()));
    MC = std::move(c);
    return kOk;
  }
  int issue_step_mega(int variant) {
    const bool actor = (variant & 1) != 0;
    cudaStream_t M = stream, sE = fork.side[0], sP = fork.side[3];
    B200_TRY(MA[variant]->launch(M));
    B200_TRY(fork.link(M, sE));
    B200_TRY(launch_enc_loss(sE));
    B200_TRY(launch_critic_loss(M));
    const bool fix_tree = use_lap && bound != nullptr;  // LAP write-back + sum-tree repair beside the backward pass
    if (fix_tree) {
      B200_TRY(fork.link(M, sP));
      B200_TRY(bound->scatter_and_fix(prio_dbg.p, act_stride, B, sP));
    }
    B200_TRY(fork.link(sE, M));
    B200_TRY(MB[variant]->launch(M));
    if (actor) {
      B200_TRY(launch_actor_loss(M));
      B200_TRY(MC->launch(M));
    }
    if (fix_tree) B200_TRY(fork.link(sP, M));
    return kOk;
  }
#endif

#define GL(tbl, st) B200_TRY(gemm_table_launch(tbl, G, st))
#define RL(tbl, st) B200_TRY(row_table_launch(tbl, G, st))
#define RLF(tbl, st) do { if (!fuse_norm) RL(tbl, st); } while (0)  // forward AvgL1Norm: fused into its producer on the tensor-core path

  int launch_enc_loss(cudaStream_t s) {
    dim3 grid((unsigned)ceil_div(B, 8), (unsigned)G), block(256);
    const Mat zn = e_z.rows_from(B, B);
    B200_LAUNCH(td7_encoder_loss_kernel, grid, block, 0, s, e_zsa.p, e_zsa.ld, zn.p, zn.ld, d_zsa.p, d_zsa.ld, B, Z, act_stride, enc_partial, enc_counter, d_info,
                info_cols, 0, d_step, (long long)kInfoRing * info_cols);
    B200_CHECK_LAUNCH();
    return kOk;
  }
  int launch_critic_loss(cudaStream_t s) {
    CriticLossArgs a;
    memset(&a, 0, sizeof(a));
    a.q1 = qv[0].p; a.q2 = qv[1].p; a.qt1 = qtv[0].p; a.qt2 = qtv[1].p;
    a.reward = rew.p; a.notdone = nd.p; a.dq1 = dq[0].p; a.dq2 = dq[1].p;
    a.y_out = y_dbg.p; a.prio_out = prio_dbg.p;
    a.range = d_range;
    a.info = d_info; a.info_stride = info_cols; a.info_col = 1; a.info_gs = (long long)kInfoRing * info_cols;
    a.st = d_step;
    a.gamma = (float)cfg.discount;
    a.B = B; a.use_lap = use_lap ? 1 : 0; a.huber_sum = 1; a.gs = act_stride;
    // (the LAP write-back itself -- last-writer-wins scatter, running max, sum-tree repair -- is a separate node beside the
    // backward pass: Replay::scatter_and_fix on the priorities left in prio_dbg)
    a.idx_gs = bound ? (long long)bound->max_batch : 0;
    B200_LAUNCH(critic_loss_kernel, dim3((unsigned)G), dim3((unsigned)round_up(B, 32)), sizeof(ScatterTable), s, a);
    B200_CHECK_LAUNCH();
    return kOk;
  }
  int launch_actor_loss(cudaStream_t s) {
    B200_LAUNCH(td7_actor_loss_kernel, dim3((unsigned)G), dim3((unsigned)round_up(B, 32)), 0, s, qpi[0].p, qpi[1].p, B, act_stride, d_info, info_cols, 2,
                (long long)kInfoRing * info_cols, d_step);
    B200_CHECK_LAUNCH();
    return kOk;
  }

  // variant: bit0 = actor step
  int issue_step(int variant) override {
#ifndef B200_EMU
    if (use_mega() && MC) return issue_step_mega(variant);  // (recorded at creation when the option was on)
#endif
    const bool actor = (variant & 1) != 0;
    cudaStream_t M = stream, sE = fork.side[0], sA = fork.side[1], sN = fork.side[2], sP = fork.side[3];
    // counters / Adam scalars / smoothing noise: off the critical path, joined where first needed
    B200_TRY(fork.link(M, sP));
    B200_TRY(launch_prep(actor ? 0x7u : 0x6u, sP));  // optimizers: 0 policy, 1 critics, 2 encoder
    B200_TRY(launch_gauss(noise, 0, sP));
    // ---- critic chain: its first launch (4 x 64 tiles) gets the machine to itself; the encoder branch, which has slack,
    // forks off behind it instead of competing for the first wave of CTAs
    GL(C[0], M);
    // ---- encoder branch
    B200_TRY(fork.link(M, sE));
    GL(E[0], sE); GL(E[1], sE); GL(E[2], sE); RLF(En[0], sE);
    GL(E[3], sE); GL(E[4], sE); GL(E[5], sE);
    B200_TRY(fork.link(sP, sE));  // info slot (loss) and Adam scalars (wgrad)
    B200_TRY(launch_enc_loss(sE));
    GL(E[6], sE); GL(We[5], sE); GL(E[7], sE); GL(We[4], sE); GL(E[8], sE); GL(We[3], sE); RL(En[1], sE);
    GL(E[9], sE); GL(We[2], sE); GL(E[10], sE); GL(We[1], sE); GL(We[0], sE);
    B200_TRY(fork.link(M, sN));
    RLF(CnQ, sN);
    GL(C2v[actor ? 1 : 0], M);
    B200_TRY(fork.link(M, sN));
    RLF(CnP[actor ? 1 : 0], sN);
    GL(C[3], M); RLF(Cn[2], M);
    B200_TRY(fork.link(sN, M));
    if (actor) {  // hoisted part of the actor step: needs zs_f, pb0n, the (unchanged) policy and fixed encoder
      B200_TRY(fork.link(M, sA));
      GL(Aa[0], sA); GL(Aa[1], sA); GL(Aa[2], sA); GL(Aa[3], sA); GL(Aa[4], sA); GL(Aa[5], sA);
    }
    GL(C[4], M); GL(C[5], M);
    B200_TRY(fork.link(sP, M));  // smoothing noise for the policy's last layer; counters for the loss / Adam
    GL(C[6], M);
    if (split_q) {  // (sP is idle between the noise and the LAP write-back)
      B200_TRY(fork.link(M, sP));
      GL(Cq[0], sP); GL(Cq[1], sP); GL(Cq[2], sP);
    }
    GL(C[7], M);
    B200_TRY(fork.link(M, sN));
    RLF(Cn[3], sN);
    GL(C[8], M); GL(C[9], M);
    B200_TRY(fork.link(sN, M));
    GL(C[10], M); GL(C[11], M); GL(C[12], M);
    if (split_q) B200_TRY(fork.link(sP, M));
    B200_TRY(launch_critic_loss(M));
    const bool fix_tree = use_lap && bound != nullptr;  // LAP write-back + sum-tree repair, off the chain
    if (fix_tree) {
      B200_TRY(fork.link(M, sP));
      B200_TRY(bound->scatter_and_fix(prio_dbg.p, act_stride, B, sP));
    }
    GL(C[13], M);
    B200_TRY(fork.link(M, sN)); GL(Wq[0], sN);  // q3 weights were last read by C[13]
    GL(C[14], M);
    B200_TRY(fork.link(M, sN)); GL(Wq[1], sN);
    GL(C[15], M);
    B200_TRY(fork.link(M, sN)); GL(Wq[2], sN);
    RL(critic_norm_bwd, M); GL(Wq[3], M);
    if (!actor) B200_TRY(fork.link(sN, M));
    if (actor) {
      B200_TRY(fork.link(sA, M));
      GL(Aa[6], M); RLF(An[0], M);   // q01 was just updated on this stream; the deeper critic layers are still being
      B200_TRY(fork.link(sN, M));   // updated on sN: joined only where the new q1 weights are first read
      GL(Aa[7], M); GL(Aa[8], M);
      B200_TRY(fork.link(M, sN));
      GL(Aa[9], sN);
      B200_TRY(launch_actor_loss(sN));
      GL(Aa[10], M); GL(Aa[11], M); GL(Aa[12], M);
      B200_TRY(fork.link(sN, M));  // (re-joins the loss branch; sN is reused below)
      B200_TRY(fork.link(M, sN));
      RL(An[1], sN);
      GL(Aa[13], M); GL(Aa[14], M);
      B200_TRY(fork.link(sN, M));
      GL(Aa[15], M); GL(Aa[16], M);
      B200_TRY(fork.link(M, sN)); GL(Wp[0], sN);  // l3 weights were last read by Aa[16]
      GL(Aa[17], M);
      B200_TRY(fork.link(M, sN)); GL(Wp[1], sN);
      GL(Aa[18], M);
      B200_TRY(fork.link(M, sN)); GL(Wp[2], sN);
      RL(An[2], M); GL(Wp[3], M);
      B200_TRY(fork.link(sN, M));
    }
    B200_TRY(fork.link(sE, M));
    if (fix_tree) B200_TRY(fork.link(sP, M));
    return kOk;
  }

  // every target_update_rate calls (td7.py:325-331, 279-285)
  int hard_update() {
    B200_TRY(copy_net(qt[0], q[0], exec));
    B200_TRY(copy_net(qt[1], q[1], exec));
    B200_TRY(copy_net(fet, fe, exec));
    B200_TRY(copy_net(fe, enc, exec));
    B200_LAUNCH(td7_range_sync_kernel, dim3((unsigned)ceil_div(G, 128)), dim3(128), 0, exec, d_range, G);
    B200_CHECK_LAUNCH();
    if (use_lap && bound) B200_TRY(bound->reset_max_priority(exec));
    return kOk;
  }

  int bind_replay(Replay* r) {
    if (r == bound) return kOk;
    if (r) {
      B200_REQUIRE(r->groups == G, "replay buffer and agent must stack the same number of seeds");
      B200_REQUIRE(r->S == S && r->A == A, "replay buffer and agent dims differ");
    }
    bound = r;
    drop_graphs();  // the priority write-back pointers are baked into the captured graphs
    return kOk;
  }

  // one train_ops call on the batch already present in the fixed input buffers
  int update(Replay* r, const float* noise_dev) {
    B200_TRY(bind_replay(use_lap ? r : bound));
    if (use_lap) B200_REQUIRE(bound != nullptr && bound->prioritized, "TD7(use_lap=True) needs a LAP replay buffer");
    const int inject = noise_dev ? 1 : 0;
    if (inject != last_inject) {
      B200_CHECK_CUDA(cudaMemcpyAsync(d_flags, &inject_vals[inject], sizeof(int), cudaMemcpyHostToDevice, exec));
      last_inject = inject;
    }
    if (noise_dev)  // caller layout: [G][B][A] contiguous
      for (int g = 0; g < G; ++g)
        B200_CHECK_CUDA(cudaMemcpyAsync(noise.p + (long long)g * act_stride, noise_dev + (long long)g * B * A, (size_t)B * A * sizeof(float), cudaMemcpyDeviceToDevice, exec));
    n_calls += 1;
    const bool actor = (n_calls % cfg.policy_freq) == 0;
    B200_TRY(run_step(actor ? 1 : 0));
    if (n_calls % cfg.target_update_rate == 0) B200_TRY(hard_update());
    return kOk;
  }
  int last_inject = 0;
  const int inject_vals[2] = {0, 1};

#undef GL
#undef RL
#undef RLF
};

}  // namespace b200rl
