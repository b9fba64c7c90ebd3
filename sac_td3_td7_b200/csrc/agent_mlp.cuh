// agent_mlp.cuh -- TD3 (rl/agent/td3.py:206-242) and SAC (rl/agent/sac.py:251-295) updates over the
// reference's 256-256 MLP actor / twin critics (rl/nn/mlp.py), as replayable launch programs.
//
// The policy forward at s (actor step) does not depend on the critic update, so it rides in the
// same grouped launches as the target-policy forward at s' and the current critics.
#pragma once
#include "agent_base.cuh"

namespace b200rl {

enum MlpAlgo : int { kAlgoTD3 = 0, kAlgoSAC = 1 };

struct MlpConfig {
  int algo;
  int S, A, H, B, G;
  double discount, policy_lr, critic_lr, tau;
  int policy_freq;                          // TD3
  double target_policy_noise, noise_clip;   // TD3
  double min_log_std, max_log_std, tmp;     // SAC (tmp < 0 -> automatic temperature)
  int use_lap, keep_grads;
};

struct MlpAgent : AgentBase {
  MlpConfig cfg;
  bool sac = false, auto_tmp = false;
  int Aout = 0;
  Net pol, q[2], qt[2];
  float *la_p = nullptr, *la_m = nullptr, *la_v = nullptr, *la_g = nullptr;  // SAC log-temperature [G]
  NormChunk* d_chunks = nullptr;  // the policy's gradient tensors, cut into chunks (norm/policy)
  double* d_norm_partial = nullptr;
  unsigned* d_norm_counter = nullptr;
  int nchunks = 0, nslices = 0;

  Mat noise, noise2;
  Mat h1, h2, pout_n, a_next, logp_n, c1[2], c2[2], qv[2], tc1[2], tc2[2], qtv[2], dq[2], d_c2[2], d_c1[2];
  Mat g1, g2, pout, a_pi, logp, ac1[2], ac2[2], qpi[2], dqa[2], dlogp, d_ac2[2], d_ac1[2], d_api, d_pout, d_g2, d_g1;
  Mat y_dbg, prio_dbg;

  GemmTable T[2][3], Tq[3], Td[2], TwL[3], U[3], Ud[3], Up[2], UwL[3];  // TwL / UwL: critic / policy weight gradients (+Adam) per layer

  void layout_acts() {
    act_cursor = 0;
    ss = alloc_act(2 * B, S, true); act_in = alloc_act(B, A, true); rew = alloc_act(B, 1); nd = alloc_act(B, 1);
    noise = alloc_act(B, A); noise2 = alloc_act(B, A);
    h1 = alloc_act(B, H); h2 = alloc_act(B, H); pout_n = alloc_act(B, Aout); a_next = alloc_act(B, A, true); logp_n = alloc_act(B, 1);
    g1 = alloc_act(B, H); g2 = alloc_act(B, H); pout = alloc_act(B, Aout); a_pi = alloc_act(B, A, true); logp = alloc_act(B, 1);
    dlogp = alloc_act(B, 1); d_api = alloc_act(B, A); d_pout = alloc_act(B, Aout); d_g2 = alloc_act(B, H); d_g1 = alloc_act(B, H);
    for (int k = 0; k < 2; ++k) {
      c1[k] = alloc_act(B, H); c2[k] = alloc_act(B, H); qv[k] = alloc_act(B, 1);
      tc1[k] = alloc_act(B, H); tc2[k] = alloc_act(B, H); qtv[k] = alloc_act(B, 1);
      dq[k] = alloc_act(B, 1); d_c2[k] = alloc_act(B, H); d_c1[k] = alloc_act(B, H);
      ac1[k] = alloc_act(B, H); ac2[k] = alloc_act(B, H); qpi[k] = alloc_act(B, 1);
      dqa[k] = alloc_act(B, 1); d_ac2[k] = alloc_act(B, H); d_ac1[k] = alloc_act(B, H);
    }
    y_dbg = alloc_act(B, 1); prio_dbg = alloc_act(B, 1);
  }

  int init(const MlpConfig& c) {
    cfg = c;
    sac = (c.algo == kAlgoSAC);
    auto_tmp = sac && c.tmp < 0.0;
    S = c.S; A = c.A; H = c.H; Z = 0; B = c.B; G = c.G;
    Aout = sac ? 2 * A : A;
    use_lap = c.use_lap != 0; keep_grads = c.keep_grads != 0;
    B200_REQUIRE(S > 0 && A > 0 && H > 0 && G > 0, "mlp agent: dims must be positive");
    B200_REQUIRE(B > 0 && B <= 1024, "mlp agent: batch size must be in 1..1024");
    B200_REQUIRE(!(sac && use_lap), "SAC with LAP is dead code in the reference (sac.py:200-207) and is not provided");
    OptimHyper hy[3] = {{c.policy_lr, 0.9, 0.999, 1e-8}, {c.critic_lr, 0.9, 0.999, 1e-8}, {c.policy_lr, 0.9, 0.999, 1e-8}};  // policy, critics, temperature
    B200_TRY(init_common(3, sac ? 6 : 3, hy));
    const bool kg = keep_grads;
    keep_grads = kg || !sac;  // TD3 logs norm/policy from the stored policy gradient
    B200_TRY(alloc_net(&pol, make_layout(kNetMlp, S, A, H, 0, S, Aout), true, d_adam + 0));
    keep_grads = kg;
    for (int k = 0; k < 2; ++k) {
      B200_TRY(alloc_net(&q[k], make_layout(kNetMlp, S, A, H, 0, S + A, 1, S), true, d_adam + 1));
      B200_TRY(alloc_net(&qt[k], make_layout(kNetMlp, S, A, H, 0, S + A, 1, S), false, nullptr));
    }
    if (sac) {
      B200_TRY(arena.alloc(&la_p, (size_t)G)); B200_TRY(arena.alloc(&la_m, (size_t)G));
      B200_TRY(arena.alloc(&la_v, (size_t)G)); B200_TRY(arena.alloc(&la_g, (size_t)G));
    } else {
      std::vector<TensorSlice> sl;
      for (const LayerDesc& d : pol.L.layers) {
        sl.push_back(TensorSlice{d.w_off, (long long)d.out * d.ld});
        sl.push_back(TensorSlice{d.b_off, (long long)d.out});
      }
      nslices = (int)sl.size();
      std::vector<NormChunk> ch;
      for (int t = 0; t < nslices; ++t)
        for (long long o = 0; o < sl[t].numel; o += kNormChunk) ch.push_back(NormChunk{sl[t].offset + o, (int)std::min<long long>(kNormChunk, sl[t].numel - o), t});
      nchunks = (int)ch.size();
      B200_TRY(arena.alloc(&d_chunks, ch.size()));
      B200_CHECK_CUDA(cudaMemcpy(d_chunks, ch.data(), ch.size() * sizeof(NormChunk), cudaMemcpyHostToDevice));
      B200_TRY(arena.alloc(&d_norm_partial, (size_t)G * nchunks));
      B200_TRY(arena.alloc(&d_norm_counter, (size_t)G));
    }
    layout_acts();
    act_stride = act_cursor;
    B200_TRY(arena.alloc(&act_base, (size_t)G * act_stride));
    layout_acts();
    B200_TRY(build_tables());
    B200_TRY(sync());
    B200_CHECK_CUDA(cudaDeviceSynchronize());  // (uploads above went through the legacy default stream; the engine runs on non-blocking ones)
    return kOk;
  }

  template <class V>
  int up(GemmTable& t, int mode, V probs) {
    std::vector<GemmProblem> v(probs);
    return gemm_table_upload(&t, mode, v, &arena, G);
  }

  int build_tables() {
    using PV = std::vector<GemmProblem>;
    const Mat s_cur = ss.rows_from(0, B), s_nxt = ss.rows_from(B, B);
    const int act_last = sac ? kActNone : kActTanh;
    for (int v = 0; v < 2; ++v) {  // v=1: also the policy forward at s (actor step)
      PV t0{fwd_problem(pol, 0, {s_nxt}, h1, kActRelu), fwd_problem(q[0], 0, {s_cur, act_in}, c1[0], kActRelu), fwd_problem(q[1], 0, {s_cur, act_in}, c1[1], kActRelu)};
      PV t1{fwd_problem(pol, 1, {h1}, h2, kActRelu), fwd_problem(q[0], 1, {c1[0]}, c2[0], kActRelu), fwd_problem(q[1], 1, {c1[1]}, c2[1], kActRelu)};
      GemmProblem pn = sac ? fwd_problem(pol, 2, {h2}, pout_n, kActNone) : fwd_problem(pol, 2, {h2}, a_next, kActTanhNoise);
      if (!sac) {
        pn.aux = noise.p; pn.aux_ld = noise.ld; pn.aux_gs = noise.gs;
        pn.noise_sigma = (float)cfg.target_policy_noise; pn.noise_clip = (float)cfg.noise_clip;
      }
      PV t2{pn, fwd_problem(q[0], 2, {c2[0]}, qv[0], kActNone), fwd_problem(q[1], 2, {c2[1]}, qv[1], kActNone)};
      if (v == 1) {
        t0.push_back(fwd_problem(pol, 0, {s_cur}, g1, kActRelu));
        t1.push_back(fwd_problem(pol, 1, {g1}, g2, kActRelu));
        t2.push_back(sac ? fwd_problem(pol, 2, {g2}, pout, kActNone) : fwd_problem(pol, 2, {g2}, a_pi, act_last));
      }
      B200_TRY(up(T[v][0], kFwd, t0)); B200_TRY(up(T[v][1], kFwd, t1)); B200_TRY(up(T[v][2], kFwd, t2));
    }
    B200_TRY(up(Tq[0], kFwd, PV{fwd_problem(qt[0], 0, {s_nxt, a_next}, tc1[0], kActRelu), fwd_problem(qt[1], 0, {s_nxt, a_next}, tc1[1], kActRelu)}));
    B200_TRY(up(Tq[1], kFwd, PV{fwd_problem(qt[0], 1, {tc1[0]}, tc2[0], kActRelu), fwd_problem(qt[1], 1, {tc1[1]}, tc2[1], kActRelu)}));
    B200_TRY(up(Tq[2], kFwd, PV{fwd_problem(qt[0], 2, {tc2[0]}, qtv[0], kActNone), fwd_problem(qt[1], 2, {tc2[1]}, qtv[1], kActNone)}));
    B200_TRY(up(Td[0], kDgrad, PV{dgrad_problem({{&q[0], 2, dq[0], 0}}, H, d_c2[0], &c2[0], kActRelu), dgrad_problem({{&q[1], 2, dq[1], 0}}, H, d_c2[1], &c2[1], kActRelu)}));
    B200_TRY(up(Td[1], kDgrad, PV{dgrad_problem({{&q[0], 1, d_c2[0], 0}}, H, d_c1[0], &c1[0], kActRelu), dgrad_problem({{&q[1], 1, d_c2[1], 0}}, H, d_c1[1], &c1[1], kActRelu)}));
    {
      // one launch per layer: layer l's gradient (and its in-place Adam step) starts as soon as the dgrad that READS layer
      // l's weights has finished
      PV w2, w1, w0;
      for (int k = 0; k < 2; ++k) {
        wgrad_problems(w2, q[k], 2, dq[k], {c2[k]}, true);
        wgrad_problems(w1, q[k], 1, d_c2[k], {c1[k]}, true);
        wgrad_problems(w0, q[k], 0, d_c1[k], {s_cur, act_in}, true);
      }
      B200_TRY(up(TwL[2], kWgrad, w2)); B200_TRY(up(TwL[1], kWgrad, w1)); B200_TRY(up(TwL[0], kWgrad, w0));
    }
    B200_TRY(up(U[0], kFwd, PV{fwd_problem(q[0], 0, {s_cur, a_pi}, ac1[0], kActRelu), fwd_problem(q[1], 0, {s_cur, a_pi}, ac1[1], kActRelu)}));
    B200_TRY(up(U[1], kFwd, PV{fwd_problem(q[0], 1, {ac1[0]}, ac2[0], kActRelu), fwd_problem(q[1], 1, {ac1[1]}, ac2[1], kActRelu)}));
    B200_TRY(up(U[2], kFwd, PV{fwd_problem(q[0], 2, {ac2[0]}, qpi[0], kActNone), fwd_problem(q[1], 2, {ac2[1]}, qpi[1], kActNone)}));
    B200_TRY(up(Ud[0], kDgrad, PV{dgrad_problem({{&q[0], 2, dqa[0], 0}}, H, d_ac2[0], &ac2[0], kActRelu), dgrad_problem({{&q[1], 2, dqa[1], 0}}, H, d_ac2[1], &ac2[1], kActRelu)}));
    B200_TRY(up(Ud[1], kDgrad, PV{dgrad_problem({{&q[0], 1, d_ac2[0], 0}}, H, d_ac1[0], &ac1[0], kActRelu), dgrad_problem({{&q[1], 1, d_ac2[1], 0}}, H, d_ac1[1], &ac1[1], kActRelu)}));
    // d a = dQ1/da + dQ2/da (action columns of the first critic layer); TD3: times tanh'
    B200_TRY(up(Ud[2], kDgrad, PV{dgrad_problem({{&q[0], 0, d_ac1[0], S}, {&q[1], 0, d_ac1[1], S}}, A, d_api, sac ? nullptr : &a_pi, kActTanh)}));
    const Mat& d_last = sac ? d_pout : d_api;
    B200_TRY(up(Up[0], kDgrad, PV{dgrad_problem({{&pol, 2, d_last, 0}}, H, d_g2, &g2, kActRelu)}));
    B200_TRY(up(Up[1], kDgrad, PV{dgrad_problem({{&pol, 1, d_g2, 0}}, H, d_g1, &g1, kActRelu)}));
    {
      PV w2, w1, w0;
      wgrad_problems(w2, pol, 2, d_last, {g2}, true);
      wgrad_problems(w1, pol, 1, d_g2, {g1}, true);
      wgrad_problems(w0, pol, 0, d_g1, {s_cur}, true);
      B200_TRY(up(UwL[2], kWgrad, w2)); B200_TRY(up(UwL[1], kWgrad, w1)); B200_TRY(up(UwL[0], kWgrad, w0));
    }
    built = true;
    return kOk;
  }

#define GL(tbl, st) B200_TRY(gemm_table_launch(tbl, G, st))

  int launch_critic_loss(cudaStream_t s) {
    CriticLossArgs a;
    memset(&a, 0, sizeof(a));
    a.q1 = qv[0].p; a.q2 = qv[1].p; a.qt1 = qtv[0].p; a.qt2 = qtv[1].p;
    a.reward = rew.p; a.notdone = nd.p; a.dq1 = dq[0].p; a.dq2 = dq[1].p;
    a.y_out = y_dbg.p; a.prio_out = prio_dbg.p;
    if (sac) {
      a.logp_next = logp_n.p;
      a.log_alpha = auto_tmp ? la_p : nullptr;
      a.alpha_const = (float)cfg.tmp;
    }
    a.info = d_info; a.info_stride = info_cols; a.info_col = 0; a.info_gs = (long long)kInfoRing * info_cols;
    a.st = d_step;
    a.gamma = (float)cfg.discount;
    a.B = B; a.use_lap = use_lap ? 1 : 0; a.gs = act_stride;
    // (the LAP write-back itself -- last-writer-wins scatter, running max, sum-tree repair -- is a separate node beside the
    // backward pass: Replay::scatter_and_fix on the priorities left in prio_dbg)
    B200_LAUNCH(critic_loss_kernel, dim3((unsigned)G), dim3((unsigned)round_up(B, 32)), sizeof(ScatterTable), s, a);
    B200_CHECK_LAUNCH();
    return kOk;
  }
  int launch_actor_loss(cudaStream_t s) {
    ActorLossArgs a;
    memset(&a, 0, sizeof(a));
    a.q1 = qpi[0].p; a.q2 = qpi[1].p; a.dq1 = dqa[0].p; a.dq2 = dqa[1].p;
    a.info = d_info; a.info_stride = info_cols; a.info_gs = (long long)kInfoRing * info_cols; a.st = d_step;
    a.B = B; a.gs = act_stride;
    a.col_tmp_loss = a.col_entropy = a.col_alpha = a.col_tmp_grad = -1;
    if (sac) {
      a.logp = logp.p; a.dlogp = dlogp.p;
      a.log_alpha = auto_tmp ? la_p : nullptr; a.alpha_const = (float)cfg.tmp;
      a.target_entropy = -(float)A;  // sac.py:67
      a.grad_log_alpha = la_g;
      a.col_policy = 3; a.col_entropy = 5;
      if (auto_tmp) { a.col_alpha = 1; a.col_tmp_grad = 2; a.col_tmp_loss = 4; }
    } else {
      a.col_policy = 1;
    }
    B200_LAUNCH(minq_actor_loss_kernel, dim3((unsigned)G), dim3((unsigned)round_up(B, 32)), 0, s, a);
    B200_CHECK_LAUNCH();
    return kOk;
  }
  int launch_sample_fwd(const Mat& out, const Mat& eps, const Mat& a, const Mat& lp, cudaStream_t s) {
    dim3 grid((unsigned)ceil_div(B, 8), (unsigned)G), block(256);
    B200_LAUNCH(sac_sample_fwd_kernel, grid, block, 0, s, out.p, out.ld, eps.p, eps.ld, a.p, a.ld, lp.p, B, A, (float)cfg.min_log_std, (float)cfg.max_log_std, act_stride,
                act_stride, act_stride, act_stride);
    B200_CHECK_LAUNCH();
    return kOk;
  }
  int launch_polyak(Net& tgt, const Net& src, cudaStream_t s) {
    const long long n = (long long)G * src.L.numel;
    B200_LAUNCH(polyak_kernel, dim3((unsigned)launch_elementwise_grid(n)), dim3(256), 0, s, tgt.p, src.p, n, (float)cfg.tau, (float)(1.0 - cfg.tau));
    B200_CHECK_LAUNCH();
    return kOk;
  }

  // variant bit0: actor step (SAC: always)
  int issue_step(int variant) override {
    const bool actor = sac || (variant & 1);
    cudaStream_t M = stream, sN = fork.side[0], sP = fork.side[3];
    unsigned mask = 0x2u | (actor ? 0x1u : 0u) | ((sac && auto_tmp) ? 0x4u : 0u);
    B200_TRY(fork.link(M, sP));  // counters / Adam scalars / noise: off the critical path
    B200_TRY(launch_prep(mask, sP));
    B200_TRY(launch_gauss(noise, 0, sP));
    if (sac) B200_TRY(launch_gauss(noise2, 1, sP));
    const int v = actor ? 1 : 0;
    GL(T[v][0], M); GL(T[v][1], M);
    B200_TRY(fork.link(sP, M));
    GL(T[v][2], M);
    if (sac) {
      B200_TRY(launch_sample_fwd(pout_n, noise, a_next, logp_n, M));
      B200_TRY(fork.link(M, sN));
      B200_TRY(launch_sample_fwd(pout, noise2, a_pi, logp, sN));
    }
    GL(Tq[0], M); GL(Tq[1], M); GL(Tq[2], M);
    B200_TRY(launch_critic_loss(M));
    const bool fix_tree = use_lap && bound != nullptr;  // LAP write-back + sum-tree repair, off the chain
    if (fix_tree) {
      B200_TRY(fork.link(M, sP));
      B200_TRY(bound->scatter_and_fix(prio_dbg.p, act_stride, B, sP));
    }
    cudaStream_t sQ = fork.side[1], sW = fork.side[2];
    GL(Td[0], M);
    B200_TRY(fork.link(M, sW)); GL(TwL[2], sW);  // layer-2 weights were last read by Td[0]
    GL(Td[1], M);
    B200_TRY(fork.link(M, sW)); GL(TwL[1], sW);
    GL(TwL[0], M);
    B200_TRY(fork.link(sW, M));
    // td3.py:239 / sac.py:293 -- target critics <- Polyak(critics): TD3 only on actor steps, SAC every call.  The critics
    // are final here, so the two target updates leave the chain and run beside the actor step.
    const bool polyak_now = sac || actor;
    if (polyak_now) {
      B200_TRY(fork.link(M, sQ));
      B200_TRY(launch_polyak(qt[0], q[0], sQ));
      B200_TRY(launch_polyak(qt[1], q[1], sQ));
    }
    if (sac) B200_TRY(fork.link(sN, M));
    if (actor) {
      GL(U[0], M); GL(U[1], M); GL(U[2], M);
      B200_TRY(launch_actor_loss(M));
      GL(Ud[0], M); GL(Ud[1], M); GL(Ud[2], M);
      if (sac) {
        dim3 grid((unsigned)ceil_div(B, 8), (unsigned)G), block(32, 8);
        B200_LAUNCH(sac_sample_bwd_kernel, grid, block, 0, M, pout.p, pout.ld, noise2.p, noise2.ld, a_pi.p, a_pi.ld, d_api.p, d_api.ld, dlogp.p, d_pout.p, d_pout.ld, B, A,
                    (float)cfg.min_log_std, (float)cfg.max_log_std, act_stride, act_stride, act_stride, act_stride);
        B200_CHECK_LAUNCH();
      }
      GL(Up[0], M);
      B200_TRY(fork.link(M, sW)); GL(UwL[2], sW);  // layer-2 weights were last read by Up[0]
      GL(Up[1], M);
      B200_TRY(fork.link(M, sW)); GL(UwL[1], sW);
      GL(UwL[0], M);
      B200_TRY(fork.link(sW, M));
      if (!sac) {
        B200_TRY(fork.link(M, sN));
        B200_LAUNCH(grad_norm_sum_kernel, dim3((unsigned)nchunks, (unsigned)G), dim3(256), 0, sN, pol.grad, pol.gs, d_chunks, nchunks, nslices, d_norm_partial, d_norm_counter,
                    d_info, info_cols, 2,
                    (long long)kInfoRing * info_cols, d_step);
        B200_CHECK_LAUNCH();
        B200_TRY(fork.link(sN, M));
      }
      if (sac && auto_tmp) {
        B200_LAUNCH(adam_blob_kernel, dim3(1), dim3(256), 0, M, la_p, la_m, la_v, la_g, (long long)G, d_adam + 2);
        B200_CHECK_LAUNCH();
      }
    }
    if (polyak_now) {
      if (!sac) B200_TRY(launch_polyak(pol, pol, M));  // target_policy aliases policy (td3.py:86): after its Adam step
      B200_TRY(fork.link(sQ, M));
    }
    if (fix_tree) B200_TRY(fork.link(sP, M));
    return kOk;
  }
  int bind_replay(Replay* r) {
    if (r == bound) return kOk;
    if (r) {
      B200_REQUIRE(r->groups == G, "replay buffer and agent must stack the same number of seeds");
      B200_REQUIRE(r->S == S && r->A == A, "replay buffer and agent dims differ");
    }
    bound = r;
    drop_graphs();
    return kOk;
  }

  int last_inject = 0;
  const int inject_vals[2] = {0, 1};
  // noise_dev: TD3 [G][B][A]; SAC [2][G][B][A] (critic draw, then actor draw)
  int update(Replay* r, const float* noise_dev) {
    B200_TRY(bind_replay(use_lap ? r : bound));
    if (use_lap) B200_REQUIRE(bound != nullptr && bound->prioritized, "use_lap=True needs a LAP replay buffer");
    const int inject = noise_dev ? 1 : 0;
    if (inject != last_inject) {
      B200_CHECK_CUDA(cudaMemcpyAsync(d_flags, &inject_vals[inject], sizeof(int), cudaMemcpyHostToDevice, exec));
      last_inject = inject;
    }
    if (noise_dev) {
      const long long n = (long long)B * A;
      for (int g = 0; g < G; ++g) {
        B200_CHECK_CUDA(cudaMemcpyAsync(noise.p + (long long)g * act_stride, noise_dev + g * n, n * sizeof(float), cudaMemcpyDeviceToDevice, exec));
        if (sac)
          B200_CHECK_CUDA(cudaMemcpyAsync(noise2.p + (long long)g * act_stride, noise_dev + (G + g) * n, n * sizeof(float), cudaMemcpyDeviceToDevice, exec));
      }
    }
    const bool actor = sac || (n_calls % cfg.policy_freq) == 0;
    n_calls += 1;
    B200_TRY(run_step(actor ? 1 : 0));
    return kOk;
  }
#undef GL
};

}  // namespace b200rl
