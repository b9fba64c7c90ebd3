// api.cuh -- extern "C" entry points declared in include/b200rl.h.
#pragma once
#include <atomic>
#include <chrono>
#include <cstddef>
#include <memory>

#include "../../include/b200rl.h"
#include "act.cuh"
#include "agent_mlp.cuh"
#include "agent_td7.cuh"

using namespace b200rl;

static_assert(sizeof(b200rl_gemm_problem) == sizeof(GemmProblem), "public/internal gemm problem layouts differ");
static_assert(sizeof(b200rl_gemm_seg) == sizeof(GemmSeg), "public/internal gemm segment layouts differ");
static_assert(offsetof(b200rl_gemm_problem, C) == offsetof(GemmProblem, C), "layout");
static_assert(offsetof(b200rl_gemm_problem, aux) == offsetof(GemmProblem, aux), "layout");
static_assert(offsetof(b200rl_gemm_problem, P) == offsetof(GemmProblem, P), "layout");
static_assert(offsetof(b200rl_gemm_problem, adam) == offsetof(GemmProblem, adam), "layout");
static_assert(offsetof(b200rl_gemm_problem, tiles_j) == offsetof(GemmProblem, tiles_j), "layout");
static_assert(sizeof(b200rl_adam_scalars) == sizeof(AdamScalars), "layout");

struct b200rl_replay {
  Replay r;
};

struct b200rl_agent {
  std::unique_ptr<AgentBase> base;
  TD7Agent* td7 = nullptr;
  MlpAgent* mlp = nullptr;
  std::vector<Net*> nets;
  // acting path (batch-1 forward of .sample())
  Mat o_s, ah1, ah2, au3, azs, ap0, ap0n, ap1, ap2, aout;
  GemmTable act_t[6];
  RowTable act_n[2];
  int n_act_t = 0;
  // fused acting kernel (act.cuh) + its pinned host slots: observation in, action out, sequence tag
  ActArgs act_args;
  float* h_obs = nullptr; float* d_obs_map = nullptr;
  float* h_act = nullptr; float* d_act_map = nullptr;
  unsigned long long* h_tag = nullptr; unsigned long long* d_tag_map = nullptr;
  unsigned* d_done = nullptr;
  unsigned long long act_seq = 0;
  int act_smem = 0;
  ~b200rl_agent() {
    if (h_obs) cudaFreeHost(h_obs);
    if (h_act) cudaFreeHost(h_act);
    if (h_tag) cudaFreeHost(h_tag);
  }
};

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static int ensure_device() {
#ifndef B200_EMU
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_last_error(std::string("no CUDA device available: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                   " (libb200rl has no CPU fallback)");
    return kErrNoDevice;
  }
#endif
  return kOk;
}

extern "C" {

const char* b200rl_last_error(void) { return last_error_ref().c_str(); }
int b200rl_version(void) { return 100; }
int b200rl_is_cuda_build(void) {
#ifdef B200_EMU
  return 0;
#else
  return 1;
#endif
}

int b200rl_set_mma(int on) {
  const int prev = use_mma() ? 1 : 0;
  mma_override() = on < 0 ? -1 : (on != 0);
  return prev;
}

int b200rl_set_mega(int on) {
#ifndef B200_EMU
  const int prev = AgentBase::mega_override();
  AgentBase::mega_override() = on < 0 ? -1 : (on != 0);
  return prev;
#else
  (void)on;
  return -1;
#endif
}

int b200rl_malloc(void** out_dev, uint64_t bytes) {
  B200_TRY(ensure_device());
  B200_CHECK_CUDA(cudaMalloc(out_dev, bytes ? bytes : 16));
  return kOk;
}
int b200rl_free(void* dev) {
  B200_CHECK_CUDA(cudaFree(dev));
  return kOk;
}
int b200rl_memcpy_h2d(void* dst_dev, const void* src_host, uint64_t bytes, void* stream) {
  B200_CHECK_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}
int b200rl_memcpy_d2h(void* dst_host, const void* src_dev, uint64_t bytes, void* stream) {
  B200_CHECK_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}
int b200rl_stream_sync(void* stream) {
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}

// ============================================================================ replay
int b200rl_replay_create(int64_t capacity, int state_dim, int action_dim, int groups, int prioritized, int share_rows, b200rl_replay** out) {
  B200_REQUIRE(out != nullptr, "out is null");
  B200_TRY(ensure_device());
  auto h = std::make_unique<b200rl_replay>();
  B200_TRY(h->r.init(capacity, state_dim, action_dim, groups, prioritized != 0, share_rows != 0));
  *out = h.release();
  return kOk;
}
int b200rl_replay_destroy(b200rl_replay* h) {
  delete h;
  return kOk;
}
int b200rl_replay_row_stride(const b200rl_replay* h) { return h ? h->r.row_stride : kErrInvalidArg; }
int b200rl_replay_row_layout(const b200rl_replay* h, int* offsets5) {
  B200_REQUIRE(h && offsets5, "null argument");
  const RowLayout& L = h->r.lay;
  offsets5[0] = L.state; offsets5[1] = L.action; offsets5[2] = L.reward; offsets5[3] = L.notdone; offsets5[4] = L.next_state;
  return kOk;
}
int b200rl_replay_append(b200rl_replay* h, int group, const float* rows_host, int64_t count, void* stream) {
  B200_REQUIRE(h && rows_host, "null argument");
  return h->r.append(group, rows_host, count, cudaMemcpyHostToDevice, as_stream(stream));
}
int b200rl_replay_append_device(b200rl_replay* h, int group, const float* rows_dev, int64_t count, void* stream) {
  B200_REQUIRE(h && rows_dev, "null argument");
  return h->r.append(group, rows_dev, count, cudaMemcpyDeviceToDevice, as_stream(stream));
}
int b200rl_replay_set_action_norm(b200rl_replay* h, const float* scale_host, const float* bias_host) {
  B200_REQUIRE(h && scale_host && bias_host, "null argument");
  return h->r.set_action_norm(scale_host, bias_host);
}
static bool is_device_pointer(const void* p) {
#ifndef B200_EMU
  cudaPointerAttributes at;
  const bool dev = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice;
  (void)cudaGetLastError();
  return dev;
#else
  (void)p;
  return false;
#endif
}
int b200rl_replay_append_raw(b200rl_replay* h, int group, const float* rows, int64_t count, void* stream) {
  B200_REQUIRE(h && rows, "null argument");
  B200_REQUIRE(count <= 64, "append_raw: at most 64 rows per call");
  return h->r.append(group, rows, count, is_device_pointer(rows) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, as_stream(stream), true);
}
int b200rl_replay_append_all(b200rl_replay* h, const float* rows, int64_t count, int raw_actions, void* stream) {
  B200_REQUIRE(h && rows, "null argument");
  return h->r.append_all(rows, count, is_device_pointer(rows) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, as_stream(stream), raw_actions != 0);
}
int b200rl_replay_append_transitions(b200rl_replay* h, int group, int n, const float* obs_host, const float* action_host, const float* reward_host,
                                     const float* next_obs_host, const float* notdone_host, int raw_actions, void* stream) {
  B200_REQUIRE(h && obs_host && action_host && reward_host && next_obs_host && notdone_host, "null argument");
  B200_REQUIRE(group >= 0 && group < h->r.groups, "bad group");
  return h->r.append_transitions(group, n, obs_host, action_host, reward_host, next_obs_host, notdone_host, raw_actions != 0, as_stream(stream));
}
int b200rl_replay_sample(b200rl_replay* h, int batch, const float* u_dev, float* state_dev, float* action_dev, float* reward_dev, float* next_state_dev,
                         float* done_dev, void* stream) {
  B200_REQUIRE(h && state_dev && action_dev && reward_dev && next_state_dev && done_dev, "null argument");
  Replay::Outputs o;
  const int S = h->r.S, A = h->r.A;
  o.state = state_dev; o.ld_state = S; o.state_gs = (long long)batch * S;
  o.action = action_dev; o.ld_action = A; o.action_gs = (long long)batch * A;
  o.reward = reward_dev; o.reward_gs = batch;
  o.next_state = next_state_dev; o.ld_next = S; o.next_gs = (long long)batch * S;
  o.done = done_dev; o.done_gs = batch;
  return h->r.sample(batch, u_dev, o, as_stream(stream));
}
int b200rl_replay_last_indices(b200rl_replay* h, int64_t* idx_host, void* stream) {
  B200_REQUIRE(h && idx_host, "null argument");
  const int B = h->r.last_batch;
  for (int g = 0; g < h->r.groups; ++g)
    B200_CHECK_CUDA(cudaMemcpyAsync(idx_host + (long long)g * B, h->r.d_idx + (long long)g * h->r.max_batch, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                    as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}
int b200rl_replay_update_priority(b200rl_replay* h, const float* priority_dev, int batch, void* stream) {
  B200_REQUIRE(h && priority_dev, "null argument");
  return h->r.update_priority(priority_dev, batch, batch, as_stream(stream));
}
int b200rl_replay_reset_max_priority(b200rl_replay* h, void* stream) {
  B200_REQUIRE(h, "null argument");
  return h->r.reset_max_priority(as_stream(stream));
}
int b200rl_replay_get_state(b200rl_replay* h, int group, int64_t* ptr, int64_t* size, float* max_priority, void* stream) {
  B200_REQUIRE(h && group >= 0 && group < h->r.groups, "bad handle or group");
  if (ptr) *ptr = h->r.ptr[group];
  if (size) *size = h->r.size[group];
  if (max_priority) {
    B200_CHECK_CUDA(cudaMemcpyAsync(max_priority, h->r.d_max_priority + group, sizeof(float), cudaMemcpyDeviceToHost, as_stream(stream)));
    B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  }
  return kOk;
}
int b200rl_replay_set_cursor(b200rl_replay* h, int group, int64_t ptr, int64_t size) {
  if (h) h->r.sums_valid = false;
  B200_REQUIRE(h && group >= 0 && group < h->r.groups, "bad handle or group");
  B200_REQUIRE(ptr >= 0 && ptr < h->r.capacity && size >= 0 && size <= h->r.capacity, "cursor out of range");
  h->r.ptr[group] = ptr;
  h->r.size[group] = size;
  return kOk;
}
int b200rl_replay_set_rng(b200rl_replay* h, uint64_t seed) {
  B200_REQUIRE(h, "null argument");
  h->r.rng_seed = seed;
  h->r.rng_counter = 0;
  return kOk;
}
int b200rl_replay_read_priority(b200rl_replay* h, int group, int64_t first, int64_t count, float* out_host, void* stream) {
  B200_REQUIRE(h && out_host && h->r.prioritized, "null argument or uniform buffer");
  B200_REQUIRE(group >= 0 && group < h->r.groups && first >= 0 && first + count <= h->r.capacity, "range");
  B200_CHECK_CUDA(cudaMemcpyAsync(out_host, h->r.d_priority + (long long)group * h->r.cap_pad + first, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost,
                                  as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}
int b200rl_replay_write_priority(b200rl_replay* h, int group, int64_t first, int64_t count, const float* in_host, float max_priority, void* stream) {
  B200_REQUIRE(h && h->r.prioritized, "null argument or uniform buffer");
  B200_REQUIRE(group >= 0 && group < h->r.groups && first >= 0 && first + count <= h->r.capacity, "range");
  if (count > 0)
    B200_CHECK_CUDA(cudaMemcpyAsync(h->r.d_priority + (long long)group * h->r.cap_pad + first, in_host, (size_t)count * sizeof(float), cudaMemcpyHostToDevice,
                                    as_stream(stream)));
  h->r.sums_valid = false;  // the next sample rebuilds the sum tree
  if (max_priority >= 0.f)
    B200_CHECK_CUDA(cudaMemcpyAsync(h->r.d_max_priority + group, &max_priority, sizeof(float), cudaMemcpyHostToDevice, as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}
int b200rl_replay_read_rows(b200rl_replay* h, int group, int64_t first, int64_t count, float* out_host, void* stream) {
  B200_REQUIRE(h && out_host, "null argument");
  B200_REQUIRE(group >= 0 && group < h->r.groups && first >= 0 && first + count <= h->r.capacity, "range");
  const float* base = h->r.d_rows + (h->r.share_rows ? 0 : (long long)group * h->r.capacity * h->r.row_stride);
  B200_CHECK_CUDA(cudaMemcpyAsync(out_host, base + first * h->r.row_stride, (size_t)count * h->r.row_stride * sizeof(float), cudaMemcpyDeviceToHost, as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}

// ============================================================================ agents
static int build_act_path(b200rl_agent* h) {
  AgentBase& a = *h->base;
  using PV = std::vector<GemmProblem>;
  // small activations for the batch-1 forward live in their own arena block
  const int S = a.S, A = a.A, H = a.H;
  const int Aout = h->mlp ? h->mlp->Aout : A;
  const int Z = h->td7 ? a.Z : 0;
  const long long per = round_up(S, 4) + 4 * (long long)round_up(H, 4) + 2 * (long long)round_up(Z > 0 ? Z : 4, 4) + 2 * (long long)round_up(H, 4) + round_up(Aout, 4);
  float* base = nullptr;
  B200_TRY(a.arena.alloc(&base, (size_t)a.G * per));
  long long cur = 0;
  auto mk = [&](int cols) { Mat m; m.p = base + cur; m.rows = 1; m.cols = cols; m.ld = cols; m.gs = per; cur += round_up(cols, 4); return m; };
  h->o_s = mk(S); h->ah1 = mk(H); h->ah2 = mk(H); h->ap0 = mk(H); h->ap0n = mk(H);
  h->au3 = mk(Z > 0 ? Z : 4); h->azs = mk(Z > 0 ? Z : 4); h->ap1 = mk(H); h->ap2 = mk(H); h->aout = mk(Aout);
  auto up = [&](GemmTable& t, PV v) { return gemm_table_upload(&t, kFwd, v, &a.arena, a.G); };
  if (h->td7) {
    TD7Agent& t = *h->td7;
    B200_TRY(up(h->act_t[0], PV{fwd_problem(t.fe, 0, {h->o_s}, h->ah1, kActElu), fwd_problem(t.pol, 0, {h->o_s}, h->ap0, kActNone)}));
    B200_TRY(up(h->act_t[1], PV{fwd_problem(t.fe, 1, {h->ah1}, h->ah2, kActElu)}));
    B200_TRY(up(h->act_t[2], PV{fwd_problem(t.fe, 2, {h->ah2}, h->au3, kActNone)}));
    std::vector<RowOp> n0{make_rowop(h->ap0.p, h->ap0.ld, per, h->ap0n.p, h->ap0n.ld, per, 1, H), make_rowop(h->au3.p, h->au3.ld, per, h->azs.p, h->azs.ld, per, 1, Z)};
    B200_TRY(row_table_upload(&h->act_n[0], false, n0, &a.arena));
    B200_TRY(up(h->act_t[3], PV{fwd_problem(t.pol, 1, {h->ap0n, h->azs}, h->ap1, kActRelu)}));
    B200_TRY(up(h->act_t[4], PV{fwd_problem(t.pol, 2, {h->ap1}, h->ap2, kActRelu)}));
    B200_TRY(up(h->act_t[5], PV{fwd_problem(t.pol, 3, {h->ap2}, h->aout, kActTanh)}));
    h->n_act_t = 6;
  } else {
    MlpAgent& m = *h->mlp;
    B200_TRY(up(h->act_t[0], PV{fwd_problem(m.pol, 0, {h->o_s}, h->ah1, kActRelu)}));
    B200_TRY(up(h->act_t[1], PV{fwd_problem(m.pol, 1, {h->ah1}, h->ah2, kActRelu)}));
    B200_TRY(up(h->act_t[2], PV{fwd_problem(m.pol, 2, {h->ah2}, h->aout, m.sac ? kActNone : kActTanh)}));
    h->n_act_t = 3;
  }
  // ---- fused path (one cluster kernel per call)
  ActArgs& aa = h->act_args;
  memset(&aa, 0, sizeof(aa));
  auto layer = [](const Net& n, int l) {
    const LayerDesc& d = n.L.layers[l];
    ActLayer r;
    r.W = n.p + d.w_off; r.b = n.p + d.b_off; r.gs = n.gs; r.ld = d.ld; r.in = d.in; r.out = d.out;
    return r;
  };
  aa.S = S; aa.H = H; aa.Z = Z; aa.Aout = Aout;
  aa.vmax = (int)round_up(std::max(H, std::max(Z, 4)), 4);
  if (h->td7) {
    TD7Agent& t = *h->td7;
    aa.kind = 0;
    for (int l = 0; l < 3; ++l) aa.L[l] = layer(t.fe, l);
    for (int l = 0; l < 4; ++l) aa.L[3 + l] = layer(t.pol, l);
  } else {
    MlpAgent& m = *h->mlp;
    aa.kind = m.sac ? 2 : 1;
    for (int l = 0; l < 3; ++l) aa.L[l] = layer(m.pol, l);
  }
  aa.total_ctas = (unsigned)(a.G * kActCluster);
  // weight slices prefetched by every CTA (act.cuh): all layers behind the first phase, if they fit beside the vectors
  {
    const int nl = h->td7 ? 7 : 3;
    long long tot = 0;
    for (int l = 0; l < 7; ++l) aa.pre_off[l] = -1;
    for (int l = 1; l < nl; ++l) {
      if (h->td7 && l == 3) continue;  // l0 reads the observation in the first phase, beside zs1
      aa.pre_off[l] = (int)tot;
      tot += (long long)ceil_div(aa.L[l].out, kActCluster) * aa.L[l].ld;
    }
    // OFF by default: measured on B200 it does not pay -- back-to-back launches 15.4 us with it, 14.4 us without (TD7), 8.2
    // against 6.2 us (TD3): a phase costs ~2.7 us whether its weights wait in shared memory or come from L2, so the
    // chain is bound by the reduction / DSMEM push / cluster-barrier sequence, not by the weight reads.
    static const bool pre = [] { const char* e = getenv("B200RL_ACT_PREFETCH"); return e ? atoi(e) != 0 : false; }();
    // (only while every cluster of the launch is resident at once: with hundreds of stacked seeds the kernel runs in
    // waves, where many small CTAs per SM hide the weight reads better than one CTA holding 165 KB)
    if (!pre || a.G * kActCluster > 148 || (round_up(S, 4) + 4ll * aa.vmax + tot) * (long long)sizeof(float) > 200 * 1024) {
      for (int l = 0; l < 7; ++l) aa.pre_off[l] = -1;
      tot = 0;
    }
    aa.pre_total = (int)tot;
    if (const char* e = getenv("B200RL_ACT_DBG")) aa.dbg = atoi(e);
  }
  h->act_smem = (int)((round_up(S, 4) + 4ll * aa.vmax + aa.pre_total) * sizeof(float));
  B200_REQUIRE(h->act_smem <= 200 * 1024, "acting path: observation / hidden sizes too large for the fused kernel");
#ifndef B200_EMU
  B200_TRY(a.arena.alloc(&h->d_done, 1));
  B200_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->h_obs), (size_t)a.G * S * sizeof(float), cudaHostAllocMapped));
  B200_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->h_act), (size_t)a.G * Aout * sizeof(float), cudaHostAllocMapped));
  B200_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->h_tag), 64, cudaHostAllocMapped));
  B200_CHECK_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->d_obs_map), h->h_obs, 0));
  B200_CHECK_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->d_act_map), h->h_act, 0));
  B200_CHECK_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->d_tag_map), h->h_tag, 0));
  *h->h_tag = 0ull;
  if (h->act_smem > 48 * 1024) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(act_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->act_smem));
    B200_CHECK_CUDA(cudaFuncSetAttribute(act_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->act_smem));
  }
#endif
  return kOk;
}

#ifndef B200_EMU
static inline bool use_fused_act() {
  static const bool on = [] { const char* e = getenv("B200RL_ACT_FUSED"); return e ? atoi(e) != 0 : true; }();  // 0: the launch-table path
  return on;
}
// inline_obs != null: one agent's observation by value in the kernel parameters
static int launch_act_fused(b200rl_agent* h, const float* obs, long long obs_gs, float* out, long long out_gs, bool tag, cudaStream_t s, const float* inline_obs = nullptr) {
  ActArgs aa = h->act_args;
  aa.obs = obs; aa.obs_gs = obs_gs; aa.out = out; aa.out_gs = out_gs;
  aa.done = h->d_done;
  aa.tag = tag ? h->d_tag_map : nullptr;
  aa.tag_value = tag ? ++h->act_seq : 0ull;
  static ActObs ob;  // (parameter block; only the first S floats are read)
  if (inline_obs) {
    memcpy(ob.v, inline_obs, (size_t)h->base->S * sizeof(float));
    B200_LAUNCH(act_fused_kernel<true>, dim3((unsigned)(h->base->G * kActCluster)), dim3(kActThreads), (size_t)h->act_smem, s, aa, ob);
  } else {
    B200_LAUNCH(act_fused_kernel<false>, dim3((unsigned)(h->base->G * kActCluster)), dim3(kActThreads), (size_t)h->act_smem, s, aa, ob);
  }
  B200_CHECK_LAUNCH();
  return kOk;
}
#endif

int b200rl_td7_create(const b200rl_td7_config* c, b200rl_agent** out) {
  B200_REQUIRE(c && out, "null argument");
  B200_TRY(ensure_device());
  auto h = std::make_unique<b200rl_agent>();
  auto* t = new TD7Agent();
  h->base.reset(t);
  h->td7 = t;
  TD7Config cfg;
  cfg.S = c->state_dim; cfg.A = c->action_dim; cfg.H = c->hdim; cfg.Z = c->zs_dim; cfg.B = c->batch; cfg.G = c->groups;
  cfg.discount = c->discount; cfg.policy_lr = c->policy_lr; cfg.critic_lr = c->critic_lr;
  cfg.target_update_rate = c->target_update_rate; cfg.policy_freq = c->policy_freq;
  cfg.target_policy_noise = c->target_policy_noise; cfg.noise_clip = c->noise_clip;
  cfg.use_lap = c->use_lap; cfg.keep_grads = c->keep_grads;
  B200_REQUIRE(cfg.target_update_rate > 0 && cfg.policy_freq > 0, "update periods must be positive");
  B200_TRY(t->init(cfg));
  h->nets = {&t->enc, &t->pol, &t->q[0], &t->q[1], &t->qt[0], &t->qt[1], &t->fe, &t->fet};
  B200_TRY(build_act_path(h.get()));
  *out = h.release();
  return kOk;
}

int b200rl_mlp_create(const b200rl_mlp_config* c, b200rl_agent** out) {
  B200_REQUIRE(c && out, "null argument");
  B200_REQUIRE(c->algo == 0 || c->algo == 1, "algo must be 0 (TD3) or 1 (SAC)");
  B200_TRY(ensure_device());
  auto h = std::make_unique<b200rl_agent>();
  auto* m = new MlpAgent();
  h->base.reset(m);
  h->mlp = m;
  MlpConfig cfg;
  cfg.algo = c->algo; cfg.S = c->state_dim; cfg.A = c->action_dim; cfg.H = c->hdim; cfg.B = c->batch; cfg.G = c->groups;
  cfg.discount = c->discount; cfg.policy_lr = c->policy_lr; cfg.critic_lr = c->critic_lr; cfg.tau = c->tau;
  cfg.policy_freq = c->policy_freq > 0 ? c->policy_freq : 1;
  cfg.target_policy_noise = c->target_policy_noise; cfg.noise_clip = c->noise_clip;
  cfg.min_log_std = c->min_log_std; cfg.max_log_std = c->max_log_std; cfg.tmp = c->tmp;
  cfg.use_lap = c->use_lap; cfg.keep_grads = c->keep_grads;
  B200_TRY(m->init(cfg));
  h->nets = {&m->pol, &m->q[0], &m->q[1], &m->qt[0], &m->qt[1]};
  B200_TRY(build_act_path(h.get()));
  *out = h.release();
  return kOk;
}

int b200rl_agent_destroy(b200rl_agent* h) {
  delete h;
  return kOk;
}

int b200rl_agent_net_count(const b200rl_agent* h) { return h ? (int)h->nets.size() + ((h->mlp && h->mlp->sac) ? 1 : 0) : kErrInvalidArg; }

static bool is_temp_net(const b200rl_agent* h, int net) { return h->mlp && h->mlp->sac && net == (int)h->nets.size(); }

int b200rl_agent_net_info(const b200rl_agent* h, int net, int64_t* numel, int* nlayers) {
  B200_REQUIRE(h && net >= 0 && net < b200rl_agent_net_count(h), "bad handle or net id");
  if (is_temp_net(h, net)) {
    if (numel) *numel = 1;
    if (nlayers) *nlayers = 0;
    return kOk;
  }
  if (numel) *numel = h->nets[net]->L.numel;
  if (nlayers) *nlayers = (int)h->nets[net]->L.layers.size();
  return kOk;
}
int b200rl_agent_layer_info(const b200rl_agent* h, int net, int layer, int64_t* w_off, int64_t* b_off, int* out_dim, int* in_dim, int* ld) {
  B200_REQUIRE(h && net >= 0 && net < (int)h->nets.size(), "bad handle or net id");
  const auto& L = h->nets[net]->L.layers;
  B200_REQUIRE(layer >= 0 && layer < (int)L.size(), "bad layer");
  if (w_off) *w_off = L[layer].w_off;
  if (b_off) *b_off = L[layer].b_off;
  if (out_dim) *out_dim = L[layer].out;
  if (in_dim) *in_dim = L[layer].in;
  if (ld) *ld = L[layer].ld;
  return kOk;
}

int b200rl_agent_layer_split(const b200rl_agent* h, int net, int layer, int* split) {
  B200_REQUIRE(h && split && net >= 0 && net < (int)h->nets.size(), "bad handle, net id or null argument");
  const auto& L = h->nets[net]->L.layers;
  B200_REQUIRE(layer >= 0 && layer < (int)L.size(), "bad layer");
  *split = L[layer].split;
  return kOk;
}

static int net_blob(b200rl_agent* h, int net, int which, int group, float** ptr, long long* numel) {
  B200_REQUIRE(h && net >= 0 && net < b200rl_agent_net_count(h), "bad handle or net id");
  B200_REQUIRE(group >= -1 && group < h->base->G, "bad group");
  B200_REQUIRE(which >= 0 && which <= 3, "which must be 0..3");
  const bool all = group < 0;  // every stacked seed: the blobs of one network are contiguous [groups][numel]
  if (all) group = 0;
  if (is_temp_net(h, net)) {
    MlpAgent& m = *h->mlp;
    float* arr[4] = {m.la_p, m.la_m, m.la_v, m.la_g};
    *ptr = arr[which] + group;
    *numel = all ? h->base->G : 1;
    return kOk;
  }
  Net* n = h->nets[net];
  float* arr[4] = {n->p, n->m, n->v, n->grad};
  if (!arr[which]) {
    set_last_error("this network has no such tensor (frozen net, or gradients not kept: create with keep_grads=1)");
    return kErrState;
  }
  *ptr = arr[which] + (long long)group * n->gs;
  *numel = all ? (long long)h->base->G * n->L.numel : n->L.numel;
  return kOk;
}
int b200rl_agent_read(b200rl_agent* h, int net, int which, int group, float* out_host, void* stream) {
  float* p; long long n;
  B200_TRY(net_blob(h, net, which, group, &p, &n));
  B200_CHECK_CUDA(cudaMemcpyAsync(out_host, p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}
int b200rl_agent_write(b200rl_agent* h, int net, int which, int group, const float* in_host, void* stream) {
  float* p; long long n;
  B200_TRY(net_blob(h, net, which, group, &p, &n));
  B200_CHECK_CUDA(cudaMemcpyAsync(p, in_host, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}

int b200rl_agent_copy_from(b200rl_agent* dst, const b200rl_agent* src, int full, void* stream) {
  B200_REQUIRE(dst && src, "null argument");
  B200_REQUIRE((dst->td7 != nullptr) == (src->td7 != nullptr) && dst->nets.size() == src->nets.size(), "agents are of different kinds");
  B200_REQUIRE(dst->base->G == src->base->G, "agents stack different numbers of seeds");
  cudaStream_t s = as_stream(stream);
  for (size_t i = 0; i < dst->nets.size(); ++i) {
    Net* d = dst->nets[i];
    const Net* c = src->nets[i];
    B200_REQUIRE(d->L.numel == c->L.numel, "network shapes differ");
    const size_t bytes = (size_t)dst->base->G * d->L.numel * sizeof(float);
    B200_CHECK_CUDA(cudaMemcpyAsync(d->p, c->p, bytes, cudaMemcpyDeviceToDevice, s));
    if (full && d->m && c->m) {
      B200_CHECK_CUDA(cudaMemcpyAsync(d->m, c->m, bytes, cudaMemcpyDeviceToDevice, s));
      B200_CHECK_CUDA(cudaMemcpyAsync(d->v, c->v, bytes, cudaMemcpyDeviceToDevice, s));
    }
  }
  if (full) {
    AgentBase& D = *dst->base;
    const AgentBase& C = *src->base;
    B200_CHECK_CUDA(cudaMemcpyAsync(D.d_step, C.d_step, sizeof(StepState), cudaMemcpyDeviceToDevice, s));
    B200_CHECK_CUDA(cudaMemcpyAsync(&D.d_step->mail, &D.d_mail, sizeof(InfoMail*), cudaMemcpyHostToDevice, s));  // keep its own mailbox
    B200_CHECK_CUDA(cudaMemcpyAsync(D.d_adam, C.d_adam, (size_t)C.nopt * sizeof(AdamScalars), cudaMemcpyDeviceToDevice, s));
    D.n_calls = C.n_calls;
    if (dst->td7) B200_CHECK_CUDA(cudaMemcpyAsync(dst->td7->d_range, src->td7->d_range, (size_t)D.G * sizeof(TD7Range), cudaMemcpyDeviceToDevice, s));
    if (dst->mlp && dst->mlp->sac) {
      const size_t b = (size_t)D.G * sizeof(float);
      B200_CHECK_CUDA(cudaMemcpyAsync(dst->mlp->la_p, src->mlp->la_p, b, cudaMemcpyDeviceToDevice, s));
      B200_CHECK_CUDA(cudaMemcpyAsync(dst->mlp->la_m, src->mlp->la_m, b, cudaMemcpyDeviceToDevice, s));
      B200_CHECK_CUDA(cudaMemcpyAsync(dst->mlp->la_v, src->mlp->la_v, b, cudaMemcpyDeviceToDevice, s));
    }
  }
  return kOk;
}

int b200rl_agent_get_counters(b200rl_agent* h, int64_t* n_runs, int64_t* adam_steps) {
  B200_REQUIRE(h, "null argument");
  StepState st;
  B200_CHECK_CUDA(cudaMemcpy(&st, h->base->d_step, sizeof(st), cudaMemcpyDeviceToHost));
  if (n_runs) *n_runs = h->base->n_calls;
  if (adam_steps)
    for (int k = 0; k < 3; ++k) adam_steps[k] = st.adam_t[k];
  return kOk;
}
int b200rl_agent_set_counters(b200rl_agent* h, int64_t n_runs, const int64_t* adam_steps, void* stream) {
  B200_REQUIRE(h && adam_steps, "null argument");
  StepState st;
  memset(&st, 0, sizeof(st));
  st.n_runs = (int)n_runs;
  std::vector<OptimHyper> hy(h->base->nopt);
  B200_CHECK_CUDA(cudaMemcpy(hy.data(), h->base->d_hyper, hy.size() * sizeof(OptimHyper), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 4; ++k) st.b1pow[k] = st.b2pow[k] = 1.0;
  for (int k = 0; k < 3 && k < h->base->nopt; ++k) {
    st.adam_t[k] = (int)adam_steps[k];
    st.b1pow[k] = pow(hy[k].beta1, (double)adam_steps[k]);
    st.b2pow[k] = pow(hy[k].beta2, (double)adam_steps[k]);
  }
  st.mail = h->base->d_mail;  // the pinned loss mailbox belongs to the handle, not to the restored counters
  B200_CHECK_CUDA(cudaMemcpyAsync(h->base->d_step, &st, sizeof(st), cudaMemcpyHostToDevice, as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  h->base->n_calls = n_runs;
  return kOk;
}
int b200rl_td7_value_range(b200rl_agent* h, int group, float* out4_host, void* stream) {
  B200_REQUIRE(h && h->td7 && out4_host && group >= -1 && group < h->base->G, "not a TD7 agent / bad group");
  B200_CHECK_CUDA(cudaMemcpyAsync(out4_host, h->td7->d_range + (group < 0 ? 0 : group), (group < 0 ? h->base->G : 1) * sizeof(TD7Range), cudaMemcpyDeviceToHost,
                                  as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}
int b200rl_td7_set_value_range(b200rl_agent* h, int group, const float* in4_host, void* stream) {
  B200_REQUIRE(h && h->td7 && in4_host && group >= -1 && group < h->base->G, "not a TD7 agent / bad group");
  B200_CHECK_CUDA(cudaMemcpyAsync(h->td7->d_range + (group < 0 ? 0 : group), in4_host, (group < 0 ? h->base->G : 1) * sizeof(TD7Range), cudaMemcpyHostToDevice,
                                  as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}
int b200rl_agent_set_rng(b200rl_agent* h, uint64_t seed) {
  B200_REQUIRE(h, "null argument");
  h->base->rng_seed = seed;
  h->base->drop_graphs();  // the seed is a baked kernel argument
  return h->base->rebuild_mega();
}

int b200rl_agent_io(b200rl_agent* h, float** ss_dev, float** action_dev, float** reward_dev, float** notdone_dev, int64_t* group_stride) {
  B200_REQUIRE(h, "null argument");
  AgentBase& a = *h->base;
  if (ss_dev) *ss_dev = a.ss.p;
  if (action_dev) *action_dev = a.act_in.p;
  if (reward_dev) *reward_dev = a.rew.p;
  if (notdone_dev) *notdone_dev = a.nd.p;
  if (group_stride) *group_stride = a.act_stride;
  return kOk;
}
int b200rl_agent_io_ld(b200rl_agent* h, int* ld_state, int* ld_action) {
  B200_REQUIRE(h, "null argument");
  if (ld_state) *ld_state = h->base->ss.ld;
  if (ld_action) *ld_action = h->base->act_in.ld;
  return kOk;
}
int b200rl_replay_sample_to_agent(b200rl_replay* replay, b200rl_agent* h, int batch, const float* u_dev, void* stream) {
  B200_REQUIRE(h && replay, "null argument");
  AgentBase& a = *h->base;
  B200_REQUIRE(replay->r.groups == a.G && replay->r.S == a.S && replay->r.A == a.A, "replay buffer and agent shapes differ");
  B200_REQUIRE(batch == a.B, "sample_to_agent: the batch size must be the engine's");
  return replay->r.sample(batch, u_dev, a.replay_outputs(), as_stream(stream));
}
int b200rl_agent_stage_batch(b200rl_agent* h, int group, const float* state_dev, const float* action_dev, const float* reward_dev, const float* next_state_dev,
                             const float* done_dev, void* stream) {
  B200_REQUIRE(h && state_dev && action_dev && reward_dev && next_state_dev && done_dev, "null argument");
  B200_REQUIRE(group >= 0 && group < h->base->G, "bad group");
  h->base->exec = as_stream(stream);
  return h->base->stage_batch(state_dev, action_dev, reward_dev, next_state_dev, done_dev, group);
}

static int agent_update_one(b200rl_agent* h, b200rl_replay* replay, const float* noise_dev) {
  Replay* r = replay ? &replay->r : nullptr;
  if (h->td7) return h->td7->update(r, noise_dev);
  return h->mlp->update(r, noise_dev);
}
int b200rl_agent_update(b200rl_agent* h, b200rl_replay* replay, const float* noise_dev, void* stream) {
  B200_REQUIRE(h, "null argument");
  h->base->exec = as_stream(stream);
  return agent_update_one(h, replay, noise_dev);
}
int b200rl_agent_update_n(b200rl_agent* h, b200rl_replay* replay, int n, void* stream) {
  B200_REQUIRE(h && replay, "null argument");
  B200_REQUIRE(n >= 0, "n must be non-negative");
  AgentBase& a = *h->base;
  B200_REQUIRE(replay->r.groups == a.G && replay->r.S == a.S && replay->r.A == a.A, "replay buffer and agent shapes differ");
  a.exec = as_stream(stream);
  const Replay::Outputs o = a.replay_outputs();
  for (int i = 0; i < n; ++i) {
    B200_TRY(replay->r.sample(a.B, nullptr, o, a.exec, &a.d_step->n_runs));
    B200_TRY(agent_update_one(h, replay, nullptr));
  }
  return kOk;
}
int b200rl_agent_info_cols(const b200rl_agent* h) { return h ? h->base->info_cols : kErrInvalidArg; }
int b200rl_agent_read_info(b200rl_agent* h, int64_t first_step, int count, int group, float* out_host, void* stream) {
  B200_REQUIRE(h && out_host, "null argument");
  AgentBase& a = *h->base;
  B200_REQUIRE(group >= -1 && group < a.G, "bad group");
  B200_REQUIRE(count >= 0 && count <= kInfoRing, "at most 1024 steps stay readable");
  B200_REQUIRE(first_step >= 1 && first_step + count - 1 <= a.n_calls && first_step > a.n_calls - kInfoRing, "step range is not (or no longer) in the info ring");
  cudaStream_t s = as_stream(stream);
  // group = -1: every stacked seed, out_host [groups][count][cols] -- one strided copy per contiguous run of the ring
  const int ngroups = group < 0 ? a.G : 1;
  const float* ring = a.d_info + (long long)(group < 0 ? 0 : group) * kInfoRing * a.info_cols;
  for (int i = 0; i < count;) {
    const int slot = (int)((first_step - 1 + i) % kInfoRing);
    const int run = std::min(count - i, kInfoRing - slot);
    B200_CHECK_CUDA(cudaMemcpy2DAsync(out_host + (long long)i * a.info_cols, (size_t)count * a.info_cols * sizeof(float), ring + (long long)slot * a.info_cols,
                                      (size_t)kInfoRing * a.info_cols * sizeof(float), (size_t)run * a.info_cols * sizeof(float), (size_t)ngroups, cudaMemcpyDeviceToHost, s));
    i += run;
  }
  B200_CHECK_CUDA(cudaStreamSynchronize(s));
  return kOk;
}

int b200rl_agent_wait_info(b200rl_agent* h, int64_t step, int col, float* out) {
  B200_REQUIRE(h && out, "null argument");
  AgentBase& a = *h->base;
  if (!a.h_mail || col < 0 || col >= kMailCols || step != a.n_calls) return 1;  // not the newest update: read the ring
  const volatile unsigned long long* m = &a.h_mail[col].bits;
  const auto t0 = std::chrono::steady_clock::now();
  long spins = 0;
  unsigned long long x;
  while ((int)((x = *m) >> 32) < (int)step) {
    if ((++spins & 0xfff) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) return 1;  // never posted: read the ring
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  if ((int)(x >> 32) != (int)step) return 1;  // a later update already posted over it
  const unsigned lo = (unsigned)(x & 0xffffffffull);
  float v;
  memcpy(&v, &lo, sizeof(v));
  *out = v;
  return kOk;
}
int b200rl_agent_act(b200rl_agent* h, const float* obs_dev, float* out_dev, void* stream) {
  B200_REQUIRE(h && obs_dev && out_dev, "null argument");
  AgentBase& a = *h->base;
  cudaStream_t s = as_stream(stream);
  const int G = a.G;
#ifndef B200_EMU
  if (use_fused_act()) return launch_act_fused(h, obs_dev, a.S, out_dev, h->aout.cols, false, s);
#endif
  for (int g = 0; g < G; ++g)
    B200_CHECK_CUDA(cudaMemcpyAsync(h->o_s.p + (long long)g * h->o_s.gs, obs_dev + (long long)g * a.S, (size_t)a.S * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (h->td7) {
    B200_TRY(gemm_table_launch(h->act_t[0], G, s));
    B200_TRY(gemm_table_launch(h->act_t[1], G, s));
    B200_TRY(gemm_table_launch(h->act_t[2], G, s));
    B200_TRY(row_table_launch(h->act_n[0], G, s));
    B200_TRY(gemm_table_launch(h->act_t[3], G, s));
    B200_TRY(gemm_table_launch(h->act_t[4], G, s));
    B200_TRY(gemm_table_launch(h->act_t[5], G, s));
  } else {
    for (int i = 0; i < 3; ++i) B200_TRY(gemm_table_launch(h->act_t[i], G, s));
  }
  const int Aout = h->aout.cols;
  for (int g = 0; g < G; ++g)
    B200_CHECK_CUDA(cudaMemcpyAsync(out_dev + (long long)g * Aout, h->aout.p + (long long)g * h->aout.gs, (size_t)Aout * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return kOk;
}

int b200rl_agent_act_host(b200rl_agent* h, const float* obs_host, float* out_host, void* stream) {
  B200_REQUIRE(h && obs_host && out_host, "null argument");
  AgentBase& a = *h->base;
  cudaStream_t s = as_stream(stream);
  const int G = a.G, Aout = h->aout.cols;
#ifndef B200_EMU
  if (use_fused_act()) {
    if (G == 1 && a.S <= kActInlineObs) {
      B200_TRY(launch_act_fused(h, nullptr, 0, h->d_act_map, Aout, true, s, obs_host));
    } else {
      memcpy(h->h_obs, obs_host, (size_t)G * a.S * sizeof(float));
      B200_TRY(launch_act_fused(h, h->d_obs_map, a.S, h->d_act_map, Aout, true, s));
    }
    const volatile unsigned long long* tag = h->h_tag;
    const unsigned long long want = h->act_seq;
    const auto t0 = std::chrono::steady_clock::now();
    long spins = 0;
    while (*tag != want) {
      if ((++spins & 0xfff) == 0) {
        if (cudaStreamQuery(s) == cudaSuccess && *tag != want) { (void)cudaGetLastError(); break; }  // ran without posting: fall through to the check
        (void)cudaGetLastError();
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) { set_last_error("agent_act_host: the acting kernel did not post its result"); return kErrCuda; }
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    if (*tag != want) { B200_CHECK_CUDA(cudaStreamSynchronize(s)); if (*tag != want) { set_last_error("agent_act_host: acting kernel failed"); return kErrCuda; } }
    std::atomic_thread_fence(std::memory_order_acquire);
    memcpy(out_host, h->h_act, (size_t)G * Aout * sizeof(float));
    return kOk;
  }
#endif
  // launch-table path: stage through device buffers
  DeviceArena tmp;
  float *o = nullptr, *y = nullptr;
  B200_TRY(tmp.alloc(&o, (size_t)G * a.S));
  B200_TRY(tmp.alloc(&y, (size_t)G * Aout));
  B200_CHECK_CUDA(cudaMemcpyAsync(o, obs_host, (size_t)G * a.S * sizeof(float), cudaMemcpyHostToDevice, s));
  B200_TRY(b200rl_agent_act(h, o, y, stream));
  B200_CHECK_CUDA(cudaMemcpyAsync(out_host, y, (size_t)G * Aout * sizeof(float), cudaMemcpyDeviceToHost, s));
  B200_CHECK_CUDA(cudaStreamSynchronize(s));
  return kOk;
}

int b200rl_agent_debug_read(b200rl_agent* h, int which, int group, float* out_host, void* stream) {
  B200_REQUIRE(h && out_host, "null argument");
  AgentBase& a = *h->base;
  B200_REQUIRE(group >= 0 && group < a.G, "bad group");
  const Mat* m = nullptr;
  if (h->td7) {
    TD7Agent& t = *h->td7;
    const Mat* tab[6] = {&t.y_dbg, &t.prio_dbg, &t.qv[0], &t.qv[1], &t.a_pi, &t.noise};
    B200_REQUIRE(which >= 0 && which < 6, "which must be 0..5");
    m = tab[which];
  } else {
    MlpAgent& t = *h->mlp;
    const Mat* tab[7] = {&t.y_dbg, &t.prio_dbg, &t.qv[0], &t.qv[1], &t.a_pi, &t.noise, &t.noise2};
    B200_REQUIRE(which >= 0 && which < (t.sac ? 7 : 6), "which must be 0..5 (SAC: 0..6)");
    m = tab[which];
  }
  B200_CHECK_CUDA(cudaMemcpy2DAsync(out_host, (size_t)m->cols * sizeof(float), m->p + (long long)group * m->gs, (size_t)m->ld * sizeof(float), (size_t)m->cols * sizeof(float),
                                    (size_t)m->rows, cudaMemcpyDeviceToHost, as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}

int b200rl_agent_launches_per_update(b200rl_agent* h, int actor_step) {
  B200_REQUIRE(h, "null argument");
  return h->base->launches[actor_step ? 1 : 0];
}

// ============================================================================ op-level
int b200rl_op_gemm(int mode, const b200rl_gemm_problem* problems_host, int nproblems, int groups, void* stream) {
  B200_REQUIRE(problems_host && nproblems > 0 && groups > 0, "bad arguments");
  B200_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0..2");
  B200_TRY(ensure_device());
  std::vector<GemmProblem> v(nproblems);
  memcpy(v.data(), problems_host, (size_t)nproblems * sizeof(GemmProblem));
  DeviceArena tmp;
  GemmTable t;
  B200_TRY(gemm_table_upload(&t, mode, v, &tmp, groups));
  B200_TRY(gemm_table_launch(t, groups, as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}
#ifdef B200RL_PROBE
// out[256][2] = {earliest CTA entry, latest CTA exit} (ns, global timer) per table since the last call; resets the slots
int b200rl_debug_timeline(unsigned long long* out256x2) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out256x2, g_tl, sizeof(unsigned long long) * 512) != cudaSuccess) return 1;
  static unsigned long long init[256][2];
  for (auto& e : init) { e[0] = ~0ull; e[1] = 0ull; }
  return cudaMemcpyToSymbol(g_tl, init, sizeof(init)) == cudaSuccess ? 0 : 1;
}
// per-tile trace of the persistent dataflow kernels: out[programs = 8][tiles = 4096][5]; resets the buffer
int b200rl_debug_mega_trace(unsigned long long* out) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, g_mega_trace, sizeof(g_mega_trace)) != cudaSuccess) return 1;
  std::vector<unsigned long long> zero(sizeof(g_mega_trace) / 8, 0ull);
  return cudaMemcpyToSymbol(g_mega_trace, zero.data(), sizeof(g_mega_trace)) == cudaSuccess ? 0 : 1;
}
int b200rl_debug_probe(long long* out64x10) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out64x10, g_probe, sizeof(long long) * 640) == cudaSuccess ? 0 : 1;
}
#endif
int b200rl_op_gemm_timed(int mode, const b200rl_gemm_problem* problems_host, int nproblems, int groups, int warmup, int iters, float* ms_per_launch, void* stream) {
  B200_REQUIRE(problems_host && nproblems > 0 && groups > 0 && iters > 0 && ms_per_launch, "bad arguments");
  B200_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0..2");
  B200_TRY(ensure_device());
  std::vector<GemmProblem> v(nproblems);
  memcpy(v.data(), problems_host, (size_t)nproblems * sizeof(GemmProblem));
  DeviceArena tmp;
  GemmTable t;
  B200_TRY(gemm_table_upload(&t, mode, v, &tmp, groups));
  cudaStream_t s = as_stream(stream);
  for (int i = 0; i < warmup; ++i) B200_TRY(gemm_table_launch(t, groups, s));
#ifndef B200_EMU
  // the `iters` launches are captured into one CUDA graph so that the measurement sees device-side
  // back-to-back execution (as inside the update graphs), not the host's launch rate
  cudaStream_t cs;
  B200_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  B200_CHECK_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
  int st = kOk;
  for (int i = 0; i < iters && st == kOk; ++i) st = gemm_table_launch(t, groups, cs);
  cudaError_t ce = cudaStreamEndCapture(cs, &graph);
  if (st != kOk) return st;
  B200_CHECK_CUDA(ce);
  B200_CHECK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  B200_CHECK_CUDA(cudaGraphLaunch(exec, s));
  B200_CHECK_CUDA(cudaStreamSynchronize(s));
  cudaEvent_t e0, e1;
  B200_CHECK_CUDA(cudaEventCreate(&e0));
  B200_CHECK_CUDA(cudaEventCreate(&e1));
  B200_CHECK_CUDA(cudaEventRecord(e0, s));
  B200_CHECK_CUDA(cudaGraphLaunch(exec, s));
  B200_CHECK_CUDA(cudaEventRecord(e1, s));
  B200_CHECK_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  B200_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  cudaStreamDestroy(cs);
  *ms_per_launch = ms / (float)iters;
#else
  for (int i = 0; i < iters; ++i) B200_TRY(gemm_table_launch(t, groups, s));
  *ms_per_launch = 0.f;
#endif
  return kOk;
}
int b200rl_replay_time_sample(b200rl_replay* h, int batch, int warmup, int iters, float* ms_out2, void* stream) {
  B200_REQUIRE(h && ms_out2 && iters > 0, "bad arguments");
  Replay& r = h->r;
  long long n;
  B200_TRY(r.common_size(&n));
  B200_REQUIRE(n > 0, "buffer is empty");
  DeviceArena tmp;
  float *st, *ac, *rw, *ns, *dn;
  const int G = r.groups;
  B200_TRY(tmp.alloc(&st, (size_t)G * batch * r.S)); B200_TRY(tmp.alloc(&ns, (size_t)G * batch * r.S));
  B200_TRY(tmp.alloc(&ac, (size_t)G * batch * r.A)); B200_TRY(tmp.alloc(&rw, (size_t)G * batch)); B200_TRY(tmp.alloc(&dn, (size_t)G * batch));
  Replay::Outputs o;
  o.state = st; o.ld_state = r.S; o.state_gs = (long long)batch * r.S;
  o.next_state = ns; o.ld_next = r.S; o.next_gs = (long long)batch * r.S;
  o.action = ac; o.ld_action = r.A; o.action_gs = (long long)batch * r.A;
  o.reward = rw; o.reward_gs = batch; o.done = dn; o.done_gs = batch;
  cudaStream_t s = as_stream(stream);
  ms_out2[0] = ms_out2[1] = 0.f;
#ifndef B200_EMU
  const int ntiles = ceil_div((int)n, kPrioTile);
  cudaEvent_t e0, e1;
  B200_CHECK_CUDA(cudaEventCreate(&e0));
  B200_CHECK_CUDA(cudaEventCreate(&e1));
  for (int i = 0; i < warmup; ++i) B200_TRY(r.sample(batch, nullptr, o, s));
  if (r.prioritized) {
    dim3 grid((unsigned)ceil_div(ntiles, kPrioTilesPerCta), (unsigned)G), block(256);
    B200_CHECK_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; ++i) B200_LAUNCH(lap_tile_sums_kernel, grid, block, 0, s, r.d_priority, r.cap_pad, (int)n, r.d_tile_sums, ntiles, (long long)r.max_tiles);
    B200_CHECK_CUDA(cudaEventRecord(e1, s));
    B200_CHECK_CUDA(cudaEventSynchronize(e1));
    B200_CHECK_CUDA(cudaEventElapsedTime(&ms_out2[0], e0, e1));
    ms_out2[0] /= (float)iters;
  }
  B200_CHECK_CUDA(cudaEventRecord(e0, s));
  for (int i = 0; i < iters; ++i) B200_TRY(r.sample(batch, nullptr, o, s));
  B200_CHECK_CUDA(cudaEventRecord(e1, s));
  B200_CHECK_CUDA(cudaEventSynchronize(e1));
  B200_CHECK_CUDA(cudaEventElapsedTime(&ms_out2[1], e0, e1));
  ms_out2[1] /= (float)iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
#else
  (void)warmup;
  for (int i = 0; i < iters; ++i) B200_TRY(r.sample(batch, nullptr, o, s));
#endif
  return kOk;
}
int b200rl_op_avgl1norm(int backward, const float* x_dev, const float* dy_dev, float* y_dev, int rows, int n, void* stream) {
  B200_REQUIRE(x_dev && y_dev && rows > 0 && n > 0 && (!backward || dy_dev), "bad arguments");
  B200_TRY(ensure_device());
  std::vector<RowOp> ops{make_rowop(x_dev, n, 0, y_dev, n, 0, rows, n, dy_dev, n, 0)};
  DeviceArena tmp;
  RowTable t;
  B200_TRY(row_table_upload(&t, backward != 0, ops, &tmp));
  B200_TRY(row_table_launch(t, 1, as_stream(stream)));
  B200_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return kOk;
}
int b200rl_op_adam(float* p_dev, float* m_dev, float* v_dev, const float* g_dev, int64_t n, double lr, double beta1, double beta2, double eps, int step, void* stream) {
  B200_REQUIRE(p_dev && m_dev && v_dev && g_dev && n > 0 && step >= 1, "bad arguments");
  B200_TRY(ensure_device());
  DeviceArena tmp;
  StepState* st; AdamScalars* sc; OptimHyper* hy;
  B200_TRY(tmp.alloc(&st, 1)); B200_TRY(tmp.alloc(&sc, 1)); B200_TRY(tmp.alloc(&hy, 1));
  StepState h0; memset(&h0, 0, sizeof(h0)); h0.adam_t[0] = step - 1;
  for (int k = 0; k < 4; ++k) h0.b1pow[k] = h0.b2pow[k] = 1.0;
  h0.b1pow[0] = pow(beta1, (double)(step - 1)); h0.b2pow[0] = pow(beta2, (double)(step - 1));
  OptimHyper hh{lr, beta1, beta2, eps};
  B200_CHECK_CUDA(cudaMemcpy(st, &h0, sizeof(h0), cudaMemcpyHostToDevice));
  B200_CHECK_CUDA(cudaMemcpy(hy, &hh, sizeof(hh), cudaMemcpyHostToDevice));
  cudaStream_t s = as_stream(stream);
  B200_LAUNCH(prep_step_kernel, dim3(1), dim3(32), 0, s, st, sc, hy, 1, 1u, kInfoRing);
  B200_LAUNCH(adam_blob_kernel, dim3((unsigned)launch_elementwise_grid(n)), dim3(256), 0, s, p_dev, m_dev, v_dev, g_dev, (long long)n, sc);
  B200_CHECK_LAUNCH();
  B200_CHECK_CUDA(cudaStreamSynchronize(s));
  return kOk;
}
int b200rl_op_polyak(float* target_dev, const float* src_dev, int64_t n, double tau, void* stream) {
  B200_REQUIRE(target_dev && src_dev && n > 0, "bad arguments");
  B200_TRY(ensure_device());
  B200_LAUNCH(polyak_kernel, dim3((unsigned)launch_elementwise_grid(n)), dim3(256), 0, as_stream(stream), target_dev, src_dev, (long long)n, (float)tau, (float)(1.0 - tau));
  B200_CHECK_LAUNCH();
  return kOk;
}
int b200rl_op_sac_sample(const float* net_out_dev, const float* eps_dev, float* action_dev, float* logp_dev, int B, int A, double min_log_std, double max_log_std,
                         void* stream) {
  B200_REQUIRE(net_out_dev && eps_dev && action_dev && logp_dev && B > 0 && A > 0, "bad arguments");
  B200_TRY(ensure_device());
  dim3 grid((unsigned)ceil_div(B, 8), 1), block(256);
  B200_LAUNCH(sac_sample_fwd_kernel, grid, block, 0, as_stream(stream), net_out_dev, 2 * A, eps_dev, A, action_dev, A, logp_dev, B, A, (float)min_log_std, (float)max_log_std,
              0LL, 0LL, 0LL, 0LL);
  B200_CHECK_LAUNCH();
  return kOk;
}
int b200rl_op_sac_sample_bwd(const float* net_out_dev, const float* eps_dev, const float* action_dev, const float* daction_dev, const float* dlogp_dev,
                             float* dnet_out_dev, int B, int A, double min_log_std, double max_log_std, void* stream) {
  B200_REQUIRE(net_out_dev && eps_dev && action_dev && daction_dev && dlogp_dev && dnet_out_dev && B > 0 && A > 0, "bad arguments");
  B200_TRY(ensure_device());
  dim3 grid((unsigned)ceil_div(B, 8), 1), block(32, 8);
  B200_LAUNCH(sac_sample_bwd_kernel, grid, block, 0, as_stream(stream), net_out_dev, 2 * A, eps_dev, A, action_dev, A, daction_dev, A, dlogp_dev, dnet_out_dev, 2 * A, B, A,
              (float)min_log_std, (float)max_log_std, 0LL, 0LL, 0LL, 0LL);
  B200_CHECK_LAUNCH();
  return kOk;
}

}  // extern "C"
