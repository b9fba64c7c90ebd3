// agent_base.cuh -- state shared by the TD7 / TD3 / SAC update engines: parameter blobs,
// activation arena, step counters, info ring, fork/join streams and CUDA-graph replay.
#pragma once
#include "nets.cuh"
#include "replay.cuh"
#include "mega.cuh"

namespace b200rl {

constexpr int kInfoRing = 1024;  // update steps whose losses stay readable without a sync

// Gaussian fill: z ~ N(0,1) from Philox(seed; counter = *step_counter, index), unless *inject != 0
// (then the buffer already holds caller-provided noise -- parity tests).
__global__ void __launch_bounds__(256) gauss_fill_kernel(float* __restrict__ out_base, long long gs, long long n, unsigned long long seed0, int stream_id,
                                                         const int* __restrict__ step_counter, const int* __restrict__ inject) {
  B200_PDL_PROLOGUE();
  if (*inject) return;
  const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pair * 2 >= n) return;
  // every stacked seed draws from its own key, so its stream equals that of a single-seed run
  float* out = out_base + (long long)blockIdx.y * gs;
  Philox ph(seed0 + 0x9E3779B97F4A7C15ULL * (unsigned long long)blockIdx.y);
  uint32_t r[4];
  ph(((unsigned long long)(unsigned)(*step_counter) << 8) | (unsigned)stream_id, (unsigned long long)pair, r);
  float z0, z1;
  gauss2_from_bits(r[0], r[1], &z0, &z1);
  out[pair * 2] = z0;
  if (pair * 2 + 1 < n) out[pair * 2 + 1] = z1;
}

// copy a caller batch (five [B,.] tensors) into the engine's fixed input buffers
struct StageArgs {
  const float* state; const float* action; const float* reward; const float* next_state; const float* done;
  float* ss; float* act; float* rew; float* nd;
  int B, S, A;
  int ld_s, ld_a;  // leading dimensions of the engine's state / action buffers
};
__global__ void __launch_bounds__(256) stage_batch_kernel(StageArgs a) {
  B200_PDL_PROLOGUE();
  const int nS = a.B * a.S, nA = a.B * a.A;
  const int total = 2 * nS + nA + 2 * a.B;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < 2 * nS) {
      const int r = i / a.S, c = i - r * a.S;  // rows 0..B-1: state, B..2B-1: next_state
      a.ss[(long long)r * a.ld_s + c] = (r < a.B) ? a.state[i] : a.next_state[i - nS];
    } else if (i < 2 * nS + nA) {
      const int k = i - 2 * nS, r = k / a.A, c = k - r * a.A;
      a.act[(long long)r * a.ld_a + c] = a.action[k];
    }
    else if (i < 2 * nS + nA + a.B) a.rew[i - 2 * nS - nA] = a.reward[i - 2 * nS - nA];
    else a.nd[i - 2 * nS - nA - a.B] = a.done[i - 2 * nS - nA - a.B];
  }
}

struct AgentBase {
  int S = 0, A = 0, H = 256, Z = 256, B = 256, G = 1;
  bool use_lap = false;
  bool keep_grads = false;
  DeviceArena arena;
  cudaStream_t stream = nullptr;  // capture / setup stream (owned)
  cudaStream_t exec = nullptr;    // caller's stream: graphs are launched and loose ops enqueued here
  Fork fork;

  StepState* d_step = nullptr;
  AdamScalars* d_adam = nullptr;
  OptimHyper* d_hyper = nullptr;
  int* d_flags = nullptr;  // [0] = noise injected by caller
  int nopt = 0;
  float* d_info = nullptr;
  int info_cols = 0;
  InfoMail* h_mail = nullptr;  // pinned, mapped: losses of seed 0 as soon as their kernel has run (ops.cuh: publish_info)
  InfoMail* d_mail = nullptr;
  long long n_calls = 0;  // host mirror of StepState.n_runs
  unsigned long long rng_seed = 0xB200B200ULL;
  unsigned low_priority_streams = 0;  // bit i: fork.side[i] carries slack work (set by the subclass before init_common)

  // activation arena (two-pass layout)
  float* act_base = nullptr;
  long long act_cursor = 0, act_stride = 0;

  // fixed input buffers
  Mat ss, act_in, rew, nd;
  Replay* bound = nullptr;
  bool built = false;
  int launches[4] = {0, 0, 0, 0};  // kernels per update, by graph variant
#ifndef B200_EMU
  cudaGraphExec_t graphs[4] = {nullptr, nullptr, nullptr, nullptr};
#endif

  virtual ~AgentBase() {
#ifndef B200_EMU
    for (auto& g : graphs)
      if (g) cudaGraphExecDestroy(g);
#endif
    fork.destroy();
    if (stream) cudaStreamDestroy(stream);
    if (h_mail) cudaFreeHost(h_mail);
  }

  // pad4: rows start every round_up(cols, 4) floats (the arena is zero-filled and nothing writes the padding), which
  // makes observation- and action-wide activations 16-byte aligned operands (TMA, 128-bit loads)
  Mat alloc_act(int rows, int cols, bool pad4 = false) {
    Mat m;
    m.rows = rows; m.cols = cols; m.ld = pad4 ? (int)round_up(cols, 4) : cols;
    m.p = act_base ? act_base + act_cursor : nullptr;
    m.gs = act_stride;
    act_cursor += round_up((long long)rows * m.ld, 4);
    return m;
  }

  int init_common(int nopt_, int info_cols_, const OptimHyper* hyper) {
    nopt = nopt_;
    info_cols = info_cols_;
    {
      int least = 0, greatest = 0;
      B200_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      B200_CHECK_CUDA(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, greatest));
    }
    B200_TRY(fork.init(stream, 4, 64, low_priority_streams));
    B200_TRY(arena.alloc(&d_step, 1));
    B200_TRY(arena.alloc(&d_adam, (size_t)nopt));
    B200_TRY(arena.alloc(&d_hyper, (size_t)nopt));
    B200_TRY(arena.alloc(&d_flags, 4));
    B200_TRY(arena.alloc(&d_info, (size_t)G * kInfoRing * info_cols));
    B200_CHECK_CUDA(cudaMemcpy(d_hyper, hyper, nopt * sizeof(OptimHyper), cudaMemcpyHostToDevice));
#ifdef B200_EMU
    B200_CHECK_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_mail), kMailCols * sizeof(InfoMail)));
    d_mail = h_mail;
#else
    B200_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_mail), kMailCols * sizeof(InfoMail), cudaHostAllocMapped));
    B200_CHECK_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_mail), h_mail, 0));
#endif
    memset(h_mail, 0, kMailCols * sizeof(InfoMail));
    StepState st0;
    memset(&st0, 0, sizeof(st0));
    st0.mail = d_mail;
    for (int k = 0; k < 4; ++k) st0.b1pow[k] = st0.b2pow[k] = 1.0;
    B200_CHECK_CUDA(cudaMemcpy(d_step, &st0, sizeof(st0), cudaMemcpyHostToDevice));
    return kOk;
  }

  int alloc_net(Net* n, const NetLayout& L, bool trained, const AdamScalars* adam) {
    n->L = L;
    n->gs = L.numel;
    B200_TRY(arena.alloc(&n->p, (size_t)G * L.numel));
    if (trained) {
      B200_TRY(arena.alloc(&n->m, (size_t)G * L.numel));
      B200_TRY(arena.alloc(&n->v, (size_t)G * L.numel));
      if (keep_grads) B200_TRY(arena.alloc(&n->grad, (size_t)G * L.numel));
      n->adam = adam;
    }
    return kOk;
  }

  int copy_net(Net& dst, const Net& src, cudaStream_t s) {
    B200_CHECK_CUDA(cudaMemcpyAsync(dst.p, src.p, (size_t)G * src.L.numel * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return kOk;
  }

#ifndef B200_EMU
  // the persistent dataflow path (mega.cuh): an option (B200RL_MEGA=1 / b200rl_set_mega), single agents only (stacked
  // seeds share the step counters, which the per-seed counters of the dataflow kernel do not order).  It is parity-green
  // but NOT the default: measured on B200 a TD7 update takes 345 us through it (382 us before the descriptor table moved to kernel parameters) against 203 us through the
  // launch-per-layer graphs -- the 32 x 32 FFMA tile sustains ~28 tiles/us machine-wide either way (20 % of the fp32
  // peak, shared-memory fragment traffic), the classic graphs already keep the machine that busy with 2.4 launches in
  // flight, and the dataflow kernel adds ~3 us of ticket / descriptor / counter traffic per tile (profiles/README.md).
  static int& mega_override() { static int v = -1; return v; }
  bool use_mega() const {
    static const bool on = [] { const char* e = getenv("B200RL_MEGA"); return e ? atoi(e) != 0 : false; }();
    return (mega_override() >= 0 ? mega_override() != 0 : on) && G == 1;
  }
  PrepArgs prep_args(unsigned step_mask) const {
    PrepArgs p;
    memset(&p, 0, sizeof(p));
    p.st = d_step; p.adam = d_adam; p.hyper = d_hyper; p.nopt = nopt; p.step_mask = step_mask; p.ring = kInfoRing;
    return p;
  }
  GaussArgs gauss_args(const Mat& noise, int stream_id) const {
    GaussArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.out_base = noise.p; ga.gs = noise.gs; ga.n = (long long)noise.rows * noise.cols; ga.seed0 = rng_seed;
    ga.step_counter = &d_step->n_runs; ga.inject = d_flags; ga.stream_id = stream_id;
    return ga;
  }
  static int sm_count() {
    static const int n = [] { int v = 148, d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v; }();
    return n;
  }
#endif
  virtual int rebuild_mega() { return kOk; }  // (re)record the dataflow programs: at build time and when a baked scalar (the RNG seed) changes

  int launch_prep(unsigned step_mask, cudaStream_t s) {
    B200_LAUNCH(prep_step_kernel, dim3(1), dim3(32), 0, s, d_step, d_adam, d_hyper, nopt, step_mask, kInfoRing);
    B200_CHECK_LAUNCH();
    return kOk;
  }

  int launch_gauss(const Mat& noise, int stream_id, cudaStream_t s) {
    const long long n = (long long)noise.rows * noise.cols;
    dim3 grid((unsigned)((n / 2 + 256) / 256), (unsigned)G), block(256);
    B200_LAUNCH(gauss_fill_kernel, grid, block, 0, s, noise.p, noise.gs, n, rng_seed, stream_id, &d_step->n_runs, d_flags);
    B200_CHECK_LAUNCH();
    return kOk;
  }

  // caller-provided batch -> fixed buffers (device pointers, single seed)
  int stage_batch(const float* state, const float* action, const float* reward, const float* next_state, const float* done, int group) {
    StageArgs a;
    const long long o = (long long)group * act_stride;
    a.state = state; a.action = action; a.reward = reward; a.next_state = next_state; a.done = done;
    a.ss = ss.p + o; a.act = act_in.p + o; a.rew = rew.p + o; a.nd = nd.p + o;
    a.B = B; a.S = S; a.A = A; a.ld_s = ss.ld; a.ld_a = act_in.ld;
    B200_LAUNCH(stage_batch_kernel, dim3(32), dim3(256), 0, exec, a);
    B200_CHECK_LAUNCH();
    return kOk;
  }

  Replay::Outputs replay_outputs() const {
    Replay::Outputs o;
    o.state = ss.p; o.ld_state = ss.ld; o.state_gs = act_stride;
    o.next_state = ss.p + (long long)B * ss.ld; o.ld_next = ss.ld; o.next_gs = act_stride;
    o.action = act_in.p; o.ld_action = act_in.ld; o.action_gs = act_stride;
    o.reward = rew.p; o.reward_gs = act_stride;
    o.done = nd.p; o.done_gs = act_stride;
    return o;
  }

  // ---- graph replay -----------------------------------------------------------------------
  virtual int issue_step(int variant) = 0;  // enqueue one update on `stream` (+ side streams)

  int run_step(int variant) {
#ifdef B200_EMU
    fork.begin_step();
    const int c0 = launch_counter();
    const int st_emu = issue_step(variant);
    launches[variant] = launch_counter() - c0;
    return st_emu;
#else
    if (!graphs[variant] && exec == stream) { set_last_error("agent: the caller stream must differ from the capture stream"); return kErrState; }
    if (!graphs[variant]) {
      cudaGraph_t g = nullptr;
      fork.begin_step();
      B200_CHECK_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
      const int c0 = launch_counter();
      int st = issue_step(variant);
      launches[variant] = launch_counter() - c0;
      cudaError_t e = cudaStreamEndCapture(stream, &g);
      if (st != kOk) { if (g) cudaGraphDestroy(g); return st; }
      B200_CHECK_CUDA(e);
      B200_CHECK_CUDA(cudaGraphInstantiate(&graphs[variant], g, 0));
      cudaGraphDestroy(g);
    }
    B200_CHECK_CUDA(cudaGraphLaunch(graphs[variant], exec));
    return kOk;
#endif
  }
  void drop_graphs() {
#ifndef B200_EMU
    for (auto& g : graphs) { if (g) cudaGraphExecDestroy(g); g = nullptr; }
#endif
  }

  int sync() {
    B200_CHECK_CUDA(cudaStreamSynchronize(stream));
    return kOk;
  }
  int sync_exec() {
    B200_CHECK_CUDA(cudaStreamSynchronize(exec));
    return kOk;
  }
};

}  // namespace b200rl
