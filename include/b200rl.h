/* b200rl.h -- C ABI of libb200rl.so: the B200-native (sm_100a) off-policy update path.
 *
 * The reference (seungju-k1m/sac-td3-td7) is pure Python and has NO FFI for this path; its
 * boundary is duck-typed Python.  This header is therefore the boundary a
 * maintainer binds with ctypes from drop-in classes of the same names
 * (sac_td3_td7_b200/rl/...).  Each entry point names the reference interface it serves.
 *
 * Conventions: every function returns 0 on success or a negative b200rl_status; the message
 * of the last failure on the calling thread is b200rl_last_error().  Pointers named *_dev are
 * device pointers owned by the caller; `stream` is a cudaStream_t passed as void* (NULL = the
 * default stream).  Handles own their persistent device state (ring buffer, parameters, Adam
 * moments, activations); nothing is allocated inside a sample / update call.  One caller thread
 * per handle.  fp32 throughout; indices are int64.
 */
#ifndef B200RL_H_
#define B200RL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B200RL_OK = 0,
  B200RL_ERR_INVALID_ARG = -1,
  B200RL_ERR_CUDA = -2,
  B200RL_ERR_NO_DEVICE = -3,
  B200RL_ERR_STATE = -4,
  B200RL_ERR_UNSUPPORTED = -5
} b200rl_status;

const char* b200rl_last_error(void);
int b200rl_version(void);
/* 1 when built by nvcc for sm_100a, 0 for the test-only host emulation build */
int b200rl_is_cuda_build(void);

/* Product loop of the small-tile GEMM kernel (single agents, narrow layers): 0 = fp32 FFMA (default), 1 =
 * error-compensated 3xTF32 on the warp-level tensor-core path (mma.sync), -1 = back to the default / B200RL_MMA.  Returns the previous
 * setting.  Takes effect for launches recorded afterwards (update graphs already captured keep theirs).  A/B
 * measurements and the parity tests of both paths. */
int b200rl_set_mma(int on);

/* Update engine of single agents created AFTERWARDS: 0 = one CUDA-graph node per grouped layer launch (default), 1 = the
 * persistent dataflow kernels of csrc/mega.cuh (a whole phase of the update as one launch, tiles ordered by counters in
 * global memory; parity-green, slower on B200 -- an experiment kept for A/B), -1 = default / B200RL_MEGA.  Returns the
 * previous override. */
int b200rl_set_mega(int on);

/* ---- raw device memory helpers (tests, benches; Python normally uses torch tensors) ---- */
int b200rl_malloc(void** out_dev, uint64_t bytes);
int b200rl_free(void* dev);
int b200rl_memcpy_h2d(void* dst_dev, const void* src_host, uint64_t bytes, void* stream);
int b200rl_memcpy_d2h(void* dst_host, const void* src_dev, uint64_t bytes, void* stream);
int b200rl_stream_sync(void* stream);

/* =============================================================================================
 * Replay memory -- rl/replay_memory/lap.py:12-76 (LAPReplayMemory), simple.py:12-65
 * (SimpleReplayMemory).  A transition is one packed fp32 row whose observation vectors start on 16-byte boundaries:
 *   [ state(S) pad->x4 | next_state(S) pad->x4 | action(A) | reward | notdone | zero pad to a multiple of 4 ]
 * (b200rl_replay_row_stride / b200rl_replay_row_layout give the stride and the field offsets).
 * `groups` > 1 stacks independent seeds; with share_rows the transitions are stored once and
 * only the priority arrays are per seed.
 * ============================================================================================= */
typedef struct b200rl_replay b200rl_replay;

/* lap.py:15-29 / simple.py:15-27 (__init__) */
int b200rl_replay_create(int64_t capacity, int state_dim, int action_dim, int groups, int prioritized, int share_rows,
                         b200rl_replay** out);
int b200rl_replay_destroy(b200rl_replay* h);
int b200rl_replay_row_stride(const b200rl_replay* h);
/* float offsets of {state, action, reward, notdone, next_state} inside a packed row */
int b200rl_replay_row_layout(const b200rl_replay* h, int* offsets5);
/* lap.py:31-43 (append): `count` packed rows (action already a/scale-bias); new rows take
 * priority = running max.  rows_host is host memory; *_device takes a device pointer. */
int b200rl_replay_append(b200rl_replay* h, int group, const float* rows_host, int64_t count, void* stream);
int b200rl_replay_append_device(b200rl_replay* h, int group, const float* rows_dev, int64_t count, void* stream);
/* lap.py:35 `action = action / action_scale - action_bias` on the device: set the per-dimension scale / bias once
 * (miscellaneous.py:58-66), then append rows whose action columns are still in ENVIRONMENT units.  The kernel divides
 * and subtracts in IEEE fp32, which is what numpy computes for the fp32 actions the agents emit.  rows: <= 64 rows in
 * pinned host memory or device memory (read by the kernel itself, in stream order). */
int b200rl_replay_set_action_norm(b200rl_replay* h, const float* scale_host, const float* bias_host);
int b200rl_replay_append_raw(b200rl_replay* h, int group, const float* rows, int64_t count, void* stream);
/* one environment step of a whole population (rl/rollout.py:45 once per stacked seed) as ONE launch: `count` (<= 64)
 * rows for every stacked seed, rows = [groups][count][row_stride] (private row stores) or [count][row_stride] (shared
 * row store), pinned host or device memory; raw_actions != 0 normalises the action columns as above. */
int b200rl_replay_append_all(b200rl_replay* h, const float* rows, int64_t count, int raw_actions, void* stream);
/* rl/rollout.py:45 -> lap.py:31-43 for one environment step, straight from the caller's arrays: n = 1 (one transition
 * for stacked seed `group`) or n = groups (one per stacked seed: obs [n][S], action [n][A], reward [n], next_obs [n][S],
 * notdone [n]; a shared row store keeps seed 0's).  The library packs the row(s) into a pinned staging slot it owns
 * (16 slots, one event each) and ONE kernel copies, normalises (raw_actions != 0), writes priorities and repairs the
 * sum tree. */
int b200rl_replay_append_transitions(b200rl_replay* h, int group, int n, const float* obs_host, const float* action_host, const float* reward_host,
                                     const float* next_obs_host, const float* notdone_host, int raw_actions, void* stream);
/* lap.py:45-64 (sample).  u_dev: [groups][B] uniforms in [0,1) or NULL (on-device Philox).
 * Outputs are [groups][B][dim] contiguous device tensors.  The sampled indices stay in the
 * handle for update_priority (lap.py:61); b200rl_replay_last_indices copies them out. */
int b200rl_replay_sample(b200rl_replay* h, int batch, const float* u_dev, float* state_dev, float* action_dev,
                         float* reward_dev, float* next_state_dev, float* done_dev, void* stream);
int b200rl_replay_last_indices(b200rl_replay* h, int64_t* idx_host /* [groups][batch] */, void* stream);
/* lap.py:66-69 (update_priority): priority_dev [groups][batch]; duplicates: last position wins */
int b200rl_replay_update_priority(b200rl_replay* h, const float* priority_dev, int batch, void* stream);
/* lap.py:71-73 (reset_max_priority) */
int b200rl_replay_reset_max_priority(b200rl_replay* h, void* stream);
/* lap.py:75-76 (__len__ == ptr), .size, .max_priority (run_w_checkpoint.py:171-172); synchronises */
int b200rl_replay_get_state(b200rl_replay* h, int group, int64_t* ptr, int64_t* size, float* max_priority, void* stream);
/* test / checkpoint access */
int b200rl_replay_set_cursor(b200rl_replay* h, int group, int64_t ptr, int64_t size);
int b200rl_replay_set_rng(b200rl_replay* h, uint64_t seed);
int b200rl_replay_read_priority(b200rl_replay* h, int group, int64_t first, int64_t count, float* out_host, void* stream);
int b200rl_replay_write_priority(b200rl_replay* h, int group, int64_t first, int64_t count, const float* in_host, float max_priority, void* stream);
int b200rl_replay_read_rows(b200rl_replay* h, int group, int64_t first, int64_t count, float* out_host, void* stream);

/* =============================================================================================
 * Agents -- rl/agent/td7.py:31 (TD7), td3.py:30 (TD3), sac.py:24 (SAC): constructor
 * hyper-parameters, train_ops, load_state_dict, the acting forward of .sample().
 * Networks are packed blobs: every nn.Linear as W[out][ld] (ld = in rounded up to 4) then b[out].
 * ============================================================================================= */
typedef struct b200rl_agent b200rl_agent;

typedef struct {
  int state_dim, action_dim, hdim, zs_dim, batch, groups;
  double discount, policy_lr, critic_lr;
  int target_update_rate, policy_freq;
  double target_policy_noise, noise_clip;
  int use_lap, keep_grads;
} b200rl_td7_config; /* td7.py:34-48 */

typedef struct {
  int algo; /* 0 = TD3 (td3.py:33-47), 1 = SAC (sac.py:27-41) */
  int state_dim, action_dim, hdim, batch, groups;
  double discount, policy_lr, critic_lr, tau;
  int policy_freq;
  double target_policy_noise, noise_clip;
  double min_log_std, max_log_std, tmp;
  int use_lap, keep_grads;
} b200rl_mlp_config;

int b200rl_td7_create(const b200rl_td7_config* cfg, b200rl_agent** out);
int b200rl_mlp_create(const b200rl_mlp_config* cfg, b200rl_agent** out);
int b200rl_agent_destroy(b200rl_agent* h);

/* network ids.  TD7: 0 encoder, 1 policy, 2 q1, 3 q2, 4 target_q1, 5 target_q2, 6 fixed_encoder,
 * 7 fixed_encoder_target.  TD3/SAC: 0 policy, 1 q1, 2 q2, 3 target_q1, 4 target_q2; SAC also
 * 5 = log-temperature (one float).  which: 0 parameters, 1 exp_avg, 2 exp_avg_sq, 3 gradient. */
int b200rl_agent_net_count(const b200rl_agent* h);
int b200rl_agent_net_info(const b200rl_agent* h, int net, int64_t* numel, int* nlayers);
int b200rl_agent_layer_info(const b200rl_agent* h, int net, int layer, int64_t* w_off, int64_t* b_off, int* out_dim, int* in_dim, int* ld);
/* A critic's first layer consumes cat[state, action] (sale.py:114, mlp.py:86-88): its weight rows hold the state columns,
 * zero padding up to the next multiple of 4 floats, then the action columns.  *split = number of state columns when the
 * layer is laid out that way (logical column c >= split lives at c + round_up(split, 4) - split), 0 for a dense row. */
int b200rl_agent_layer_split(const b200rl_agent* h, int net, int layer, int* split);
/* group = -1 addresses every stacked seed at once: [groups][numel] contiguous (population checkpoints: one copy per
 * tensor kind instead of one per seed); the TD7 value-range calls take group = -1 the same way ([groups][4]). */
int b200rl_agent_read(b200rl_agent* h, int net, int which, int group, float* out_host, void* stream);
int b200rl_agent_write(b200rl_agent* h, int net, int which, int group, const float* in_host, void* stream);
/* td7.py:114-125 / td3.py:90-96 / sac.py:101-107 (load_state_dict(other_agent)); full=1 also copies
 * optimizer moments, counters and value range (deepcopy) */
int b200rl_agent_copy_from(b200rl_agent* dst, const b200rl_agent* src, int full, void* stream);
/* counters: n_runs and Adam step counts (host mirrors); value range for TD7 (device read, syncs) */
int b200rl_agent_get_counters(b200rl_agent* h, int64_t* n_runs, int64_t* adam_steps /*[3]*/);
int b200rl_agent_set_counters(b200rl_agent* h, int64_t n_runs, const int64_t* adam_steps /*[3]*/, void* stream);
int b200rl_td7_value_range(b200rl_agent* h, int group, float* out4_host, void* stream);
int b200rl_td7_set_value_range(b200rl_agent* h, int group, const float* in4_host, void* stream);
int b200rl_agent_set_rng(b200rl_agent* h, uint64_t seed);

/* fixed device input buffers of the engine ([groups] x {ss[2B][S], action[B][A], reward[B], notdone[B]});
 * group_stride is in floats.  replay.sample can write here directly (zero-copy train_ops). */
int b200rl_agent_io(b200rl_agent* h, float** ss_dev, float** action_dev, float** reward_dev, float** notdone_dev, int64_t* group_stride);
/* row strides (floats) of the ss and action buffers above: widths rounded up to 4 floats, zero padding */
int b200rl_agent_io_ld(b200rl_agent* h, int* ld_state, int* ld_action);
/* lap.py:45-64 / simple.py sample(batch) straight into the engine's input buffers (what b200rl_agent_update_n does per
 * step); the tensors returned by the Python sample() alias these buffers, so train_ops needs no staging copy */
int b200rl_replay_sample_to_agent(b200rl_replay* h, b200rl_agent* agent, int batch, const float* u_dev, void* stream);
/* copy a caller batch (five device tensors, rl/utils/annotation.py:23 BATCH) into the input buffers */
int b200rl_agent_stage_batch(b200rl_agent* h, int group, const float* state_dev, const float* action_dev, const float* reward_dev,
                             const float* next_state_dev, const float* done_dev, void* stream);
/* train_ops(batch, replay_buffer) on the staged batch: td7.py:287-332 / td3.py:206-242 / sac.py:251-295.
 * replay may be NULL when use_lap == 0.  noise_dev: injected N(0,1) draws ([groups][B][A]; SAC:
 * [2][groups][B][A]) or NULL for on-device Philox. */
int b200rl_agent_update(b200rl_agent* h, b200rl_replay* replay, const float* noise_dev, void* stream);
/* run_train_ops(n_ops=n) (rl/runner/run.py:87-96): n x { replay.sample(batch) ; train_ops } entirely on
 * the device, no host round trip */
int b200rl_agent_update_n(b200rl_agent* h, b200rl_replay* replay, int n, void* stream);
/* the info dict of update number `step` (1-based count of train_ops calls) while it is still in the
 * 1024-entry ring; columns: TD7 {train/encoder, train/q_fn, train/policy}; TD3 {train/q_fn,
 * train/policy, norm/policy}; SAC {train/q_fn, tmp, norm/tmp, train/policy, train/tmp, entropy}.
 * Synchronises `stream`.  group = -1 reads every stacked seed: out_host [groups][count][cols]. */
int b200rl_agent_info_cols(const b200rl_agent* h);
int b200rl_agent_read_info(b200rl_agent* h, int64_t first_step, int count, int group, float* out_host, void* stream);
/* acting forward of .sample() for one observation per stacked seed: obs_dev [groups][S] -> out_dev
 * [groups][A] (TD7/TD3: tanh policy output; SAC: [groups][2A] mean | raw log_std) --
 * td7.py:158-162, td3.py:137-141, sac.py:154-159 */
int b200rl_agent_act(b200rl_agent* h, const float* obs_dev, float* out_dev, void* stream);
/* the same forward host to host (rl/rollout.py:42 -> td7.py:141-156): obs_host [groups][S] is copied into the engine's
 * pinned slot, ONE kernel (a thread-block cluster per stacked seed) reads it over the bus, writes the action(s) into a
 * pinned mailbox followed by a sequence tag, and the call returns as soon as the tag arrives -- no stream
 * synchronisation, no copy-engine round trips.  out_host [groups][A] (SAC: [groups][2A]). */
int b200rl_agent_act_host(b200rl_agent* h, const float* obs_host, float* out_host, void* stream);
/* parity probes of the last update: which = 0 q_target[B], 1 priority[B], 2 q1[B], 3 q2[B], 4 pi_action[B][A],
 * 5 the N(0,1) draws of the update [B][A] (target smoothing / SAC critic draw), 6 SAC's actor draw [B][A] */
int b200rl_agent_debug_read(b200rl_agent* h, int which, int group, float* out_host, void* stream);
/* the info value `col` of seed 0 for update `step` as soon as the kernel that computes it has run (it is posted to a
 * pinned host mailbox), without waiting for the rest of the update -- what `float(loss)` (td7.py:303,315,324) needs.
 * Returns 0 and the value, or 1 when `step` is not the newest update / the value was not posted: read the ring then. */
int b200rl_agent_wait_info(b200rl_agent* h, int64_t step, int col, float* out);
/* number of kernel launches issued by the last update (bench: gpu_launches) */
int b200rl_agent_launches_per_update(b200rl_agent* h, int actor_step);

/* =============================================================================================
 * Op-level entry points (kernel parity tests; rl/nn forward modules)
 * ============================================================================================= */
typedef struct {
  const float* A; const float* B;
  int64_t a_gs, b_gs;
  int a_ld, b_ld, len, b_ones;
} b200rl_gemm_seg;

typedef struct {
  float w1, beta2, w2, bc2_sqrt, neg_step, eps;
  int step, pad;
} b200rl_adam_scalars;

typedef struct {
  b200rl_gemm_seg seg[3];
  int nseg, I, J;
  float* C; int64_t c_gs; int c_ld; int act;
  const float* bias; int64_t bias_gs;
  const float* aux; int64_t aux_gs; int aux_ld; int accumulate;
  float* P; float* M; float* V; int64_t p_gs; int p_ld;
  float noise_sigma, noise_clip;
  const b200rl_adam_scalars* adam;
  int tile_begin, tiles_j;
} b200rl_gemm_problem;

/* mode: 0 forward (Y = X W^T + b, activation), 1 dgrad (dX = dZ W, times act'), 2 wgrad (dW = dZ^T X,
 * optional fused Adam).  act: 0 none, 1 ReLU, 2 ELU, 3 tanh, 4 tanh + clipped noise + clamp. */
int b200rl_op_gemm(int mode, const b200rl_gemm_problem* problems_host, int nproblems, int groups, void* stream);
/* same launch repeated `iters` times after `warmup` launches, timed with CUDA events on `stream`;
 * returns the mean device time of one launch in milliseconds (bench.py roofline) */
int b200rl_op_gemm_timed(int mode, const b200rl_gemm_problem* problems_host, int nproblems, int groups, int warmup, int iters, float* ms_per_launch,
                         void* stream);
/* mean device time (ms) of the two sampler kernels for the current buffer contents: [0] tile sums, [1] search+gather */
int b200rl_replay_time_sample(b200rl_replay* h, int batch, int warmup, int iters, float* ms_out2, void* stream);
/* rl/nn/sale.py:11-13 AvgL1Norm forward (y = x / max(mean|x|, 1e-8)) and its backward */
int b200rl_op_avgl1norm(int backward, const float* x_dev, const float* dy_dev, float* y_dev, int rows, int n, void* stream);
/* torch.optim.Adam single-tensor step `step` (1-based) on n floats */
int b200rl_op_adam(float* p_dev, float* m_dev, float* v_dev, const float* g_dev, int64_t n, double lr, double beta1, double beta2, double eps,
                   int step, void* stream);
/* td3.py:195-204 / sac.py:244-249: target = tau*src + target*(1-tau) */
int b200rl_op_polyak(float* target_dev, const float* src_dev, int64_t n, double tau, void* stream);
/* sac.py:154-172: net_out [B][2A], eps [B][A] -> action [B][A], logp [B] */
int b200rl_op_sac_sample(const float* net_out_dev, const float* eps_dev, float* action_dev, float* logp_dev, int B, int A, double min_log_std,
                         double max_log_std, void* stream);
int b200rl_op_sac_sample_bwd(const float* net_out_dev, const float* eps_dev, const float* action_dev, const float* daction_dev, const float* dlogp_dev,
                             float* dnet_out_dev, int B, int A, double min_log_std, double max_log_std, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200RL_H_ */
