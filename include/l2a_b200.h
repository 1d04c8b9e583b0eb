/*
 * l2a_b200.h -- C ABI of the B200-native MPC planning engine (libl2a_b200.so).
 *
 * This is the drop-in boundary for the one hot path of iclavera/learning_to_adapt:
 *   MPCController.get_actions  ->  {random shooting | CEM}  ->  H x dynamics_model.predict  ->
 *   env.reward  ->  argmax,   plus GrBAL's one-gradient-step MetaMLPDynamicsModel.adapt.
 *
 * The reference has no FFI layer (pure Python over TF1); each entry point below names the reference
 * Python call site (file:line, relative to the upstream tree) it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types in any signature.
 *   - every function returns 0 on success or a negative l2a_status; the message of the last failure
 *     on the calling thread is returned by l2a_last_error().  No exceptions cross the boundary.
 *   - unless stated otherwise every data pointer is a DEVICE pointer to float32 (or int32) memory owned by
 *     the caller (e.g. torch tensor .data_ptr()); the library only borrows it for the duration of the
 *     stream-ordered call.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - calls on one l2a_ctx are not re-entrant; use one ctx per (process, device).
 */
#ifndef L2A_B200_H
#define L2A_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L2A_MAX_LAYERS 8          /* dense layers incl. the output layer */

typedef enum {
  L2A_OK = 0,
  L2A_ERR_INVALID = -1,           /* bad argument / unsupported shape */
  L2A_ERR_CUDA = -2,              /* a CUDA runtime call failed */
  L2A_ERR_UNSUPPORTED = -3,       /* no kernel for this configuration (there is NO CPU fallback) */
  L2A_ERR_NO_DEVICE = -4
} l2a_status;

/* reward families = the closed forms of the reference env classes */
typedef enum {
  L2A_REWARD_HALF_CHEETAH = 0,    /* envs/half_cheetah_env.py:58-65 (+ _blocks_env.py:56-63, _hfield_env.py:59-66) */
  L2A_REWARD_ANT = 1,             /* envs/ant_env.py:56-66 */
  L2A_REWARD_ARM = 2              /* envs/arm_7dof_env.py:91-99 */
} l2a_reward_kind;

/* which weight set a candidate row is stepped through */
typedef enum {
  L2A_SETS_SHARED = 0,            /* every row uses set `first_set`            (mlp_dynamics.py:204-222) */
  L2A_SETS_PER_ENV = 1,           /* rows of env k use set first_set + k       (meta_mlp_dynamics.py:296-306) */
  L2A_SETS_ENSEMBLE_MEAN = 2      /* every row goes through n_sets sets, deltas averaged (BASELINE "ensemble") */
} l2a_set_mode;

typedef enum {
  L2A_KERNEL_AUTO = 0,            /* tcgen05 when every hidden width is a multiple of 128, else SIMT */
  L2A_KERNEL_SIMT = 1,            /* fp32 FFMA kernel, any shape */
  L2A_KERNEL_TCGEN05 = 2,         /* split-bf16 (3-pass) tcgen05.mma kernel, fp32 accumulate in TMEM */
  L2A_KERNEL_TCGEN05_PAIR = 3     /* the same on CTA pairs (tcgen05.mma.cta_group::2, M = 256): hidden widths multiples of 256 */
} l2a_kernel_choice;

typedef struct l2a_ctx l2a_ctx;
typedef struct l2a_model l2a_model;

/* Shape of the dynamics MLP: [obs_dim + act_dim] -> hidden[0] -> ... -> hidden[n_hidden-1] -> obs_dim,
 * ReLU hidden layers, linear output (dynamics/core/utils.py:75-142).  n_sets = weight sets resident at once
 * (set 0 = prior theta; GrBAL adapted sets / ensemble members use the others). */
typedef struct {
  int32_t obs_dim;
  int32_t act_dim;
  int32_t n_hidden;
  int32_t hidden[L2A_MAX_LAYERS - 1];
  int32_t n_sets;
} l2a_mlp_desc;

typedef struct {
  int32_t n_candidates;           /* N: candidates per env                        (mpc_controller.py:109) */
  int32_t n_envs;                 /* m: observations planned in this call         (mpc_controller.py:110) */
  int32_t horizon;                /* H                                            (mpc_controller.py:111) */
  int32_t set_mode;               /* l2a_set_mode */
  int32_t first_set;
  int32_t n_sets;                 /* used by L2A_SETS_ENSEMBLE_MEAN */
  int32_t reward_kind;            /* l2a_reward_kind */
  float   dt;                     /* env.dt */
  int64_t act_stride_t;           /* element strides of the candidate-action tensor: action j of row r at */
  int64_t act_stride_row;         /*   step t is actions[t*act_stride_t + r*act_stride_row + j]           */
  int32_t kernel;                 /* l2a_kernel_choice */
  int32_t reserved;
} l2a_rollout_params;

const char* l2a_last_error(void);
int l2a_version(void);

/* ---- context (replaces tf.Session creation, trainers/mb_trainer.py:46-48) ------------------------------ */
/* Threading / streams: a context is NOT re-entrant and serves ONE stream at a time -- its per-env reduction workspace, the
 * ensemble exchange scratch and the adapt workspace are shared by every launch made through it, so two l2a_rollout / l2a_adapt
 * calls in flight on different streams of the same context would race.  Use one context per (process, device, stream); the
 * host-buffer plans (l2a_plan_*) run on their own private stream and synchronise before returning. */
int l2a_ctx_create(int device, l2a_ctx** out);
int l2a_ctx_destroy(l2a_ctx* ctx);
/* number of this library's kernels launched on the ctx since creation (bench.py's gpu_launches) */
int64_t l2a_ctx_launch_count(const l2a_ctx* ctx);

/* ---- model storage (replaces the TF variables of dynamics/core/layers.py:142-171) --------------------- */
int l2a_model_create(l2a_ctx* ctx, const l2a_mlp_desc* desc, l2a_model** out);
int l2a_model_destroy(l2a_ctx* ctx, l2a_model* model);
/* W[l]: device fp32 [in_l, out_l] row-major ("hidden_l/kernel"), b[l]: device fp32 [out_l]; W and b are HOST arrays
 * of n_hidden+1 device pointers.  Values are copied (and re-tiled for the tensor-core path) before return of the
 * stream-ordered work.  Replaces MLP.set_params (layers.py:81-101). */
int l2a_model_set_params(l2a_ctx* ctx, l2a_model* model, int set, const float* const* W, const float* const* b,
                         void* stream);
/* copies set `set` out into caller-provided device buffers (MLP.get_param_values, layers.py:71-79) */
int l2a_model_get_params(l2a_ctx* ctx, l2a_model* model, int set, float* const* W, float* const* b, void* stream);
/* Resident parameters in place (on-device training; replaces the sess.run(train_op) variable updates of
 * mlp_dynamics.py:150-160 / meta_mlp_dynamics.py:218-226): device pointer to the fp32 block of weight set `set`
 * (floats_per_set floats; kernel of layer l at w_off[l] as [in_l, out_l] row-major, bias at b_off[l]; w_off / b_off: HOST
 * int32[n_hidden+1] or NULL).  The caller updates the values on the device and then calls l2a_model_refresh, which re-tiles
 * sets [first_set, first_set + n_sets) for the tensor-core rollout. */
int l2a_model_param_block(l2a_ctx* ctx, l2a_model* model, int set, float** ptr_out, int64_t* floats_per_set, int32_t* w_off,
                          int32_t* b_off);
int l2a_model_refresh(l2a_ctx* ctx, l2a_model* model, int first_set, int n_sets, void* stream);
/* mean / denominators of the reference's normalize()/denormalize() (mlp_dynamics.py:265-270):
 * x_n = (x - mean) / den with den = std + 1e-10 already folded in by the caller (in float64), and
 * delta = y * delta_scale + delta_mean with delta_scale = std_delta + 1e-10.  Six device fp32 arrays. */
int l2a_model_set_normalization(l2a_ctx* ctx, l2a_model* model, const float* obs_mean, const float* obs_den,
                                const float* act_mean, const float* act_den, const float* delta_mean,
                                const float* delta_scale, void* stream);

/* ---- K1: fused H-step rollout + reward + per-env argmax ------------------------------------------------
 * Replaces the body of MPCController.get_rs_action (policies/mpc_controller.py:116-129): H x predict
 * (mlp_dynamics.py:204-222 / meta_mlp_dynamics.py:276-306), env.reward, discounted accumulate, np.argmax.
 *   obs0          [m, D]       observation per env
 *   actions       strided      candidate actions (see l2a_rollout_params); row r belongs to env r / N
 *   discount_pow  [H]          discount**t, computed by the caller in float64 and rounded
 *   returns       [m, N] or NULL
 *   best_ret [m], best_idx [m] (int32, first max wins like np.argmax), best_act [m, A] = actions at t=0 of the winner */
int l2a_rollout(l2a_ctx* ctx, l2a_model* model, const l2a_rollout_params* p, const float* obs0,
                const float* actions, const float* discount_pow, float* returns, float* best_ret,
                int32_t* best_idx, float* best_act, void* stream);

/* ---- host-buffer planning call: the reference-facing call --------------------------------------------------------------
 * Replaces MPCController.get_actions for random shooting as its caller sees it (policies/mpc_controller.py:59-65, 108-129,
 * incl. get_random_action :67-69): HOST observations in, HOST actions out.  A plan fixes p (N, m, H, weight sets, reward
 * family, dt, kernel; the action strides are ignored -- the plan owns the [H, m*N, A] candidate tensor), the discount, the
 * action bounds low/high (HOST float32 [A]) and the Philox seed.
 * l2a_plan_run: obs HOST float64 [m, D] -> act_out HOST float64 [m, A] (+ optional ret_out float32 [m], idx_out int32 [m]).
 *   Per call: H2D copy of obs from pinned memory -> candidate sampling U[low, high) on the device (Philox4x32-10, counter =
 *   (element, call index)) -> K1 -> D2H of the result to pinned memory; after the first call the four stream operations
 *   are one captured CUDA graph (env L2A_NO_GRAPH=1 issues them individually).  Work queued on `stream` before the call is
 *   waited for; the call returns when the outputs are in the host buffers. */
typedef struct l2a_plan l2a_plan;
int l2a_plan_create(l2a_ctx* ctx, l2a_model* model, const l2a_rollout_params* p, float discount, const float* low,
                    const float* high, uint64_t seed, l2a_plan** out);
int l2a_plan_run(l2a_ctx* ctx, l2a_plan* plan, const double* obs, double* act_out, float* ret_out, int32_t* idx_out,
                 void* stream);
int l2a_plan_destroy(l2a_ctx* ctx, l2a_plan* plan);
int l2a_plan_uses_graph(const l2a_plan* plan);   /* 1 once the call sequence has been captured and is being replayed */
/* the candidate tensor the most recent l2a_plan_run drew, [H, m*N, A] float32, to a HOST buffer (tests / diagnostics) */
int l2a_plan_copy_candidates(l2a_ctx* ctx, l2a_plan* plan, float* host_out);
/* CEM plans: the returns [m, N] float32 of the LAST iteration's rollout (what mpc_controller.py:100 holds), to a HOST buffer */
int l2a_plan_copy_returns(l2a_ctx* ctx, l2a_plan* plan, float* host_out);
/* bytes the plan copies host->device / device->host inside every l2a_plan_run[_ex] call (its pinned input / output blocks) */
int l2a_plan_io_bytes(const l2a_plan* plan, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

/* ---- the general host-buffer planning call ------------------------------------------------------------------------------
 * l2a_plan_create_ex / l2a_plan_run_ex extend the call above along two axes, still ONE C call and one CUDA-graph replay per
 * MPCController.get_actions:
 *  sampler  L2A_SAMPLER_PHILOX   device Philox stream (throughput mode, as l2a_plan_create).
 *           L2A_SAMPLER_MT19937  the reference's OWN draw (policies/mpc_controller.py:67-69,114: np.random.uniform(low, high,
 *                                (H*N*m, A)) from numpy's global MT19937 stream), regenerated bit-exactly on the device from the
 *                                generator state the caller passes in (np.random.get_state(): 624 key words + position);
 *                                the advanced state is returned for np.random.set_state().  No host RNG, no candidate upload;
 *                                act_out carries the float64 candidate values the reference would return.
 *  shard    shard_world > 1: this process rolls candidates [shard_offset, shard_offset + p->n_candidates) of every env's
 *           n_candidates_total (one process per GPU); the per-env (return, global index, action) records are
 *           exchanged over peer memory (NVLink) inside the same graph: every rank writes its record into every peer's exchange
 *           buffer and raises a sequence flag, waits for all flags in its own buffer and selects with np.argmax semantics over
 *           the concatenated candidates.  With MT19937 every rank regenerates the reference's full stream and materialises
 *           only its slice, so a seeded G-GPU run returns exactly the single-GPU / reference actions.
 *           Peer plumbing: l2a_plan_exchange_buffer -> l2a_ipc_get_handle -> (host exchanges the 64-byte handles, e.g.
 *           torch.distributed.all_gather_object) -> l2a_ipc_open_handle -> l2a_plan_attach_peers.  Every rank must make the
 *           same sequence of l2a_plan_run_ex calls (collective semantics). */
enum { L2A_SAMPLER_PHILOX = 0, L2A_SAMPLER_MT19937 = 1 };
/*  planner  L2A_PLANNER_RS   random shooting (policies/mpc_controller.py:108-129).
 *           L2A_PLANNER_CEM  the cross-entropy planner (:71-106), ALL cem_iters iterations inside the one call / one graph:
 *                            per iteration  normal draw z [N, m, H*A] (MT19937: numpy's legacy polar Box-Muller stream incl. its
 *                            cached second value, float64; PHILOX: Box-Muller on the device stream) -> a = mean + z*std, clipped
 *                            copy (:86-87) -> K1 on the UNclipped samples viewed as (N*m, H, A) (:88-99) -> rank -> elite refit
 *                            (:101-104; cem_compat != 0 reproduces the reference's rank-mask defect, 0 = true top-k).  mean = 0,
 *                            std = 1 at the start of every call (:79-80).  act_out = float64 first action of the last iteration's
 *                            best row (:106).  cem_alpha = 0 gives the RNN controller's variant (rnn_mpc_controller.py:107-108).
 *                            Not combinable with shard_world > 1. */
enum { L2A_PLANNER_RS = 0, L2A_PLANNER_CEM = 1 };
typedef struct {
  int32_t sampler;              /* L2A_SAMPLER_* */
  int32_t shard_rank;           /* 0 .. shard_world-1 */
  int32_t shard_world;          /* <= 1: not sharded */
  int32_t n_candidates_total;   /* N of the whole job (0 = p->n_candidates) */
  int64_t shard_offset;         /* global index of this rank's first candidate of every env */
  uint64_t seed;                /* Philox key (each rank's stream is offset by its rank) */
  int32_t planner;              /* L2A_PLANNER_* */
  int32_t cem_iters;            /* num_cem_iters */
  int32_t cem_num_elites;       /* max(int(N * percent_elites), 1) (:78) */
  int32_t cem_compat;           /* 1 = the reference's elite mask */
  double cem_alpha;             /* mean' = alpha*mean + (1-alpha)*mean(elites) (:103) */
} l2a_plan_opts;
typedef struct {
  uint32_t* mt_key;             /* in/out HOST uint32[624]: MT19937 key (np.random.get_state()[1])          -- MT19937 only */
  int32_t* mt_pos;              /* in/out HOST int32:       position in [0, 624] (np.random.get_state()[2]) -- MT19937 only */
  double* act_out;              /* out HOST float64 [m, A]: chosen first actions */
  float* ret_out;               /* out HOST float32 [m] or NULL: return of the winner */
  int64_t* idx_out;             /* out HOST int64 [m] or NULL: global candidate index of the winner */
  int32_t* mt_has_gauss;        /* in/out HOST int32:  np.random.get_state()[3] -- MT19937 + CEM only */
  double* mt_cached;            /* in/out HOST float64: np.random.get_state()[4] -- MT19937 + CEM only */
  double* cem_mean_out;         /* out HOST float64 [m, H*A] or NULL: CEM mean after the last refit */
  double* cem_std_out;          /* out HOST float64 [m, H*A] or NULL */
  int32_t flags;                /* L2A_PLAN_* bits of THIS call (need l2a_plan_attach_window) */
  int32_t reserved;
} l2a_plan_io;
/*  GrBAL env step (samplers/sampler.py:81-91) inside the same call / graph, for a plan with an attached adaptation window:
 *   L2A_PLAN_ADAPT  before planning: gather the envs' last M transitions from the device window, K2 adapt from weight set
 *                   src_set into sets dst_first_set.. (= dynamics_model.switch_to_pre_adapt(); dynamics_model.adapt(...)) and
 *                   re-tile them; the plan's own p->set_mode / first_set say which sets the planner then uses.
 *   L2A_PLAN_PUSH   after planning: append (observation, chosen action) of every env to the window in float64
 *                   (running_paths[idx]["observations"/"actions"].append, :109-110); the window's host length mirror follows. */
enum { L2A_PLAN_ADAPT = 1, L2A_PLAN_PUSH = 2 };
int l2a_plan_create_ex(l2a_ctx* ctx, l2a_model* model, const l2a_rollout_params* p, double discount, const double* low,
                       const double* high, const l2a_plan_opts* opts, l2a_plan** out);
int l2a_plan_run_ex(l2a_ctx* ctx, l2a_plan* plan, const double* obs, l2a_plan_io* io, void* stream);
int l2a_plan_exchange_buffer(l2a_plan* plan, void** ptr_out, uint64_t* bytes_out);
int l2a_plan_attach_peers(l2a_ctx* ctx, l2a_plan* plan, void* const* peer_bufs /* HOST array [shard_world] of device pointers */);
/* the exchange step alone on DEVICE-resident per-rank results of l2a_rollout (best_ret [m], best_idx [m] local, best_act [m, A]):
 * final_rec_out [m, 2 + A] float64 DEVICE = (return, global candidate index, action) of the winner over all ranks, on every rank */
int l2a_plan_exchange_resident(l2a_ctx* ctx, l2a_plan* plan, const float* best_ret, const int32_t* best_idx, const float* best_act,
                               double* final_rec_out, void* stream);
int l2a_ipc_get_handle(l2a_ctx* ctx, void* dev_ptr, void* handle64_out);
int l2a_ipc_open_handle(l2a_ctx* ctx, const void* handle64, void** dev_ptr_out);
int l2a_ipc_close_handle(l2a_ctx* ctx, void* dev_ptr);

/* Candidate sampling alone (DEVICE pointers): out[rows, A] = U[low, high) from Philox4x32-10 keyed by seed, counter =
 * (element block, call_index).  What l2a_plan_run uses internally; the multi-GPU candidate shard calls it with a per-rank seed.
 * Replaces MPCController.get_random_action (policies/mpc_controller.py:67-69) in throughput mode. */
int l2a_sample_uniform(l2a_ctx* ctx, const float* low, const float* high, float* out, int64_t rows, int A, uint64_t seed,
                       uint64_t call_index, void* stream);

/* ---- K4: one dynamics step (API compatibility) --------------------------------------------------------
 * Replaces (Meta)MLPDynamicsModel.predict (mlp_dynamics.py:204-222, meta_mlp_dynamics.py:276-306).
 * obs [n, D], act [n, A] raw (un-normalised).  set_mode SHARED: all rows use first_set; PER_ENV: n must be
 * divisible by n_sets and row chunk k uses set first_set + k; ENSEMBLE_MEAN: mean over n_sets.
 * Writes the denormalised delta [n, D] and/or next_obs = obs + delta [n, D] (either may be NULL). */
int l2a_predict(l2a_ctx* ctx, l2a_model* model, int set_mode, int first_set, int n_sets, const float* obs,
                const float* act, int n, float* delta_out, float* next_out, int kernel, void* stream);

/* ---- K2: GrBAL one-step inner adaptation ---------------------------------------------------------------
 * Replaces MetaMLPDynamicsModel.adapt (meta_mlp_dynamics.py:321-345) + _adapt_sym (409-421): for task k < K,
 *   theta'_k = theta_src - inner_lr * d/dtheta mean_{M x D}( (target_k - f_theta(x_k))^2 ).
 * x [K, M, D+A], target [K, M, D]: already normalised by the caller exactly as the reference does on the
 * host (:334-339).  Result goes to weight sets dst_first_set .. dst_first_set+K-1. */
int l2a_adapt(l2a_ctx* ctx, l2a_model* model, const float* x, const float* target, int K, int M, float inner_lr,
              int src_set, int dst_first_set, void* stream);

/* ---- f3: device-resident adaptation window ----------------------------------------------------
 * Replaces the running-path list slicing of Sampler.obtain_samples (samplers/sampler.py:82-90: obs[-M-1:-1], act[-M-1:-1],
 * obs[-M:] per env, np.stack, then the float64 normalisation + float32 feed of meta_mlp_dynamics.py:334-339) with a ring of
 * the last M+1 (observation, action) pairs per env in HBM.  All calls are stream-ordered and launch only kernels (no host
 * synchronisation), so window_push -> adapt_from_window -> l2a_rollout form one dependent chain on `stream`.
 *   l2a_window_create:  n_envs rings of M+1 pairs; float64 storage (what the env returns).
 *   l2a_window_set_normalization: HOST float64 means / stds (NOT std + 1e-10: the kernel adds it like mlp_dynamics.py:266).
 *   l2a_window_push:    append one (observation, action) pair per env (sampler.py:109-110); obs [n_envs, D], act [n_envs, A]
 *                       float64 DEVICE pointers.
 *   l2a_window_reset:   path of `env` ended (sampler.py:128); env < 0 resets every env.
 *   l2a_window_length:  host mirror of env's path length (appends since the last reset) -- the reference's
 *                       `len(running_paths[0]['observations']) > M + 1` test (sampler.py:82) reads this for env 0.
 *   l2a_window_gather:  write the normalised windows x [n_envs, M, D+A], target [n_envs, M, D] (float32, DEVICE) -- the
 *                       arrays the reference feeds obs_ph|act_ph and delta_ph with (meta_mlp_dynamics.py:334-345).
 *   l2a_adapt_from_window: l2a_window_gather into the window's own buffers, then exactly l2a_adapt with K = n_envs.
 *                       Both: L2A_ERR_INVALID if any env's path is shorter than M+1 (the reference's np.stack would be ragged). */
typedef struct l2a_window l2a_window;
int l2a_window_create(l2a_ctx* ctx, int n_envs, int M, int obs_dim, int act_dim, l2a_window** out);
int l2a_window_destroy(l2a_ctx* ctx, l2a_window* w);
int l2a_window_set_normalization(l2a_ctx* ctx, l2a_window* w, const double* obs_mean, const double* obs_std,
                                 const double* act_mean, const double* act_std, const double* delta_mean,
                                 const double* delta_std, void* stream);
int l2a_window_push(l2a_ctx* ctx, l2a_window* w, const double* obs, const double* act, void* stream);
int l2a_window_reset(l2a_ctx* ctx, l2a_window* w, int env, void* stream);
int l2a_window_length(l2a_ctx* ctx, const l2a_window* w, int env);
int l2a_window_gather(l2a_ctx* ctx, l2a_window* w, float* x, float* target, void* stream);
int l2a_adapt_from_window(l2a_ctx* ctx, l2a_model* model, l2a_window* w, float inner_lr, int src_set, int dst_first_set,
                          void* stream);
/* attach the window to a host-buffer plan (see L2A_PLAN_ADAPT / L2A_PLAN_PUSH): window envs / dims must match the plan's */
int l2a_plan_attach_window(l2a_ctx* ctx, l2a_plan* plan, l2a_window* w, float inner_lr, int src_set, int dst_first_set);

/* ---- K1c: CEM sampling / refit (policies/mpc_controller.py:84-104) -------------------------------------
 * l2a_cem_sample: a = mean + z*std (:86) -> samples [n, m, H*A] fp32 (rolled out UNclipped, :88-89) and
 *   clipped copy (:87).  mean/std are float64 [m, H*A]; z fp32 [n, m, H*A]; clip_low/high fp32 [H*A].
 * l2a_cem_refit: elite selection + mean/std update (:101-104) from returns [m, n].
 *   compat != 0 reproduces the reference's rank-mask defect (:101); compat == 0 takes the true top-k (m == 1 only:
 *   L2A_ERR_UNSUPPORTED for m > 1, where the reference's sample -> env layout is itself inconsistent).
 *   rank_scratch: int32 [m, n]. */
int l2a_cem_sample(l2a_ctx* ctx, const float* z, const double* mean, const double* std, const float* clip_low,
                   const float* clip_high, int n, int m, int ha, float* samples, float* clipped, void* stream);
int l2a_cem_refit(l2a_ctx* ctx, const float* returns, const float* clipped, int n, int m, int ha, int num_elites,
                  double alpha, int compat, int32_t* rank_scratch, double* mean, double* std, void* stream);

/* ---- ReBAL: recurrent (single-layer LSTM) dynamics model and its planner rollout -----------
 * Replaces RNNDynamicsModel.predict (dynamics/rnn_dynamics.py:233-252; cell = tf.nn.rnn_cell.LSTMCell, forget_bias 1, tanh,
 * dynamics/core/utils.py:193-198, + dense output :227-234) and the H-step loop of RNNMPCController.get_rs_action
 * (policies/rnn_mpc_controller.py:112-134) incl. repeat_hidden (:165-187): every candidate of env g starts from
 * (hidden_c[g], hidden_h[g]).  cell_kernel [(D+A+Hs), 4*Hs] with gate order i, j, f, o; out_kernel [Hs, D].  fp32 SIMT. */
typedef struct l2a_rnn_model l2a_rnn_model;
int l2a_rnn_model_create(l2a_ctx* ctx, int obs_dim, int act_dim, int hidden, l2a_rnn_model** out);
int l2a_rnn_model_destroy(l2a_ctx* ctx, l2a_rnn_model* model);
int l2a_rnn_model_set_params(l2a_ctx* ctx, l2a_rnn_model* model, const float* cell_kernel, const float* cell_bias,
                             const float* out_kernel, const float* out_bias, void* stream);
int l2a_rnn_model_set_normalization(l2a_ctx* ctx, l2a_rnn_model* model, const float* obs_mean, const float* obs_den,
                                    const float* act_mean, const float* act_den, const float* delta_mean,
                                    const float* delta_scale, void* stream);
/* p->set_mode / first_set / n_sets are ignored; p->kernel: AUTO = tcgen05 when hidden is 128 or 256 (split-bf16 MMAs, gates of a
 * unit combined in one thread), else the fp32 SIMT kernel.  hidden_c, hidden_h: [m, Hs]. */
int l2a_rnn_rollout(l2a_ctx* ctx, l2a_rnn_model* model, const l2a_rollout_params* p, const float* obs0, const float* hidden_c,
                    const float* hidden_h, const float* actions, const float* discount_pow, float* returns, float* best_ret,
                    int32_t* best_idx, float* best_act, void* stream);
/* one step for n rows: delta_out [n, D] (denormalised), next state c_out, h_out [n, Hs] */
int l2a_rnn_predict(l2a_ctx* ctx, l2a_rnn_model* model, const float* obs, const float* act, const float* hidden_c,
                    const float* hidden_h, int n, float* delta_out, float* c_out, float* h_out, void* stream);

/* ---- K3: multi-GPU candidate shard (new; the reference is single-process) -------------------------
 * Each rank rolls its slice of every env's candidates with l2a_rollout; the only exchange is ONE all-gather (NCCL,
 * issued by the host through torch.distributed) of the packed per-env triple.
 * l2a_shard_pack  : packed[m, 3+A] = (best_ret, (best_idx + idx_offset) >> 16, & 0xFFFF, best_act[A])   (fp32, exact)
 * l2a_shard_select: gathered[G, m, 3+A] -> winner per env, np.argmax semantics over the concatenated candidates
 *                   (policies/mpc_controller.py:128-129): max return, NaN wins, ties -> lowest global index. */
int l2a_shard_pack(l2a_ctx* ctx, const float* best_ret, const int32_t* best_idx, const float* best_act, int64_t idx_offset,
                   int m, int A, float* packed, void* stream);
int l2a_shard_select(l2a_ctx* ctx, const float* gathered, int G, int m, int A, float* best_ret, int64_t* best_idx,
                     float* best_act, void* stream);

/* Host-only (no device needed): the tensor-core tiling a model of this shape gets.  out8[0] = 1 if the tcgen05 rollout supports
 * the shape (hidden widths multiples of 128 and <= 512, obs_dim <= 48, act_dim <= 16, pad8(obs_dim) + act_dim <= 64), then
 * hidden_pairs (32 KB ring stages of the hidden layers), out_n (MMA N of the output layer), out_kcs (K chunks per output
 * stage), out_stages, stages_per_set, set_bytes (low and high 32 bits).  Used by the CPU-side tests of the host logic. */
int l2a_tc_plan_query(const l2a_mlp_desc* desc, int32_t* out8);
/* Same for the CTA-pair kernel (tcgen05.mma.cta_group::2; hidden widths multiples of 256 and <= 512): out12[0] = 1 if supported, then
 * hidden_stages (64 KB ring stages = 2 x 32 KB, one half per CTA), l0_packed, out_n, out_kcs, out_stages, stages_per_set, set_bytes
 * (low, high 32 bits); and the launch geometry the tile choice gives `n_candidates` x `n_envs` with `n_members` ensemble members on
 * a device with `num_sms` SMs: out12[9] = candidates per CTA, out12[10] = candidate tiles per env, out12[11] = CTAs. */
int l2a_tc2_plan_query(const l2a_mlp_desc* desc, int32_t n_candidates, int32_t n_envs, int32_t n_members, int32_t num_sms, int32_t* out12);

/* Diagnostics and microbenchmarks: only in the DEBUG build of the library (-DL2A_DEBUG_KERNELS -> lib/libl2a_b200_debug.so,
 * `python -m learning_to_adapt_b200.build --debug`); the product library does not contain them. */
#ifdef L2A_DEBUG_KERNELS
/* ---- diagnostics -----------------------------------------------------------------------------------------
 * Single tcgen05 GEMM tile through the same descriptor / TMEM code as the rollout kernel:
 * C[128, n] = A[128, k] * B[n, k]^T with A, B fp32 split into bf16 hi/lo on the device. */
int l2a_debug_umma_tile(l2a_ctx* ctx, const float* A, const float* B, float* C, int n, int k, int variant,
                        void* stream);

/* Weight-stream pipeline microbenchmark: every CTA of `grid` streams n_tiles_per_pass tiles of tile_bytes from `blob`
 * `passes` times through a `stages`-deep TMA/mbarrier ring (consumer holds each tile hold_cycles; `producers` / `consumers`
 * = 1 or 2 threads in different warps taking alternate tiles); cycles_out[grid] (device int64) receives the SM cycles each
 * CTA took. */
int l2a_debug_stream(l2a_ctx* ctx, const void* blob, int n_tiles_per_pass, int passes, int stages, int tile_bytes,
                     int hold_cycles, int producers, int consumers, int grid, long long* cycles_out, void* stream);

/* Tensor-pipe rate microbenchmark: SM cycles for `iters` tile pairs (12 split-bf16 MMAs, N = nc) with the A operand from
 * shared memory (mode 0), from tensor memory (1), tcgen05.cp only (2), cp + TS-mode MMA pipelined (3). cycles_out: device int64[1]. */
int l2a_debug_mma_rate(l2a_ctx* ctx, int nc, int mode, int iters, long long* cycles_out, void* stream);

/* CTA-pair probes for the next step of K1, one cluster of 2 CTAs, NC = 80 candidates per CTA.
 * mode 0: `iters` tile pairs of tcgen05.mma.cta_group::2 (M = 256, N = 160, 12 split-bf16 MMAs each): cycles_out[0] = total SM
 * cycles, [1] = cycles until the last MMA was issued.  mode 1: `iters` DSMEM bulk copies of copy_bytes (16 .. 16384, multiple
 * of 16) into the peer CTA in both directions: cycles_out[2], [3] = cycles each receiver waited.  cycles_out: device int64[4]. */
int l2a_debug_pair(l2a_ctx* ctx, int mode, int iters, int copy_bytes, long long* cycles_out, void* stream);

/* When set (device int64[128]), CTA 0 of the tcgen05 rollout records clock64() stamps of its pipeline events during
 * horizon step 1 (see L2A_STAMP slots in csrc/rollout_tc.cuh).  NULL switches it off. */
int l2a_debug_set_timeline(l2a_ctx* ctx, long long* buf128);

#endif /* L2A_DEBUG_KERNELS */

#ifdef __cplusplus
}
#endif
#endif /* L2A_B200_H */
