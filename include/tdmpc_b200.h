/*
 * tdmpc_b200 -- C ABI of the B200-native TD-MPC planner (libtdmpc_b200.so).
 *
 * The reference (Slienteagle-wyb/tdmpc) is pure Python/PyTorch and has no FFI of its own;
 * the boundary it offers is the Python object surface `TDMPC.plan(obs, eval_mode, step, t0)`
 * and `TOLD.h/next/pi/Q`.  Each entry point below names the reference function it replaces
 * (paths relative to the reference's src/).  A ctypes binding (tdmpc_b200/_lib.py) sits on
 * top; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every pointer marked "device" is a CUDA device pointer owned by the
 *     caller; `stream` is a cudaStream_t passed as void*.
 *   - all calls are stream-ordered and asynchronous; nothing synchronises the host.
 *   - no allocation inside tdmpc_plan / tdmpc_estimate_value (CUDA-graph capturable) once
 *     tdmpc_reserve() has been called for the shapes in use.
 *   - return 0 on success, a negative tdmpc_status otherwise; tdmpc_last_error() gives text.
 *   - there is NO CPU fallback: on a machine without an sm_100 GPU tdmpc_create() fails.
 *   - tensors are fp32 row-major unless stated.  Row layout of action tensors is
 *     [problem][trajectory][t][a]  (the reference uses [t][trajectory][a]; the Python wrapper
 *     converts where the reference layout is user-visible, i.e. `_prev_mean` [H,A]).
 */
#ifndef TDMPC_B200_H
#define TDMPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDMPC_MAX_ITERS 16

#if defined(__GNUC__)
#define TDMPC_API __attribute__((visibility("default")))
#else
#define TDMPC_API
#endif

typedef enum {
  TDMPC_OK = 0,
  TDMPC_ERR_INVALID = -1,     /* bad argument / unsupported shape            */
  TDMPC_ERR_CUDA = -2,        /* a CUDA runtime call failed                  */
  TDMPC_ERR_ARCH = -3,        /* device is not sm_100 (no fallback path)     */
  TDMPC_ERR_WORKSPACE = -4,   /* tdmpc_reserve() not called / too small      */
  TDMPC_ERR_STATE = -5        /* weights not set                             */
} tdmpc_status;

typedef enum {
  TDMPC_PREC_FP32 = 0,   /* CUDA-core fp32 everywhere: the 1e-5 parity mode              */
  TDMPC_PREC_BF16 = 1,   /* tcgen05 bf16 operands, fp32 accumulate: the throughput mode.  The Q heads' tanh (rounded to
                          * bf16 right afterwards) is the one-MUFU tanh.approx.f32 (2^-11 relative): +6 % throughput, median
                          * trajectory-value deviation from the quantised oracle <= 2e-5 (north_star tolerance: 1e-3)            */
  TDMPC_PREC_BF16_EXACT = 2  /* same kernel with every activation at fp32 accuracy (two-MUFU tanh): median deviation 2e-7 .. 1e-6 */
} tdmpc_precision;

typedef enum {
  TDMPC_DYN_MLP = 0,     /* TOLD._dynamics = helper.mlp (algorithm/tdmpc.py:16)                          */
  TDMPC_DYN_DRNN = 1     /* DGruDyna: NormGRUCell + BN prior MLP (models/gru_dyna.py:11-29)              */
} tdmpc_dynamics;

/* Static model description: algorithm/tdmpc.py:12-23 (TOLD.__init__) + cfg fields it reads. */
typedef struct {
  int32_t obs_dim;        /* cfg.obs_shape[0]                                  */
  int32_t action_dim;     /* cfg.action_dim     (<= 64; bf16 paths: latent_dim + action_dim <= 192; GRU policy rollout on tensor cores: <= 8) */
  int32_t latent_dim;     /* cfg.latent_dim     (bf16 paths: <= 128)           */
  int32_t mlp_dim;        /* cfg.mlp_dim        (bf16 paths and tdmpc_update_forward: 512; fp32 paths and tdmpc_update: any) */
  int32_t enc_dim;        /* cfg.enc_dim                                       */
  int32_t hidden_dim;     /* cfg.hidden_dim     (DRNN only; bf16 path: 128).  Shapes outside these limits make tdmpc_create /
                           * tdmpc_set_weights return TDMPC_ERR_INVALID with the violated limit in tdmpc_last_error -- there is no
                           * silent fallback to another path. */
  int32_t enc_layernorm;  /* 1: helper.dmlab_enc_norm state branch (helper.py:161-165) */
  int32_t dynamics;       /* tdmpc_dynamics                                    */
  int32_t precision;      /* tdmpc_precision                                   */
  int32_t device;         /* CUDA device ordinal                               */
} tdmpc_model_cfg;

/* fp32 device pointers to the reference's parameters, nn.Linear layout [out, in].
 * Names follow the reference state_dict (tdmpc.py:12-23, helper.py:131-133,169-176,197-201). */
typedef struct {
  const float* enc_w[2];      /* _encoder.0.weight [E,D]; _encoder.{2|3}.weight [L,E]              */
  const float* enc_b[2];
  const float* enc_ln_g;      /* _encoder.1.{weight,bias} [E] when enc_layernorm, else NULL       */
  const float* enc_ln_b;
  const float* dyn_w[3];      /* _dynamics.{0,2,4}.weight [M,L+A] [M,M] [L,M]                     */
  const float* dyn_b[3];
  const float* rew_w[3];      /* _reward.{0,2,4}.weight   [M,L+A] [M,M] [1,M]                     */
  const float* rew_b[3];
  const float* pi_w[3];       /* _pi.{0,2,4}.weight       [M,L]   [M,M] [A,M]                     */
  const float* pi_b[3];
  const float* q_w[2][3];     /* _Q{1,2}.{0,3,6}.weight   [M,L+A] [M,M] [1,M]                     */
  const float* q_b[2][3];
  const float* q_ln_g[2][2];  /* _Q{1,2}.{1,4}.weight [M]                                          */
  const float* q_ln_b[2][2];
  /* TDMPC_DYN_DRNN only (tdmpc_icem_similarity_drnn.py:14-36, models/gru_dyna.py:11-23, models/rnns.py:8-17).
   * dyn_w / dyn_b are unused; rew_w[0] is [M, hidden_dim] (the reward reads the GRU state). */
  const float* gru_w_ih;      /* _dynamics.gru_cell.weight_ih.weight [3*Hd, L+A] (no bias); chunks reset, update, newval */
  const float* gru_w_hh;      /* _dynamics.gru_cell.weight_hh.weight [3*Hd, Hd]  (no bias)         */
  const float* gru_ln_g[3];   /* _dynamics.gru_cell.ln_{reset,update,newval}.weight [Hd] (eps 1e-3) */
  const float* gru_ln_b[3];
  const float* prior_w[2];    /* _dynamics.prior_mlp.{0,3}.weight [M,Hd] [L,M]                     */
  const float* prior_b[2];
  const float* prior_bn[4];   /* _dynamics.prior_mlp.1.{weight, bias, running_mean, running_var} [M] (BatchNorm1d, eval) */
} tdmpc_weights;

/* One call of TDMPC.plan (algorithm/tdmpc.py:94-163) for `batch` independent problems. */
typedef struct {
  int32_t batch;            /* B: independent planning problems (the reference has B = 1)        */
  int32_t horizon;          /* H = int(min(cfg.horizon, linear_schedule(horizon_schedule, step)))  tdmpc.py:111 */
  int32_t iterations;       /* cfg.iterations (<= TDMPC_MAX_ITERS)                                */
  int32_t num_elites;       /* cfg.num_elites                                                     */
  int32_t num_pi_trajs;     /* P of the policy rollout: int(mixture_coef * num_samples)  tdmpc.py:112 */
  /* rows per iteration, in the reference's concatenation order [sampled | reused elites | pi]
   * (tdmpc.py:134; tdmpc_icem_similarity_mlp.py:230-233).  TDMPC: n_samples[i]=N, n_reuse[i]=0, n_pi[i]=P. */
  int32_t n_samples[TDMPC_MAX_ITERS];
  int32_t n_reuse[TDMPC_MAX_ITERS];
  int32_t n_pi[TDMPC_MAX_ITERS];
  double discount;          /* cfg.discount (double: the reference accumulates gamma^t in Python floats) */
  double momentum;          /* cfg.momentum (double: 1 - momentum is formed in double, tdmpc.py:149)     */
  float temperature;        /* cfg.temperature                                                    */
  float min_std;            /* cfg.min_std: policy noise in rollouts (tdmpc.py:91,117)            */
  float std_floor;          /* agent.std: lower clamp of the CEM std (tdmpc.py:148)               */
  float std_max;            /* upper clamp, 2 (tdmpc.py:148)                                      */
  float init_std;           /* 2 for TDMPC (tdmpc.py:123), 0.5 for iCEM                           */
  int32_t eval_mode;        /* 1: no exploration noise on the returned action (tdmpc.py:157)      */
  int32_t t0_all;           /* 1/0: every problem is / is not at the first step; -1: read t0[]    */
  int32_t warm_last;        /* iCEM: mean[-1] = prev_mean[-1] (tdmpc_icem_similarity_mlp.py:180)  */
  int32_t best_a_last;      /* iCEM: sample 0 of the last iteration := mean (:213-214)            */
  int32_t shift_elites;     /* iCEM: iteration-0 reuse = previous step's elites shifted by one (:217-224) */
  int32_t has_prev;         /* 0: prev_mean holds nothing yet (hasattr(self,'_prev_mean') False)  */
  int32_t elite_horizon;    /* iCEM: horizon the stored elite_actions were planned with (== horizon unless the
                               plan horizon just grew, `extend_horizon`, tdmpc_icem_similarity_mlp.py:170-173,220-221) */
  /* iCEM colored action noise for the built-in generator (tdmpc_icem_similarity_mlp.py:126-157).  Ignored where
   * noise is injected (injected buffers hold final unit-variance noise).  color_len = length of the synthesised
   * series (cfg.horizon), 0 = white noise only.  color_mats: device [3][color_len][color_len] fp32 synthesis
   * matrices (colored = S . white along the horizon): [0] second class ("pink"), [1] third class ("brown"),
   * [2] the tail sampler of the shifted elites (sample_action_sequence, cfg.noise_beta).  Per iteration the first
   * n_white[i] sampled rows are white, the next n_color1[i] use mats[0], the rest mats[1]. */
  int32_t color_len;
  const float* color_mats;
  int32_t n_white[TDMPC_MAX_ITERS];
  int32_t n_color1[TDMPC_MAX_ITERS];

  const float* obs;         /* device [B, obs_dim]                                                */
  const float* z0;          /* device [B, latent_dim] or NULL; when set the encoder is skipped    */
  const uint8_t* t0;        /* device [B] or NULL (t0_all != -1)                                  */
  float* prev_mean;         /* device [B, H, A] in/out: agent._prev_mean (tdmpc.py:125,154)       */
  float* elite_actions;     /* device [B, K, H, A] in/out or NULL: iCEM agent._elite_actions      */
  const float* hidden;      /* DRNN: device [B, hidden_dim] GRU state in (NULL = zeros)  tdmpc_icem_similarity_drnn.py:168 */
  float* hidden_out;        /* DRNN: device [B, hidden_dim] state after stepping the model with the returned action (:267) */

  /* randomness: either injected draws (parity tests; consumption order = SURVEY Appendix B)
   * or, where a pointer is NULL, the built-in Philox4x32-10 stream keyed by (seed, call).        */
  uint64_t seed;
  uint64_t call;            /* plan-call counter: distinct noise per call                         */
  uint64_t* call_dev;       /* device [1] or NULL: added to `call` by every kernel of the plan and incremented by the
                             * library at the end of it, so that replays of a CUDA graph that captured this call draw
                             * fresh noise (the batched stepping loop: train.py:94-99 calls plan once per env step)  */
  const float* noise_pi;    /* device [B, H, P, A]           TruncatedNormal draws of the pi rollout  */
  const float* noise_cem;   /* device [B, I, H, Nmax, A]     torch.randn(H,N,A) of each iteration     */
  const float* noise_term;  /* device [B, I, Tmax, A]        terminal-policy draws in estimate_value  */
  const float* noise_out;   /* device [B, A]                 exploration noise (tdmpc.py:158)         */
  const float* noise_shift; /* device [B, H, E0, A]          iCEM: tail of the shifted elites (E0 = n_reuse[0]) */
  const double* u_pick;     /* device [B]                    uniform of np.random.choice (:153)       */

  float* action;            /* device [B, A] out                                                  */
  float* metrics;           /* device [B, 2] out: external_reward_mean, current_std (tdmpc.py:160-161) */

  /* optional debug taps (NULL to skip) */
  float* tap_values;        /* device [I, B, Tmax]                                                */
  int32_t* tap_elite_idx;   /* device [I, B, K]                                                   */
  float* tap_mean;          /* device [I, B, H, A]                                                */
  float* tap_std;           /* device [I, B, H, A]                                                */
  float* tap_z0;            /* device [B, L]                                                      */
} tdmpc_plan_args;

typedef struct tdmpc_handle tdmpc_handle;

/* Replaces: TDMPC.__init__ / TOLD.__init__ as far as planning needs them (tdmpc.py:12-23,56-66). */
TDMPC_API int tdmpc_create(const tdmpc_model_cfg* cfg, tdmpc_handle** out);
TDMPC_API int tdmpc_destroy(tdmpc_handle* h);
TDMPC_API const char* tdmpc_last_error(const tdmpc_handle* h);   /* h may be NULL: error of the last failed create */

/* Grow the library-owned scratch (trajectory actions, values, packed weights ...) so that plans
 * with up to `batch` problems x `max_rows` trajectories x `max_horizon` steps need no allocation. */
TDMPC_API int tdmpc_reserve(tdmpc_handle* h, int32_t batch, int32_t max_rows, int32_t max_horizon, int32_t max_elites);

/* Re-pack the reference's fp32 parameters for the kernels (transposed fp32 for the CUDA-core
 * path, bf16 swizzled 16 KB tiles in consumption order for the tcgen05 path).  Call after every
 * agent.update()/load() (tdmpc.py:80,230 mutate the parameters). Idempotent. */
TDMPC_API int tdmpc_set_weights(tdmpc_handle* h, const tdmpc_weights* w, void* stream);

/* Replaces TOLD.h (tdmpc.py:30-32): z[B,L] = encoder(obs[B,D]). */
TDMPC_API int tdmpc_encode(tdmpc_handle* h, const float* obs, int32_t batch, float* z_out, void* stream);

/* Replaces TDMPC.estimate_value (tdmpc.py:83-92) for B problems x T trajectories:
 *   values[b,j] = nan_to_num( sum_t gamma^t r_t + gamma^H min(Q1,Q2)(z_H, pi(z_H, min_std)) )
 * z0 [B,L]; actions [B,T,H,A]; noise_term [B,T,A] (or NULL -> Philox (seed, call, iteration 0));
 * values [B,T]; last_reward [B,T] (reward of the final step; may be NULL). */
TDMPC_API int tdmpc_estimate_value(tdmpc_handle* h, const float* z0, const float* actions, const float* noise_term,
                         int32_t batch, int32_t rows, int32_t horizon, double discount, float min_std,
                         uint64_t seed, uint64_t call, float* values, float* last_reward, void* stream);

/* Same for the GRU latent dynamics: TdICemSimDssm.estimate_value(z, actions, hidden) (tdmpc_icem_similarity_drnn.py:124-133).
 * hidden [B, hidden_dim] is the GRU state the B problems start from (NULL = zeros, which is also what
 * tdmpc_estimate_value uses on a TDMPC_DYN_DRNN handle). */
TDMPC_API int tdmpc_estimate_value_drnn(tdmpc_handle* h, const float* z0, const float* hidden, const float* actions,
                              const float* noise_term, int32_t batch, int32_t rows, int32_t horizon, double discount,
                              float min_std, uint64_t seed, uint64_t call, float* values, float* last_reward, void* stream);

/* Replaces the policy rollout of TDMPC.plan (tdmpc.py:113-118):
 * for t: a_t = pi(z, min_std); z = next(z, a_t).  pi_actions [B,P,H,A]; noise_pi [B,H,P,A] or NULL. */
TDMPC_API int tdmpc_pi_rollout(tdmpc_handle* h, const float* z0, const float* noise_pi, int32_t batch, int32_t rows,
                     int32_t horizon, float min_std, uint64_t seed, uint64_t call, float* pi_actions, void* stream);

/* Replaces tdmpc.py:138-149 for B problems: top-K (sorted descending, ties -> lower index),
 * softmax weights, weighted mean/std, clamp, momentum.  values [B,T]; actions [B,T,H,A];
 * mean/std [B,H,A] in/out; elite_idx [B,K] out; score [B,K] out (may be NULL). */
TDMPC_API int tdmpc_select_update(tdmpc_handle* h, const float* values, const float* actions, int32_t batch, int32_t rows,
                        int32_t horizon, int32_t num_elites, float temperature, double momentum, float std_floor,
                        float std_max, float* mean, float* std, int32_t* elite_idx, float* score, void* stream);

/* Replaces TDMPC.plan (tdmpc.py:94-163) -- and, with the iCEM fields, TdICemSimMlp.plan
 * (tdmpc_icem_similarity_mlp.py:159-265) -- for `batch` problems. */
TDMPC_API int tdmpc_plan(tdmpc_handle* h, const tdmpc_plan_args* args, void* stream);

/* The built-in noise stream, exposed so tests can feed the SAME draws to the oracle:
 * out[i] = i-th standard normal of stream `stream_id` (0 pi, 1 cem, 2 term, 3 out, 5 shift) for (seed, call);
 * stream 4 writes uniforms in [0,1) as float64 (u_pick). */
TDMPC_API int tdmpc_fill_noise(tdmpc_handle* h, int32_t stream_id, uint64_t seed, uint64_t call, uint64_t count,
                     void* out, void* stream);

/* ---- prioritised replay sampling: ReplayBuffer.sample, algorithm/helper.py:506-534 ----
 * Handle-less, stream-ordered, no allocation; errors go to tdmpc_last_error(NULL).
 *
 * tdmpc_replay_sample replaces helper.py:507-513:
 *   probs = priorities[:total] ** per_alpha; probs /= probs.sum()                      (fp32)
 *   idxs  = np.random.choice(total, batch, p=probs, replace=replace)                   (float64 cdf, searchsorted 'right';
 *           replace=0: numpy's redraw rounds -- first occurrences kept in draw order, found entries zeroed, the rest redrawn)
 *   weights = (total * probs[idxs]) ** (-per_beta); weights /= weights.max()
 * without moving the priority vector to the host. */
typedef struct {
  const float* priorities;   /* device [total]: ReplayBuffer._priorities[:total]                                   */
  int64_t total;             /* len(probs): capacity once the buffer is full, else ReplayBuffer.idx                */
  int32_t batch;             /* cfg.batch_size, 1..4096                                                            */
  int32_t replace;           /* np.random.choice(replace=...); the reference passes `not self._full`               */
  float per_alpha;           /* cfg.per_alpha                                                                       */
  float per_beta;            /* cfg.per_beta                                                                        */
  const double* uniforms;    /* device, optional: U[0,1) float64 draws in the order numpy's random_sample yields
                              * them (round r consumes batch - found_so_far values); NULL = built-in Philox stream  */
  int64_t n_uniforms;        /* length of `uniforms`; round r (counted over resumed calls) is only run while
                              * batch*(r+1) values are available                                                    */
  uint64_t seed, call;       /* Philox key of the built-in stream                                                   */
  int32_t max_rounds;        /* replace=0: redraw rounds to enqueue in THIS call (0 = 8); see status                */
  int32_t resume;            /* 0: fresh draw.  r > 0: continue the draw left in `workspace` by earlier calls that ran
                              * r rounds in total (same arguments) -- lets a caller enqueue 2 rounds, read `status`,
                              * and only pay for more when numpy's loop would really have gone on                   */
  void* workspace;           /* device, >= tdmpc_replay_workspace_bytes(total, batch)                               */
  int64_t workspace_bytes;
  int64_t* idxs;             /* out device [batch] int64 (torch.long); -1 where max_rounds was not enough           */
  float* weights;            /* out device [batch]                                                                   */
  int32_t* status;           /* out device [2], optional: {indices found, rounds run so far}                        */
} tdmpc_replay_sample_args;

/* tdmpc_replay_gather replaces helper.py:515-528 (state modality): the horizon+1 step windows behind each index,
 * the episode-final observation substituted on the last step (helper.py:526-527). */
typedef struct {
  const int64_t* idxs;       /* device [batch]                                                  */
  int32_t batch, horizon;    /* ReplayBuffer.batch_size, ReplayBuffer.horizon                   */
  int32_t obs_dim, action_dim;
  int64_t episode_length;    /* cfg.episode_length                                              */
  const float* obs_buf;      /* device ReplayBuffer._obs      [capacity + 1, obs_dim]           */
  const float* last_obs_buf; /* device ReplayBuffer._last_obs [capacity / episode_length, obs_dim] */
  const float* action_buf;   /* device ReplayBuffer._action   [capacity, action_dim]            */
  const float* reward_buf;   /* device ReplayBuffer._reward   [capacity]                        */
  float* obs;                /* out [batch, obs_dim]                                            */
  float* next_obs;           /* out [horizon + 1, batch, obs_dim]                               */
  float* action;             /* out [horizon + 1, batch, action_dim]                            */
  float* reward;             /* out [horizon + 1, batch] (the reference returns reward.unsqueeze(2)) */
} tdmpc_replay_gather_args;

TDMPC_API int64_t tdmpc_replay_workspace_bytes(int64_t total, int32_t batch);
TDMPC_API int tdmpc_replay_sample(const tdmpc_replay_sample_args* args, void* stream);
TDMPC_API int tdmpc_replay_gather(const tdmpc_replay_gather_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-step forward: the loss terms of TDMPC.update (algorithm/tdmpc.py:191-228) with the
 * TD target of TDMPC._td_target (tdmpc.py:180-189) for one replay batch, fp32.  `online` and `target` are two handles of the
 * same model shape holding agent.model and agent.model_target (MLP dynamics, precision any: the fp32 parameter arena is used).
 * Per sample b (the reference's [B, 1] tensors):
 *   consistency[b] = sum_t rho^t mean_L (z_{t+1} - h_target(next_obs_t))^2      reward_loss[b] = sum_t rho^t (r_t - reward_t)^2
 *   value[b]       = sum_t rho^t ((Q1 - td)^2 + (Q2 - td)^2)                    priority[b]    = sum_t rho^t (|Q1 - td| + |Q2 - td|)
 *   total[b]       = c_cons min(consistency, 1e4) + c_rew min(reward_loss, 1e4) + c_val min(value, 1e4)
 * with td = reward_t + discount * min(Q_target(h(next_obs_t), pi(h(next_obs_t), min_std))).  The mean / weighted reductions
 * over the batch, the backward pass and the optimiser stay with the caller (not implemented here). */
typedef struct {
  int32_t batch, horizon;
  double discount, rho;            /* cfg.discount, cfg.rho (python floats in the reference)                       */
  float min_std;
  float consistency_coef, reward_coef, value_coef;
  const float* obs;                /* [batch, obs_dim]                 replay_buffer.sample()[0]                   */
  const float* next_obses;         /* [horizon, batch, obs_dim]        ...[1][:horizon]                            */
  const float* action;             /* [horizon, batch, action_dim]     ...[2][:horizon]                            */
  const float* reward;             /* [horizon, batch]                 ...[3][:horizon].squeeze(2)                 */
  const float* noise_pi;           /* [horizon, batch, action_dim] standard-normal draws of pi() inside _td_target, or NULL:
                                    * Philox (seed, call)                                                           */
  uint64_t seed, call;
  float* consistency_loss;         /* out [batch]                                                                   */
  float* reward_loss;              /* out [batch]                                                                   */
  float* value_loss;               /* out [batch]                                                                   */
  float* priority_loss;            /* out [batch]  (what update() hands to update_priorities, before its clamp)     */
  float* total_loss;               /* out [batch]                                                                   */
  float* zs;                       /* out [horizon + 1, batch, latent_dim] or NULL: the latents update_pi consumes   */
  float* td_target;                /* out [horizon, batch] or NULL (debug taps)                                      */
  float* q1;                       /* out [horizon, batch] or NULL                                                   */
  float* q2;                       /* out [horizon, batch] or NULL                                                   */
  float* reward_pred;              /* out [horizon, batch] or NULL                                                   */
} tdmpc_update_args;

TDMPC_API int tdmpc_update_forward(tdmpc_handle* online, tdmpc_handle* target, const tdmpc_update_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The whole training step: TDMPC.update (algorithm/tdmpc.py:191-245) with _td_target (:180-189),
 * update_pi (:165-178), torch.nn.utils.clip_grad_norm_, torch.optim.Adam.step (both optimisers, default betas / eps, no weight
 * decay) and helper.ema (helper.py:48-52) for one replay batch -- forward, backward through the H-step latent unroll, optimiser
 * and target update -- fp32 on the device.  State observations, MLP dynamics, plain encoder (no LayerNorm variant).
 * The GEMMs run on the tensor cores at fp32 accuracy (every fp32 operand split into hi + lo bf16, three tcgen05 MMAs per product
 * into an fp32 accumulator: ~3e-6 relative; environment TDMPC_TRAIN_GEMM=simt selects the fp32 CUDA-core GEMM instead), and the
 * step's launch sequence is captured once per (shapes, hyper-parameters, persistent pointers) and replayed as a CUDA graph
 * (TDMPC_TRAIN_GRAPH=0: direct launches).  The caller's input tensors may live anywhere on the device and change every call: they are
 * copied into the workspace first.  Not thread-safe across concurrent calls on the same device (one graph cache per process).
 *
 * The PARAMETERS ARE UPDATED IN PLACE through the pointers of `online` (model) and `target` (model_target): the const in
 * tdmpc_weights is cast away.  After the call every handle that packed these parameters must be given them again
 * (tdmpc_set_weights).  Flat arenas (`adam_*`, `grads`) hold one float per parameter in state_dict order:
 *   _encoder.0.{weight,bias} _encoder.2.{weight,bias} | _dynamics.{0,2,4}.{weight,bias} | _reward.{0,2,4}.{weight,bias} |
 *   _pi.{0,2,4}.{weight,bias} | _Q1.0.{weight,bias} _Q1.1.{weight,bias} _Q1.3.{weight,bias} _Q1.4.{weight,bias} _Q1.6.{weight,bias} | _Q2...
 * (tdmpc_train_param_count returns the total and the [begin, end) range of the _pi block).  The model optimiser never touches
 * the _pi block of its arenas (those gradients are None in the reference), the policy optimiser has arenas of the _pi size. */
typedef struct {
  tdmpc_weights online, target;    /* device pointers of agent.model / agent.model_target; updated in place                     */
  int32_t batch, horizon;
  double discount, rho;            /* cfg.discount, cfg.rho                                                                    */
  float min_std;
  float consistency_coef, reward_coef, value_coef;
  double lr, beta1, beta2, adam_eps;   /* torch.optim.Adam(lr=cfg.lr): betas (0.9, 0.999), eps 1e-8                            */
  double grad_clip_norm;           /* cfg.grad_clip_norm                                                                       */
  double tau;                      /* cfg.tau                                                                                  */
  int32_t update_target;           /* step % cfg.update_freq == 0 (tdmpc.py:236)                                               */
  int64_t adam_step, pi_adam_step; /* optimiser step counts BEFORE this call (state['step']); this call is step + 1            */
  float* adam_m; float* adam_v;    /* [param_count] exp_avg / exp_avg_sq of agent.optim                                        */
  float* pi_adam_m; float* pi_adam_v;  /* [pi_count] of agent.pi_optim                                                         */
  float* grads;                    /* [param_count] out: the clipped gradients of the model step; _pi block: the clipped policy gradients */
  const float* obs;                /* [batch, obs_dim]                                                                          */
  const float* next_obses;         /* [horizon, batch, obs_dim]                                                                 */
  const float* action;             /* [horizon, batch, action_dim]                                                              */
  const float* reward;             /* [horizon, batch]                                                                          */
  const float* weights;            /* [batch] importance weights of the prioritised sample                                      */
  const float* noise_pi;           /* [horizon, batch, action_dim] draws of pi() inside _td_target, or NULL: Philox stream 0    */
  const float* noise_pi_update;    /* [horizon + 1, batch, action_dim] draws of pi() inside update_pi, or NULL: Philox stream 1 */
  uint64_t seed, call;
  float* metrics;                  /* out [8] device: consistency_loss, reward_loss, value_loss, total_loss, weighted_loss (the
                                    * means the reference returns), grad_norm (before clipping), pi_loss, pi grad norm         */
  float* priorities;               /* out [batch]: priority_loss.clamp(max=1e4), what update() hands to update_priorities       */
  void* workspace; int64_t workspace_bytes;   /* device scratch of tdmpc_train_workspace_bytes(cfg, batch, horizon)            */
} tdmpc_train_args;

TDMPC_API int64_t tdmpc_train_param_count(const tdmpc_model_cfg* cfg, int64_t* pi_begin, int64_t* pi_end);
TDMPC_API int64_t tdmpc_train_workspace_bytes(const tdmpc_model_cfg* cfg, int32_t batch, int32_t horizon);
/* `h`: any handle of the model's shape on the device (error reporting, launch counting); its packed weights are not used. */
TDMPC_API int tdmpc_update(tdmpc_handle* h, const tdmpc_train_args* args, void* stream);

/* Number of kernels of THIS library launched by the handle so far (bench.py's gpu_launches). */
TDMPC_API uint64_t tdmpc_launch_count(const tdmpc_handle* h);

/* Layout self-check for FFI bindings: writes up to n of {sizeof(tdmpc_model_cfg), sizeof(tdmpc_weights),
 * sizeof(tdmpc_plan_args), offsetof(plan_args, obs), offsetof(.., seed), offsetof(.., action), offsetof(.., tap_z0),
 * sizeof(tdmpc_replay_sample_args), sizeof(tdmpc_replay_gather_args), sizeof(tdmpc_update_args), sizeof(tdmpc_train_args)} and returns
 * how many values exist. */
TDMPC_API int32_t tdmpc_abi_sizes(int32_t* out, int32_t n);

/* Build information: "sm_100a;tcgen05=1;..." */
TDMPC_API const char* tdmpc_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* TDMPC_B200_H */
