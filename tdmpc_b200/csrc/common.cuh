// Shared declarations of the tdmpc_b200 planner library (internal; the public ABI is include/tdmpc_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>
#include "../../include/tdmpc_b200.h"

#define TDMPC_MAX_H 16

// ------------------------------------------------------------------------------------------
// Packed parameters, CUDA-core fp32 path: every Linear stored transposed, Wt[k][n] (K padded to
// a multiple of 4 with zero rows) so that lanes sweep n contiguously.
struct LinearT {
  const float* wt;   // [Kp][N]
  const float* b;    // [N] or nullptr
  int K, Kp, N;
  const float* post_scale;   // optional per-output affine applied after the bias (eval-mode BatchNorm1d)
  const float* post_shift;
};
struct LayerNormP { const float* g; const float* b; };

struct ModelFp32 {
  LinearT enc[2];  LayerNormP enc_ln; int enc_has_ln;
  LinearT dyn[3];
  LinearT rew[3];
  LinearT pi[3];          // pi[0] is padded to K = L+A with zero rows for the action inputs
  LinearT q[2][3];
  LayerNormP q_ln[2][2];
  // DRNN dynamics (models/gru_dyna.py:11-29): NormGRUCell + BatchNorm prior MLP; reward reads the GRU state
  int drnn, Hd;
  LinearT gru_ih, gru_hh;        // [L+A -> 3Hd], [Hd -> 3Hd], no bias
  LayerNormP gru_ln[3];          // reset, update, newval (eps 1e-3)
  LinearT prior[2];              // Hd -> M (with BN affine) -> L
};

// ------------------------------------------------------------------------------------------
// Noise source: injected buffer (parity tests) or Philox4x32-10 keyed by (seed, call, stream).
// Both are addressed by the same flat element index, see include/tdmpc_b200.h for the layouts.
enum NoiseStream { NS_PI = 0, NS_CEM = 1, NS_TERM = 2, NS_OUT = 3, NS_PICK = 4, NS_SHIFT = 5, NS_REPLAY = 6 };

struct NoiseSrc {
  const float* inj;      // nullptr -> Philox
  uint64_t seed, call;
  int stream;
  const uint64_t* call_dev;   // optional device counter added to `call` (CUDA-graph replays: tdmpc_plan_args.call_dev)
};

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void philox_block(const NoiseSrc& s, uint64_t e4, uint32_t out[4]) {
  const uint64_t call = s.call + (s.call_dev ? __ldg(reinterpret_cast<const unsigned long long*>(s.call_dev)) : 0ull);
  philox4x32_10((uint32_t)e4, (uint32_t)(e4 >> 32), (uint32_t)s.stream, (uint32_t)call,
                (uint32_t)s.seed, (uint32_t)(s.seed >> 32) ^ (uint32_t)(call >> 32), out);
}

// i-th standard normal of the stream (Box-Muller on the Philox block holding element e).
__device__ __forceinline__ float philox_normal(const NoiseSrc& s, uint64_t e) {
  uint32_t r[4];
  philox_block(s, e >> 2, r);
  const int pair = (int)(e & 2);
  const float u1 = ((float)r[pair] + 1.0f) * 2.3283064365386963e-10f;   // (0,1]
  const float u2 = (float)r[pair + 1] * 2.3283064365386963e-10f;        // [0,1)
  const float rad = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return (e & 1) ? rad * sn : rad * cs;
}

// The four standard normals of one Philox block (elements 4*e4 .. 4*e4+3), bit-identical to four philox_normal calls.
__device__ __forceinline__ void philox_normal4(const NoiseSrc& s, uint64_t e4, float z[4]) {
  uint32_t r[4];
  philox_block(s, e4, r);
#pragma unroll
  for (int pair = 0; pair < 4; pair += 2) {
    const float u1 = ((float)r[pair] + 1.0f) * 2.3283064365386963e-10f;
    const float u2 = (float)r[pair + 1] * 2.3283064365386963e-10f;
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    z[pair] = rad * cs;
    z[pair + 1] = rad * sn;
  }
}

__device__ __forceinline__ float noise_at(const NoiseSrc& s, uint64_t e) {
  return s.inj ? __ldg(s.inj + e) : philox_normal(s, e);
}

__device__ __forceinline__ double philox_uniform64(const NoiseSrc& s, uint64_t e) {
  uint32_t r[4];
  philox_block(s, e, r);
  const uint64_t bits = (((uint64_t)r[0] << 32) | r[1]) >> 11;   // 53 random bits, like numpy's random_sample
  return (double)bits * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------------------------------------
// Activations with the reference's semantics.
__device__ __forceinline__ float elu1(float x) { return x > 0.0f ? x : expm1f(x); }

// nan_to_num_(0): NaN -> 0, +inf -> FLT_MAX, -inf -> -FLT_MAX (tdmpc.py:92)
__device__ __forceinline__ float nan_to_num0(float x) {
  if (x != x) return 0.0f;
  if (isinf(x)) return x > 0 ? 3.402823466e+38f : -3.402823466e+38f;
  return x;
}

// torch.min propagates NaN (fminf would drop it); the NaN is then zeroed by nan_to_num0.
__device__ __forceinline__ float min_nanprop(float a, float b) {
  return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b);
}

// TruncatedNormal(mu, std).sample(clip=0.3) then clamp to (-1+1e-6, 1-1e-6)  (helper.py:86-96)
__device__ __forceinline__ float policy_sample(float mu, float std, float eps) {
  if (!(std > 0.0f)) return mu;
  float e = eps * std;
  e = fminf(fmaxf(e, -0.3f), 0.3f);
  const float x = mu + e;
  return fminf(fmaxf(x, -1.0f + 1e-6f), 1.0f - 1e-6f);
}

// ------------------------------------------------------------------------------------------
// Trajectory-rollout launch parameters, shared by the fp32 and tcgen05 kernels.
enum RolloutMode { MODE_VALUE = 0, MODE_PI = 1, MODE_STEP = 2 };   // MODE_STEP: one DRNN transition with given actions -> hidden_out

struct RolloutParams {
  int mode;
  int B, T, H;             // problems, trajectories per problem, horizon
  int Tcap;                // row pitch (per problem) of actions / pi_actions / values / last_reward
  int L, A, M;
  const float* z0;         // [B, L]
  const float* hidden0;    // DRNN: [B, Hd] initial GRU state (nullptr = zeros)
  float* hidden_out;       // MODE_STEP: [B*T, Hd]
  const float* actions;    // MODE_VALUE: [B, Tcap, H, A] read
  float* pi_actions;       // MODE_PI:    [B, Tcap, H, A] written
  float* values;           // [B, Tcap]
  float* last_reward;      // [B, Tcap] or nullptr
  NoiseSrc noise;          // term noise (VALUE) / pi noise (PI)
  uint64_t noise_base;     // element offset of this iteration inside the stream
  uint64_t noise_bstride;  // elements between consecutive problems
  uint64_t noise_tstride;  // MODE_PI: elements between consecutive steps (P*A)
  float min_std;
  float disc[TDMPC_MAX_H + 1];   // (float)gamma^t, gamma^t accumulated in double like the reference
};

// Sampling of one iteration's action rows, in the reference's concatenation order
// [sampled | reused elites | pi]  (tdmpc.py:130-134; tdmpc_icem_similarity_mlp.py:139-157,200-233).
struct SampleSpec {
  int N, E, P;               // rows of the iteration
  int best_a;                // iCEM best-a: sample 0 := mean (tdmpc_icem_similarity_mlp.py:213-214)
  NoiseSrc cem;              // unit-variance noise [.., Trows, cem_rows, A]; injected buffers are used as they are
  uint64_t cem_base, cem_bstride; int cem_rows;
  // colored noise for the built-in generator (ignored when cem.inj is set): rows [0,n_white) white,
  // [n_white, n_white+n_pink) mats[0], the rest mats[1]; mats are [color_len][color_len] synthesis
  // matrices (colored = S * white along the horizon axis)
  const float* color_mats; int color_len; int n_white, n_pink;
};

// Selection / distribution update / next-iteration sampling.
struct SelectParams {
  int B, T, H, A, K;
  int iter, n_iters;
  const float* values;       // [B, Tcap]
  const float* last_reward;  // [B, Tcap]
  float* actions;            // [B, Tcap, H, A] (row pitch = Tcap rows); read for elites, written for the next iteration
  int Tcap;
  float* mean; float* std;   // [B, H, A]
  float temperature, momentum, one_minus_momentum, std_floor, std_max;
  int do_sample, do_output;  // stage switches: sample the next iteration's rows / write the plan outputs
  int stage;                 // set by launch_select: the K elite rows are gathered into shared memory first
  int* elite_idx;            // [B, K] scratch/out
  float* score;              // [B, K] scratch/out
  float* elite_actions;      // [B, K, H, A] out (iCEM carry-over) or nullptr
  // next-iteration sampling (do_sample)
  SampleSpec next;
  const float* pi_actions;    // [B, P0, H, A]
  int P0;
  // final outputs (last iteration)
  int eval_mode;
  NoiseSrc out_noise; NoiseSrc pick;
  const double* u_pick;
  float* action; float* metrics; float* prev_mean;
  // taps
  int* tap_elite_idx; float* tap_mean; float* tap_std;
};

// Stand-alone sampling of one iteration's rows from (mean, std) in global memory (sample_kernel, grid (B, nparts)).
struct SampleParams {
  int B, H, A, Tcap;
  SampleSpec spec;
  const float* mean; const float* std;   // [B, H, A]
  float* actions;
  const float* pi_actions; int P0;
  const float* reuse; int reuse_rows;    // reused elite rows [B, reuse_rows, H, A] or nullptr (already in place)
};

struct InitParams {
  int B, H, A;
  int Tcap;
  int skip_sample;            // 1: iteration-0 rows are drawn by sample_kernel instead
  SampleSpec first;           // rows of iteration 0
  float init_std;
  int t0_all; const uint8_t* t0; int has_prev; int warm_last;
  const float* prev_mean;     // [B,H,A]
  float* mean; float* std;
  float* actions;
  const float* pi_actions; int P0;
  // iCEM: iteration-0 reuse = previous step's elites shifted by one step, tail re-sampled
  // (tdmpc_icem_similarity_mlp.py:217-224).  elite_h = horizon the stored elites were planned with.
  float* elite_actions; int K; int shift_elites; int elite_h;
  NoiseSrc shift; const float* shift_mat;     // tail noise [B, Trows, E, A]; synthesis matrix or nullptr
};

// kernels implemented in the .cu files -----------------------------------------------------
void launch_encode(const ModelFp32& m, const float* obs, int B, int D, int E, int L, float* z, cudaStream_t st);
void launch_rollout_fp32(const ModelFp32& m, const RolloutParams& p, cudaStream_t st);
int  launch_rollout_tc(const struct TcModel& m, const RolloutParams& p, int num_sms, cudaStream_t st);
void launch_select(const SelectParams& p, int small_batch, cudaStream_t st);   // small_batch: fewer CTAs than 2 per SM
int select_configure();                                                       // once per device (tdmpc_create)
void launch_init(const InitParams& p, cudaStream_t st);
void launch_sample(const SampleParams& p, int nparts, cudaStream_t st);
void launch_bump_counter(uint64_t* c, cudaStream_t st);     // *c += 1 (tdmpc_plan_args.call_dev)
void launch_fill_noise(int stream_id, uint64_t seed, uint64_t call, uint64_t count, void* out, cudaStream_t st);
void launch_transpose_pack(const float* w, int N, int K, int Kp, int Kdst_off, float* wt, cudaStream_t st);

// ------------------------------------------------------------------------------------------
// Training-step forward (update_fwd.cu): TDMPC.update's loss terms for one replay batch.
struct UpdateFwdParams {
  int B, H, D, E, L, A, M;
  const float* obs;          // [B][D]
  const float* next_obses;   // [H][B][D]
  const float* action;       // [H][B][A]
  const float* reward;       // [H][B]
  NoiseSrc noise;            // policy draws inside _td_target: element (t * B + b) * A + a
  float min_std, discount;
  float rho_pow[TDMPC_MAX_H];
  float c_cons, c_rew, c_val;
  float *consistency, *reward_loss, *value, *priority, *total;      // [B]
  float *zs;                 // [H + 1][B][L] or nullptr (the latents update_pi consumes, tdmpc.py:201,213)
  float *td_target, *q1, *q2, *reward_pred;                          // [H][B] or nullptr (taps)
};

size_t update_fwd_smem(const ModelFp32& m, int D, int L, int A, int M, int E);
int launch_update_fwd(const ModelFp32& m, const ModelFp32& mt, const UpdateFwdParams& p, cudaStream_t st);
// whole training step (train.cu)
long long train_param_layout(const tdmpc_model_cfg& c, long long* pi_begin, long long* pi_end);
long long train_workspace_bytes(const tdmpc_model_cfg& c, int B, int H);
int launch_train_step(const tdmpc_model_cfg& cfg, const tdmpc_train_args& a, cudaStream_t st);     // launches (> 0) or -1
