// C-ABI entry points of libtdmpc_b200.so (declared in include/tdmpc_b200.h): handle management,
// weight packing, scratch reservation and the host-side orchestration of one TDMPC.plan call
// (algorithm/tdmpc.py:94-163) as a fixed sequence of stream-ordered kernel launches.
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "tc_model.cuh"

int select_max_rows();
int select_max_elites();
int select_max_ha();
void launch_bn_affine(const float* g, const float* b, const float* rm, const float* rv, float eps, int n, float* scale, float* shift, cudaStream_t st);

struct tdmpc_handle {
  tdmpc_model_cfg cfg;
  int num_sms = 0;
  std::string err;
  // packed parameters
  float* arena32 = nullptr; size_t arena32_floats = 0;
  ModelFp32 m32{};
  TcModel tc{};
  void* arena_tc = nullptr; size_t arena_tc_bytes = 0;
  bool weights_set = false;
  // scratch
  int capB = 0, capT = 0, capH = 0, capK = 0;
  float *z0 = nullptr, *actions = nullptr, *pi_actions = nullptr, *values = nullptr, *last_reward = nullptr;
  float *mean = nullptr, *std = nullptr, *score = nullptr, *elites = nullptr;
  int* elite_idx = nullptr;
  uint64_t launches = 0;
};

static std::string g_create_err;
void tdmpc_set_global_error(const char* msg) { g_create_err = msg ? msg : ""; }   // handle-less entry points (replay.cu)

#define FAIL(h, code, ...)                                   \
  do {                                                       \
    char _b[512]; snprintf(_b, sizeof(_b), __VA_ARGS__);     \
    (h)->err = _b; return (code);                            \
  } while (0)

#define CUDA_OK(h, call)                                                                        \
  do {                                                                                          \
    cudaError_t _e = (call);                                                                    \
    if (_e != cudaSuccess) FAIL(h, TDMPC_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); \
  } while (0)

static int check_launch(tdmpc_handle* h, const char* what, int n = 1) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) FAIL(h, TDMPC_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  h->launches += n;
  return 0;
}

// ---------------------------------------------------------------------------------------------
static size_t lin_floats(int N, int K) { return (size_t)((K + 3) & ~3) * N + (size_t)((N + 3) & ~3); }

static size_t arena32_size(const tdmpc_model_cfg& c) {
  const int D = c.obs_dim, A = c.action_dim, L = c.latent_dim, M = c.mlp_dim, E = c.enc_dim;
  size_t n = lin_floats(E, D) + lin_floats(L, E) + 2 * (size_t)E;
  n += lin_floats(M, L + A) + lin_floats(M, M) + lin_floats(L, M);        // dyn
  n += lin_floats(M, L + A) + lin_floats(M, M) + lin_floats(1, M);        // rew
  n += lin_floats(M, L + A) + lin_floats(M, M) + lin_floats(A, M);        // pi (K padded to L+A)
  n += 2 * (lin_floats(M, L + A) + lin_floats(M, M) + lin_floats(1, M) + 4 * (size_t)M);
  if (c.dynamics == TDMPC_DYN_DRNN) {
    const int Hd = c.hidden_dim;
    n += lin_floats(3 * Hd, L + A) + lin_floats(3 * Hd, Hd) + 6 * (size_t)Hd;     // GRU gates + 3 LayerNorms
    n += lin_floats(M, Hd) + lin_floats(L, M) + 2 * (size_t)M;                      // prior MLP + BN scale/shift
    n += lin_floats(M, Hd);                                                         // reward layer 1 reads the GRU state
  }
  return n + 64;
}

extern "C" int tdmpc_create(const tdmpc_model_cfg* cfg, tdmpc_handle** out) {
  if (!cfg || !out) { g_create_err = "null argument"; return TDMPC_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_err = "no CUDA device: tdmpc_b200 has no CPU fallback";
    return TDMPC_ERR_ARCH;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "bad device ordinal"; return TDMPC_ERR_INVALID; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) { g_create_err = "cudaGetDeviceProperties failed"; return TDMPC_ERR_CUDA; }
  if (prop.major != 10) {
    char b[256];
    snprintf(b, sizeof(b), "device %d is sm_%d%d; this library is built for sm_100a only (no fallback)", cfg->device, prop.major, prop.minor);
    g_create_err = b;
    return TDMPC_ERR_ARCH;
  }
  if (cfg->obs_dim <= 0 || cfg->action_dim <= 0 || cfg->latent_dim <= 0 || cfg->mlp_dim <= 0 || cfg->enc_dim <= 0) {
    g_create_err = "model dims must be positive"; return TDMPC_ERR_INVALID;
  }
  if (cfg->action_dim > 64) { g_create_err = "action_dim > 64 unsupported"; return TDMPC_ERR_INVALID; }
  if (cfg->dynamics != TDMPC_DYN_MLP && cfg->dynamics != TDMPC_DYN_DRNN) { g_create_err = "unknown dynamics type"; return TDMPC_ERR_INVALID; }
  if (cfg->dynamics == TDMPC_DYN_DRNN) {
    if (cfg->hidden_dim <= 0 || cfg->hidden_dim > 128 || cfg->hidden_dim % 4) { g_create_err = "DRNN: hidden_dim must be a multiple of 4 in (0, 128]"; return TDMPC_ERR_INVALID; }
    if (3 * cfg->hidden_dim > cfg->mlp_dim) { g_create_err = "DRNN: 3*hidden_dim must not exceed mlp_dim"; return TDMPC_ERR_INVALID; }
  }
  if (cfg->precision == TDMPC_PREC_BF16 || cfg->precision == TDMPC_PREC_BF16_EXACT) {
    const char* why = tc_supported(*cfg);
    if (why) { g_create_err = std::string("bf16 tcgen05 path: ") + why; return TDMPC_ERR_INVALID; }
  } else if (cfg->precision != TDMPC_PREC_FP32) { g_create_err = "unknown precision"; return TDMPC_ERR_INVALID; }
  tdmpc_handle* h = new tdmpc_handle();
  h->cfg = *cfg;
  h->num_sms = prop.multiProcessorCount;
  cudaSetDevice(cfg->device);
  if (select_configure() != 0) { g_create_err = "cudaFuncSetAttribute(select_kernel) failed"; delete h; return TDMPC_ERR_CUDA; }
  h->arena32_floats = arena32_size(*cfg);
  if (cudaMalloc(&h->arena32, h->arena32_floats * sizeof(float)) != cudaSuccess) { g_create_err = "cudaMalloc(arena32) failed"; delete h; return TDMPC_ERR_CUDA; }
  if (cfg->precision != TDMPC_PREC_FP32) {
    h->arena_tc_bytes = tc_arena_bytes(*cfg);
    if (cudaMalloc(&h->arena_tc, h->arena_tc_bytes) != cudaSuccess) { g_create_err = "cudaMalloc(arena_tc) failed"; cudaFree(h->arena32); delete h; return TDMPC_ERR_CUDA; }
  }
  *out = h;
  return TDMPC_OK;
}

static void free_scratch(tdmpc_handle* h) {
  cudaFree(h->z0); cudaFree(h->actions); cudaFree(h->pi_actions); cudaFree(h->values); cudaFree(h->last_reward);
  cudaFree(h->mean); cudaFree(h->std); cudaFree(h->score); cudaFree(h->elites); cudaFree(h->elite_idx);
  h->z0 = h->actions = h->pi_actions = h->values = h->last_reward = h->mean = h->std = h->score = h->elites = nullptr;
  h->elite_idx = nullptr;
}

extern "C" int tdmpc_destroy(tdmpc_handle* h) {
  if (!h) return TDMPC_OK;
  cudaSetDevice(h->cfg.device);
  free_scratch(h);
  cudaFree(h->arena32);
  cudaFree(h->arena_tc);
  tc_release_host(&h->tc);
  delete h;
  return TDMPC_OK;
}

extern "C" const char* tdmpc_last_error(const tdmpc_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }
void tc_set_trace(unsigned long long* dev_buf);
int tc_trace_slots();
// Debug hook (not part of the public header): device buffer of tc_trace_slots() uint64 that receives
// clock64 stamps [role][job][start,end] of one tile of the tcgen05 rollout kernel; NULL disables.
extern "C" __attribute__((visibility("default"))) int tdmpc_debug_set_trace(void* dev_buf) { tc_set_trace((unsigned long long*)dev_buf); return tc_trace_slots(); }

// Debug hook (not part of the public header): copies the job program of (mode, horizon) to the host; returns its length.
extern "C" __attribute__((visibility("default"))) int tdmpc_debug_jobs(tdmpc_handle* h, int mode, int horizon, uint32_t* out, int cap) {
  if (!h || !h->weights_set || h->cfg.precision == TDMPC_PREC_FP32 || mode < 0 || mode > 1 || horizon < 1 || horizon > TDMPC_MAX_H) return -1;
  const int n = h->tc.njobs[mode][horizon];
  if (out && cap >= n) cudaMemcpy(out, h->tc.jobs[mode][horizon], sizeof(uint32_t) * n, cudaMemcpyDeviceToHost);
  return n;
}
extern "C" int32_t tdmpc_abi_sizes(int32_t* out, int32_t n) {
  const int32_t v[] = {(int32_t)sizeof(tdmpc_model_cfg), (int32_t)sizeof(tdmpc_weights), (int32_t)sizeof(tdmpc_plan_args),
                       (int32_t)offsetof(tdmpc_plan_args, obs), (int32_t)offsetof(tdmpc_plan_args, seed),
                       (int32_t)offsetof(tdmpc_plan_args, action), (int32_t)offsetof(tdmpc_plan_args, tap_z0),
                       (int32_t)sizeof(tdmpc_replay_sample_args), (int32_t)sizeof(tdmpc_replay_gather_args), (int32_t)sizeof(tdmpc_update_args),
                       (int32_t)sizeof(tdmpc_train_args)};
  const int32_t m = (int32_t)(sizeof(v) / sizeof(v[0]));
  for (int32_t i = 0; i < n && i < m; ++i) out[i] = v[i];
  return m;
}
extern "C" uint64_t tdmpc_launch_count(const tdmpc_handle* h) { return h ? h->launches : 0; }
extern "C" const char* tdmpc_build_info(void) { return "tdmpc_b200;arch=sm_100a;rollout=tcgen05+tmem+bulk-tma|fp32-cuda-core;dynamics=mlp,gru;select=bitonic;replay=per-sampler"; }

extern "C" int tdmpc_reserve(tdmpc_handle* h, int32_t batch, int32_t max_rows, int32_t max_horizon, int32_t max_elites) {
  if (!h) return TDMPC_ERR_INVALID;
  if (batch <= 0 || max_rows <= 0 || max_horizon <= 0 || max_elites <= 0) FAIL(h, TDMPC_ERR_INVALID, "reserve: non-positive size");
  if (max_rows > select_max_rows()) FAIL(h, TDMPC_ERR_INVALID, "reserve: %d trajectories per problem > %d", max_rows, select_max_rows());
  if (max_horizon > TDMPC_MAX_H) FAIL(h, TDMPC_ERR_INVALID, "reserve: horizon %d > %d", max_horizon, TDMPC_MAX_H);
  if (max_elites > select_max_elites()) FAIL(h, TDMPC_ERR_INVALID, "reserve: num_elites %d > %d", max_elites, select_max_elites());
  if (batch <= h->capB && max_rows <= h->capT && max_horizon <= h->capH && max_elites <= h->capK) return TDMPC_OK;
  cudaSetDevice(h->cfg.device);
  CUDA_OK(h, cudaDeviceSynchronize());
  free_scratch(h);
  const int B = batch > h->capB ? batch : h->capB, T = max_rows > h->capT ? max_rows : h->capT;
  const int H = max_horizon > h->capH ? max_horizon : h->capH, K = max_elites > h->capK ? max_elites : h->capK;
  const size_t A = h->cfg.action_dim, L = h->cfg.latent_dim;
  CUDA_OK(h, cudaMalloc(&h->z0, sizeof(float) * B * L));
  CUDA_OK(h, cudaMalloc(&h->actions, sizeof(float) * (size_t)B * T * H * A));
  CUDA_OK(h, cudaMalloc(&h->pi_actions, sizeof(float) * (size_t)B * T * H * A));
  CUDA_OK(h, cudaMalloc(&h->values, sizeof(float) * (size_t)B * T));
  CUDA_OK(h, cudaMalloc(&h->last_reward, sizeof(float) * (size_t)B * T));
  CUDA_OK(h, cudaMalloc(&h->mean, sizeof(float) * (size_t)B * H * A));
  CUDA_OK(h, cudaMalloc(&h->std, sizeof(float) * (size_t)B * H * A));
  CUDA_OK(h, cudaMalloc(&h->score, sizeof(float) * (size_t)B * K));
  CUDA_OK(h, cudaMalloc(&h->elites, sizeof(float) * (size_t)B * K * H * A));
  CUDA_OK(h, cudaMalloc(&h->elite_idx, sizeof(int) * (size_t)B * K));
  h->capB = B; h->capT = T; h->capH = H; h->capK = K;
  return TDMPC_OK;
}

// ---------------------------------------------------------------------------------------------
struct ArenaCursor { float* base; size_t off; };

static int pack_linear(tdmpc_handle* h, ArenaCursor& c, const float* w, const float* b, int N, int K, int Kfull,
                       LinearT* out, cudaStream_t st) {
  // Wt [Kp][N] with Kp = roundup4(Kfull); only the first K rows are filled (rest stay zero).
  const int Kp = (Kfull + 3) & ~3;
  float* wt = c.base + c.off; c.off += (size_t)Kp * N;
  float* bp = c.base + c.off; c.off += (size_t)((N + 3) & ~3);
  if (!w) FAIL(h, TDMPC_ERR_INVALID, "set_weights: a weight pointer is NULL");
  launch_transpose_pack(w, N, K, Kp, 0, wt, st);
  if (int r = check_launch(h, "transpose_pack")) return r;
  if (b) CUDA_OK(h, cudaMemcpyAsync(bp, b, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
  out->wt = wt; out->b = b ? bp : nullptr; out->K = Kfull; out->Kp = Kp; out->N = N;
  out->post_scale = nullptr; out->post_shift = nullptr;
  return 0;
}

static int pack_ln(tdmpc_handle* h, ArenaCursor& c, const float* g, const float* b, int n, LayerNormP* out, cudaStream_t st) {
  if (!g || !b) FAIL(h, TDMPC_ERR_INVALID, "set_weights: a LayerNorm pointer is NULL");
  float* gp = c.base + c.off; c.off += n;
  float* bp = c.base + c.off; c.off += n;
  CUDA_OK(h, cudaMemcpyAsync(gp, g, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  CUDA_OK(h, cudaMemcpyAsync(bp, b, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  out->g = gp; out->b = bp;
  return 0;
}

extern "C" int tdmpc_set_weights(tdmpc_handle* h, const tdmpc_weights* w, void* stream) {
  if (!h || !w) return TDMPC_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  cudaSetDevice(h->cfg.device);
  const tdmpc_model_cfg& c = h->cfg;
  const int D = c.obs_dim, A = c.action_dim, L = c.latent_dim, M = c.mlp_dim, E = c.enc_dim;
  CUDA_OK(h, cudaMemsetAsync(h->arena32, 0, h->arena32_floats * sizeof(float), st));
  ArenaCursor cur{h->arena32, 0};
  ModelFp32& m = h->m32;
  int r;
#define PL(...) if ((r = pack_linear(h, cur, __VA_ARGS__, st))) return r
  PL(w->enc_w[0], w->enc_b[0], E, D, D, &m.enc[0]);
  PL(w->enc_w[1], w->enc_b[1], L, E, E, &m.enc[1]);
  m.enc_has_ln = c.enc_layernorm;
  if (c.enc_layernorm) { if ((r = pack_ln(h, cur, w->enc_ln_g, w->enc_ln_b, E, &m.enc_ln, st))) return r; }
  m.drnn = c.dynamics == TDMPC_DYN_DRNN;
  m.Hd = c.hidden_dim;
  if (!m.drnn) {
    PL(w->dyn_w[0], w->dyn_b[0], M, L + A, L + A, &m.dyn[0]);
    PL(w->dyn_w[1], w->dyn_b[1], M, M, M, &m.dyn[1]);
    PL(w->dyn_w[2], w->dyn_b[2], L, M, M, &m.dyn[2]);
    PL(w->rew_w[0], w->rew_b[0], M, L + A, L + A, &m.rew[0]);
  } else {
    const int Hd = c.hidden_dim;
    PL(w->gru_w_ih, nullptr, 3 * Hd, L + A, L + A, &m.gru_ih);
    PL(w->gru_w_hh, nullptr, 3 * Hd, Hd, Hd, &m.gru_hh);
    for (int g = 0; g < 3; ++g)
      if ((r = pack_ln(h, cur, w->gru_ln_g[g], w->gru_ln_b[g], Hd, &m.gru_ln[g], st))) return r;
    PL(w->prior_w[0], w->prior_b[0], M, Hd, Hd, &m.prior[0]);
    PL(w->prior_w[1], w->prior_b[1], L, M, M, &m.prior[1]);
    for (int i = 0; i < 4; ++i) if (!w->prior_bn[i]) FAIL(h, TDMPC_ERR_INVALID, "set_weights: a BatchNorm pointer is NULL");
    float* sc = cur.base + cur.off; cur.off += M;
    float* sh = cur.base + cur.off; cur.off += M;
    launch_bn_affine(w->prior_bn[0], w->prior_bn[1], w->prior_bn[2], w->prior_bn[3], 1e-5f, M, sc, sh, st);
    if ((r = check_launch(h, "bn_affine"))) return r;
    m.prior[0].post_scale = sc; m.prior[0].post_shift = sh;
    PL(w->rew_w[0], w->rew_b[0], M, Hd, Hd, &m.rew[0]);
  }
  PL(w->rew_w[1], w->rew_b[1], M, M, M, &m.rew[1]);
  PL(w->rew_w[2], w->rew_b[2], 1, M, M, &m.rew[2]);
  PL(w->pi_w[0], w->pi_b[0], M, L, L + A, &m.pi[0]);     // action rows stay zero
  PL(w->pi_w[1], w->pi_b[1], M, M, M, &m.pi[1]);
  PL(w->pi_w[2], w->pi_b[2], A, M, M, &m.pi[2]);
  for (int q = 0; q < 2; ++q) {
    PL(w->q_w[q][0], w->q_b[q][0], M, L + A, L + A, &m.q[q][0]);
    PL(w->q_w[q][1], w->q_b[q][1], M, M, M, &m.q[q][1]);
    PL(w->q_w[q][2], w->q_b[q][2], 1, M, M, &m.q[q][2]);
    for (int j = 0; j < 2; ++j)
      if ((r = pack_ln(h, cur, w->q_ln_g[q][j], w->q_ln_b[q][j], M, &m.q_ln[q][j], st))) return r;
  }
#undef PL
  if (cur.off > h->arena32_floats) FAIL(h, TDMPC_ERR_INVALID, "internal: fp32 arena overflow");
  if (c.precision != TDMPC_PREC_FP32) {
    int n = tc_pack(c, *w, h->m32, h->arena_tc, &h->tc, st);
    if (n < 0) FAIL(h, TDMPC_ERR_CUDA, "tcgen05 weight packing failed");
    if ((r = check_launch(h, "tc_pack", n))) return r;
  }
  h->weights_set = true;
  return TDMPC_OK;
}

// ---------------------------------------------------------------------------------------------
static void fill_disc(double gamma, int H, float* disc) {
  double d = 1.0;                      // G, discount = 0, 1 ... discount *= cfg.discount  (tdmpc.py:86-90)
  for (int t = 0; t <= H; ++t) { disc[t] = (float)d; d *= gamma; }
}

static int run_rollout(tdmpc_handle* h, const RolloutParams& p, cudaStream_t st) {
  if ((long long)p.B * p.T <= 0) return 0;
  // DRNN in bf16 mode: value and policy rollouts run on tcgen05 (the policy rollout needs action_dim <= 8, else its job
  // list is empty); the single post-plan transition keeps the fp32 kernel
  const bool tc = h->cfg.precision != TDMPC_PREC_FP32 && p.mode != MODE_STEP && h->tc.njobs[p.mode][p.H] > 0;
  if (tc) {
    int n = launch_rollout_tc(h->tc, p, h->num_sms, st);
    if (n < 0) FAIL(h, TDMPC_ERR_CUDA, "tcgen05 rollout launch failed");
    return check_launch(h, "rollout_tc", n);
  }
  launch_rollout_fp32(h->m32, p, st);
  return check_launch(h, "rollout_fp32");
}

extern "C" int tdmpc_encode(tdmpc_handle* h, const float* obs, int32_t batch, float* z_out, void* stream) {
  if (!h || !obs || !z_out || batch <= 0) return TDMPC_ERR_INVALID;
  if (!h->weights_set) FAIL(h, TDMPC_ERR_STATE, "encode: weights not set");
  cudaSetDevice(h->cfg.device);
  launch_encode(h->m32, obs, batch, h->cfg.obs_dim, h->cfg.enc_dim, h->cfg.latent_dim, z_out, (cudaStream_t)stream);
  return check_launch(h, "encode");
}

static int estimate_value_impl(tdmpc_handle* h, const float* z0, const float* hidden, const float* actions, const float* noise_term,
                               int32_t batch, int32_t rows, int32_t horizon, double discount, float min_std,
                               uint64_t seed, uint64_t call, float* values, float* last_reward, void* stream) {
  if (!h || !z0 || !actions || !values) return TDMPC_ERR_INVALID;
  if (!h->weights_set) FAIL(h, TDMPC_ERR_STATE, "estimate_value: weights not set");
  if (batch <= 0 || rows <= 0 || horizon <= 0 || horizon > TDMPC_MAX_H) FAIL(h, TDMPC_ERR_INVALID, "estimate_value: bad shape");
  if (hidden && h->cfg.dynamics != TDMPC_DYN_DRNN) FAIL(h, TDMPC_ERR_INVALID, "estimate_value: hidden state given to a model without GRU dynamics");
  cudaSetDevice(h->cfg.device);
  RolloutParams p{};
  p.mode = MODE_VALUE; p.B = batch; p.T = rows; p.Tcap = rows; p.H = horizon;
  p.L = h->cfg.latent_dim; p.A = h->cfg.action_dim; p.M = h->cfg.mlp_dim;
  p.z0 = z0; p.hidden0 = hidden; p.actions = actions; p.values = values; p.last_reward = last_reward;
  p.noise = NoiseSrc{noise_term, seed, call, NS_TERM};
  p.noise_base = 0; p.noise_bstride = (uint64_t)rows * p.A; p.noise_tstride = 0;
  p.min_std = min_std;
  fill_disc(discount, horizon, p.disc);
  return run_rollout(h, p, (cudaStream_t)stream);
}

extern "C" int tdmpc_estimate_value(tdmpc_handle* h, const float* z0, const float* actions, const float* noise_term,
                                    int32_t batch, int32_t rows, int32_t horizon, double discount, float min_std,
                                    uint64_t seed, uint64_t call, float* values, float* last_reward, void* stream) {
  return estimate_value_impl(h, z0, nullptr, actions, noise_term, batch, rows, horizon, discount, min_std, seed, call, values, last_reward, stream);
}

extern "C" int tdmpc_estimate_value_drnn(tdmpc_handle* h, const float* z0, const float* hidden, const float* actions,
                                         const float* noise_term, int32_t batch, int32_t rows, int32_t horizon, double discount,
                                         float min_std, uint64_t seed, uint64_t call, float* values, float* last_reward, void* stream) {
  return estimate_value_impl(h, z0, hidden, actions, noise_term, batch, rows, horizon, discount, min_std, seed, call, values, last_reward, stream);
}

extern "C" int tdmpc_pi_rollout(tdmpc_handle* h, const float* z0, const float* noise_pi, int32_t batch, int32_t rows,
                                int32_t horizon, float min_std, uint64_t seed, uint64_t call, float* pi_actions, void* stream) {
  if (!h || !z0 || !pi_actions) return TDMPC_ERR_INVALID;
  if (!h->weights_set) FAIL(h, TDMPC_ERR_STATE, "pi_rollout: weights not set");
  if (batch <= 0 || rows <= 0 || horizon <= 0 || horizon > TDMPC_MAX_H) FAIL(h, TDMPC_ERR_INVALID, "pi_rollout: bad shape");
  cudaSetDevice(h->cfg.device);
  RolloutParams p{};
  p.mode = MODE_PI; p.B = batch; p.T = rows; p.Tcap = rows; p.H = horizon;
  p.L = h->cfg.latent_dim; p.A = h->cfg.action_dim; p.M = h->cfg.mlp_dim;
  p.z0 = z0; p.pi_actions = pi_actions;
  p.noise = NoiseSrc{noise_pi, seed, call, NS_PI};
  p.noise_base = 0; p.noise_bstride = (uint64_t)horizon * rows * p.A; p.noise_tstride = (uint64_t)rows * p.A;
  p.min_std = min_std;
  return run_rollout(h, p, (cudaStream_t)stream);
}

extern "C" int tdmpc_select_update(tdmpc_handle* h, const float* values, const float* actions, int32_t batch, int32_t rows,
                                   int32_t horizon, int32_t num_elites, float temperature, double momentum, float std_floor,
                                   float std_max, float* mean, float* std, int32_t* elite_idx, float* score, void* stream) {
  if (!h || !values || !actions || !mean || !std || !elite_idx) return TDMPC_ERR_INVALID;
  if (batch <= 0 || rows <= 0 || rows > select_max_rows() || horizon <= 0 || horizon > TDMPC_MAX_H ||
      num_elites <= 0 || num_elites > rows || num_elites > select_max_elites())
    FAIL(h, TDMPC_ERR_INVALID, "select_update: bad shape (rows=%d K=%d H=%d)", rows, num_elites, horizon);
  cudaSetDevice(h->cfg.device);
  SelectParams p{};
  p.B = batch; p.T = rows; p.Tcap = rows; p.H = horizon; p.A = h->cfg.action_dim; p.K = num_elites;
  p.iter = 0; p.n_iters = 1; p.values = values; p.actions = const_cast<float*>(actions);
  p.mean = mean; p.std = std; p.temperature = temperature; p.momentum = (float)momentum;
  p.one_minus_momentum = (float)(1.0 - momentum); p.std_floor = std_floor; p.std_max = std_max;
  p.elite_idx = elite_idx; p.score = score;
  launch_select(p, batch < 2 * h->num_sms, (cudaStream_t)stream);
  return check_launch(h, "select");
}

extern "C" int tdmpc_update_forward(tdmpc_handle* h, tdmpc_handle* ht, const tdmpc_update_args* a, void* stream) {
  if (!h || !ht || !a) return TDMPC_ERR_INVALID;
  if (!h->weights_set || !ht->weights_set) FAIL(h, TDMPC_ERR_STATE, "update_forward: weights of the online / target handle not set");
  const tdmpc_model_cfg &c = h->cfg, &ct = ht->cfg;
  if (c.obs_dim != ct.obs_dim || c.action_dim != ct.action_dim || c.latent_dim != ct.latent_dim || c.mlp_dim != ct.mlp_dim ||
      c.enc_dim != ct.enc_dim || c.enc_layernorm != ct.enc_layernorm || c.dynamics != ct.dynamics || c.device != ct.device)
    FAIL(h, TDMPC_ERR_INVALID, "update_forward: online and target handles differ in shape or device");
  if (c.dynamics != TDMPC_DYN_MLP) FAIL(h, TDMPC_ERR_INVALID, "update_forward: MLP dynamics only");
  if (a->batch <= 0 || a->horizon <= 0 || a->horizon > TDMPC_MAX_H) FAIL(h, TDMPC_ERR_INVALID, "update_forward: bad batch / horizon");
  if (!a->obs || !a->next_obses || !a->action || !a->reward || !a->consistency_loss || !a->reward_loss || !a->value_loss ||
      !a->priority_loss || !a->total_loss)
    FAIL(h, TDMPC_ERR_INVALID, "update_forward: a required pointer is NULL");
  if (c.mlp_dim != 512 || c.enc_dim > c.mlp_dim) FAIL(h, TDMPC_ERR_INVALID, "update_forward: mlp_dim must be 512 and enc_dim <= mlp_dim");
  cudaSetDevice(c.device);
  UpdateFwdParams p{};
  p.B = a->batch; p.H = a->horizon; p.D = c.obs_dim; p.E = c.enc_dim; p.L = c.latent_dim; p.A = c.action_dim; p.M = c.mlp_dim;
  p.obs = a->obs; p.next_obses = a->next_obses; p.action = a->action; p.reward = a->reward;
  p.noise = NoiseSrc{a->noise_pi, a->seed, a->call, NS_PI};
  p.min_std = a->min_std; p.discount = (float)a->discount;
  double r = 1.0;                       // rho = cfg.rho ** t: a python float, rounded once when it multiplies the fp32 tensor
  for (int t = 0; t < a->horizon; ++t) { p.rho_pow[t] = (float)r; r *= a->rho; }
  p.c_cons = a->consistency_coef; p.c_rew = a->reward_coef; p.c_val = a->value_coef;
  p.consistency = a->consistency_loss; p.reward_loss = a->reward_loss; p.value = a->value_loss; p.priority = a->priority_loss;
  p.total = a->total_loss; p.zs = a->zs; p.td_target = a->td_target; p.q1 = a->q1; p.q2 = a->q2; p.reward_pred = a->reward_pred;
  if (launch_update_fwd(h->m32, ht->m32, p, (cudaStream_t)stream) < 0) FAIL(h, TDMPC_ERR_INVALID, "update_forward: shape does not fit in shared memory");
  return check_launch(h, "update_fwd");
}

extern "C" int64_t tdmpc_train_param_count(const tdmpc_model_cfg* cfg, int64_t* pi_begin, int64_t* pi_end) {
  if (!cfg) return -1;
  long long lo = 0, hi = 0;
  const long long n = train_param_layout(*cfg, &lo, &hi);
  if (pi_begin) *pi_begin = lo;
  if (pi_end) *pi_end = hi;
  return n;
}
extern "C" int64_t tdmpc_train_workspace_bytes(const tdmpc_model_cfg* cfg, int32_t batch, int32_t horizon) {
  if (!cfg || batch <= 0 || horizon <= 0) return -1;
  return train_workspace_bytes(*cfg, batch, horizon);
}
extern "C" int tdmpc_update(tdmpc_handle* h, const tdmpc_train_args* a, void* stream) {
  if (!h || !a) return TDMPC_ERR_INVALID;
  const tdmpc_model_cfg& c = h->cfg;
  if (c.dynamics != TDMPC_DYN_MLP || c.enc_layernorm) FAIL(h, TDMPC_ERR_INVALID, "update: MLP dynamics and the plain state encoder only");
  if (a->batch <= 0 || a->horizon <= 0 || a->horizon >= TDMPC_MAX_H) FAIL(h, TDMPC_ERR_INVALID, "update: bad batch / horizon");
  if (!a->obs || !a->next_obses || !a->action || !a->reward || !a->weights || !a->metrics || !a->priorities || !a->grads || !a->adam_m ||
      !a->adam_v || !a->pi_adam_m || !a->pi_adam_v || !a->workspace)
    FAIL(h, TDMPC_ERR_INVALID, "update: a required pointer is NULL");
  const tdmpc_weights* ws[2] = {&a->online, &a->target};
  for (const tdmpc_weights* w : ws) {
    bool ok = w->enc_w[0] && w->enc_w[1] && w->enc_b[0] && w->enc_b[1];
    for (int i = 0; i < 3; ++i)
      ok = ok && w->dyn_w[i] && w->dyn_b[i] && w->rew_w[i] && w->rew_b[i] && w->pi_w[i] && w->pi_b[i] && w->q_w[0][i] && w->q_b[0][i] && w->q_w[1][i] && w->q_b[1][i];
    for (int i = 0; i < 2; ++i) ok = ok && w->q_ln_g[0][i] && w->q_ln_b[0][i] && w->q_ln_g[1][i] && w->q_ln_b[1][i];
    if (!ok) FAIL(h, TDMPC_ERR_INVALID, "update: a parameter pointer is NULL");
  }
  if (a->workspace_bytes < train_workspace_bytes(c, a->batch, a->horizon)) FAIL(h, TDMPC_ERR_INVALID, "update: workspace too small");
  cudaSetDevice(c.device);
  const int n = launch_train_step(c, *a, (cudaStream_t)stream);
  if (n < 0) FAIL(h, TDMPC_ERR_CUDA, "update: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  h->launches += (uint64_t)n;
  return TDMPC_OK;
}

extern "C" int tdmpc_fill_noise(tdmpc_handle* h, int32_t stream_id, uint64_t seed, uint64_t call, uint64_t count,
                                void* out, void* stream) {
  if (!h || !out || stream_id < 0 || stream_id > NS_SHIFT) return TDMPC_ERR_INVALID;
  cudaSetDevice(h->cfg.device);
  launch_fill_noise(stream_id, seed, call, count, out, (cudaStream_t)stream);
  return check_launch(h, "fill_noise");
}

extern "C" int tdmpc_plan(tdmpc_handle* h, const tdmpc_plan_args* a, void* stream) {
  if (!h || !a) return TDMPC_ERR_INVALID;
  if (!h->weights_set) FAIL(h, TDMPC_ERR_STATE, "plan: weights not set (call tdmpc_set_weights)");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = a->batch, H = a->horizon, I = a->iterations, K = a->num_elites, P0 = a->num_pi_trajs;
  const int A = h->cfg.action_dim, L = h->cfg.latent_dim;
  if (B <= 0 || H <= 0 || H > TDMPC_MAX_H || I <= 0 || I > TDMPC_MAX_ITERS || K <= 0 || P0 < 0)
    FAIL(h, TDMPC_ERR_INVALID, "plan: bad sizes B=%d H=%d I=%d K=%d P=%d", B, H, I, K, P0);
  if (H * A > select_max_ha()) FAIL(h, TDMPC_ERR_INVALID, "plan: H*A=%d too large", H * A);
  int Tmax = 0, Nmax = 0;
  for (int i = 0; i < I; ++i) {
    if (a->n_samples[i] < 0 || a->n_reuse[i] < 0 || a->n_pi[i] < 0 || a->n_pi[i] > P0 || a->n_reuse[i] > K)
      FAIL(h, TDMPC_ERR_INVALID, "plan: bad row counts at iteration %d", i);
    const int T = a->n_samples[i] + a->n_reuse[i] + a->n_pi[i];
    if (T < K) FAIL(h, TDMPC_ERR_INVALID, "plan: iteration %d has %d trajectories < num_elites %d", i, T, K);
    Tmax = T > Tmax ? T : Tmax;
    Nmax = a->n_samples[i] > Nmax ? a->n_samples[i] : Nmax;
  }
  if (B > h->capB || Tmax > h->capT || P0 > h->capT || H > h->capH || K > h->capK)
    FAIL(h, TDMPC_ERR_WORKSPACE, "plan: call tdmpc_reserve(%d, %d, %d, %d) first", B, Tmax > P0 ? Tmax : P0, H, K);
  if (!a->z0 && !a->obs) FAIL(h, TDMPC_ERR_INVALID, "plan: obs and z0 both NULL");
  if (!a->action || !a->metrics || !a->prev_mean) FAIL(h, TDMPC_ERR_INVALID, "plan: action/metrics/prev_mean must be set");
  if (a->t0_all < 0 && !a->t0) FAIL(h, TDMPC_ERR_INVALID, "plan: t0_all=-1 needs t0[]");
  cudaSetDevice(h->cfg.device);

  const float* z0 = a->z0;
  if (!z0) {
    launch_encode(h->m32, a->obs, B, h->cfg.obs_dim, h->cfg.enc_dim, L, h->z0, st);
    if (int r = check_launch(h, "encode")) return r;
    z0 = h->z0;
  }
  if (a->tap_z0) CUDA_OK(h, cudaMemcpyAsync(a->tap_z0, z0, sizeof(float) * B * L, cudaMemcpyDeviceToDevice, st));

  RolloutParams rp{};
  rp.B = B; rp.H = H; rp.L = L; rp.A = A; rp.M = h->cfg.mlp_dim; rp.z0 = z0; rp.min_std = a->min_std;
  rp.hidden0 = h->cfg.dynamics == TDMPC_DYN_DRNN ? a->hidden : nullptr;
  fill_disc(a->discount, H, rp.disc);

  // policy trajectories (tdmpc.py:113-118)
  if (P0 > 0) {
    RolloutParams pp = rp;
    pp.mode = MODE_PI; pp.T = P0; pp.Tcap = P0; pp.pi_actions = h->pi_actions;
    pp.noise = NoiseSrc{a->noise_pi, a->seed, a->call, NS_PI, a->call_dev};
    pp.noise_base = 0; pp.noise_bstride = (uint64_t)H * P0 * A; pp.noise_tstride = (uint64_t)P0 * A;
    if (int r = run_rollout(h, pp, st)) return r;
  }

  float* elites = a->elite_actions ? a->elite_actions : h->elites;
  bool any_reuse = false;
  for (int i = 0; i < I; ++i) any_reuse = any_reuse || a->n_reuse[i] > 0;
  if (a->n_reuse[0] > 0 && !a->elite_actions) FAIL(h, TDMPC_ERR_INVALID, "plan: n_reuse[0] > 0 needs elite_actions from the previous step");
  const int CL = a->color_len;
  if (CL < 0 || CL > TDMPC_MAX_H || (CL > 0 && (!a->color_mats || CL < H))) FAIL(h, TDMPC_ERR_INVALID, "plan: bad colored-noise spec");
  const int eh = a->elite_horizon > 0 ? a->elite_horizon : H;
  if (eh > H || eh < 1) FAIL(h, TDMPC_ERR_INVALID, "plan: elite_horizon %d out of range", eh);
  // time rows of the cem noise stream: injected buffers are [.., H, Nmax, A]; the built-in colored
  // generator needs color_len white rows per trajectory
  const int Trows = (a->noise_cem || CL == 0) ? H : CL;
  const NoiseSrc cem{a->noise_cem, a->seed, a->call, NS_CEM, a->call_dev};
  const uint64_t cem_bstride = (uint64_t)I * Trows * Nmax * A;
  auto spec = [&](int i) {
    SampleSpec sp{};
    sp.N = a->n_samples[i]; sp.E = a->n_reuse[i]; sp.P = a->n_pi[i];
    sp.best_a = (a->best_a_last && i == I - 1) ? 1 : 0;
    sp.cem = cem; sp.cem_base = (uint64_t)i * Trows * Nmax * A; sp.cem_bstride = cem_bstride; sp.cem_rows = Nmax;
    sp.color_mats = a->color_mats; sp.color_len = CL;
    sp.n_white = CL > 0 ? a->n_white[i] : sp.N; sp.n_pink = CL > 0 ? a->n_color1[i] : 0;
    return sp;
  };

  // With one CTA per problem a small batch leaves the sampling of N*H*A normals per iteration to a handful of SMs (B = 1, dog
  // dims: 250 us per iteration): below two CTAs per SM the rows are drawn by sample_kernel, `nparts` CTAs per problem.
  int nparts = 1;
  if (B < 2 * h->num_sms) {
    int most = 1;
    for (int i = 0; i < I; ++i) most = a->n_samples[i] > most ? a->n_samples[i] : most;
    const int want = (2 * h->num_sms + B - 1) / B, useful = (int)(((size_t)most * H * A + 1023) / 1024);
    nparts = want < useful ? want : useful;
    nparts = nparts > 64 ? 64 : (nparts < 8 ? 1 : nparts);     // under 8 k normals per iteration the extra launch costs more than it saves
  }
  auto sample_split = [&](const SampleSpec& s_, const float* reuse) -> int {
    SampleParams q{};
    q.B = B; q.H = H; q.A = A; q.Tcap = Tmax; q.spec = s_; q.mean = h->mean; q.std = h->std; q.actions = h->actions;
    q.pi_actions = h->pi_actions; q.P0 = P0; q.reuse = reuse; q.reuse_rows = K;
    launch_sample(q, nparts, st);
    return check_launch(h, "sample");
  };

  InitParams ip{};
  ip.B = B; ip.H = H; ip.A = A; ip.Tcap = Tmax; ip.first = spec(0); ip.skip_sample = nparts > 1;
  ip.init_std = a->init_std; ip.t0_all = a->t0_all; ip.t0 = a->t0; ip.has_prev = a->has_prev; ip.warm_last = a->warm_last;
  ip.prev_mean = a->prev_mean; ip.mean = h->mean; ip.std = h->std; ip.actions = h->actions;
  ip.pi_actions = h->pi_actions; ip.P0 = P0; ip.elite_actions = elites; ip.K = K; ip.shift_elites = a->shift_elites; ip.elite_h = eh;
  ip.shift = NoiseSrc{a->noise_shift, a->seed, a->call, NS_SHIFT, a->call_dev};
  ip.shift_mat = (CL > 0 && !a->noise_shift) ? a->color_mats + 2 * (size_t)CL * CL : nullptr;
  launch_init(ip, st);
  if (int r = check_launch(h, "init")) return r;
  if (nparts > 1) if (int r = sample_split(ip.first, nullptr)) return r;

  for (int i = 0; i < I; ++i) {
    const int T = a->n_samples[i] + a->n_reuse[i] + a->n_pi[i];
    float* values = a->tap_values ? a->tap_values + (size_t)i * B * Tmax : h->values;
    RolloutParams vp = rp;
    vp.mode = MODE_VALUE; vp.T = T; vp.Tcap = Tmax; vp.actions = h->actions; vp.values = values; vp.last_reward = h->last_reward;
    vp.noise = NoiseSrc{a->noise_term, a->seed, a->call, NS_TERM, a->call_dev};
    vp.noise_base = (uint64_t)i * Tmax * A; vp.noise_bstride = (uint64_t)I * Tmax * A;
    if (int r = run_rollout(h, vp, st)) return r;

    SelectParams sp{};
    sp.B = B; sp.T = T; sp.Tcap = Tmax; sp.H = H; sp.A = A; sp.K = K; sp.iter = i; sp.n_iters = I;
    sp.values = values; sp.last_reward = h->last_reward; sp.actions = h->actions;
    sp.mean = h->mean; sp.std = h->std;
    sp.temperature = a->temperature; sp.momentum = (float)a->momentum; sp.one_minus_momentum = (float)(1.0 - a->momentum);
    sp.std_floor = a->std_floor; sp.std_max = a->std_max;
    sp.elite_idx = h->elite_idx; sp.score = h->score;
    sp.elite_actions = (a->elite_actions || any_reuse) ? elites : nullptr;
    const bool last = i == I - 1;
    sp.do_sample = !last && nparts == 1; sp.do_output = last;
    if (!last) {
      sp.next = spec(i + 1);
      sp.pi_actions = h->pi_actions; sp.P0 = P0;
    }
    sp.eval_mode = a->eval_mode;
    sp.out_noise = NoiseSrc{a->noise_out, a->seed, a->call, NS_OUT, a->call_dev};
    sp.pick = NoiseSrc{nullptr, a->seed, a->call, NS_PICK, a->call_dev};
    sp.u_pick = a->u_pick;
    sp.action = a->action; sp.metrics = a->metrics; sp.prev_mean = a->prev_mean;
    sp.tap_elite_idx = a->tap_elite_idx; sp.tap_mean = a->tap_mean; sp.tap_std = a->tap_std;
    launch_select(sp, B < 2 * h->num_sms, st);
    if (int r = check_launch(h, "select")) return r;
    if (!last && nparts > 1) if (int r = sample_split(sp.next, sp.elite_actions)) return r;
  }
  if (h->cfg.dynamics == TDMPC_DYN_DRNN && a->hidden_out) {
    // step the model with the chosen action to get the next GRU state (tdmpc_icem_similarity_drnn.py:267)
    RolloutParams sp = rp;
    sp.mode = MODE_STEP; sp.T = 1; sp.Tcap = 1; sp.H = 1; sp.actions = a->action; sp.hidden_out = a->hidden_out;
    if (int r = run_rollout(h, sp, st)) return r;
  }
  if (a->call_dev) {
    launch_bump_counter(a->call_dev, st);
    if (int r = check_launch(h, "call counter")) return r;
  }
  return TDMPC_OK;
}
