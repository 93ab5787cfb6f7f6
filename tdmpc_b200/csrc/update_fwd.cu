// Forward pass of the TOLD training step: the loss terms of TDMPC.update
// (algorithm/tdmpc.py:191-228) and the TD target of TDMPC._td_target (tdmpc.py:180-189) for one replay batch, fp32 on the
// CUDA cores (the 1e-5 parity mode; same building blocks as the fp32 rollout kernel).
//
// One CTA owns a tile of 32 replay samples and unrolls the latent model over the horizon with the activations resident in
// shared memory, exactly as the reference does per batch:
//     z = h(obs)
//     for t < H:  Q1, Q2 = Q(z, a_t);  z, r = next(z, a_t)
//                 next_z = h_target(next_obs_t);  td = reward_t + gamma * min(Q_target(h(next_obs_t), pi(h(next_obs_t), min_std)))
//                 consistency += rho^t mean((z - next_z)^2);  reward += rho^t (r - reward_t)^2
//                 value += rho^t ((Q1 - td)^2 + (Q2 - td)^2);  priority += rho^t (|Q1 - td| + |Q2 - td|)
//     total = c_c min(consistency, 1e4) + c_r min(reward, 1e4) + c_v min(value, 1e4)
// (state observations: the reference's augmentation is the identity, helper.py:256-262).  Two parameter sets take part -- the
// online model and the slow-moving target -- so the entry point takes two handles.  Backward / optimiser are not implemented.
#include "common.cuh"
#include "fp32_layers.cuh"

namespace {

// TOLD.h on a tile: Linear(D, E) -> (LayerNorm) -> ELU -> Linear(E, L)  (helper.py:131-133, 161-165); `hid` is scratch [R][ldh]
__device__ void encode_tile(const ModelFp32& m, const float* x, int ldx, float* hid, int ldh, float* out, int ldo) {
  if (m.enc_has_ln) {
    layer<ACT_NONE>(x, ldx, m.enc[0], hid, ldh);
    layernorm_act<false>(hid, ldh, m.enc[0].N, m.enc_ln, 1e-5f);
  } else {
    layer<ACT_ELU>(x, ldx, m.enc[0], hid, ldh);
  }
  layer<ACT_NONE>(hid, ldh, m.enc[1], out, ldo);
}

// both Q heads of `m` on the [z, a] tile (helper.py:197-201): q1o / q2o [R]
__device__ void q_heads(const ModelFp32& m, const float* za, int ldz, float* xa, float* xb, int M, float* q1o, float* q2o) {
  for (int h = 0; h < 2; ++h) {
    layer<ACT_NONE>(za, ldz, m.q[h][0], xa, M);
    layernorm_act<true>(xa, M, M, m.q_ln[h][0], 1e-5f);
    layer<ACT_NONE>(xa, M, m.q[h][1], xb, M);
    layernorm_act<false>(xb, M, M, m.q_ln[h][1], 1e-5f);
    head_dot(xb, M, m.q[h][2], h == 0 ? q1o : q2o);
  }
}

__global__ void __launch_bounds__(NT, 1) update_fwd_kernel(ModelFp32 m, ModelFp32 mt, UpdateFwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int L = p.L, A = p.A, M = p.M, D = p.D;
  const int ldz = m.pi[0].Kp;                  // padded L + A
  const int ldn = (L + 3) & ~3;
  const int ldo = (D + 3) & ~3;
  float* xa = smem;                            // [R][M]   (also the encoder's hidden layer, E <= M)
  float* xb = xa + R * M;                      // [R][M]
  float* za = xb + R * M;                      // [R][ldz]  [z | a_t | 0]
  float* zb = za + R * ldz;                    // [R][ldz]  [h(next_obs) | pi(.) | 0]
  float* zn = zb + R * ldz;                    // [R][ldn]  z' = dynamics([z, a])
  float* zt = zn + R * ldn;                    // [R][ldn]  h_target(next_obs)
  float* ob = zt + R * ldn;                    // [R][ldo]  observation tile
  float* act = ob + R * ldo;                   // [R][A]
  float* q1 = act + R * A;                     // [R] each
  float* q2 = q1 + R;
  float* t1 = q2 + R;
  float* t2 = t1 + R;
  float* rp = t2 + R;

  const int tid = threadIdx.x;
  const long long g0 = (long long)blockIdx.x * R;
  auto load_obs = [&](const float* src) {      // src: [B][D] slice
    for (int idx = tid; idx < R * ldo; idx += NT) {
      const int r = idx / ldo, c = idx - r * ldo;
      const long long g = g0 + r;
      ob[idx] = (g < p.B && c < D) ? __ldg(src + (size_t)g * D + c) : 0.0f;
    }
    __syncthreads();
  };
  for (int idx = tid; idx < R * ldz; idx += NT) { za[idx] = 0.0f; zb[idx] = 0.0f; }
  __syncthreads();

  // z = h(obs)
  load_obs(p.obs);
  encode_tile(m, ob, ldo, xa, M, za, ldz);
  if (p.zs)
    for (int idx = tid; idx < R * L; idx += NT) {
      const int r = idx / L, c = idx - r * L;
      if (g0 + r < p.B) p.zs[(size_t)(g0 + r) * L + c] = za[r * ldz + c];
    }

  float cons = 0.0f, rew = 0.0f, val = 0.0f, pri = 0.0f;        // per-sample running sums: threads tid < R
  for (int t = 0; t < p.H; ++t) {
    for (int idx = tid; idx < R * A; idx += NT) {
      const int r = idx / A, a = idx - r * A;
      const long long g = g0 + r;
      za[r * ldz + L + a] = g < p.B ? __ldg(p.action + ((size_t)t * p.B + g) * A + a) : 0.0f;
    }
    __syncthreads();
    // Q1, Q2 = Q(z, a_t); z', r = next(z, a_t)
    q_heads(m, za, ldz, xa, xb, M, q1, q2);
    layer<ACT_ELU>(za, ldz, m.dyn[0], xa, M);
    layer<ACT_ELU>(xa, M, m.dyn[1], xb, M);
    layer<ACT_NONE>(xb, M, m.dyn[2], zn, ldn);
    layer<ACT_ELU>(za, ldz, m.rew[0], xa, M);
    layer<ACT_ELU>(xa, M, m.rew[1], xb, M);
    head_dot(xb, M, m.rew[2], rp);
    // next_z = h_target(next_obs_t);  td target through the ONLINE encoder and policy and the TARGET Q heads (tdmpc.py:184-188)
    load_obs(p.next_obses + (size_t)t * p.B * D);
    encode_tile(mt, ob, ldo, xa, M, zt, ldn);
    encode_tile(m, ob, ldo, xa, M, zb, ldz);
    for (int idx = tid; idx < R * A; idx += NT) zb[(idx / A) * ldz + L + idx % A] = 0.0f;      // the policy's first Linear has zero rows there
    __syncthreads();
    layer<ACT_ELU>(zb, ldz, m.pi[0], xa, M);
    layer<ACT_ELU>(xa, M, m.pi[1], xb, M);
    layer<ACT_NONE>(xb, M, m.pi[2], act, A);
    for (int idx = tid; idx < R * A; idx += NT) {
      const int r = idx / A, a = idx - r * A;
      const long long g = g0 + r;
      float v = 0.0f;
      if (g < p.B) v = policy_sample(tanhf(act[idx]), p.min_std, noise_at(p.noise, ((uint64_t)t * p.B + g) * A + a));
      zb[r * ldz + L + a] = v;
    }
    __syncthreads();
    q_heads(mt, zb, ldz, xa, xb, M, t1, t2);
    if (tid < R && g0 + tid < p.B) {
      const long long g = g0 + tid;
      const float rt = __ldg(p.reward + (size_t)t * p.B + g);
      const float td = __fadd_rn(rt, __fmul_rn(p.discount, fminf(t1[tid], t2[tid])));          // reward + discount * min(Q_target)
      float ms = 0.0f;
      for (int c = 0; c < L; ++c) { const float d = zn[tid * ldn + c] - zt[tid * ldn + c]; ms = fmaf(d, d, ms); }
      const float rho = p.rho_pow[t];
      const float d1 = q1[tid] - td, d2 = q2[tid] - td, dr = rp[tid] - rt;
      cons = __fadd_rn(cons, __fmul_rn(rho, ms / (float)L));
      rew = __fadd_rn(rew, __fmul_rn(rho, dr * dr));
      val = __fadd_rn(val, __fmul_rn(rho, __fadd_rn(d1 * d1, d2 * d2)));
      pri = __fadd_rn(pri, __fmul_rn(rho, __fadd_rn(fabsf(d1), fabsf(d2))));
      if (p.td_target) p.td_target[(size_t)t * p.B + g] = td;
      if (p.q1) p.q1[(size_t)t * p.B + g] = q1[tid];
      if (p.q2) p.q2[(size_t)t * p.B + g] = q2[tid];
      if (p.reward_pred) p.reward_pred[(size_t)t * p.B + g] = rp[tid];
    }
    __syncthreads();
    for (int idx = tid; idx < R * L; idx += NT) {
      const int r = idx / L, c = idx - r * L;
      const float v = zn[r * ldn + c];
      za[r * ldz + c] = v;
      if (p.zs && g0 + r < p.B) p.zs[((size_t)(t + 1) * p.B + g0 + r) * L + c] = v;
    }
    __syncthreads();
  }
  if (tid < R && g0 + tid < p.B) {
    const long long g = g0 + tid;
    p.consistency[g] = cons; p.reward_loss[g] = rew; p.value[g] = val; p.priority[g] = pri;
    p.total[g] = __fadd_rn(__fadd_rn(__fmul_rn(p.c_cons, fminf(cons, 1e4f)), __fmul_rn(p.c_rew, fminf(rew, 1e4f))), __fmul_rn(p.c_val, fminf(val, 1e4f)));
  }
}

}  // namespace

// returns the dynamic shared-memory bytes, or 0 if the shape does not fit
size_t update_fwd_smem(const ModelFp32& m, int D, int L, int A, int M, int E) {
  if (E > M) return 0;
  const int ldz = m.pi[0].Kp, ldn = (L + 3) & ~3, ldo = (D + 3) & ~3;
  return sizeof(float) * ((size_t)2 * R * M + 2 * (size_t)R * ldz + 2 * (size_t)R * ldn + (size_t)R * ldo + (size_t)R * A + 5 * R);
}

int launch_update_fwd(const ModelFp32& m, const ModelFp32& mt, const UpdateFwdParams& p, cudaStream_t st) {
  const size_t smem = update_fwd_smem(m, p.D, p.L, p.A, p.M, p.E);
  if (smem == 0 || smem > 227 * 1024) return -1;
  if (cudaFuncSetAttribute(update_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  update_fwd_kernel<<<(p.B + R - 1) / R, NT, smem, st>>>(m, mt, p);
  return 1;
}
