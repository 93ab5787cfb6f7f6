// CUDA-core fp32 trajectory rollout: the "fp32 mode" of the planner (1e-5 parity with the
// reference).  One CTA owns a tile of 32 trajectories and walks them through the whole horizon
// with the activations resident in shared memory; weights stream from L2 (they are shared by
// every CTA).  Restates TDMPC.estimate_value (algorithm/tdmpc.py:83-92) and the policy rollout
// of TDMPC.plan (tdmpc.py:113-118) on top of TOLD.next / pi / Q (tdmpc.py:34-50).
#include "common.cuh"
#include "fp32_layers.cuh"

namespace {

// NormGRUCell gates (models/rnns.py:19-29), warp per trajectory.  gi, gh: [R][3*Hd] pre-activations of
// weight_ih([z,a]) and weight_hh(h); chunk order reset, update, newval; LayerNorm eps 1e-3;
//   r = sigmoid(LN_r(gi_r + gh_r)); u = sigmoid(LN_u(gi_u + gh_u)); n = tanh(LN_n(gi_n + r * gh_n));
//   h' = u * n + (1 - u) * h          (note: the opposite convention of torch.nn.GRUCell)
__device__ void gru_gates(const float* __restrict__ gi, const float* __restrict__ gh, int ldg, float* __restrict__ hid, int Hd,
                          const LayerNormP* ln) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv = 1.0f / (float)Hd;
  auto ln_stats = [&](auto&& val, float& mu, float& rstd) {
    float s = 0.0f;
    for (int k = lane; k < Hd; k += 32) s += val(k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mu = s * inv;
    float v = 0.0f;
    for (int k = lane; k < Hd; k += 32) { const float d = val(k) - mu; v = fmaf(d, d, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    rstd = 1.0f / sqrtf(v * inv + 1e-3f);
  };
  for (int r = warp; r < R; r += NW) {
    const float* gir = gi + r * ldg;
    const float* ghr = gh + r * ldg;
    float* h = hid + r * Hd;
    float mu, rstd;
    auto xr = [&](int k) { return gir[k] + ghr[k]; };
    ln_stats(xr, mu, rstd);
    // reset gate kept in registers: Hd <= 128 -> at most 4 values per lane
    float rg[4];
    for (int i = 0, k = lane; k < Hd; k += 32, ++i)
      rg[i] = 1.0f / (1.0f + expf(-((xr(k) - mu) * rstd * __ldg(ln[0].g + k) + __ldg(ln[0].b + k))));
    auto xu = [&](int k) { return gir[Hd + k] + ghr[Hd + k]; };
    float mu_u, rstd_u;
    ln_stats(xu, mu_u, rstd_u);
    // newval pre-activation needs the lane's own reset values only (same k)
    float nv[4];
    for (int i = 0, k = lane; k < Hd; k += 32, ++i) nv[i] = gir[2 * Hd + k] + rg[i] * ghr[2 * Hd + k];
    float s = 0.0f;
    for (int i = 0, k = lane; k < Hd; k += 32, ++i) s += nv[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu_n = s * inv;
    float v = 0.0f;
    for (int i = 0, k = lane; k < Hd; k += 32, ++i) { const float d = nv[i] - mu_n; v = fmaf(d, d, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd_n = 1.0f / sqrtf(v * inv + 1e-3f);
    for (int i = 0, k = lane; k < Hd; k += 32, ++i) {
      const float u = 1.0f / (1.0f + expf(-((xu(k) - mu_u) * rstd_u * __ldg(ln[1].g + k) + __ldg(ln[1].b + k))));
      const float n = tanhf((nv[i] - mu_n) * rstd_n * __ldg(ln[2].g + k) + __ldg(ln[2].b + k));
      h[k] = u * n + (1.0f - u) * h[k];
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT, 1) rollout_fp32_kernel(ModelFp32 m, RolloutParams p) {
  extern __shared__ __align__(16) float smem[];
  const int L = p.L, A = p.A, M = p.M;
  const int ldz = m.pi[0].Kp;                  // padded L+A (pi[0] is packed with the action rows zeroed)
  const int ldn = (L + 3) & ~3;
  float* xa = smem;                            // [R][M]
  float* xb = xa + R * M;                      // [R][M]
  float* za = xb + R * M;                      // [R][ldz]   [z | a | 0 pad]
  float* zn = za + R * ldz;                    // [R][ldn]
  float* act = zn + R * ldn;                   // [R][A]
  float* G = act + R * A;                      // [R]
  float* rr = G + R;                           // [R]
  float* q1 = rr + R;                          // [R]
  float* q2 = q1 + R;                          // [R]
  float* hid = q2 + R;                         // [R][Hd] GRU state (DRNN dynamics only)

  const long long total = (long long)p.B * p.T;
  const long long g0 = (long long)blockIdx.x * R;
  const int tid = threadIdx.x;

  for (int idx = tid; idx < R * ldz; idx += NT) {
    const int r = idx / ldz, c = idx - r * ldz;
    const long long g = g0 + r;
    float v = 0.0f;
    if (g < total && c < L) v = __ldg(p.z0 + (size_t)(g / p.T) * L + c);
    za[idx] = v;
  }
  if (tid < R) { G[tid] = 0.0f; rr[tid] = 0.0f; }
  if (m.drnn)
    for (int idx = tid; idx < R * m.Hd; idx += NT) {
      const long long g = g0 + idx / m.Hd;
      hid[idx] = (p.hidden0 && g < total) ? __ldg(p.hidden0 + (size_t)(g / p.T) * m.Hd + idx % m.Hd) : 0.0f;
    }
  __syncthreads();

  // a = pi(z, min_std): tanh(MLP(z)) + truncated noise; written into za[:, L:L+A]
  auto policy = [&](uint64_t noise_off) {
    layer<ACT_ELU>(za, ldz, m.pi[0], xa, M);
    layer<ACT_ELU>(xa, M, m.pi[1], xb, M);
    layer<ACT_NONE>(xb, M, m.pi[2], act, A);
    for (int idx = tid; idx < R * A; idx += NT) {
      const int r = idx / A, a = idx - r * A;
      const long long g = g0 + r;
      float v = 0.0f;
      if (g < total) {
        const long long b = g / p.T, j = g - b * p.T;
        const float eps = noise_at(p.noise, p.noise_base + noise_off + (uint64_t)b * p.noise_bstride + (uint64_t)j * A + a);
        v = policy_sample(tanhf(act[idx]), p.min_std, eps);
      }
      za[r * ldz + L + a] = v;
    }
    __syncthreads();
  };

  for (int t = 0; t < p.H; ++t) {
    if (p.mode != MODE_PI) {
      for (int idx = tid; idx < R * A; idx += NT) {
        const int r = idx / A, a = idx - r * A;
        const long long g = g0 + r;
        float v = 0.0f;
        if (g < total) {
          const long long b = g / p.T, j = g - b * p.T;
          v = __ldg(p.actions + ((size_t)(b * p.Tcap + j) * p.H + t) * A + a);
        }
        za[r * ldz + L + a] = v;
      }
      __syncthreads();
    } else {
      policy((uint64_t)t * p.noise_tstride);
      for (int idx = tid; idx < R * A; idx += NT) {
        const int r = idx / A, a = idx - r * A;
        const long long g = g0 + r;
        if (g < total) {
          const long long b = g / p.T, j = g - b * p.T;
          p.pi_actions[((size_t)(b * p.Tcap + j) * p.H + t) * A + a] = za[r * ldz + L + a];
        }
      }
    }
    if (!m.drnn) {
      // z' = dynamics([z, a])
      layer<ACT_ELU>(za, ldz, m.dyn[0], xa, M);
      layer<ACT_ELU>(xa, M, m.dyn[1], xb, M);
      layer<ACT_NONE>(xb, M, m.dyn[2], zn, ldn);
      if (p.mode == MODE_VALUE) {
        // r = reward([z, a]);  G += gamma^t r
        layer<ACT_ELU>(za, ldz, m.rew[0], xa, M);
        layer<ACT_ELU>(xa, M, m.rew[1], xb, M);
        head_dot(xb, M, m.rew[2], rr);
        if (tid < R) G[tid] = __fadd_rn(G[tid], __fmul_rn(p.disc[t], rr[tid]));   // mul, then add, as torch does
      }
    } else {
      // DSSM.next (tdmpc_icem_similarity_drnn.py:56-59): h' = NormGRU([z,a], h); z' = prior_mlp(h'); r = reward(h')
      const int G3 = 3 * m.Hd;
      layer<ACT_NONE>(za, ldz, m.gru_ih, xa, G3);
      layer<ACT_NONE>(hid, m.Hd, m.gru_hh, xb, G3);
      gru_gates(xa, xb, G3, hid, m.Hd, m.gru_ln);
      layer<ACT_ELU>(hid, m.Hd, m.prior[0], xa, M);          // Linear + eval BatchNorm (affine) + ELU
      layer<ACT_NONE>(xa, M, m.prior[1], zn, ldn);
      if (p.mode == MODE_VALUE) {
        layer<ACT_ELU>(hid, m.Hd, m.rew[0], xa, M);
        layer<ACT_ELU>(xa, M, m.rew[1], xb, M);
        head_dot(xb, M, m.rew[2], rr);
        if (tid < R) G[tid] = __fadd_rn(G[tid], __fmul_rn(p.disc[t], rr[tid]));
      }
    }
    for (int idx = tid; idx < R * L; idx += NT) {
      const int r = idx / L, c = idx - r * L;
      za[r * ldz + c] = zn[r * ldn + c];
    }
    __syncthreads();
  }

  if (p.mode == MODE_STEP && m.drnn && p.hidden_out)
    for (int idx = tid; idx < R * m.Hd; idx += NT) {
      const long long g = g0 + idx / m.Hd;
      if (g < total) p.hidden_out[(size_t)g * m.Hd + idx % m.Hd] = hid[idx];
    }
  if (p.mode == MODE_VALUE) {
    policy(0);
    for (int h = 0; h < 2; ++h) {
      layer<ACT_NONE>(za, ldz, m.q[h][0], xa, M);
      layernorm_act<true>(xa, M, M, m.q_ln[h][0], 1e-5f);
      layer<ACT_NONE>(xa, M, m.q[h][1], xb, M);
      layernorm_act<false>(xb, M, M, m.q_ln[h][1], 1e-5f);
      head_dot(xb, M, m.q[h][2], h == 0 ? q1 : q2);
    }
    if (tid < R) {
      const long long g = g0 + tid;
      if (g < total) {
        const float v = __fadd_rn(G[tid], __fmul_rn(p.disc[p.H], min_nanprop(q1[tid], q2[tid])));
        const long long b = g / p.T, j = g - b * p.T;
        p.values[b * p.Tcap + j] = nan_to_num0(v);
        if (p.last_reward) p.last_reward[b * p.Tcap + j] = rr[tid];
      }
    }
  }
}

}  // namespace

void launch_rollout_fp32(const ModelFp32& m, const RolloutParams& p, cudaStream_t st) {
  const long long total = (long long)p.B * p.T;
  const int tiles = (int)((total + R - 1) / R);
  const int ldz = m.pi[0].Kp, ldn = (p.L + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)2 * R * p.M + (size_t)R * ldz + (size_t)R * ldn + (size_t)R * p.A + 4 * R +
                                       (m.drnn ? (size_t)R * m.Hd : 0));
  // per function AND per device: remember what each device of this process was raised to
  static size_t configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    cudaFuncSetAttribute(rollout_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  rollout_fp32_kernel<<<tiles, NT, smem, st>>>(m, p);
}
