// TOLD.h (algorithm/tdmpc.py:30-32): state encoder Linear(D,E) -> [LayerNorm(E)] -> ELU -> Linear(E,L)
// (helper.py:131-133 / helper.py:161-165).  One CTA per observation; fp32 in every precision mode
// (2*(D*E+E*L) FLOPs per problem -- negligible next to the rollout).
#include "common.cuh"

namespace {

constexpr int NT = 256;

__global__ void __launch_bounds__(NT) encode_kernel(ModelFp32 m, const float* __restrict__ obs, int D, int E, int L,
                                                    float* __restrict__ z) {
  extern __shared__ float sm[];
  float* x = sm;          // [D]
  float* h = sm + D;      // [E]
  __shared__ float red[2];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < D; i += NT) x[i] = obs[(size_t)b * D + i];
  __syncthreads();
  for (int n = threadIdx.x; n < E; n += NT) {
    float acc = 0.0f;
    for (int k = 0; k < D; ++k) acc = fmaf(x[k], __ldg(m.enc[0].wt + (size_t)k * E + n), acc);
    h[n] = acc + __ldg(m.enc[0].b + n);
  }
  __syncthreads();
  if (m.enc_has_ln) {
    if (threadIdx.x < 32) {
      float s = 0.0f;
      for (int k = threadIdx.x; k < E; k += 32) s += h[k];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mu = s / (float)E;
      float v = 0.0f;
      for (int k = threadIdx.x; k < E; k += 32) { const float d = h[k] - mu; v = fmaf(d, d, v); }
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (threadIdx.x == 0) { red[0] = mu; red[1] = 1.0f / sqrtf(v / (float)E + 1e-5f); }
    }
    __syncthreads();
    for (int n = threadIdx.x; n < E; n += NT)
      h[n] = (h[n] - red[0]) * red[1] * __ldg(m.enc_ln.g + n) + __ldg(m.enc_ln.b + n);
    __syncthreads();
  }
  for (int n = threadIdx.x; n < E; n += NT) h[n] = elu1(h[n]);
  __syncthreads();
  for (int n = threadIdx.x; n < L; n += NT) {
    float acc = 0.0f;
    for (int k = 0; k < E; ++k) acc = fmaf(h[k], __ldg(m.enc[1].wt + (size_t)k * L + n), acc);
    z[(size_t)b * L + n] = acc + __ldg(m.enc[1].b + n);
  }
}

}  // namespace

void launch_encode(const ModelFp32& m, const float* obs, int B, int D, int E, int L, float* z, cudaStream_t st) {
  encode_kernel<<<B, NT, sizeof(float) * (D + E), st>>>(m, obs, D, E, L, z);
}
