// TOLD.h (algorithm/tdmpc.py:30-32): state encoder Linear(D,E) -> [LayerNorm(E)] -> ELU -> Linear(E,L)
// (helper.py:131-133 / helper.py:161-165).  One CTA per observation; fp32 in every precision mode
// (2*(D*E+E*L) FLOPs per problem -- negligible next to the rollout).
#include "common.cuh"

namespace {

constexpr int NT = 1024;

// y[n] = sum_k x[k] * wt[k][n] + bias[n] for one observation.  A planning call encodes ONE row per environment, so with one
// thread per output every dot product is a chain of K dependent L2 round trips (E = 256: 25 us for a 26 kFLOP layer).  The K
// range is cut into S = min(NT / N, K) slices, one thread per (slice, output) with eight loads in flight; the slice sums are
// added in slice order.  S depends on the model dims only, never on the batch: a batched call equals one-at-a-time calls bit for bit.
__device__ void dense_split(const float* x, int K, const float* __restrict__ wt, const float* __restrict__ bias, int N, float* part, float* out) {
  int S = N <= NT ? NT / N : 1;
  S = S > K ? K : S;
  const int chunk = (K + S - 1) / S;
  for (int idx = threadIdx.x; idx < S * N; idx += NT) {
    const int s = idx / N, n = idx - s * N;
    const int k1 = min(K, (s + 1) * chunk);
    float acc = 0.0f;
    int k = s * chunk;
    for (; k + 8 <= k1; k += 8) {
      float wv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) wv[i] = __ldg(wt + (size_t)(k + i) * N + n);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(x[k + i], wv[i], acc);
    }
    for (; k < k1; ++k) acc = fmaf(x[k], __ldg(wt + (size_t)k * N + n), acc);
    part[idx] = acc;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += NT) {
    float acc = part[n];
    for (int s = 1; s < S; ++s) acc += part[s * N + n];
    out[n] = acc + __ldg(bias + n);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT) encode_kernel(ModelFp32 m, const float* __restrict__ obs, int D, int E, int L,
                                                    float* __restrict__ z) {
  extern __shared__ float sm[];
  float* x = sm;          // [D]
  float* h = sm + D;      // [E]
  float* part = h + E;    // [max(NT, E, L)] slice sums
  __shared__ float red[2];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < D; i += NT) x[i] = obs[(size_t)b * D + i];
  __syncthreads();
  dense_split(x, D, m.enc[0].wt, m.enc[0].b, E, part, h);
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
  dense_split(h, E, m.enc[1].wt, m.enc[1].b, L, part, z + (size_t)b * L);
}

}  // namespace

void launch_encode(const ModelFp32& m, const float* obs, int B, int D, int E, int L, float* z, cudaStream_t st) {
  const int np = NT > E ? (NT > L ? NT : L) : (E > L ? E : L);
  encode_kernel<<<B, NT, sizeof(float) * (D + E + np), st>>>(m, obs, D, E, L, z);
}
