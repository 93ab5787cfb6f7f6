// fp32 CUDA-core building blocks shared by the fp32 rollout kernel (rollout_fp32.cu) and the training-step forward
// (update_fwd.cu): dense layers over a tile of R rows resident in shared memory, the N = 1 heads, LayerNorm + activation.
// nn.Linear / ELU / LayerNorm semantics of algorithm/helper.py:169-176,197-201.
#pragma once
#include "common.cuh"

namespace {

constexpr int R = 32;          // trajectories per CTA
constexpr int NT = 256;        // threads per CTA
constexpr int NW = NT / 32;

enum Act { ACT_NONE = 0, ACT_ELU = 1 };

// out[r][n] = act( sum_k x[r][k] * Wt[k][n] + b[n] ),  N == 512: 8 rows x 8 cols per thread.
template <int ACT>
__device__ void layer512(const float* __restrict__ x, int ldx, const LinearT& lin, float* __restrict__ out, int ldo) {
  const int tid = threadIdx.x;
  const int c0 = (tid & 63) * 8;
  const int r0 = (tid >> 6) * 8;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  const float* wt = lin.wt + c0;
  for (int k = 0; k < lin.Kp; k += 4) {
    float4 xv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(x + (r0 + i) * ldx + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(k + kk) * 512));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(k + kk) * 512 + 4));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xs = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
        acc[i][0] = fmaf(xs, w0.x, acc[i][0]); acc[i][1] = fmaf(xs, w0.y, acc[i][1]);
        acc[i][2] = fmaf(xs, w0.z, acc[i][2]); acc[i][3] = fmaf(xs, w0.w, acc[i][3]);
        acc[i][4] = fmaf(xs, w1.x, acc[i][4]); acc[i][5] = fmaf(xs, w1.y, acc[i][5]);
        acc[i][6] = fmaf(xs, w1.z, acc[i][6]); acc[i][7] = fmaf(xs, w1.w, acc[i][7]);
      }
    }
  }
  float bv[8], ps[8], pt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bv[j] = lin.b ? __ldg(lin.b + c0 + j) : 0.0f;
    ps[j] = lin.post_scale ? __ldg(lin.post_scale + c0 + j) : 1.0f;
    pt[j] = lin.post_scale ? __ldg(lin.post_shift + c0 + j) : 0.0f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = acc[i][j] + bv[j];
      if (lin.post_scale) v[j] = __fadd_rn(__fmul_rn(v[j], ps[j]), pt[j]);   // eval-mode BatchNorm1d as an affine map
      if (ACT == ACT_ELU) v[j] = elu1(v[j]);
    }
    float* o = out + (r0 + i) * ldo + c0;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// Generic (any N, K) layer: one output per thread-iteration, lanes sweep n.
template <int ACT>
__device__ void layer_any(const float* __restrict__ x, int ldx, const LinearT& lin, float* __restrict__ out, int ldo) {
  const int N = lin.N;
  for (int idx = threadIdx.x; idx < R * N; idx += NT) {
    const int r = idx / N, n = idx - r * N;
    float acc = 0.0f;
    const float* xr = x + r * ldx;
    for (int k = 0; k < lin.K; ++k) acc = fmaf(xr[k], __ldg(lin.wt + (size_t)k * N + n), acc);
    acc += lin.b ? __ldg(lin.b + n) : 0.0f;
    if (lin.post_scale) acc = __fadd_rn(__fmul_rn(acc, __ldg(lin.post_scale + n)), __ldg(lin.post_shift + n));
    if (ACT == ACT_ELU) acc = elu1(acc);
    out[r * ldo + n] = acc;
  }
}

template <int ACT>
__device__ void layer(const float* x, int ldx, const LinearT& lin, float* out, int ldo) {
  if (lin.N == 512) layer512<ACT>(x, ldx, lin, out, ldo);
  else layer_any<ACT>(x, ldx, lin, out, ldo);
  __syncthreads();
}

// N == 1 head: out[r] = x[r][:] . w + b  (warp per row)
__device__ void head_dot(const float* __restrict__ x, int ldx, const LinearT& lin, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += NW) {
    float acc = 0.0f;
    for (int k = lane; k < lin.K; k += 32) acc = fmaf(x[r * ldx + k], __ldg(lin.wt + k), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[r] = acc + (lin.b ? __ldg(lin.b) : 0.0f);
  }
  __syncthreads();
}

// nn.LayerNorm over the last dim (biased variance, eps), then tanh or ELU, in place. Warp per row.
template <bool TANH>
__device__ void layernorm_act(float* __restrict__ x, int ldx, int n, const LayerNormP& ln, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += NW) {
    float* xr = x + r * ldx;
    float s = 0.0f;
    for (int k = lane; k < n; k += 32) s += xr[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s / (float)n;
    float v = 0.0f;
    for (int k = lane; k < n; k += 32) { const float d = xr[k] - mu; v = fmaf(d, d, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.0f / sqrtf(v / (float)n + eps);
    for (int k = lane; k < n; k += 32) {
      const float y = (xr[k] - mu) * rstd * __ldg(ln.g + k) + __ldg(ln.b + k);
      xr[k] = TANH ? tanhf(y) : elu1(y);
    }
  }
  __syncthreads();
}

}  // namespace
