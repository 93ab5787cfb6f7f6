// Which pipe bounds the activation epilogue of rollout_tc_kernel?  One CTA per SM, W warps per scheduler, each thread runs the
// per-element chain of the layer-1 epilogue (add bias, scale, ex2, subtract, select, bf16x2 pack) on register data, in several
// variants.  Prints cycles per warp-element for every (variant, warps per scheduler).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/epi_rate tools/epi_rate.cu && build/epi_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t pack_int(float lo, float hi) {      // RNE by integer arithmetic (finite inputs)
  uint32_t a = __float_as_uint(lo), b = __float_as_uint(hi);
  a += 0x7FFFu + ((a >> 16) & 1u); b += 0x7FFFu + ((b >> 16) & 1u);
  return __byte_perm(a, b, 0x7632);
}

// 2^t on the FMA / ALU pipes (no MUFU): round-to-nearest split t = j + f, |f| <= 0.5, degree-5 polynomial for 2^f, exponent added
// with an integer shift-add.  ~1.5e-7 relative.
__device__ __forceinline__ float ex2_poly(float t) {
  t = fmaxf(t, -125.0f);
  const float r = t + 12582912.0f;              // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = t - (r - 12582912.0f);
  float p = 1.3333558e-3f;
  p = fmaf(p, f, 9.6181291e-3f);
  p = fmaf(p, f, 5.5504109e-2f);
  p = fmaf(p, f, 2.4022651e-1f);
  p = fmaf(p, f, 6.9314718e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

template <int VARIANT>
__global__ void __launch_bounds__(1024, 1) k(const float* __restrict__ bias, uint32_t* out, long long* cyc, int iters) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = (float)((threadIdx.x * 37 + i * 11) % 64) * 0.05f - 1.6f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float y0 = v[2 * i] + __ldg(bias + 2 * i), y1 = v[2 * i + 1] + __ldg(bias + 2 * i + 1);
      float r0, r1;
      if (VARIANT >= 5) {                                              // ELU with every (VARIANT-3)-th pair on the polynomial
        const bool poly = (i % ((VARIANT - 3) / 2)) == 0;
        const float e0 = (poly ? ex2_poly(y0 * 1.4426950408889634f) : ex2a(y0 * 1.4426950408889634f)) - 1.0f;
        const float e1 = ((poly && (VARIANT & 1) == 0) ? ex2_poly(y1 * 1.4426950408889634f) : ex2a(y1 * 1.4426950408889634f)) - 1.0f;
        r0 = y0 > 0.0f ? y0 : e0; r1 = y1 > 0.0f ? y1 : e1;
      } else if (VARIANT == 0 || VARIANT == 3 || VARIANT == 4) {          // full ELU
        const float e0 = ex2a(y0 * 1.4426950408889634f) - 1.0f, e1 = ex2a(y1 * 1.4426950408889634f) - 1.0f;
        r0 = y0 > 0.0f ? y0 : e0; r1 = y1 > 0.0f ? y1 : e1;
      } else if (VARIANT == 1) {                                      // no MUFU: a multiply in its place
        const float e0 = y0 * 1.4426950408889634f * 0.7f - 1.0f, e1 = y1 * 1.4426950408889634f * 0.7f - 1.0f;
        r0 = y0 > 0.0f ? y0 : e0; r1 = y1 > 0.0f ? y1 : e1;
      } else {                                                        // MUFU only
        r0 = ex2a(y0); r1 = ex2a(y1);
      }
      if (VARIANT == 3) o[i] = __float_as_uint(r0) ^ __float_as_uint(r1);       // no conversion
      else if (VARIANT == 4) o[i] = pack_int(r0, r1);                             // integer RNE pack
      else o[i] = pack(r0, r1);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= o[i];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += __uint_as_float((acc & 1u) << 20);     // loop-carried dependence, keeps everything alive
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// The same ELU chain written in three STAGES over 32 elements, each stage in its own basic block behind an opaque branch (`one` is
// a kernel argument that is always 1): ptxas schedules inside basic blocks only, so this is the one way to keep it from
// re-interleaving the stages into the 2-3-element pipeline it prefers.
__global__ void __launch_bounds__(512, 1) k_staged(const float* __restrict__ bias, uint32_t* out, long long* cyc, int iters, int one) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = (float)((threadIdx.x * 37 + i * 11) % 64) * 0.05f - 1.6f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float y[32], t[32];
    uint32_t o[16];
#pragma unroll
    for (int i = 0; i < 32; ++i) { y[i] = v[i] + __ldg(bias + i); t[i] = y[i] * 1.4426950408889634f; }
    if (one) {
#pragma unroll
      for (int i = 0; i < 32; ++i) t[i] = ex2a(t[i]);
    }
    if (one > -1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float r0 = y[2 * i] > 0.0f ? y[2 * i] : t[2 * i] - 1.0f, r1 = y[2 * i + 1] > 0.0f ? y[2 * i + 1] : t[2 * i + 1] - 1.0f;
        o[i] = pack(r0, r1);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= o[i];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += __uint_as_float((acc & 1u) << 20);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  float* bias; uint32_t* out; long long* cyc;
  cudaMalloc(&bias, 256); cudaMemset(bias, 0, 256);
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  const char* names[10] = {"full ELU + cvt.bf16x2", "no MUFU (FMUL instead) + cvt", "MUFU only + cvt", "full ELU, no cvt", "full ELU + integer RNE pack",
                           "ELU, 1 of 2 elements polynomial", "?", "ELU, 1 of 4 polynomial", "?", "ELU, 1 of 6 polynomial"};
  const int vars[] = {0};
  for (int var : vars)
    for (int wps = 1; wps <= 8; wps *= 2) {
      const int threads = 128 * wps;
      for (int rep = 0; rep < 2; ++rep) {
        switch (var) {
          case 0: k<0><<<148, threads>>>(bias, out, cyc, iters); break;
          case 1: k<1><<<148, threads>>>(bias, out, cyc, iters); break;
          case 2: k<2><<<148, threads>>>(bias, out, cyc, iters); break;
          case 3: k<3><<<148, threads>>>(bias, out, cyc, iters); break;
          case 4: k<4><<<148, threads>>>(bias, out, cyc, iters); break;
          case 5: k<5><<<148, threads>>>(bias, out, cyc, iters); break;      // pair period 1, one of the pair: 1 of 2
          case 7: k<7><<<148, threads>>>(bias, out, cyc, iters); break;      // period 2, one of the pair: 1 of 4
          case 9: k<9><<<148, threads>>>(bias, out, cyc, iters); break;      // period 3: 1 of 6
          default: k<11><<<148, threads>>>(bias, out, cyc, iters); break;    // period 4: 1 of 8
        }
        cudaDeviceSynchronize();
      }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = (double)h[0] / iters / 32.0;      // cycles per element of ONE warp
      printf("%-34s warps/scheduler %d: %6.2f cycles per element per warp -> %6.2f cycles per warp-element per scheduler\n", var < 10 ? names[var] : "ELU, 1 of 8 polynomial", wps, c, c / wps);
    }
  for (int wps = 1; wps <= 4; wps *= 2) {
    for (int rep = 0; rep < 2; ++rep) { k_staged<<<148, 128 * wps>>>(bias, out, cyc, iters, 1); cudaDeviceSynchronize(); }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = (double)h[0] / iters / 32.0;
    printf("%-34s warps/scheduler %d: %6.2f cycles per element per warp -> %6.2f cycles per warp-element per scheduler\n", "full ELU + cvt, STAGED blocks", wps, c, c / wps);
  }
  return 0;
}
