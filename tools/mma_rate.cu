// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) issued back-to-back by one thread,
// for N in {64,128,256}, SS vs TS operands, accumulating into ONE accumulator (dependent chain) or
// alternating between TWO accumulators.  Operands are whatever is in smem/TMEM (values irrelevant).
#include <cstdio>
#include <cstdlib>
#include "../tdmpc_b200/csrc/tc_prims.cuh"
using namespace tc;

// variant with loop-invariant operands and no predicate register: the barest possible issue sequence
__device__ __forceinline__ void umma_ts_bare(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc) {
  asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc) : "memory");
}
__global__ void __launch_bounds__(128, 1) bare_kernel(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, N);
      const uint64_t bd = umma_desc_make(umma_desc_lo(smem_u32(smem + 16384)));
      const uint32_t a = tmem + 384;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_ts_bare(tmem, a, bd, idesc);
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncwarp();
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int ts, int two_acc, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem)), b_lo = umma_desc_lo(smem_u32(smem + 16384));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t d = tmem + ((two_acc && (ks & 1)) ? 256u : 0u);
        if (ts) umma_ts(d, tmem + 384 + (ks & 3) * 8, umma_desc_make(b_lo + 2 * (ks & 3)), idesc, 1u);
        else umma_ss(d, umma_desc_make(a_lo + 2 * (ks & 3)), umma_desc_make(b_lo + 2 * (ks & 3)), idesc, 1u);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}


// Variant: the whole warp walks the issue loop converged and the election predicate sits INSIDE the asm statement, so the
// operands are computed in uniform control flow (ptxas can keep them in uniform registers: no R2UR per MMA).
__device__ __forceinline__ void umma_ts_pred(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t lead) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t}"
               ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(lead) : "memory");
}
__device__ __forceinline__ void umma_commit_pred(uint64_t* bar, uint32_t lead) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
               ::"r"(smem_u32(bar)), "r"(lead) : "memory");
}
__global__ void __launch_bounds__(128, 1) uni_kernel(int N, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[5];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 5; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t lead = elect_one() ? 1u : 0u;
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t b_lo = umma_desc_lo(smem_u32(smem + 16384));
    const long long t0 = clock64();
    if (mode == 0) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_ts_pred(tmem + ((ks & 4) ? 128u : 0u), tmem + 384 + (ks & 3) * 8, umma_desc_make(b_lo + 2 * (ks & 3)), idesc, lead);
      }
    } else {
      // as in the kernel: a commit to a ring barrier every 4 MMAs (no waits)
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          umma_ts_pred(tmem + ((ks & 4) ? 128u : 0u), tmem + 384 + (ks & 3) * 8, umma_desc_make(b_lo + 2 * (ks & 3)), idesc, lead);
          if ((ks & 3) == 3) umma_commit_pred(&bar[1 + (ks >> 2)], lead);
        }
      }
    }
    const long long t1 = clock64();
    umma_commit_pred(&bar[0], lead);
    mbar_wait(&bar[0], 0);
    const long long t2 = clock64();
    if (lead) { out[0] = t1 - t0; out[1] = t2 - t0; }
    __syncwarp();
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// Variant: NO election at all -- every lane of the converged warp executes the (warp-level, uniform-datapath) instruction, with
// operands that change every MMA like in the real kernel (32 distinct A addresses, 16 B descriptors, 2 accumulators).
__device__ __forceinline__ void umma_ts_all(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc) {
  asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(128, 1) vary_kernel(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[5];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 5; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  if (warp == 1) {
    const uint32_t lead = elect_one() ? 1u : 0u;
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t b_lo = umma_desc_lo(smem_u32(smem));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = (it & 1) * 128;
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t a = 256 + kb * 32 + ks * 8;
          const uint64_t bd = umma_desc_make(b_lo + (kb & 3) * 1024 + 2 * ks);
          if (MODE == 0) umma_ts_all(d, a, bd, idesc);
          else if (MODE == 1) umma_ts_pred(d, a, bd, idesc, lead);
          else { if (lead) umma_ts_bare(d, a, bd, idesc); }
        }
        if (MODE == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1 + (kb & 3)])) : "memory");
        else if (MODE == 1) umma_commit_pred(&bar[1 + (kb & 3)], lead);
        else { if (lead) umma_commit(&bar[1 + (kb & 3)]); }
      }
    }
    const long long t1 = clock64();
    umma_commit_pred(&bar[0], lead);
    mbar_wait(&bar[0], 0);
    const long long t2 = clock64();
    if (lead) { out[0] = t1 - t0; out[1] = t2 - t0; }
    __syncwarp();
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(0);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 256;
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {64, 128, 256})
      for (int two = 0; two < 2; ++two) {
        if (two && N == 256 && 256 + N > 384) { /* second accumulator at col 256 overlaps the TS A region only for N=256: skip */ if (ts) continue; }
        rate_kernel<<<1, 128, 64 * 1024>>>(N, ts, two, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s N=%3d %s: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d) %s\n", ts ? "TS" : "SS", N, two ? "two accumulators" : "one accumulator ",
               h[0] / (8.0 * iters), h[1] / (8.0 * iters), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  cudaFuncSetAttribute(bare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {32, 64, 128, 256}) {
    bare_kernel<<<1, 128, 64 * 1024>>>(N, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("bare TS N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d) %s\n", N, h[0] / (8.0 * iters), h[1] / (8.0 * iters), N / 2,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  cudaFuncSetAttribute(uni_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {64, 128}) {
      uni_kernel<<<1, 128, 64 * 1024>>>(N, mode, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("uniform-flow TS N=%3d %s: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d) %s\n", N, mode ? "with ring commits" : "plain", h[0] / (8.0 * iters),
             h[1] / (8.0 * iters), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  {
    auto run = [&](auto kern, const char* name) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      kern<<<1, 128, 64 * 1024>>>(128, 64, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("varying operands, %s: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal 64) %s\n", name, h[0] / (32.0 * 64), h[1] / (32.0 * 64),
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    run(vary_kernel<2>, "divergent if (lead)      ");
    run(vary_kernel<1>, "predicate inside the asm ");
    run(vary_kernel<0>, "whole warp, no election  ");
  }
  return 0;
}
