// Microbenchmark: can TWO warps take turns issuing tcgen05.mma bursts so that one warp's per-job overhead (barrier checks,
// descriptor set-up, the ~150-cycle start-up of the first MMA of a burst) hides behind the other warp's burst?
//   mode 0: one warp, J jobs of k MMAs each, `ovh` cycles of busy work between jobs
//   mode 1: two warps alternate jobs; a warp issues its job only after the other has finished ISSUING the previous one (token
//           = mbarrier with one arrival), does its busy work BEFORE waiting for the token
//   mode 2: two warps, no token (hardware arbitration), J/2 jobs each
// M=128, N=128, K=16 bf16; SS operands for k <= 8 (layer-1-like jobs), TS for larger k; two accumulators alternate.
#include <cstdio>
#include <cstdlib>
#include "../tdmpc_b200/csrc/tc_prims.cuh"
using namespace tc;

__device__ __forceinline__ void spin(int cycles) { const long long t0 = clock64(); while (clock64() - t0 < cycles) { } }

__global__ void __launch_bounds__(128, 1) dual_kernel(int mode, int k, int J, int ovh, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[4];          // 0, 1: job done (per issuer); 2, 3: tokens
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = slot;
  const int ni = mode == 0 ? 1 : 2;
  if (warp >= 1 && warp <= ni) {
    const int id = warp - 1;
    const uint32_t lead = elect_one() ? 1u : 0u;
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem)), b_lo = umma_desc_lo(smem_u32(smem + 16384));
    const bool ts = k > 8;
    uint32_t tok_phase = 0, done_phase = 0;
    const long long t0 = clock64();
    for (int j = id; j < J; j += ni) {
      spin(ovh);                                             // this job's overhead (checks, set-up)
      if (mode == 1 && j > 0) { mbar_wait(&bar[2 + id], tok_phase); tok_phase ^= 1; }
      const uint32_t d = tmem + (j & 1) * 128;
      for (int i = 0; i < k; ++i) {
        if (ts) umma_ts_p(d, tmem + 384 + (i & 3) * 8, umma_desc_make(b_lo + 2 * (i & 3)), idesc, i ? 1u : 0u, lead);
        else umma_ss_p(d, umma_desc_make(a_lo + 2 * (i & 3)), umma_desc_make(b_lo + 2 * (i & 3)), idesc, i ? 1u : 0u, lead);
      }
      umma_commit_p(&bar[id], lead);
      if (mode == 1 && lead) mbar_arrive(&bar[2 + (id ^ 1)]);
      __syncwarp();
      // as in the kernel, a job's accumulator is reused two jobs later: wait for my previous commit before going on
      if (j >= ni * 1) { }
      mbar_wait(&bar[id], done_phase); done_phase ^= 1;      // (keeps at most one job of this issuer in flight)
    }
    const long long t1 = clock64();
    if (lead) out[id] = t1 - t0;
    __syncwarp();
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int J = 256;
  for (int k : {4, 32})
    for (int ovh : {0, 300, 600})
      for (int mode = 0; mode < 3; ++mode) {
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 2; ++rep) {
          cudaMemset(d, 0, 16);
          dual_kernel<<<1, 128, 64 * 1024>>>(mode, k, J, ovh, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        const long long t = h[0] > h[1] ? h[0] : h[1];
        printf("k=%2d overhead=%3d mode=%d (%s): %6.1f cycles per job (ideal pipe %d)\n", k, ovh, mode,
               mode == 0 ? "one issuer" : mode == 1 ? "two issuers, token" : "two issuers, free", (double)t / J, k * 64);
      }
  return 0;
}
