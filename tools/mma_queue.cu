// Microbenchmark: how far can one thread run ahead of the tensor pipe?  Issues a burst of n tcgen05.mma (M=128, N=128, K=16, TS or SS)
// into an idle pipe and records the issue-side time of the burst and the time until the commit lands, for n = 1..64; then the cost of the
// other instructions of the issuing warp's per-job path: a successful mbarrier.try_wait, tcgen05.fence::after_thread_sync,
// tcgen05.commit (issue side) and the commit -> mbarrier latency of an empty pipe.
#include <cstdio>
#include <cstdlib>
#include "../tdmpc_b200/csrc/tc_prims.cuh"
using namespace tc;

__global__ void __launch_bounds__(128, 1) queue_kernel(int ts, int nmax, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t lead = elect_one() ? 1u : 0u;
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem)), b_lo = umma_desc_lo(smem_u32(smem + 16384));
    uint32_t ph = 0;
    for (int n = 1; n <= nmax; ++n) {
      const long long t0 = clock64();
      for (int i = 0; i < n; ++i) {
        if (ts) umma_ts_p(tmem + (i & 1) * 128, tmem + 384 + (i & 3) * 8, umma_desc_make(b_lo + 2 * (i & 3)), idesc, 1u, lead);
        else umma_ss_p(tmem + (i & 1) * 128, umma_desc_make(a_lo + 2 * (i & 3)), umma_desc_make(b_lo + 2 * (i & 3)), idesc, 1u, lead);
      }
      const long long t1 = clock64();
      umma_commit_p(&bar[0], lead);
      const long long t2 = clock64();
      mbar_wait(&bar[0], ph); ph ^= 1;
      const long long t3 = clock64();
      if (lead) { out[(n - 1) * 3] = t1 - t0; out[(n - 1) * 3 + 1] = t2 - t1; out[(n - 1) * 3 + 2] = t3 - t0; }
      __syncwarp();
    }
    // op costs, pipe idle
    long long acc[4] = {0, 0, 0, 0};
    uint32_t ph1 = 0;
    for (int r = 0; r < 64; ++r) {
      long long t0 = clock64();
      umma_commit_p(&bar[1], lead);                       // empty pipe: commit -> barrier latency
      long long t1 = clock64();
      mbar_wait(&bar[1], ph1); ph1 ^= 1;
      long long t2 = clock64();
      const bool ok = mbar_try_wait(&bar[1], ph1 ^ 1);    // a try_wait that succeeds at once (phase already complete)
      long long t3 = clock64();
      tcgen05_fence_after();
      long long t4 = clock64();
      acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2 + (ok ? 0 : 1000000); acc[3] += t4 - t3;
    }
    if (lead) for (int i = 0; i < 4; ++i) out[3 * nmax + i] = acc[i] / 64;
    __syncwarp();
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  const int nmax = 48;
  long long* d; cudaMalloc(&d, (3 * nmax + 4) * 8);
  cudaFuncSetAttribute(queue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  static long long h[3 * 48 + 4];
  for (int ts = 0; ts < 2; ++ts) {
    for (int rep = 0; rep < 2; ++rep) {                    // second run: warm instruction cache
      queue_kernel<<<1, 128, 64 * 1024>>>(ts, nmax, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s N=128: burst n -> issue cycles / commit issue / until complete\n", ts ? "TS" : "SS");
    for (int n = 1; n <= nmax; ++n)
      if (n <= 12 || n % 4 == 0) printf("  n=%2d issue %5lld commit %4lld complete %6lld\n", n, h[(n - 1) * 3], h[(n - 1) * 3 + 1], h[(n - 1) * 3 + 2]);
    printf("  op costs (cycles, incl. ~2 clock reads): commit issue %lld, commit->barrier wait %lld, successful try_wait %lld, fence::after %lld\n",
           h[3 * nmax], h[3 * nmax + 1], h[3 * nmax + 2], h[3 * nmax + 3]);
  }
  return 0;
}
