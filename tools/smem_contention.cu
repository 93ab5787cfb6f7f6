// Microbenchmark: does the weight stream (bulk-TMA writes into shared memory) slow down tcgen05.mma?
// One CTA per SM on every SM.  Warp 1 issues back-to-back TS MMAs (M=128, N=128, K=16; A in TMEM, B = one
// fixed 16 KB tile in smem) -- 64 cycles each when alone.  Warp 0 optionally keeps `depth` bulk copies of 16 KB
// (global/L2 -> a separate smem ring) in flight, `tiles_per_4mma` tiles per 4 MMAs worth of time is not enforced:
// it simply streams as fast as completions allow, and the achieved bytes/clk is reported next to cycles/MMA.
#include <cstdio>
#include <cstdlib>
#include "../tdmpc_b200/csrc/tc_prims.cuh"
using namespace tc;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity, int mode) {
  if (mode == 0) mbar_wait(bar, parity);
  else while (!mbar_test_wait(bar, parity)) {}
}
__device__ __forceinline__ void arrive_tx_relaxed(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void arm(uint64_t* bar, uint32_t bytes, int mode) {
  if (mode == 2) arrive_tx_relaxed(bar, bytes); else mbar_arrive_expect_tx(bar, bytes);
}
__device__ __forceinline__ void umma_ts_bare(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc) {
  asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_ss_bare(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(const uint8_t* gsrc, int depth, int ss, int iters, int tile_bytes, int nprod, int wmode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar, full[8];
  __shared__ long long cnt[4];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  __shared__ long long copied;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1); stop = 0; copied = 0; cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint64_t bd = umma_desc_make(umma_desc_lo(smem_u32(smem + 16384)));
      const uint64_t ad = umma_desc_make(umma_desc_lo(smem_u32(smem)));
      const uint32_t a = tmem + 384;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) { if (ss) umma_ss_bare(tmem, ad, bd, idesc); else umma_ts_bare(tmem, a, bd, idesc); }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      stop = 1;
      out[blockIdx.x * 2] = t1 - t0;
    }
    __syncwarp();
  } else if (depth > 0) {
    const int pw = warp == 0 ? 0 : warp - 1;      // producer index 0..2
    if ((threadIdx.x & 31) == 0 && pw < nprod) {
      const int d = depth / nprod;                // slots owned by this producer
      uint8_t* ring = smem + 32768 + pw * d * 16384;
      uint64_t* fb = full + pw * d;
      uint32_t ph = 0; long long n = 0;
      for (int i = 0; i < d; ++i) { mbar_arrive_expect_tx(&fb[i], tile_bytes); bulk_g2s(ring + i * tile_bytes, gsrc + (size_t)((blockIdx.x * 7 + i) % 32) * 16384, tile_bytes, &fb[i]); }
      int s = 0;
      while (!stop) {
        wait_mode(&fb[s], ph, wmode == 1 ? 1 : 0);
        ++n;
        if (wmode == 3) {
          bulk_g2s(ring + s * tile_bytes, gsrc + (size_t)((blockIdx.x * 7 + n) % 32) * 16384, tile_bytes, &fb[s]);
          mbar_arrive_expect_tx(&fb[s], tile_bytes);
        } else {
          arm(&fb[s], tile_bytes, wmode);
          bulk_g2s(ring + s * tile_bytes, gsrc + (size_t)((blockIdx.x * 7 + n) % 32) * 16384, tile_bytes, &fb[s]);
        }
        if (++s == d) { s = 0; ph ^= 1; }
      }
      for (int i = 0; i < d; ++i) { mbar_wait(&fb[s], ph); if (++s == d) { s = 0; ph ^= 1; } }
      cnt[pw] = n;
    }
  }
  tcgen05_fence_before(); __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x * 2 + 1] = cnt[0] + cnt[1] + cnt[2];
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  int nsm = 148;
  long long* d; cudaMalloc(&d, sizeof(long long) * 2 * nsm);
  uint8_t* g; cudaMalloc(&g, 64 * 16384 + 32768); cudaMemset(g, 0x3c, 64 * 16384 + 32768);
  const int smem = 32768 + 8 * 16384;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2048;
  struct Cfg { int ss, tb, depth, nprod, wmode; };
  const Cfg cfgs[] = {{0, 16384, 4, 1, 0}, {0, 16384, 4, 1, 2}, {0, 16384, 4, 1, 3}, {0, 16384, 4, 2, 2}, {0, 16384, 4, 2, 3}};
  for (const Cfg& c : cfgs) {
    cudaMemset(d, 0, sizeof(long long) * 2 * nsm);
    k<<<nsm, 128, smem>>>(g, c.depth, c.ss, iters, c.tb, c.nprod, c.wmode, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2 * 148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = 0, tiles = 0;
    for (int i = 0; i < nsm; ++i) { cyc += h[2 * i]; tiles += h[2 * i + 1]; }
    cyc /= nsm; tiles /= nsm;
    printf("%s MMA N=128 | %s | stream: %d producer(s), %d slots x %5d B: %.1f cyc/MMA, %.1f B/clk/SM, %.0f cycles per copy %s\n", c.ss ? "SS" : "TS", c.wmode == 0 ? "arrive.release,copy" : c.wmode == 2 ? "arrive.relaxed,copy" : c.wmode == 3 ? "copy,arrive.release" : "poll", c.nprod, c.depth, c.tb,
           cyc / (8.0 * iters), tiles * c.tb / cyc, tiles > 0 ? cyc / tiles : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
