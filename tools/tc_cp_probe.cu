// Probe: does tcgen05.cp.128x256b with the SAME shared-memory descriptor an SS-mode A operand uses (K-major, 128-byte
// swizzle, one K=16 step = 128 rows x 32 bytes) land the data in tensor memory in the layout a TS-mode A operand
// expects (lane = row, 32-bit column = two consecutive bf16 of K, low half = lower k)?
#include <cstdio>
#include <cstdlib>
#include "../tdmpc_b200/csrc/tc_prims.cuh"
using namespace tc;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tile = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 64; i += 128) {
    const int row = i >> 6, k = i & 63;
    *reinterpret_cast<uint16_t*>(tile + sw128_offset(row, k)) = (uint16_t)(row * 64 + k);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && elect_one()) {
    const uint32_t lo = umma_desc_lo(smem_u32(tile));
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) tmem_cp_128x256b(tmem + 256 + ks * 8, umma_desc_make(lo + 2 * ks));
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  uint32_t v[32];
  tmem_ld32(tmem + (((uint32_t)(warp * 32)) << 16) + 256, v);
  tmem_wait_ld_dep(v);
  const int row = warp * 32 + lane;
  for (int c = 0; c < 32; ++c) out[row * 32 + c] = v[c];
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  uint32_t* d; cudaMalloc(&d, 128 * 32 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[128 * 32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int row = 0; row < 128; ++row)
    for (int c = 0; c < 32; ++c) {
      const uint32_t want = (uint32_t)(row * 64 + 2 * c) | ((uint32_t)(row * 64 + 2 * c + 1) << 16);
      if (h[row * 32 + c] != want) { if (bad < 12) printf("row %d col %d: got lo=%u hi=%u (row %u k %u / row %u k %u), want k %d,%d\n", row, c, h[row*32+c] & 0xffff, h[row*32+c] >> 16, (h[row*32+c] & 0xffff) / 64, (h[row*32+c] & 0xffff) % 64, (h[row*32+c] >> 16) / 64, (h[row*32+c] >> 16) % 64, 2*c, 2*c+1); ++bad; }
    }
  printf("tcgen05.cp.128x256b SW128 K-major -> TS A-operand layout: %s (%d mismatches of 4096)\n", bad ? "MISMATCH" : "OK", bad);
  return bad != 0;
}
