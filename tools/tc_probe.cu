// Hardware probe for the tcgen05 building blocks used by tdmpc_b200/csrc/rollout_tc.cu.
// Runs one 128 x N x 128 GEMM per case on ONE CTA and checks it against a host fp32 reference:
//   mode 0: SS  -- A and B in shared memory (128B-swizzled K-major tiles staged by bulk TMA)
//   mode 1: TS  -- A in tensor memory (bf16x2 packed, written with tcgen05.st), B in shared memory
// Usage: tc_probe [variant]   variant bit0: clear the descriptor "version" field
//                              variant bit1: swap the bf16 pair order in TMEM (TS mode)
// Exit code 0 = every case within tolerance.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../tdmpc_b200/csrc/tc_prims.cuh"

using namespace tc;

constexpr int M = 128, K = 128, KB = K / 64;

struct ProbeArgs {
  const uint8_t* a_tiles;   // KB tiles of [128 x 64] bf16, swizzled (16 KB each)
  const uint8_t* b_tiles;   // KB tiles of [tile_rows x 64] bf16, swizzled
  const float* a_plain;     // [128][K] fp32 (already bf16-representable) for TS mode
  float* d;                 // [128][N]
  int N, tile_rows, mode, variant;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                          // KB * 16 KB
  uint8_t* sB = smem + KB * 16384;             // KB * tile_rows * 128
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b_tile_bytes = (uint32_t)p.tile_rows * 128;

  if (threadIdx.x == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, KB * 16384 + KB * b_tile_bytes);
    for (int kb = 0; kb < KB; ++kb) {
      bulk_g2s(sA + kb * 16384, p.a_tiles + (size_t)kb * 16384, 16384, &bar_load);
      bulk_g2s(sB + kb * b_tile_bytes, p.b_tiles + (size_t)kb * b_tile_bytes, b_tile_bytes, &bar_load);
    }
  }
  mbar_wait(&bar_load, 0);

  const uint32_t TM_A = 256;   // TMEM columns of the A operand in TS mode (K/2 = 64 columns)
  if (p.mode == 1) {
    const int row = threadIdx.x;
    for (int g = 0; g < K / 32; ++g) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float lo = p.a_plain[row * K + g * 32 + 2 * i], hi = p.a_plain[row * K + g * 32 + 2 * i + 1];
        v[i] = (p.variant & 2) ? pack_bf16x2(hi, lo) : pack_bf16x2(lo, hi);
      }
      tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + TM_A + g * 16, v);
    }
    tmem_wait_st();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(M, p.N);
    const uint64_t vmask = (p.variant & 1) ? ~((uint64_t)3 << 46) : ~(uint64_t)0;
    for (int kb = 0; kb < KB; ++kb)
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + kb * b_tile_bytes) + ks * 32) & vmask;
        const uint32_t acc = (kb | ks) ? 1u : 0u;
        if (p.mode == 0) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + kb * 16384) + ks * 32) & vmask;
          umma_ss(tmem, adesc, bdesc, idesc, acc);
        } else {
          umma_ts(tmem, tmem + TM_A + kb * 32 + ks * 8, bdesc, idesc, acc);
        }
      }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tcgen05_fence_after();

  for (int c = 0; c < p.N; c += 32) {
    uint32_t v[32];
    if (p.N - c >= 32) {
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_wait_ld();
      for (int i = 0; i < 32; ++i) p.d[(warp * 32 + lane) * p.N + c + i] = __uint_as_float(v[i]);
    } else {
      uint32_t w[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, w);
      tmem_wait_ld();
      for (int i = 0; i < 16 && c + i < p.N; ++i) p.d[(warp * 32 + lane) * p.N + c + i] = __uint_as_float(w[i]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

static void swizzle_tiles(const std::vector<float>& w, int rows, int tile_rows, std::vector<uint8_t>& out) {
  out.assign((size_t)KB * tile_rows * 128, 0);
  for (int kb = 0; kb < KB; ++kb)
    for (int r = 0; r < rows; ++r)
      for (int kk = 0; kk < 64; ++kk) {
        __nv_bfloat16 b = __float2bfloat16(w[(size_t)r * K + kb * 64 + kk]);
        memcpy(&out[(size_t)kb * tile_rows * 128 + sw128_offset(r, kk)], &b, 2);
      }
}

static int run_case(int N, int tile_rows, int mode, int variant) {
  std::vector<float> A((size_t)M * K), B((size_t)N * K), D((size_t)M * N, 0.f);
  srand(1234 + N + mode);
  for (auto& x : A) x = bf16_round((rand() / (float)RAND_MAX - 0.5f) * 2.0f);
  for (auto& x : B) x = bf16_round((rand() / (float)RAND_MAX - 0.5f) * 2.0f);
  std::vector<uint8_t> at, bt;
  swizzle_tiles(A, M, 128, at);
  swizzle_tiles(B, N, tile_rows, bt);
  uint8_t *da, *db; float *dp, *dd;
  cudaMalloc(&da, at.size()); cudaMalloc(&db, bt.size()); cudaMalloc(&dp, A.size() * 4); cudaMalloc(&dd, D.size() * 4);
  cudaMemcpy(da, at.data(), at.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db, bt.data(), bt.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dp, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0, D.size() * 4);
  ProbeArgs p{da, db, dp, dd, N, tile_rows, mode, variant};
  const size_t smem = (size_t)KB * 16384 + (size_t)KB * tile_rows * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("case N=%d mode=%d: CUDA error %s\n", N, mode, cudaGetErrorString(e)); return 2; }
  cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      maxerr = fmax(maxerr, fabs(ref - D[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  const bool ok = maxerr <= 1e-3 * maxref;
  printf("case N=%3d tile_rows=%3d mode=%s variant=%d: max|err|=%.3e (max|ref|=%.2f) %s\n", N, tile_rows,
         mode ? "TS" : "SS", variant, maxerr, maxref, ok ? "OK" : "MISMATCH");
  cudaFree(da); cudaFree(db); cudaFree(dp); cudaFree(dd);
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  int bad = 0;
  bad |= run_case(128, 128, 0, variant);
  bad |= run_case(64, 64, 0, variant);
  bad |= run_case(48, 64, 0, variant);
  bad |= run_case(16, 64, 0, variant);
  bad |= run_case(128, 128, 1, variant);
  bad |= run_case(256, 256, 1, variant);
  printf("tc_probe variant %d: %s\n", variant, bad ? "FAILED" : "ALL OK");
  return bad;
}
