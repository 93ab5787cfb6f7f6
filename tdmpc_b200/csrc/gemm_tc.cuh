// fp32-in / fp32-out GEMM on the 5th-generation tensor cores for the training step (train.cu): C = op(A) op(B) (+ beta C) (+ bias) (ELU).
//
// The reference's update() is fp32 (algorithm/tdmpc.py:191-245) and the parity gate of the training step is 2e-5 on the losses and
// 5e-5 on the gradients, so the operands cannot simply be rounded to bf16.  Each fp32 operand is split in the staging pass into
// x = hi + lo (+ r), hi = bf16(x), lo = bf16(x - hi), |r| <= 2^-17 |x|, and the product is formed as three bf16 MMAs into the same
// fp32 TMEM accumulator:  A B  ~=  Ahi Bhi + Ahi Blo + Alo Bhi  (the dropped Alo Blo term is 2^-18 relative).  Per-product error
// 2^-16, i.e. the same order as fp32 accumulation-order noise over k = 512; three MMAs of a 128 x BN x 64 block are 3 x 4 x (BN / 4)
// tensor-pipe cycles, still far below what the operand traffic costs, so the kernel is L2-bandwidth / latency bound, not MMA bound.
//
// One CTA = one 128 x BN output tile (UMMA M = 128, N = BN): eight producer / epilogue warps and one MMA-issuing warp, two
// shared-memory stages of one 64-wide k-block:
//   * the producers load fp32 elements from global memory into registers two k-blocks ahead (coalesced along whichever index is
//     contiguous: TA / TB only change the thread -> element map), split them and write hi / lo as K-major rows of 64 bf16 with the
//     128-byte swizzle (the layout the rollout kernel's operands use, tc_prims.cuh) -- the transposition of a [K, M] or [K, N]
//     operand happens here, for free -- and arrive on the stage's `full` mbarrier;
//   * one lane of the ninth warp waits for `full`, issues the 12 MMAs of the block and commits them to the stage's `free` mbarrier
//     (the stage is rewritten two blocks later);
//   * epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31 (thread = row), columns (w / 4) BN / 2 .. + BN / 2, into a padded fp32 tile in
//     the idle stages; then every warp writes whole rows of C (128-byte segments), applying bias / beta / ELU in the SIMT kernel's order.
// Split-k (blockIdx.z, k slices that are multiples of 64) writes partials like sgemm_kernel; train.cu's fixed-order reduction adds them.
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"

namespace gemm_tc {

constexpr int BM = 128, BK = 64, THREADS = 256;

template <int BN> struct Smem {
  static constexpr int A_BYTES = BM * BK * 2;            // one [128 x 64] bf16 block: 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;   // Ahi, Alo, Bhi, Blo
  static constexpr int TOTAL = 2 * STAGE + 1024 /* alignment slack */ + 128 /* barriers, tmem slot */;
};

// 8 consecutive k of one operand row -> 16 B of hi and 16 B of lo.  Both roundings are the packed cvt.rn.bf16x2.f32 (F2FP, the ALU
// pipe); a scalar __float2bfloat16_rn is an F2F on the quarter-rate XU pipe, 64 of them per thread and k-block.
__device__ __forceinline__ void split_store(const float (&x)[8], uint32_t hi_addr, uint32_t lo_addr) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = x[2 * j], b = x[2 * j + 1];
    h[j] = tc::pack_bf16x2(a, b);                                    // low half: bf16(a), high half: bf16(b)
    l[j] = tc::pack_bf16x2(a - __uint_as_float(h[j] << 16), b - __uint_as_float(h[j] & 0xffff0000u));
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hi_addr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}

// Operand tile loader: ROWS rows (m or n index) x 64 k.  KC: k is the contiguous index in memory (P[row * ld + k]), else the row
// index is (P[k * ld + row]).  TASKS = ROWS * 8 / 256 chunks of 8 k per thread.
// Every load is UNCONDITIONAL from a clamped address and the out-of-range elements are zeroed when the registers are written to
// shared memory: a predicated `ok ? __ldg(p) : 0` compiles to a load plus a move of its result, and the move makes each of a
// thread's loads wait for the one before it (measured in the rollout kernel's action loads: DESIGN 4.1, third session).  Blocks
// with all 64 k in range (CTA-uniform test) take address arithmetic without the k clamp, tiles with all rows in range skip the masks.
template <int ROWS, bool KC> struct Loader {
  static constexpr int TASKS = ROWS * 8 / THREADS;
  float v[TASKS][8];
  int klen;          // valid k of the block held in v (CTA-uniform)
  __device__ __forceinline__ static void map(int i, int& row, int& chunk) {
    const int id = (int)threadIdx.x + i * THREADS;
    if (KC) { chunk = id & 7; row = id >> 3; } else { row = id % ROWS; chunk = id / ROWS; }
  }
  __device__ __forceinline__ void load(const float* __restrict__ P, int ld, int row0, int nrows, int k0, int kend, bool vec_ok) {
    klen = kend - k0;
    const bool full = klen >= BK;                                // CTA-uniform: no divergence, no predicated loads
#pragma unroll
    for (int i = 0; i < TASKS; ++i) {
      int row, chunk;
      map(i, row, chunk);
      const int gr = min(row0 + row, nrows - 1), gk = k0 + chunk * 8;
      if (KC) {
        const float* p = P + (size_t)gr * ld + gk;
        if (full && vec_ok) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
          v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w; v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
        } else if (full) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] = __ldg(p + j);
        } else {
          const int last = kend - 1 - gk;                          // may be negative: the whole chunk is masked, any valid address does
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] = __ldg(p + min(j, last));
        }
      } else {
        const float* p = P + (size_t)gk * ld + gr;
        if (full) {
          const float* q = p;                                      // running pointer: one 64-bit add per element instead of a multiply + LEA pair
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[i][j] = __ldg(q); q += ld; }
        } else {
          const int last = kend - 1 - gk;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] = __ldg(p + (ptrdiff_t)min(j, last) * ld);
        }
      }
    }
  }
  // SUM: also adds the (masked) elements this thread stages to `acc` -- for a [K, M] operand every task of a thread has the same row
  // m = tid % ROWS, so acc is a partial row sum over k (the bias gradient rides along with the weight-gradient GEMM)
  template <bool SUM>
  __device__ __forceinline__ void store(uint32_t hi_base, uint32_t lo_base, int row0, int nrows, float& acc) const {
    const bool interior = klen >= BK && row0 + ROWS <= nrows;      // CTA-uniform
#pragma unroll
    for (int i = 0; i < TASKS; ++i) {
      int row, chunk;
      map(i, row, chunk);
      const uint32_t off = tc::sw128_offset(row, chunk * 8);
      if (interior) {
        if (SUM) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc += v[i][j];
        }
        split_store(v[i], hi_base + off, lo_base + off);
      } else {
        const bool rv = row0 + row < nrows;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (rv && chunk * 8 + j < klen) ? v[i][j] : 0.0f;
        if (SUM) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc += x[j];
        }
        split_store(x, hi_base + off, lo_base + off);
      }
    }
  }
};

template <bool TA, bool TB, int BN>
__global__ void __launch_bounds__(THREADS + 32, 1)
gemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc, float beta,
            const float* __restrict__ bias, int elu, int kslice, long long zstride, float* __restrict__ bsum) {
  using S = Smem<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;          // swizzled blocks need 1024-byte alignment
  uint8_t* aligned = smem_raw + (base - tc::smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(aligned + 2 * S::STAGE);      // [0], [1]: stage full; [2], [3]: stage free; [4]: accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  int kbeg = 0;
  if (kslice > 0) { kbeg = blockIdx.z * kslice; K = min(K, kbeg + kslice); C += (size_t)blockIdx.z * zstride; if (bsum) bsum += (size_t)blockIdx.z * zstride; }
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the role branch a uniform branch and the global
  // loads' memory descriptor in a uniform register (without it: two R2UR moves in front of every LDG of the producer loop)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr int MMA_WARP = THREADS / 32;
  const int nkb = (K - kbeg + BK - 1) / BK;

  if (tid == 0) {
    tc::mbar_init(&bars[0], THREADS); tc::mbar_init(&bars[1], THREADS);
    tc::mbar_init(&bars[2], 1); tc::mbar_init(&bars[3], 1); tc::mbar_init(&bars[4], 1);
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == MMA_WARP) {
    // ---- MMA issuer: one lane; the producers never wait for the issue, only for the commit that frees a stage
    if (lane == 0) {
      constexpr uint32_t IDESC = tc::umma_idesc_bf16(BM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        const uint32_t st = base + s * S::STAGE;
        const uint32_t ah = tc::umma_desc_lo(st), al = tc::umma_desc_lo(st + S::A_BYTES), bh = tc::umma_desc_lo(st + 2 * S::A_BYTES),
                       bl = tc::umma_desc_lo(st + 2 * S::A_BYTES + S::B_BYTES);
        tc::mbar_wait(&bars[s], (kb >> 1) & 1);
        tc::tcgen05_fence_after();
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {                                // + 2 = 32 bytes = 16 bf16 along k inside the swizzled row
          tc::umma_ss(tmem, tc::umma_desc_make(ah + 2 * ks), tc::umma_desc_make(bh + 2 * ks), IDESC, (kb | ks) != 0);
          tc::umma_ss(tmem, tc::umma_desc_make(ah + 2 * ks), tc::umma_desc_make(bl + 2 * ks), IDESC, 1);
          tc::umma_ss(tmem, tc::umma_desc_make(al + 2 * ks), tc::umma_desc_make(bh + 2 * ks), IDESC, 1);
        }
        tc::umma_commit(&bars[2 + s]);
      }
      tc::umma_commit(&bars[4]);
    }
    __syncwarp();
  } else {
    // ---- producers: registers hold the blocks kb + 1 and kb + 2 while block kb is split and written (two loaders, loop unrolled by 2)
    const bool a_vec = !TA && (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (kbeg % 4 == 0);   // float4 along k
    const bool b_vec = TB && (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (kbeg % 4 == 0);
    Loader<BM, !TA> la0, la1;
    Loader<BN, TB> lb0, lb1;
    // bias gradient (TA products only): the first column tile's CTAs also sum their A tile over k (CTA-uniform switch)
    const bool do_sum = TA && bsum != nullptr && blockIdx.x == 0;
    float rsum = 0.0f;
    la0.load(A, lda, m0, M, kbeg, K, a_vec);
    lb0.load(B, ldb, n0, N, kbeg, K, b_vec);
    if (nkb > 1) {
      la1.load(A, lda, m0, M, kbeg + BK, K, a_vec);
      lb1.load(B, ldb, n0, N, kbeg + BK, K, b_vec);
    }
    auto step = [&](int kb, Loader<BM, !TA>& la, Loader<BN, TB>& lb) {
      const int s = kb & 1;
      const uint32_t st = base + s * S::STAGE;
      if (kb >= 2) tc::mbar_wait(&bars[2 + s], ((kb >> 1) - 1) & 1);          // the MMAs of block kb - 2 have read this stage
      if (TA && do_sum) la.template store<true>(st, st + S::A_BYTES, m0, M, rsum);
      else la.template store<false>(st, st + S::A_BYTES, m0, M, rsum);
      lb.template store<false>(st + 2 * S::A_BYTES, st + 2 * S::A_BYTES + S::B_BYTES, n0, N, rsum);
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&bars[s]);
      if (kb + 2 < nkb) {
        const int k0 = kbeg + (kb + 2) * BK;
        la.load(A, lda, m0, M, k0, K, a_vec);
        lb.load(B, ldb, n0, N, k0, K, b_vec);
      }
    };
#pragma unroll 1
    for (int kb = 0; kb < nkb; kb += 2) {
      step(kb, la0, lb0);
      if (kb + 1 < nkb) step(kb + 1, la1, lb1);
    }

    // ---- epilogue: TMEM (thread = row) -> padded fp32 tile in the (now idle) stages -> coalesced rows of C
    tc::mbar_wait(&bars[4], 0);
    tc::tcgen05_fence_after();
    constexpr int LDT = BN + 1;
    const uint32_t xch = base + 2 * S::STAGE - 1024;                          // 256 floats behind the staging tile
    if (TA && do_sum) tc::sts1(xch + tid * 4u, rsum);
    {
      constexpr int HALF = BN / 2;                                            // columns per warp: 32 (BN = 64) or 64 (BN = 128)
      const int q = warp & 3, half = warp >> 2;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HALF);
      const uint32_t trow = base + (uint32_t)((q * 32 + lane) * LDT + half * HALF) * 4u;
#pragma unroll 1
      for (int g = 0; g < HALF / 32; ++g) {
        uint32_t v[32];
        tc::tmem_ld32(taddr + g * 32, v);
        tc::tmem_wait_ld_dep(v);
#pragma unroll
        for (int j = 0; j < 32; ++j) tc::sts1(trow + (uint32_t)(g * 32 + j) * 4u, __uint_as_float(v[j]));
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (TA && do_sum && tid < BM && m0 + tid < M)                             // threads tid and tid + 128 staged the two halves of row tid's chunks
      bsum[m0 + tid] = tc::lds1(xch + tid * 4u) + tc::lds1(xch + (tid + BM) * 4u);
    // four rows x BN / 32 column groups per pass: 8 - 16 independent load -> bias -> ELU -> store chains per thread (one row at a time
    // left the two warps per scheduler waiting on each chain's latency: 15 us of a 16 us one-k-block launch were this loop)
    constexpr int CG = BN / 32;
    float bv[CG];
#pragma unroll
    for (int cc = 0; cc < CG; ++cc) bv[cc] = bias ? __ldg(bias + min(n0 + cc * 32 + lane, N - 1)) : 0.0f;
#pragma unroll 1
    for (int r0 = warp; r0 < BM; r0 += 4 * (THREADS / 32)) {
      float x[4][CG];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int cc = 0; cc < CG; ++cc) x[u][cc] = tc::lds1(base + (uint32_t)((r0 + u * (THREADS / 32)) * LDT + cc * 32 + lane) * 4u) + bv[cc];
      if (beta != 0.0f) {                                                     // CTA-uniform
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int cc = 0; cc < CG; ++cc) {
            const int gm = m0 + r0 + u * (THREADS / 32), gn = n0 + cc * 32 + lane;
            if (gm < M && gn < N) x[u][cc] += beta * C[(size_t)gm * ldc + gn];
          }
      }
      if (elu) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int cc = 0; cc < CG; ++cc) x[u][cc] = elu1(x[u][cc]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int cc = 0; cc < CG; ++cc) {
          const int gm = m0 + r0 + u * (THREADS / 32), gn = n0 + cc * 32 + lane;
          if (gm < M && gn < N) C[(size_t)gm * ldc + gn] = x[u][cc];
        }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tc::tmem_dealloc<TMEM_COLS>(tmem);
}

// true: launched.  false: shape not covered (caller falls back to the SIMT kernel).
template <int BN>
inline bool launch_bn(bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, float beta,
                      const float* bias, bool elu, int kslice, int nz, cudaStream_t st, long long zstride = 0, float* bsum = nullptr) {
  using S = Smem<BN>;
  if (zstride == 0) zstride = (long long)M * ldc;
  static bool configured[16][4] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  auto conf = [&](auto kern, int slot) {
    if (dev < 16 && configured[dev][slot]) return true;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess) return false;
    if (dev < 16) configured[dev][slot] = true;
    return true;
  };
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, nz);
#define TDMPC_GEMM_TC_LAUNCH(TA_, TB_, SLOT_)                                                                                    \
  {                                                                                                                               \
    auto kern = gemm_kernel<TA_, TB_, BN>;                                                                                       \
    if (!conf(kern, SLOT_)) return false;                                                                                        \
    kern<<<grid, THREADS + 32, S::TOTAL, st>>>(M, N, K, A, lda, B, ldb, C, ldc, beta, bias, elu ? 1 : 0, kslice, zstride, bsum);                     \
    return true;                                                                                                                  \
  }
  if (!ta && tb) TDMPC_GEMM_TC_LAUNCH(false, true, 0)
  if (!ta && !tb) TDMPC_GEMM_TC_LAUNCH(false, false, 1)
  if (ta && !tb) TDMPC_GEMM_TC_LAUNCH(true, false, 2)
#undef TDMPC_GEMM_TC_LAUNCH
  return false;
}

}  // namespace gemm_tc
