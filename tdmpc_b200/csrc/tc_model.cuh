// Packed model + job program of the tcgen05 rollout kernel (internal).
#pragma once
#include "common.cuh"

// Weight sets ("chains"): each is one 3-layer MLP of TOLD (tdmpc.py:15-19).
enum TcChainId { CH_DYN = 0, CH_REW = 1, CH_PI = 2, CH_Q1 = 3, CH_Q2 = 4, CH_COUNT = 5 };

// Job = one 128-column accumulator's worth of MMAs followed by one epilogue.  The three warp roles
// (bulk-TMA producer, MMA issuer, epilogue) all walk the same job list, which is what keeps their
// mbarrier phases in lock-step.  `type` says which MMAs are issued, `epi` what is done with the result.
enum TcJobType {
  J_L1 = 0,    // A = [z,a] tile (smem ZA)                      x W1 chunk (128 rows) -> 128 cols
  J_L2 = 1,    // A = h1 (TMEM, or smem XR with F_A_SMEM)       x W2 chunk (128 rows) -> 128 cols
  J_L3 = 2     // A = h2 (smem XS)                              x W3                   -> n3 cols
};
enum TcEpi {
  E_H1 = 0,    // +b, ELU -> bf16x2 -> TMEM h1 chunk (A operand of the TS layer-2 MMAs)
  E_XR = 1,    // +b, ELU -> bf16 -> smem XR chunk   (A operand of the SS layer-2 MMAs of the interleaved reward chain)
  E_XS = 2,    // +b, ELU -> bf16 -> smem XS chunk   (A operand of layer 3)
  E_DOT = 3,   // +b, ELU, fp32 dot with w3 -> reward (N = 1 head), G += gamma^t r
  E_Z = 4,     // z' = acc + b3 -> [z', a_{t+1}] row of ZA
  E_PI = 5,    // a = policy sample(tanh(acc + b3)) -> action columns of ZA / pi_actions
  E_STAT = 6,  // +b, accumulate LayerNorm statistics of the 512-wide row (Q heads); chunk 3: mu, rstd
  E_QN = 7,    // recomputed chunk: +b, LayerNorm, tanh -> bf16x2 -> TMEM h1 chunk   (Q head layer 1)
  E_QDOT = 8   // recomputed chunk: +b, LayerNorm, ELU, dot with w3 -> Q               (Q head layer 2)
};
enum TcJobFlags {
  F_WAIT_ZA = 1,     // MMA: wait until the [z,a] tile has been rewritten
  F_WAIT_A = 2,      // MMA: first consumer of a freshly produced A operand (h1 / XR / XS): wait for its chunks as they arrive
  F_LAST_STEP = 4,
  F_TERMINAL = 8,
  F_A_SMEM = 16      // J_L2: A operand in smem XR instead of TMEM
};

// job word: type [0,2) | epi [2,6) | chain [6,9) | chunk [9,11) | t [11,16) | flags [16,24)
__host__ __device__ inline uint32_t tc_job(int type, int epi, int chain, int chunk, int t, int flags) {
  return (uint32_t)type | ((uint32_t)epi << 2) | ((uint32_t)chain << 6) | ((uint32_t)chunk << 9) | ((uint32_t)t << 11) |
         ((uint32_t)flags << 16);
}

struct TcChain {
  uint32_t l1_tile, l2_tile, l3_tile;     // first 16 KB tile of each layer inside the arena
  const float* b1; const float* b2; const float* b3;   // fp32 biases
  const float* w3;                         // fp32 head weights [512] (reward / Q heads: N = 1)
  const float* ln_g[2]; const float* ln_b[2];
};

struct TcModel {
  const uint8_t* tiles;        // bf16 weight tiles, 16 KB each, 128B-swizzled K-major, in consumption order
  TcChain chain[CH_COUNT];
  int L, A;
  int K1p;                     // L+A rounded up to 16 (UMMA_K)
  int nkb1;                    // 64-wide k-blocks of the [z,a] operand
  int n3p[CH_COUNT];           // L3 MMA N (multiple of 16) for dyn / pi
  int l3rows[CH_COUNT];        // rows per L3 weight k-block tile: 64 or 128
  const uint32_t* jobs[2][TDMPC_MAX_H + 1];   // [mode][H] device job lists
  int njobs[2][TDMPC_MAX_H + 1];
  int use_xr;                  // 1: smem has room for XR -> reward chain interleaved with the dynamics chain
  int stages;                  // weight ring depth that fits in shared memory
  size_t smem_bytes;
};

const char* tc_supported(const tdmpc_model_cfg& c);     // nullptr when the shape is supported
size_t tc_arena_bytes(const tdmpc_model_cfg& c);
// returns number of kernels launched (>= 0) or -1
int tc_pack(const tdmpc_model_cfg& c, const tdmpc_weights& w, const ModelFp32& m32, void* arena, TcModel* out, cudaStream_t st);
