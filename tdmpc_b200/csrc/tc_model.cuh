// Packed model + job program of the tcgen05 rollout kernel (internal).
#pragma once
#include "common.cuh"

// Weight sets ("chains"): each is one 3-layer MLP of TOLD (tdmpc.py:15-19).
enum TcChainId { CH_DYN = 0, CH_REW = 1, CH_PI = 2, CH_Q1 = 3, CH_Q2 = 4, CH_COUNT = 5 };

// Job = one accumulator's worth of MMAs followed by one epilogue.  The three warp roles
// (bulk-TMA producer, MMA issuer, epilogue) all walk the same job list, which is what keeps
// their mbarrier phases in lock-step.
enum TcJobType {
  J_L1 = 0,    // [z,a] (smem)  x W1 chunk -> 128 cols;  epilogue: +b, ELU  -> h1 chunk in TMEM (A operand)
  J_L2 = 1,    // h1 (tmem)     x W2 chunk -> 128 cols;  epilogue: +b, ELU  -> XS chunk (dyn, pi) | dot with w3 (reward)
  J_L3 = 2,    // h2 (smem XS)  x W3       -> n3 cols;   epilogue: z' -> ZA (dyn) | tanh+noise -> action -> ZA (pi)
  J_QL1 = 3,   // [z,a] (smem)  x W1       -> 512 cols;  epilogue: +b, LayerNorm, tanh -> XS
  J_QL2 = 4    // h1 (smem XS)  x W2       -> 512 cols;  epilogue: +b, LayerNorm, ELU, dot with w3 -> Q
};
enum TcJobFlags { F_WAIT_ZA = 1, F_LAST_STEP = 2, F_TERMINAL = 4 };

// job word: type [0,3) | chain [3,6) | chunk [6,8) | t [8,13) | flags [13,16)
__host__ __device__ inline uint32_t tc_job(int type, int chain, int chunk, int t, int flags) {
  return (uint32_t)type | ((uint32_t)chain << 3) | ((uint32_t)chunk << 6) | ((uint32_t)t << 8) | ((uint32_t)flags << 13);
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
  int stages;                  // weight ring depth that fits in shared memory
  size_t smem_bytes;
};

const char* tc_supported(const tdmpc_model_cfg& c);     // nullptr when the shape is supported
size_t tc_arena_bytes(const tdmpc_model_cfg& c);
// returns number of kernels launched (>= 0) or -1
int tc_pack(const tdmpc_model_cfg& c, const tdmpc_weights& w, const ModelFp32& m32, void* arena, TcModel* out, cudaStream_t st);
