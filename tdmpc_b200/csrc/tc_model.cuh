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
  J_QL2 = 4,   // h1 (smem XS)  x W2       -> 512 cols;  epilogue: +b, LayerNorm, ELU, dot with w3 -> Q
  // DRNN (GRU latent dynamics, models/rnns.py:19-29): the gate pre-activations.  chunk = gate (0 reset, 1 update).
  J_G = 5,     // [z,a] x W_ih[g] + h (smem HS) x W_hh[g] -> one accumulator; epilogue: LayerNorm (in-thread), sigmoid,
               //   written back over the accumulator (r or u)
  J_GH = 6,    // gh_n = h x W_hh[newval]     -> X (TMEM columns 256-383 of the idle h1 region); no epilogue
  J_GI = 7     // gi_n = [z,a] x W_ih[newval] -> Y (columns 384-511); epilogue: n = tanh(LN(gi_n + r gh_n)),
               //   h' = u n + (1-u) h -> fp32 state + bf16 operand HS
};
constexpr int PAIR_MAX_TILES = 16;        // CTA-pair mode: launches of at most this many 128-trajectory tiles
enum TcJobFlags {
  F_WAIT_ZA = 1, F_LAST_STEP = 2, F_TERMINAL = 4,
  F_WAIT_HS = 8,     // MMA: first consumer of a freshly written hidden-state operand HS
  F_A_HS = 16,       // J_L1: A operand = HS (K = hidden_dim = 128, two k-blocks) instead of [z,a]
  F_A_TMEM = 32,     // J_L3: A operand = h1 in TMEM (TS) instead of XS (the DRNN prior MLP has two layers only)
  F_BN = 64,         // J_L1 epilogue: eval-mode BatchNorm folded to y = acc * scale + shift before the ELU
  F_PIDOT = 128,     // J_L2 of the policy chain: the last Linear (A <= 8 outputs) as fp32 dot products in the epilogue (DRNN
                     //   policy rollout: XS holds the GRU state, so there is no h2 operand for a layer-3 MMA)
  // Overlapped value-mode schedule (MLP dynamics): the reward chain keeps its hidden activation in shared memory (XS) so
  // that its 512x512 layer can run while the NEXT step's layer-1 epilogue fills the TMEM h1 buffer.
  F_TO_XS = 256,     // J_L1: epilogue writes the XS chunk (bf16, swizzled) instead of the TMEM h1 chunk
  F_A_XS = 512,      // J_L2: A operand = XS in shared memory (SS) instead of h1 in TMEM
  F_PARK = 2048,     // J_L3: MMAs only; the accumulator is parked in the (idle) h1 region instead of a ping-pong buffer
  F_EPI = 4096,      // J_L3: epilogue only (group 0), on the parked accumulator once `d3_full` has completed
  F_BUF0 = 8192,     // this job uses accumulator buffer 0 (the ping-pong continues from there)
  F_D3 = 16384,      // J_L1: after this job's MMAs the MMA warp also commits d3_full
  // Split layer 2: chunk 0 and chunk 1 of a 512 x 512 layer first accumulate over the k-blocks of h1 chunks 0-1 only
  // (F_KLO, no epilogue), which are written first, and later over the rest (F_KHI, epilogue as usual).  The tensor pipe
  // then has two half-jobs of work while the layer-1 epilogue is still producing h1 chunks 2-3.
  F_KLO = 32768,     // J_L2: k-blocks 0-3 only; acquires the accumulator, no commit, no epilogue
  F_KHI = 65536,     // J_L2: k-blocks 4-7 onto the accumulator its F_KLO half started; commit + epilogue
  F_TILE_END = 131072  // J_QL2, last job of a value-mode tile: epilogue group 0 writes the NEXT tile's [z0, a0] rows while this
                     //   job's MMAs run, and each group hands its accumulator buffer back right after reading its first chunk
};

// job word: type [0,3) | chain [3,6) | chunk [6,8) | t [8,13) | flags [13,31)
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
  int za_compact;              // 1: the last k-block holds ONE k-step (16 columns) and is stored unswizzled in 4 KB instead of 16 KB
  int za_bytes;                // shared-memory size of the [z,a] operand
  int n3p[CH_COUNT];           // L3 MMA N (multiple of 16) for dyn / pi
  int l3rows[CH_COUNT];        // rows per L3 weight k-block tile: 64 or 128
  const uint32_t* jobs[3][TDMPC_MAX_H + 1];   // [mode][H] device job lists; mode 2: value mode split over a CTA pair (rank 0 | rank 1)
  int njobs[3][TDMPC_MAX_H + 1];
  int pair_n0[TDMPC_MAX_H + 1];                // pair programs: jobs of rank 0 (rank 1's follow)
  int stages;                  // weight ring depth that fits in shared memory
  size_t smem_bytes;
  // host-side launch state, fixed when the weights are packed (environment switches are read THERE, not per launch)
  const uint32_t* jobs_host;   // host copy of every job program (owned by the handle: tc_release_host)
  int jobs_off[3][TDMPC_MAX_H + 1];
  int debug;                   // TDMPC_TC_DEBUG (experiments only)
  int producers;               // producer threads of the weight ring (divides `stages`)
  int exact_act;               // TDMPC_PREC_BF16_EXACT: two-MUFU tanh in the Q heads instead of tanh.approx
  unsigned int* tile_counter;  // device word of the dynamic tile scheduler (arena tail; zero between launches)
  // CTA-pair mode for launches of a few tiles (single-environment plans): mailbox between the dynamics CTA and the reward CTA
  float* pair_rows;            // [PAIR_MAX_TILES][TDMPC_MAX_H + 1][128][K1p] fp32 rows [z_t | a_t] published by rank 0; then [PAIR_MAX_TILES][128] Q1
  unsigned int* pair_flags;    // [PAIR_MAX_TILES][32]: slot t published (1) / consumed (0); [TDMPC_MAX_H + 1]: terminal action, [TDMPC_MAX_H + 2]: Q1
  int pair_mode;               // 0 off (TDMPC_TC_NOPAIR), 1 on
  int alt_sched;               // an alternative value-mode schedule is selected (TDMPC_TC_OVERLAP / TDMPC_TC_SPLIT): ALT instantiation
  // DRNN (TDMPC_DYN_DRNN): chain CH_DYN is the prior MLP (layer 1: K = hidden_dim, layer 3 slot = its second Linear),
  // chain CH_REW's layer 1 reads the hidden state; gate weights live in their own tiles.
  int drnn;
  uint32_t g_tile[2];          // J_G  gate g: nkb1 tiles of W_ih[g] followed by 2 tiles of W_hh[g]
  uint32_t gh_tile;            // J_GH: 2 tiles of W_hh[newval]
  uint32_t gi_tile;            // J_GI: nkb1 tiles of W_ih[newval] (+ gi_pad padding tiles that keep the ring aligned)
  int gi_pad;
  const float* gln_g[3]; const float* gln_b[3];    // LayerNorm of the reset / update / newval gates (eps 1e-3)
  const float* bn_scale; const float* bn_shift;    // prior_mlp.1 folded: y = acc * scale + (b * scale + shift)
  const float* pi_w3q;         // [8][512] fp32 copy of the bf16-rounded last policy Linear (rows >= A zero), F_PIDOT
};

const char* tc_supported(const tdmpc_model_cfg& c);     // nullptr when the shape is supported
size_t tc_arena_bytes(const tdmpc_model_cfg& c);
// returns number of kernels launched (>= 0) or -1
int tc_pack(const tdmpc_model_cfg& c, const tdmpc_weights& w, const ModelFp32& m32, void* arena, TcModel* out, cudaStream_t st);
void tc_release_host(TcModel* m);    // frees jobs_host
