// tcgen05 trajectory rollout: the bf16-operand / fp32-accumulate throughput mode of the planner.
//
// One persistent CTA per SM owns tiles of 128 trajectories (TMEM lane = trajectory) and walks
// each tile through the whole horizon -- TOLD.next (dynamics + reward) per step, then the
// terminal pi / min(Q1,Q2) (algorithm/tdmpc.py:83-92), or pi + dynamics for the policy rollout
// (tdmpc.py:113-118) -- without the latent state ever leaving the SM:
//
//   warps 0,2,3 producers: stream the pre-swizzled bf16 weight tiles (16 KB each, laid out in global
//            memory in exactly the order they are consumed) from L2 into a shared-memory ring with
//            bulk-TMA copies (cp.async.bulk ... mbarrier::complete_tx).
//   warp 1   MMA issuer: one thread issues tcgen05.mma (M=128, N=128, K=16) per k-step.
//            layer 1: A = [z,a] tile in shared memory (SS);  layer 2: A = h1 in TENSOR MEMORY (TS),
//            so the 128x512 hidden activation never touches shared memory; layer 3: A = h2 in smem.
//   warps 4-11 epilogue (two warpgroups, one per accumulator buffer): thread = trajectory.
//            tcgen05.ld the fp32 accumulator (software pipelined), add bias, ELU / LayerNorm+tanh /
//            tanh, then write the next layer's bf16 operand (tcgen05.st to TMEM or 128B-swizzled
//            st.shared), or reduce the N=1 reward / Q heads as fp32 dot products and accumulate
//            the discounted return in registers.  setmaxnreg moves registers from warps 0-3 to them.
//
// TMEM map (512 columns): [0,128) [128,256) accumulator ping-pong, [256,512) h1 as bf16x2 (K=512).
// The Q heads need the whole 512-wide pre-LayerNorm row, so their jobs use all 512 columns.
// All three roles walk the same job list (tc_model.cuh), which keeps the mbarrier phases aligned.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_model.cuh"
#include "tc_prims.cuh"

using namespace tc;

namespace {

constexpr int TILE_ROWS = 128;
constexpr int NTHREADS = 384;              // warps 0-3: producer / MMA / 2 idle; warps 4-7, 8-11: epilogue groups 0 and 1
// setmaxnreg: the control warpgroup frees (168-C)*128 registers, the two epilogue groups take (E-168)*256 -- exactly.  The MLP
// kernel's MMA warp spills its barrier phases at 56 (LDL/STL inside the issue loop); the GRU kernel's epilogue needs the 224.
template <bool DRNN> struct Regs { static constexpr int CTRL = DRNN ? 56 : 72, EPI = DRNN ? 224 : 216; };
constexpr int XCH_FLOATS = 3 * 2 * 128;   // epilogue-group exchange: {sum, sum-of-squares, dot} x 2 groups x 128 rows
constexpr int QC_FLOATS = 4 * 512;         // per-column constants of the current Q-head layer: bias, LN gamma, LN beta, w3
constexpr int TILE_BYTES = 16384;
constexpr int XS_BYTES = 8 * TILE_BYTES;      // 128 x 512 bf16
constexpr int MAX_JOBS = 22 * TDMPC_MAX_H + 20;
constexpr uint32_t TM_ACC = 0, TM_A = 256;
constexpr uint32_t TM_X = 256, TM_Y = 384;  // DRNN: gh_n and gi_n of the newval gate, inside the (then idle) h1 region
constexpr int MAX_STAGES = 6;
constexpr int SMEM_LIMIT = 232448;

struct KParams {
  TcModel m;
  RolloutParams p;
  const uint32_t* jobs;
  int njobs;
  int stages;
  int num_tiles;
  int producers;               // producer threads (warps 0, 2, 3); divides the ring depth
  int debug;                   // experiments only (TDMPC_TC_DEBUG; wrong results): bit0 = producers skip the weight copies, bit1 = accumulators
                               //   are released after the activation math, bit2 = the activation math of the MLP layers is skipped
  unsigned long long* trace;   // optional [3 roles][MAX_JOBS][2] clock64 stamps of block 0's third tile (tools/trace_rollout.py), then 128
                               //   end-of-tile clock64 + 128 globaltimer stamps of block 0, then per block {clock64 start, end, tiles, smid}
  // The job program again, in the kernel-parameter (constant) bank: the MMA warp reads it with uniform loads, which keeps its
  // whole control flow -- and with it every MMA operand -- in the uniform datapath.
  uint32_t jw[MAX_JOBS];
  int one;                     // always 1 (an opaque branch condition: tc_prims.cuh, tmem_drain_aux)
  int pair;                    // 1: value mode over CTA pairs (tile = blockIdx.x / 2, rank = blockIdx.x & 1): tc_model.cuh, build_jobs mode 2
  int njobs1, j1off;           //    rank 1's job count and its offset inside jobs / jw (rank 0: njobs, offset 0)
  unsigned int* tile_counter;  // dynamic tile scheduler: tickets handed out so far in this launch (zero between launches)
};
static_assert(sizeof(KParams) <= 4096, "kernel parameter space");

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float elu_fast(float x) {
  const float e = ex2_approx(x * 1.4426950408889634f) - 1.0f;
  return x > 0.0f ? x : e;
}
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = ex2_approx(x * 2.8853900817779268f);       // exp(2x)
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}
// tanh of the Q-head layer-1 activation (rounded to bf16 right afterwards).  EXACT = false: the one-MUFU tanh.approx.f32
// (2^-11 relative, a quarter of the bf16 rounding step that follows).  It is a DENSE perturbation -- one activation in eight
// changes its bf16 rounding -- that lifts the median trajectory-value deviation from the quantised oracle from 4e-7 to 1.5e-5 at
// dog-run dims (profiles/r2_bf16_error_probe.txt), 60 x inside north_star's 1e-3, and it is worth 6 % of the kernel (331 k vs
// 311 k cycles per tile: the two-MUFU form adds 1024 MUFU + 4096 FP instructions per trajectory to an XU-bound epilogue).
// Both forms are compiled; TDMPC_PREC_BF16 runs the approximate one, TDMPC_PREC_BF16_EXACT the two-MUFU one.
template <bool EXACT>
__device__ __forceinline__ float tanh_bf16(float x) {
  if (EXACT) return tanh_fast(x);
  float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float4 ldg4(const float* p, int i) { return __ldg(reinterpret_cast<const float4*>(p) + i); }

__device__ __forceinline__ int job_tiles(const TcModel& m, uint32_t job) {
  const int type = job & 7, chain = (job >> 3) & 7;
  switch (type) {
    case J_L1: return ((job >> 13) & F_A_HS) ? 2 : m.nkb1;
    case J_G: return m.nkb1 + 2;
    case J_GH: return 2;
    case J_GI: return m.nkb1 + m.gi_pad;
    case J_L2: return ((job >> 13) & (F_KLO | F_KHI)) ? 4 : 8;
    case J_L3: return ((job >> 13) & F_EPI) ? 0 : (m.l3rows[chain] == 128 ? 8 : 4);
    case J_QL1: return 4 * m.nkb1;
    default: return 32;
  }
}
__device__ __forceinline__ uint32_t job_first_tile(const TcModel& m, uint32_t job) {
  const int type = job & 7, chain = (job >> 3) & 7, chunk = (job >> 6) & 3;
  const TcChain& c = m.chain[chain];
  switch (type) {
    case J_L1: return c.l1_tile + chunk * (((job >> 13) & F_A_HS) ? 2 : m.nkb1);
    case J_G: return m.g_tile[chunk];
    case J_GH: return m.gh_tile;
    case J_GI: return m.gi_tile;
    case J_L2: return c.l2_tile + chunk * 8 + (((job >> 13) & F_KHI) ? 4 : 0);
    case J_L3: return c.l3_tile;
    case J_QL1: return c.l1_tile;
    default: return c.l2_tile;
  }
}

// bf16 store of one element of the [z,a] operand (128B-swizzled k-blocks of 64); explicit shared-space stores
// (ckb: index of the compact last k-block, see tc_prims.cuh umma_desc_k16, or -1)
__device__ __forceinline__ uint32_t za_offset(int row, int k, int ckb) {
  const int kb = k >> 6;
  return kb == ckb ? kb * TILE_BYTES + k16_offset(row, k & 63) : kb * TILE_BYTES + sw128_offset(row, k & 63);
}
__device__ __forceinline__ void za_store1(uint8_t* za, int row, int k, float v, int ckb) {
  const uint32_t a = smem_u32(za) + za_offset(row, k, ckb);
  const unsigned short b = __bfloat16_as_ushort(__float2bfloat16(v));
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(b) : "memory");
}
// 8 consecutive k (k0 % 8 == 0) as one 16-byte store
__device__ __forceinline__ void sw_store8(uint8_t* base, int row, int k0, const float* v) {
  const uint32_t a = smem_u32(base) + (k0 >> 6) * TILE_BYTES + sw128_offset(row, k0 & 63);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(v[0], v[1])), "r"(pack_bf16x2(v[2], v[3])),
               "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7])) : "memory");
}

// the same for the [z,a] operand, whose last k-block may be compact
__device__ __forceinline__ void za_store8(uint8_t* za, int row, int k0, const float* v, int ckb) {
  const uint32_t a = smem_u32(za) + za_offset(row, k0, ckb);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(v[0], v[1])), "r"(pack_bf16x2(v[2], v[3])),
               "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7])) : "memory");
}

// PAIR: the CTA-pair variant (kp.pair launches) is its own instantiation -- merely branching around its mailbox code costs the
// ordinary kernel registers and spills in the epilogues (measured: 3 - 5 %).
// ALT: the alternative value-mode schedules behind TDMPC_TC_OVERLAP / TDMPC_TC_SPLIT (parked layer 3, reward activation in XS, split-k
// layer 2) -- their branches too are compiled into their own instantiation only (ALTS below), not into the kernel the bench times.
// TRACE: the clock64 stamps of tools/trace_rollout.py (a per-job 64-bit stamp pointer lives across the epilogue math otherwise).
// WIDE: more than 8 action dimensions (dog-run: 38).  Their code -- the action-pair path of the z' epilogue and the batched
// unconditional loads of the row fills -- lives in its own instantiations: compiled into the kernel the bench times it cost 3.5 %
// at quadrotor dims (same-box A/B, profiles/r2c_ab_wide.txt), the epilogue's register allocation being what it is.
template <bool DRNN, bool EXACT, bool PAIR, bool ALT, bool TRACE, bool WIDE>
__global__ void __launch_bounds__(NTHREADS, 1) rollout_tc_kernel(const __grid_constant__ KParams kp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const TcModel& m = kp.m;
  const RolloutParams& p = kp.p;
  const int S = kp.stages;
  unsigned long long* const ktrace = TRACE ? kp.trace : nullptr;

  // 1024-byte aligned carve-up (the swizzle atoms need it)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xs = smem;
  uint8_t* za = xs + XS_BYTES;
  uint8_t* wring = za + m.za_bytes;
  const int ckb = m.za_compact ? m.nkb1 - 1 : -1;       // compact last k-block of [z,a]
  float* xch = reinterpret_cast<float*>(wring + S * TILE_BYTES);     // [3][2 groups][128 rows] epilogue-group exchange
  float* qcs = xch + XCH_FLOATS;                                       // [4][512] Q-head layer constants (16-byte aligned)
  uint32_t* sjobs = reinterpret_cast<uint32_t*>(qcs + QC_FLOATS);
  // CTA-pair mode: rank 0 (even block) dynamics / policy / Q1, rank 1 reward / Q2 / value.  Macros, not locals: without PAIR they
  // must fold to exactly the expressions the single-CTA kernel always used (a local here costs the epilogues a register).
  constexpr bool ALTS = !DRNN && (ALT || PAIR);          // job flags F_BUF0 / F_D3 / F_KLO / F_KHI / F_A_XS / F_TO_XS / F_PARK / F_EPI can occur
#define TC_NJOBS ((PAIR && (blockIdx.x & 1u)) ? kp.njobs1 : kp.njobs)
#define TC_JOFF ((PAIR && (blockIdx.x & 1u)) ? kp.j1off : 0)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sjobs + ((kp.njobs + (PAIR ? kp.njobs1 : 0) + 3) & ~3));
  uint64_t* w_full = bars;
  uint64_t* w_empty = w_full + MAX_STAGES;
  uint64_t* acc_full = w_empty + MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* h1_full = acc_empty + 2;     // h1 chunk c written to TMEM (layer-1 epilogue -> layer-2 MMAs)
  uint64_t* xs_full = h1_full + 4;       // XS chunk c written to smem  (layer-2 / QL1 epilogue -> layer-3 / QL2 MMAs)
  uint64_t* za_full = xs_full + 4;
  uint64_t* q_full = za_full + 1;        // Q-head accumulator chunk c complete (one barrier per chunk: never two phases ahead)
  // DRNN only: hidden-state operand written / newval-gate products X, Y complete / reset gate stored / update gate stored
  uint64_t* hs_full = q_full + 4;
  uint64_t* x_full = hs_full + 1;
  uint64_t* r_ready = x_full + 1;
  uint64_t* u_ready = r_ready + 1;
  uint64_t* d3_full = u_ready + 1;       // parked layer-3 accumulator complete (and the [z,a] tile no longer read)
  // Dynamic tile scheduler.  The first tile of a CTA is blockIdx.x; every further one is a ticket from a global counter, so that
  // SMs that run their tiles faster (the per-tile time differs by +-2.5 % across the chip: profiles/r2_tile_stats.txt) take more of
  // them instead of idling at the end.  Producer thread 0 -- the role that runs furthest ahead -- draws the ticket for tile n + 1
  // when it starts on tile n and publishes it through a 4-entry ring (tq_full: one arrival per entry use); all roles read the
  // ring in order.  They stay within a tile of each other (weight ring, accumulator and operand barriers), so an entry is
  // never overwritten before its last reader is done.
  uint64_t* tq_full = d3_full + 1;
  uint32_t* tile_q = reinterpret_cast<uint32_t*>(tq_full + 4);
  uint32_t* tmem_slot = tile_q + 4;
  // DRNN: during the steps XS is idle (the prior MLP runs layer 2 from TMEM, the reward head needs no h2 operand), so it
  // holds the GRU state: fp32 [column][row] (conflict-free for thread = row) and the bf16 K-major operand HS (2 k-blocks).
  float* hf = reinterpret_cast<float*>(xs);
  const uint32_t hf_addr = smem_u32(hf);
  uint8_t* hs = xs + 128 * 128 * 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < TC_NJOBS; i += NTHREADS) sjobs[i] = kp.jobs[TC_JOFF + i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    for (int i = 0; i < 4; ++i) { mbar_init(&h1_full[i], 128); mbar_init(&xs_full[i], 128); }
    mbar_init(za_full, 128);
    for (int i = 0; i < 4; ++i) mbar_init(&q_full[i], 1);
    mbar_init(hs_full, 256); mbar_init(x_full, 1); mbar_init(r_ready, 128); mbar_init(u_ready, 128);
    mbar_init(d3_full, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&tq_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (*tmem_slot != 0) asm volatile("trap;");   // all 512 columns of the SM are ours: the base is 0, so TMEM addresses are immediates
  constexpr uint32_t tmem = 0;
  // n-th tile of this CTA (n = 0, 1, ...) or -1 when the launch has no more tiles; every role calls it once per tile, in order
  uint32_t tq_n = 0;
  auto next_tile = [&]() -> int {
    const uint32_t e = tq_n & 3u;
    mbar_wait(&tq_full[e], (tq_n >> 2) & 1u, 16);
    const uint32_t t = *reinterpret_cast<volatile uint32_t*>(&tile_q[e]);
    ++tq_n;
    return t < (uint32_t)kp.num_tiles ? (int)t : -1;
  };
  int tiles_done = 0;
  const long long t_start = ktrace ? clock64() : 0;

  if (warp < 4) {
    // warps 0-3 form one warpgroup: producer, MMA issuer, two idle warps.  They hand most of their
    // registers to the two epilogue warpgroups.
    setmaxnreg_dec<Regs<DRNN>::CTRL>();
    if (warp != 1) {
      // ============================== producers: weight tiles, L2 -> smem ring ==============================
      // One thread can only issue a bulk copy every ~400 cycles (measured: tools/smem_contention.cu), which caps a
      // single-producer weight stream at ~40 B/clk -- less than the MMAs consume.  So up to three warps (0, 2, 3)
      // run a producer thread each; producer k of P owns the tiles (and, because P divides the ring depth, the
      // ring slots) k mod P.
      const int P = kp.producers;
      const int pk = warp == 0 ? 0 : warp - 1;
      if (lane == 0 && pk < P) {
        uint32_t stage = 0, phase = 0, mine = 0;    // mine: tile counter mod P
        if (pk == 0) { tile_q[0] = PAIR ? (blockIdx.x >> 1) : blockIdx.x; mbar_arrive(&tq_full[0]); }
        for (uint32_t n = 0; next_tile() >= 0; ++n) {
          if (PAIR && pk == 0) {                          // one tile per CTA pair
            tile_q[(n + 1) & 3u] = 0xFFFFFFFFu;
            mbar_arrive(&tq_full[(n + 1) & 3u]);
          } else if (pk == 0) {
            // the ticket for tile n + 1, drawn now so that the weight stream never waits for it.  Exactly num_tiles tickets are
            // drawn per launch (one per tile started, CTA-first tiles included through the failing last draw of every CTA), so
            // the holder of the last one puts the counter back to zero for the next launch.
            const unsigned ticket = atomicAdd(kp.tile_counter, 1u);
            if (ticket == (unsigned)kp.num_tiles - 1u) atomicExch(kp.tile_counter, 0u);
            tile_q[(n + 1) & 3u] = ticket + gridDim.x;
            mbar_arrive(&tq_full[(n + 1) & 3u]);
          }
          for (int j = 0; j < TC_NJOBS; ++j) {
            const uint32_t job = sjobs[j];
            const int nt = job_tiles(m, job);
            const uint8_t* src = m.tiles + (size_t)job_first_tile(m, job) * TILE_BYTES;
            for (int i = 0; i < nt; ++i) {
              if (mine == (uint32_t)pk) {
                const int ti = i;
                mbar_wait(&w_empty[stage], phase ^ 1, 1);
                if (TRACE && (kp.debug & 1)) mbar_arrive(&w_full[stage]);
                else {
                  mbar_arrive_expect_tx(&w_full[stage], TILE_BYTES);
                  bulk_g2s(wring + stage * TILE_BYTES, src + (size_t)ti * TILE_BYTES, TILE_BYTES, &w_full[stage]);
                }
              }
              if (++mine == (uint32_t)P) mine = 0;
              if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ============================== MMA issuer ==============================
      // The whole warp walks the job list converged (waits are warp-wide); one elected lane issues.
      // Elected once while converged: the same lane issues every MMA and commit.  The warp stays converged and the election
      // predicate sits inside the asm statements (tc_prims.cuh, umma_*_p), which keeps the operands in uniform registers.
      const uint32_t lead = elect_one() ? 1u : 0u;
      uint32_t stage = 0, wphase = 0;
      uint32_t ecnt[2] = {0, 0}, nbuf = 0, za_phase = 0;
      // bit c = parity to wait for.  h1 and XS have separate barriers: a shared one would let the
      // epilogue (reward layer 1 -> h1) complete a second phase before the MMA warp has waited for
      // the first (dynamics layer 2 -> XS), and a parity wait cannot tell the two apart.
      uint32_t h1_phase = 0, xs_phase = 0;
      const uint32_t idesc128 = umma_idesc_bf16(128, 128);
      const uint32_t za_addr = smem_u32(za), xs_addr = smem_u32(xs), ring_addr = smem_u32(wring);
      const int ks1_total = m.K1p / 16;
      const uint32_t ring_lo[4] = {umma_desc_lo(ring_addr), umma_desc_lo(ring_addr + TILE_BYTES), umma_desc_lo(ring_addr + 2 * TILE_BYTES),
                                   umma_desc_lo(ring_addr + 3 * TILE_BYTES)};

      // Weight-ring consumer.  tile_acquire() returns the smem address of the current tile and already
      // advances the ring and probes the NEXT tile's barrier, so that probe's latency overlaps the MMA
      // issue; tile_release(slot) commits the slot that was just consumed.
      uint32_t ready = 0;                    // result of the early probe of w_full[stage]
      uint32_t cur_stage = 0;
      auto tile_acquire = [&]() -> uint32_t {
        if (!ready) mbar_wait(&w_full[stage], wphase, 2);
        tcgen05_fence_after();
        cur_stage = stage;
        if (++stage == (uint32_t)S) { stage = 0; wphase ^= 1; }
        ready = mbar_try_wait(&w_full[stage], wphase) ? 1u : 0u;
        return ring_addr + cur_stage * TILE_BYTES;
      };
      auto tile_release = [&]() {            // commit -> frees the slot when the MMAs that read it retire
        umma_commit_p(&w_empty[cur_stage], lead);
      };
      auto acc_acquire = [&](int b) {
        mbar_wait(&acc_empty[b], (ecnt[b] & 1) ^ 1, 3);
        ++ecnt[b];
      };
      auto h1_wait = [&](int c) {
        mbar_wait(&h1_full[c], (h1_phase >> c) & 1, 4);
        h1_phase ^= 1u << c;
      };
      auto xs_wait = [&](int c) {
        mbar_wait(&xs_full[c], (xs_phase >> c) & 1, 5);
        xs_phase ^= 1u << c;
      };
      uint32_t hs_phase = 0;
      const uint32_t hs_addr = smem_u32(hs);
      // one k-block (<= 4 k-steps of 16) with both operands in shared memory
      auto kblock_ss = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, int nks, bool first) {
        const uint32_t alo = umma_desc_lo(a_addr), blo = umma_desc_lo(b_addr);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          if (ks < nks) umma_ss_p(d, umma_desc_make(alo + 2 * ks), umma_desc_make(blo + 2 * ks), idesc, (first && ks == 0) ? 0u : 1u, lead);
      };

      // k-block kb of the [z,a] operand against the weight tile at b_addr
      auto kblock_za = [&](uint32_t d, int kb, uint32_t b_addr, uint32_t idesc, bool first) {
        if (kb == ckb) umma_ss_p(d, umma_desc_k16(za_addr + kb * TILE_BYTES), umma_desc_make(umma_desc_lo(b_addr)), idesc, first ? 0u : 1u, lead);
        else kblock_ss(d, za_addr + kb * TILE_BYTES, b_addr, idesc, min(4, ks1_total - kb * 4), first);
      };

      for (int it = 0; next_tile() >= 0; ++it)
        for (int j = 0; j < TC_NJOBS; ++j) {
          const uint32_t job = kp.jw[TC_JOFF + j];
          const int type = job & 7, chain = (job >> 3) & 7, chunk = (job >> 6) & 3, flags = job >> 13;
          const bool tr = ktrace && blockIdx.x == 0 && it == 2 && lane == 0;
          if (tr) ktrace[(0 * MAX_JOBS + j) * 2] = clock64();
          if (flags & F_WAIT_ZA) { mbar_wait(za_full, za_phase, 6); za_phase ^= 1; }
          if (DRNN && (flags & F_WAIT_HS)) { mbar_wait(hs_full, hs_phase, 10); hs_phase ^= 1; tcgen05_fence_after(); }
          if (ALTS && (flags & F_BUF0)) nbuf = 0;
          if (type == J_L1) {
            const int b = nbuf; nbuf ^= 1;
            acc_acquire(b);
            tcgen05_fence_after();
            const uint32_t d = tmem + TM_ACC + b * 128;
            if (DRNN && (flags & F_A_HS)) {
              // prior / reward layer 1 of the DRNN model: A = the hidden state h' (K = 128)
              for (int kb = 0; kb < 2; ++kb) {
                const uint32_t wt = tile_acquire();
                kblock_ss(d, hs_addr + kb * TILE_BYTES, wt, idesc128, 4, kb == 0);
                tile_release();
              }
            } else {
              for (int kb = 0; kb < m.nkb1; ++kb) {
                const uint32_t wt = tile_acquire();
                kblock_za(d, kb, wt, idesc128, kb == 0);
                tile_release();
              }
            }
            umma_commit_p(&acc_full[b], lead);
            // every MMA issued so far -- the parked layer 3 and all layer-1 jobs that read this step's [z,a] -- complete
            if (ALTS && (flags & F_D3)) umma_commit_p(d3_full, lead);
          } else if (DRNN && type == J_G) {
            // gate pre-activation: [z,a] W_ih[g]^T + h W_hh[g]^T accumulated into one buffer (no biases, rnns.py:13-14)
            const int b = nbuf; nbuf ^= 1;
            acc_acquire(b);
            tcgen05_fence_after();
            const uint32_t d = tmem + TM_ACC + b * 128;
            for (int kb = 0; kb < m.nkb1; ++kb) {
              const uint32_t wt = tile_acquire();
              kblock_za(d, kb, wt, idesc128, kb == 0);
              tile_release();
            }
            for (int kb = 0; kb < 2; ++kb) {
              const uint32_t wt = tile_acquire();
              kblock_ss(d, hs_addr + kb * TILE_BYTES, wt, idesc128, 4, false);
              tile_release();
            }
            umma_commit_p(&acc_full[b], lead);
          } else if (DRNN && type == J_GH) {
            // gh_n = h W_hh[newval]^T -> X
            for (int kb = 0; kb < 2; ++kb) {
              const uint32_t wt = tile_acquire();
              kblock_ss(tmem + TM_X, hs_addr + kb * TILE_BYTES, wt, idesc128, 4, kb == 0);
              tile_release();
            }
          } else if (DRNN && type == J_GI) {
            // gi_n = [z,a] W_ih[newval]^T -> Y; the n-gate epilogue forms gi_n + r * gh_n in registers (rnns.py:27)
            for (int kb = 0; kb < m.nkb1; ++kb) {
              const uint32_t wt = tile_acquire();
              kblock_za(tmem + TM_Y, kb, wt, idesc128, kb == 0);
              tile_release();
            }
            for (int i = 0; i < m.gi_pad; ++i) { tile_acquire(); tile_release(); }     // padding tiles keep the ring aligned
            umma_commit_p(x_full, lead);                                            // covers J_GH's MMAs too
          } else if (ALTS && type == J_L2 && (flags & (F_KLO | F_KHI))) {
            // half of a layer-2 chunk (see F_KLO): 4 k-blocks = 4 tiles = one turn of the ring
            const bool hi = (flags & F_KHI) != 0;
            const int b = nbuf; nbuf ^= 1;                 // KLO c0, KLO c1, KHI c0, KHI c1: the ping-pong lands on the same buffers
            if (!hi) acc_acquire(b);
            tcgen05_fence_after();
            const uint32_t d = tmem + TM_ACC + b * 128;
            const int kb0 = hi ? 4 : 0;
            if (S == 4 && stage == 0) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const int kb = kb0 + k4;
                if (chunk == 0 && (kb & 1) == 0) { h1_wait(kb >> 1); tcgen05_fence_after(); }
                mbar_wait(&w_full[k4], wphase, 2);
                tcgen05_fence_after();
                const uint32_t blo = ring_lo[k4];
                const uint32_t a = tmem + TM_A + kb * 32;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts_p(d, a + ks * 8, umma_desc_make(blo + 2 * ks), idesc128, (kb | ks) ? 1u : 0u, lead);
                umma_commit_p(&w_empty[k4], lead);
              }
              wphase ^= 1;                     // one full turn
              ready = 0;
            } else {
              for (int k4 = 0; k4 < 4; ++k4) {
                const int kb = kb0 + k4;
                if (chunk == 0 && (kb & 1) == 0) { h1_wait(kb >> 1); tcgen05_fence_after(); }
                const uint32_t wt = tile_acquire();
                const uint32_t blo = umma_desc_lo(wt);
                const uint32_t a = tmem + TM_A + kb * 32;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts_p(d, a + ks * 8, umma_desc_make(blo + 2 * ks), idesc128, (kb | ks) ? 1u : 0u, lead);
                tile_release();
              }
            }
            if (hi) umma_commit_p(&acc_full[b], lead);
          } else if (type == J_L2) {
            const int b = nbuf; nbuf ^= 1;
            acc_acquire(b);
            tcgen05_fence_after();
            const uint32_t d = tmem + TM_ACC + b * 128;
            if (ALTS && (flags & F_A_XS)) {
              // reward chain of the overlapped schedule: A = its hidden activation in XS (both operands in shared memory)
#pragma unroll
              for (int kb = 0; kb < 8; ++kb) {
                if (chunk == 0 && (kb & 1) == 0) { xs_wait(kb >> 1); tcgen05_fence_after(); }
                const uint32_t wt = tile_acquire();
                kblock_ss(d, xs_addr + kb * TILE_BYTES, wt, idesc128, 4, kb == 0);
                tile_release();
              }
            } else if (S == 4 && stage == 0) {
              // Static fast path.  With a 4-slot ring every layer consumes a multiple of 4 tiles, so a layer-2 job
              // always starts at slot 0 and its 8 tiles are two full turns of the ring: slot, barrier and
              // descriptor of every k-block are compile-time choices among 4 precomputed values, which keeps the
              // single issuing thread's instruction stream short enough to stay ahead of the tensor pipe.
#pragma unroll
              for (int kb = 0; kb < 8; ++kb) {
                const int st = kb & 3;
                if (chunk == 0 && (kb & 1) == 0) { h1_wait(kb >> 1); tcgen05_fence_after(); }
                mbar_wait(&w_full[st], wphase ^ (uint32_t)(kb >> 2), 2);
                tcgen05_fence_after();
                {
                  const uint32_t blo = ring_lo[st];
                  const uint32_t a = tmem + TM_A + kb * 32;
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) umma_ts_p(d, a + ks * 8, umma_desc_make(blo + 2 * ks), idesc128, (kb | ks) ? 1u : 0u, lead);
                  umma_commit_p(&w_empty[st], lead);
                }
              }
              ready = 0;                       // slot 0 again, same parity: the early probe must be redone
            } else {
#pragma unroll
              for (int kb = 0; kb < 8; ++kb) {
                if (chunk == 0 && (kb & 1) == 0) { h1_wait(kb >> 1); tcgen05_fence_after(); }
                const uint32_t wt = tile_acquire();
                {
                  const uint32_t blo = umma_desc_lo(wt);
                  const uint32_t a = tmem + TM_A + kb * 32;
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) umma_ts_p(d, a + ks * 8, umma_desc_make(blo + 2 * ks), idesc128, (kb | ks) ? 1u : 0u, lead);
                }
                tile_release();
              }
            }
            umma_commit_p(&acc_full[b], lead);
          } else if (ALTS && type == J_L3 && (flags & F_EPI)) {
            // epilogue-only job
          } else if (type == J_L3) {
            const bool park = ALTS && (flags & F_PARK);
            int b = 0;
            if (!park) { b = nbuf; nbuf ^= 1; acc_acquire(b); }
            tcgen05_fence_after();
            // parked: the accumulator lives in the first columns of the h1 region, which this step's layer-2 MMAs (issued
            // earlier, executed in order) were the last to read and nobody writes before the parked result is consumed
            const uint32_t d = park ? tmem + TM_A : tmem + TM_ACC + b * 128;
            const uint32_t idesc = umma_idesc_bf16(128, m.n3p[chain]);
            const bool half = m.l3rows[chain] == 64;
            if (DRNN && (flags & F_A_TMEM)) {
              // second (last) Linear of the DRNN prior MLP: A = h1 in TMEM
              uint32_t wt = 0;
#pragma unroll
              for (int kb = 0; kb < 8; ++kb) {
                if ((kb & 1) == 0) { h1_wait(kb >> 1); tcgen05_fence_after(); }
                if (!half || (kb & 1) == 0) wt = tile_acquire();
                {
                  const uint32_t blo = umma_desc_lo(wt + ((half && (kb & 1)) ? 8192u : 0u));
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) umma_ts_p(d, tmem + TM_A + kb * 32 + ks * 8, umma_desc_make(blo + 2 * ks), idesc, (kb | ks) ? 1u : 0u, lead);
                }
                if (!half || (kb & 1)) tile_release();
              }
            } else if (half && S == 4 && stage == 0) {
              // static fast path: 4 tiles (two k-blocks of 64 weight rows each) = one full turn of the ring
#pragma unroll
              for (int kb = 0; kb < 8; ++kb) {
                const int st = kb >> 1;
                if ((kb & 1) == 0) {
                  xs_wait(kb >> 1); tcgen05_fence_after();
                  mbar_wait(&w_full[st], wphase, 2);
                  tcgen05_fence_after();
                }
                {
                  const uint32_t alo = umma_desc_lo(xs_addr + kb * TILE_BYTES), blo = ring_lo[st] + ((kb & 1) ? (8192u >> 4) : 0u);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_ss_p(d, umma_desc_make(alo + 2 * ks), umma_desc_make(blo + 2 * ks), idesc, (kb | ks) ? 1u : 0u, lead);
                  if (kb & 1) umma_commit_p(&w_empty[st], lead);
                }
              }
              wphase ^= 1;                     // one full turn
              ready = 0;
            } else {
              uint32_t wt = 0;
#pragma unroll
              for (int kb = 0; kb < 8; ++kb) {
                if ((kb & 1) == 0) { xs_wait(kb >> 1); tcgen05_fence_after(); }
                if (!half || (kb & 1) == 0) wt = tile_acquire();
                kblock_ss(d, xs_addr + kb * TILE_BYTES, wt + ((half && (kb & 1)) ? 8192u : 0u), idesc, 4, kb == 0);
                if (!half || (kb & 1)) tile_release();
              }
            }
            if (!park) umma_commit_p(&acc_full[b], lead);
          } else {
            // Q-head layers: one 512-column accumulator = both ping-pong buffers + the h1 region.
            // Each 128-column chunk is committed to its own barrier so the epilogue can accumulate the
            // LayerNorm statistics of chunk c while chunk c+1 is computed.
            acc_acquire(0); acc_acquire(1);
            tcgen05_fence_after();
            const bool first = type == J_QL1;
            const int nkb = first ? m.nkb1 : 8;
            if (!first && S == 4 && stage == 0) {
              // static fast path (see layer 2): 4 chunks x 8 tiles = 8 full turns of the ring.  (N = 256 instructions over
              // chunk pairs were tried -- an SS MMA is bound by its operand reads and a 256-wide one reads the A slice once
              // for two chunks, 171 vs 2 x 107 cycles in tools/mma_rate.cu -- but the LayerNorm statistics pass then starts
              // a chunk later and the tile got 1 % slower.)
              for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int kb = 0; kb < 8; ++kb) {
                  const int st = kb & 3;
                  if (c == 0 && (kb & 1) == 0) { xs_wait(kb >> 1); tcgen05_fence_after(); }
                  mbar_wait(&w_full[st], wphase ^ (uint32_t)(kb >> 2), 2);
                  tcgen05_fence_after();
                  {
                    const uint32_t alo = umma_desc_lo(xs_addr + kb * TILE_BYTES), blo = ring_lo[st];
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                      umma_ss_p(tmem + c * 128, umma_desc_make(alo + 2 * ks), umma_desc_make(blo + 2 * ks), idesc128, (kb | ks) ? 1u : 0u, lead);
                    umma_commit_p(&w_empty[st], lead);
                  }
                }
                umma_commit_p(&q_full[c], lead);
              }
              ready = 0;
            } else
            for (int c = 0; c < 4; ++c) {
              for (int kb = 0; kb < nkb; ++kb) {
                if (!first && c == 0 && (kb & 1) == 0) { xs_wait(kb >> 1); tcgen05_fence_after(); }
                const uint32_t wt = tile_acquire();
                if (first) kblock_za(tmem + c * 128, kb, wt, idesc128, kb == 0);
                else kblock_ss(tmem + c * 128, xs_addr + kb * TILE_BYTES, wt, idesc128, 4, kb == 0);
                tile_release();
              }
              umma_commit_p(&q_full[c], lead);
            }
          }
          if (tr) ktrace[(0 * MAX_JOBS + j) * 2 + 1] = clock64();
          __syncwarp();
        }
    }
  } else {
    // ============================== epilogue: two warpgroups, thread = trajectory ==============================
    // Group g (warps 4-7: g=0, warps 8-11: g=1) owns accumulator buffer g: jobs alternate buffers, so
    // the two groups drain consecutive 128-column chunks concurrently (two warps per SM sub-partition
    // hide each other's TMEM / L1 / MUFU latencies).  Row-wise reductions that span chunks (reward
    // dot product, LayerNorm statistics, Q dot product) are combined through `xch` in shared memory.
    setmaxnreg_inc<Regs<DRNN>::EPI>();
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const int L = m.L, A = m.A;
    const long long total = (long long)p.B * p.T;
    uint32_t fcnt = 0, nbuf = 0, q_phase = 0;
    uint32_t d3_phase = 0;                          // overlapped schedule: phase of d3_full (group 0 only)
    uint32_t x_phase = 0, ga = 0;                  // DRNN: phase of x_full / r_ready / u_ready, group that owns the reset gate of this step
    // exchange slots of my row as 32-bit shared-space addresses (explicit ld.shared / st.shared: a generic float* costs a
    // register pair each and takes the slower generic path): x_* the OTHER group's slot (read side), m_* mine (write side)
    const uint32_t xch_a = smem_u32(xch) + (uint32_t)row * 4u;
    const uint32_t x_sum = xch_a + (0 * 2 + (grp ^ 1)) * 512, x_ssq = xch_a + (1 * 2 + (grp ^ 1)) * 512, x_dot = xch_a + (2 * 2 + (grp ^ 1)) * 512;
    const uint32_t m_sum = xch_a + (0 * 2 + grp) * 512, m_ssq = xch_a + (1 * 2 + grp) * 512, m_dot = xch_a + (2 * 2 + grp) * 512;

    auto acc_wait = [&]() {
      mbar_wait(&acc_full[grp], fcnt & 1, 7);
      ++fcnt;
    };
    // early release: the accumulator goes back to the MMA warp as soon as it sits in registers (tmem_drain_aux)
    const int go_math = (TRACE && (kp.debug & 4)) ? 0 : kp.one;     // experiment (bit 2): skip the activation math of the MLP layers (power probe, wrong results)
    const bool late_release = TRACE && (kp.debug & 2) != 0;      // experiment: release after the activation math, as before
    auto acc_release = [&]() {
      if (!late_release) { tcgen05_fence_before(); mbar_arrive(&acc_empty[grp]); }
    };
    auto acc_release_late = [&]() {
      if (late_release) { tcgen05_fence_before(); mbar_arrive(&acc_empty[grp]); }
    };
    auto load8 = [&](const float* ptr, int idx, float4 (&x)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = ldg4(ptr, idx * 8 + i);
    };

    // the tile after the current one (-1: none), without consuming the queue entry
    auto peek_tile = [&]() -> int {
      const uint32_t e = tq_n & 3u;
      mbar_wait(&tq_full[e], (tq_n >> 2) & 1u, 16);
      const uint32_t t = *reinterpret_cast<volatile uint32_t*>(&tile_q[e]);
      return t < (uint32_t)kp.num_tiles ? (int)t : -1;
    };
    // [z0, a0, 0...] rows of tile `tile` into the [z,a] operand (my row), 16 loads in flight per round.  Called at the start of a
    // tile, or -- F_TILE_END -- for the NEXT tile from inside the last Q-head job of the current one, while its MMAs run and
    // this group would only wait: the tensor pipe then does not idle through the fill at every tile boundary.
    auto fill_za = [&](int tile) {
      const uint32_t g2 = (uint32_t)tile * TILE_ROWS + row;                       // B * T < 2^31 (checked at launch)
      const bool v2 = (long long)g2 < total;
      const uint32_t pb2 = v2 ? g2 / (uint32_t)p.T : 0u, pj2 = v2 ? g2 - pb2 * (uint32_t)p.T : 0u;
      const float* ar = (p.actions && p.mode == MODE_VALUE) ? p.actions + (((size_t)pb2 * p.Tcap + pj2) * p.H) * A : nullptr;
      const float* zr = p.z0 + (size_t)pb2 * L;
      for (int k0 = 0; k0 < (ckb >= 0 ? m.K1p : m.nkb1 * 64); k0 += 16) {
        float v[16];
        if (WIDE) {
          // unconditional loads (padding columns and rows past the end read z0 of problem 0), zeroed afterwards: a load under a
          // per-element condition is followed by a move of its result, which serialises the sixteen round trips
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = k0 + i;
            const float* src = (k >= L && k < L + A && ar) ? ar + (k - L) : zr + (k < L ? k : 0);
            v[i] = __ldg(src);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = k0 + i;
            if (!(v2 && (k < L || (k < L + A && ar)))) v[i] = 0.0f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = k0 + i;
            float x = 0.0f;
            if (v2) {
              if (k < L) x = __ldg(zr + k);
              else if (k < L + A && ar) x = __ldg(ar + (k - L));
            }
            v[i] = x;
          }
        }
        za_store8(za, row, k0, v, ckb);
        za_store8(za, row, k0 + 8, v + 8, ckb);
      }
      fence_proxy_async_smem();
      mbar_arrive(za_full);
    };
    bool prefilled = false;
    int tile_idx;
    for (int it = 0; (tile_idx = next_tile()) >= 0; ++it, ++tiles_done) {
      const long long g = (long long)tile_idx * TILE_ROWS + row;
      const bool valid = g < total;
      const int pb = valid ? (int)(g / p.T) : 0;               // B * T < 2^31 (checked at launch)
      const int pj = valid ? (int)(g - (long long)pb * p.T) : 0;
      const float* act_row = p.actions ? p.actions + (((size_t)pb * p.Tcap + pj) * p.H) * A : nullptr;
      float* pi_row = p.pi_actions ? p.pi_actions + (((size_t)pb * p.Tcap + pj) * p.H) * A : nullptr;
      float G = 0.0f, rlast = 0.0f, dot = 0.0f, q1v = 0.0f;
      float pd[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // DRNN policy rollout: partial outputs of the last policy Linear

      if (DRNN) {
        // initial GRU state of the problem (tdmpc_icem_similarity_drnn.py:202): fp32 copy + bf16 operand, 64 columns per
        // group.  XS was last read by the previous tile's Q-head MMAs, which both groups have seen complete.
        const float* h0 = p.hidden0 ? p.hidden0 + (size_t)pb * 128 : nullptr;
        const bool hz = !(h0 && valid);                       // zero state
        const float* hp = hz ? p.z0 : h0;                       // (unconditional loads: see fill_za)
        for (int k0 = grp * 64; k0 < grp * 64 + 64; k0 += 16) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __ldg(hp + (hz ? 0 : k0 + i));
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (hz) v[i] = 0.0f;
            sts1(hf_addr + ((k0 + i) * 128 + row) * 4, v[i]);
          }
          sw_store8(hs, row, k0, v);
          sw_store8(hs, row, k0 + 8, v + 8);
        }
        fence_proxy_async_smem();
        mbar_arrive(hs_full);
      }

      // initial [z0, a0, 0...] row (group 0) -- unless the previous tile's last Q-head job already wrote it (fill_za below)
      if (grp == 0 && !prefilled) fill_za(tile_idx);
      prefilled = false;
      // The action rows of the NEXT tile into L2 (group 1, idle here): a tile's rows were written by the sampling kernel a whole
      // launch ago (176 MB at 4096 environments: HBM), and every step's layer-3 epilogue reads its slice of them with one row per
      // thread -- at dog-run dims (38 actions) four dependent rounds of uncoalesced HBM latency per step without this.
      if (grp == 1 && p.mode == MODE_VALUE) {
        const int nt = peek_tile();
        const uint32_t g2 = (uint32_t)nt * TILE_ROWS + row;
        if (nt >= 0 && (long long)g2 < total) {
          const uint32_t pb2 = g2 / (uint32_t)p.T, pj2 = g2 - pb2 * (uint32_t)p.T;
          const char* pa = reinterpret_cast<const char*>(p.actions + (((size_t)pb2 * p.Tcap + pj2) * p.H) * A);
          const int nbytes = p.H * A * 4;
          for (int o = 0; o < nbytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + o));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + nbytes - 4));
        }
      }

      // CTA-pair mailbox (global memory): rows [z_t | a_t] of this tile, Q1, and one flag word per slot
      float* mb_rows = PAIR ? kp.m.pair_rows + (size_t)tile_idx * (TDMPC_MAX_H + 1) * 128 * m.K1p : nullptr;
      float* mb_q1 = PAIR ? kp.m.pair_rows + (size_t)PAIR_MAX_TILES * (TDMPC_MAX_H + 1) * 128 * m.K1p + (size_t)tile_idx * 128 : nullptr;
      unsigned int* mb_flags = PAIR ? kp.m.pair_flags + tile_idx * 32 : nullptr;
      auto mb_publish = [&](int idx) {                 // my group's 128 threads have written their rows: make them visible, raise the flag
        __threadfence();
        asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
        if ((threadIdx.x & 127) == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(mb_flags + idx), "r"(1u) : "memory");
      };
      auto mb_wait = [&](int idx) {
        const long long t0 = clock64();
        unsigned int v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(mb_flags + idx) : "memory");
          if (!v && clock64() - t0 > 4000000000LL) asm volatile("trap;");
        } while (!v);
      };
      auto mb_consume = [&](int idx, int idx2) {       // every thread of my group has read the slot: lower the flag(s) for the next launch
        asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
        if ((threadIdx.x & 127) == 0) { mb_flags[idx] = 0u; if (idx2 >= 0) mb_flags[idx2] = 0u; }
      };
      for (int j = 0; j < TC_NJOBS; ++j) {
        const uint32_t job = sjobs[j];
        const int type = job & 7, chain = (job >> 3) & 7, chunk = (job >> 6) & 3, t = (job >> 8) & 31, flags = job >> 13;
        const TcChain& ch = m.chain[chain];
        const bool tr = ktrace && blockIdx.x == 0 && it == 2 && threadIdx.x % 128 == 0;
        unsigned long long* trp = ktrace + ((1 + grp) * MAX_JOBS + j) * 2;

        if (DRNN && type >= J_G) {
          // ---- NormGRUCell (models/rnns.py:19-29): thread = trajectory, LayerNorm over the 128 gate values in-thread ----
          if (type == J_GH) continue;                                  // MMAs only
          if (type == J_G) {
            const int b = nbuf; nbuf ^= 1;
            if (chunk == 0) ga = (uint32_t)b;
            if (b != grp) continue;
            if (tr) trp[0] = clock64();
            const uint32_t acc = tlane + TM_ACC + b * 128;
            acc_wait();
            tcgen05_fence_after();
            float sum = 0.0f, ssq = 0.0f;
            tmem_stream(acc, 4, [&](int, uint32_t (&v)[32]) {
#pragma unroll
              for (int i = 0; i < 32; ++i) { const float y = __uint_as_float(v[i]); sum += y; ssq = fmaf(y, y, ssq); }
            });
            const float mu = sum * (1.0f / 128.0f);
            const float rstd = 1.0f / sqrtf(fmaxf(ssq * (1.0f / 128.0f) - mu * mu, 0.0f) + 1e-3f);
            const float* lg = m.gln_g[chunk];
            const float* lb = m.gln_b[chunk];
            // reset / update gate: sigmoid(LN(.)) written back over its own accumulator (rnns.py:25-26); the n-gate
            // epilogue reads both and then releases the buffers
#pragma unroll 1
            for (int sg = 0; sg < 4; ++sg) {
              uint32_t a[32];
              tmem_ld32(acc + sg * 32, a);
              tmem_wait_ld_dep(a);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 gg = ldg4(lg, sg * 8 + i), bb = ldg4(lb, sg * 8 + i);
                a[4 * i] = __float_as_uint(sigmoid_fast((__uint_as_float(a[4 * i]) - mu) * rstd * gg.x + bb.x));
                a[4 * i + 1] = __float_as_uint(sigmoid_fast((__uint_as_float(a[4 * i + 1]) - mu) * rstd * gg.y + bb.y));
                a[4 * i + 2] = __float_as_uint(sigmoid_fast((__uint_as_float(a[4 * i + 2]) - mu) * rstd * gg.z + bb.z));
                a[4 * i + 3] = __float_as_uint(sigmoid_fast((__uint_as_float(a[4 * i + 3]) - mu) * rstd * gg.w + bb.w));
              }
              tmem_st32(acc + sg * 32, a);
            }
            tmem_wait_st();
            tcgen05_fence_before();
            mbar_arrive(chunk == 0 ? r_ready : u_ready);
            if (tr) trp[1] = clock64();
            continue;
          }
          // J_GI: n = tanh(LN_n(gi_n + r * gh_n)); h' = u * n + (1 - u) * h  (rnns.py:27-28).  Both groups work on it, 64
          // columns each (it sits on the step's critical path); x = gi_n + r * gh_n stays in registers between the
          // statistics pass and the normalising pass; the LayerNorm statistics are exchanged through `xch`.
          const uint32_t gph = x_phase;                       // one phase of x_full / r_ready / u_ready per step
          x_phase ^= 1;
          if (tr) trp[0] = clock64();
          const int c0 = grp * 64;
          const uint32_t accr = tlane + TM_ACC + ga * 128 + c0, accu = tlane + TM_ACC + (ga ^ 1u) * 128 + c0;
          mbar_wait(x_full, gph, 13);
          mbar_wait(r_ready, gph, 12);
          tcgen05_fence_after();
          float xv[64];
          float sum = 0.0f, ssq = 0.0f;
#pragma unroll
          for (int sg = 0; sg < 2; ++sg) {
            uint32_t r[32], gi[32], gh[32];
            tmem_ld32(accr + sg * 32, r);
            tmem_ld32(tlane + TM_Y + c0 + sg * 32, gi);
            tmem_ld32(tlane + TM_X + c0 + sg * 32, gh);
            tmem_wait_ld_dep(r); tmem_wait_ld_dep(gi); tmem_wait_ld_dep(gh);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float y = fmaf(__uint_as_float(r[i]), __uint_as_float(gh[i]), __uint_as_float(gi[i]));
              xv[sg * 32 + i] = y; sum += y; ssq = fmaf(y, y, ssq);
            }
          }
          sts1(m_sum, sum); sts1(m_ssq, ssq);
          epi_bar_sync();
          const float os = lds1(x_sum), oq = lds1(x_ssq);
          const float s0 = grp == 0 ? sum : os, s1 = grp == 0 ? os : sum;        // same order in both groups
          const float q0 = grp == 0 ? ssq : oq, q1 = grp == 0 ? oq : ssq;
          const float mu = (s0 + s1) * (1.0f / 128.0f);
          const float rstd = 1.0f / sqrtf(fmaxf((q0 + q1) * (1.0f / 128.0f) - mu * mu, 0.0f) + 1e-3f);
          mbar_wait(u_ready, gph, 14);
          tcgen05_fence_after();
#pragma unroll
          for (int sg = 0; sg < 2; ++sg) {
            uint32_t u[32];
            tmem_ld32(accu + sg * 32, u);
            tmem_wait_ld_dep(u);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              const int k8 = c0 + sg * 32 + c8 * 8;
              const float4 g0 = ldg4(m.gln_g[2], k8 >> 2), g1 = ldg4(m.gln_g[2], (k8 >> 2) + 1);
              const float4 e0 = ldg4(m.gln_b[2], k8 >> 2), e1 = ldg4(m.gln_b[2], (k8 >> 2) + 1);
              const float gs[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
              const float es[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float nv = tanh_fast((xv[sg * 32 + c8 * 8 + i] - mu) * rstd * gs[i] + es[i]);
                const float uu = __uint_as_float(u[c8 * 8 + i]);
                const uint32_t ha = hf_addr + ((k8 + i) * 128 + row) * 4;
                const float hn = uu * nv + (1.0f - uu) * lds1(ha);
                sts1(ha, hn);
                o[i] = hn;
              }
              sw_store8(hs, row, k8, o);
            }
          }
          fence_proxy_async_smem();
          mbar_arrive(hs_full);
          tcgen05_fence_before();
          epi_bar_sync();                                   // both halves have read the gates (and the exchange slots)
          if (ga == (uint32_t)grp) { mbar_arrive(&acc_empty[0]); mbar_arrive(&acc_empty[1]); }
          if (tr) trp[1] = clock64();
          continue;
        }
        if (type <= J_L3) {
          if (ALTS && type == J_L3 && (flags & F_PARK)) continue;             // MMAs only (overlapped schedule)
          if (ALTS && type == J_L2 && (flags & F_KLO)) continue;              // first half of a split layer-2 chunk: MMAs only
          const bool epi_only = ALTS && type == J_L3 && (flags & F_EPI);      // epilogue of the parked layer 3: group 0
          if (epi_only && grp != 0) continue;
          if (ALTS && (flags & F_BUF0)) nbuf = 0;
          int b = grp;
          if (!epi_only) { b = nbuf; nbuf ^= 1; }
          const bool mine = b == grp;
          if (tr && (mine || (type == J_L2 && chain == CH_REW && chunk == 3))) trp[0] = clock64();
          if (type == J_L1) {
            if (!mine) continue;
            if (ALTS && (flags & F_TO_XS)) {
              // reward chain of the overlapped schedule: h1 chunk = ELU(acc + b1) -> bf16 -> XS (the layer-3 MMAs that read
              // XS were issued before this job's MMAs, so the accumulator barrier covers them)
              const float* bias = ch.b1 + chunk * 128;
              float4 x0[8];
              load8(bias, 0, x0);
              acc_wait();
              tcgen05_fence_after();
              tmem_drain_aux(tlane + TM_ACC + b * 128, x0, go_math, acc_release, [&](int s, float4 (&x)[8]) { load8(bias, s, x); },
                             [&](int s, uint32_t (&v)[32], float4 (&x)[8]) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float h[8];
                  h[0] = elu_fast(__uint_as_float(v[8 * i]) + x[2 * i].x); h[1] = elu_fast(__uint_as_float(v[8 * i + 1]) + x[2 * i].y);
                  h[2] = elu_fast(__uint_as_float(v[8 * i + 2]) + x[2 * i].z); h[3] = elu_fast(__uint_as_float(v[8 * i + 3]) + x[2 * i].w);
                  h[4] = elu_fast(__uint_as_float(v[8 * i + 4]) + x[2 * i + 1].x); h[5] = elu_fast(__uint_as_float(v[8 * i + 5]) + x[2 * i + 1].y);
                  h[6] = elu_fast(__uint_as_float(v[8 * i + 6]) + x[2 * i + 1].z); h[7] = elu_fast(__uint_as_float(v[8 * i + 7]) + x[2 * i + 1].w);
                  sw_store8(xs, row, chunk * 128 + s * 32 + i * 8, h);
                }
              });
              fence_proxy_async_smem();
              mbar_arrive(&xs_full[chunk]);
              acc_release_late();
              if (tr) trp[1] = clock64();
              continue;
            }
            if (DRNN && (flags & F_BN)) {
              // prior MLP layer 1: Linear, eval-mode BatchNorm1d (folded: y = acc * scale + shift'), ELU -> TMEM h1
              const float* sc = m.bn_scale + chunk * 128;
              const float* sh = m.bn_shift + chunk * 128;
              const uint32_t dst = tlane + TM_A + chunk * 64;
              float4 x0[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) x0[i] = ldg4(sc, i);
              acc_wait();
              tcgen05_fence_after();
              tmem_drain_aux(tlane + TM_ACC + b * 128, x0, go_math, acc_release, [&](int sg, float4 (&x)[8]) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = ldg4(sc, sg * 8 + i);
              }, [&](int sg, uint32_t (&v)[32], float4 (&x)[8]) {
                uint32_t o[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 t4 = ldg4(sh, sg * 8 + i);
                  o[2 * i] = pack_bf16x2(elu_fast(fmaf(__uint_as_float(v[4 * i]), x[i].x, t4.x)), elu_fast(fmaf(__uint_as_float(v[4 * i + 1]), x[i].y, t4.y)));
                  o[2 * i + 1] = pack_bf16x2(elu_fast(fmaf(__uint_as_float(v[4 * i + 2]), x[i].z, t4.z)), elu_fast(fmaf(__uint_as_float(v[4 * i + 3]), x[i].w, t4.w)));
                }
                tmem_st16(dst + sg * 16, o);
              });
              tmem_wait_st();
              tcgen05_fence_before();
              mbar_arrive(&h1_full[chunk]);
              acc_release_late();
              continue;
            }
            // h1 chunk = ELU(acc + b1) -> bf16x2 -> TMEM A-operand columns
            const float* bias = ch.b1 + chunk * 128;
            const uint32_t dst = tlane + TM_A + chunk * 64;
            float4 x0[8];
            load8(bias, 0, x0);
            acc_wait();
            tcgen05_fence_after();
            tmem_drain_aux(tlane + TM_ACC + b * 128, x0, go_math, acc_release, [&](int s, float4 (&x)[8]) { load8(bias, s, x); },
                           [&](int s, uint32_t (&v)[32], float4 (&x)[8]) {
              uint32_t o[16];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                o[2 * i] = pack_bf16x2(elu_fast(__uint_as_float(v[4 * i]) + x[i].x), elu_fast(__uint_as_float(v[4 * i + 1]) + x[i].y));
                o[2 * i + 1] = pack_bf16x2(elu_fast(__uint_as_float(v[4 * i + 2]) + x[i].z), elu_fast(__uint_as_float(v[4 * i + 3]) + x[i].w));
              }
              tmem_st16(dst + s * 16, o);
            });
            tmem_wait_st();
            tcgen05_fence_before();
            mbar_arrive(&h1_full[chunk]);
            acc_release_late();
            if (tr) trp[1] = clock64();
          } else if (type == J_L2) {
            const bool pidot = DRNN && (flags & F_PIDOT);
            const bool to_xs = chain != CH_REW && !pidot;
            if (mine) {
              const float* bias = ch.b2 + chunk * 128;
              float4 x0[8];
              load8(bias, 0, x0);
              acc_wait();
              tcgen05_fence_after();
              if (pidot) {
                // last policy Linear (A <= 8 rows) as dot products: h2 and W3 rounded to bf16 like the MMA operands would be
                const float* w3q = m.pi_w3q + chunk * 128;
                tmem_drain_aux(tlane + TM_ACC + b * 128, x0, go_math, acc_release, [&](int s, float4 (&x)[8]) { load8(bias, s, x); },
                                [&](int s, uint32_t (&v)[32], float4 (&x)[8]) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float h0 = __bfloat162float(__float2bfloat16(elu_fast(__uint_as_float(v[4 * i]) + x[i].x)));
                    const float h1 = __bfloat162float(__float2bfloat16(elu_fast(__uint_as_float(v[4 * i + 1]) + x[i].y)));
                    const float h2 = __bfloat162float(__float2bfloat16(elu_fast(__uint_as_float(v[4 * i + 2]) + x[i].z)));
                    const float h3 = __bfloat162float(__float2bfloat16(elu_fast(__uint_as_float(v[4 * i + 3]) + x[i].w)));
#pragma unroll
                    for (int a = 0; a < 8; ++a)
                      if (a < A) {
                        const float4 w = ldg4(w3q + a * 512, s * 8 + i);
                        pd[a] = fmaf(h0, w.x, pd[a]); pd[a] = fmaf(h1, w.y, pd[a]); pd[a] = fmaf(h2, w.z, pd[a]); pd[a] = fmaf(h3, w.w, pd[a]);
                      }
                  }
                });
              } else if (to_xs) {
                tmem_drain_aux(tlane + TM_ACC + b * 128, x0, go_math, acc_release, [&](int s, float4 (&x)[8]) { load8(bias, s, x); },
                                [&](int s, uint32_t (&v)[32], float4 (&x)[8]) {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float h[8];
                    h[0] = elu_fast(__uint_as_float(v[8 * i]) + x[2 * i].x); h[1] = elu_fast(__uint_as_float(v[8 * i + 1]) + x[2 * i].y);
                    h[2] = elu_fast(__uint_as_float(v[8 * i + 2]) + x[2 * i].z); h[3] = elu_fast(__uint_as_float(v[8 * i + 3]) + x[2 * i].w);
                    h[4] = elu_fast(__uint_as_float(v[8 * i + 4]) + x[2 * i + 1].x); h[5] = elu_fast(__uint_as_float(v[8 * i + 5]) + x[2 * i + 1].y);
                    h[6] = elu_fast(__uint_as_float(v[8 * i + 6]) + x[2 * i + 1].z); h[7] = elu_fast(__uint_as_float(v[8 * i + 7]) + x[2 * i + 1].w);
                    sw_store8(xs, row, chunk * 128 + s * 32 + i * 8, h);
                  }
                });
                fence_proxy_async_smem();
                mbar_arrive(&xs_full[chunk]);
              } else {
                // reward head (N = 1): fp32 dot product with w3 (helper.py:169-176 last Linear)
                const float* w3 = ch.w3 + chunk * 128;
                tmem_drain_aux(tlane + TM_ACC + b * 128, x0, go_math, acc_release, [&](int s, float4 (&x)[8]) { load8(bias, s, x); },
                                [&](int s, uint32_t (&v)[32], float4 (&x)[8]) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float4 w = ldg4(w3, s * 8 + i);
                    dot = fmaf(elu_fast(__uint_as_float(v[4 * i]) + x[i].x), w.x, dot);
                    dot = fmaf(elu_fast(__uint_as_float(v[4 * i + 1]) + x[i].y), w.y, dot);
                    dot = fmaf(elu_fast(__uint_as_float(v[4 * i + 2]) + x[i].z), w.z, dot);
                    dot = fmaf(elu_fast(__uint_as_float(v[4 * i + 3]) + x[i].w), w.w, dot);
                  }
                });
              }
              acc_release_late();
            }
            if (pidot) {
              if (chunk == 3) {
                // a = clamp(tanh(W3 h2 + b3) + clip(eps * std)): combine the two groups' partial sums in group 0 (qcs is idle
                // outside the Q phase), write the action into [z,a] and the output row  (tdmpc.py:39-45, helper.py:86-96)
#pragma unroll
                for (int a = 0; a < 8; ++a) if (a < A) qcs[(a * 2 + grp) * 128 + row] = pd[a];
                epi_bar_sync();
                if (grp == 0) {
                  const uint64_t nbase = p.noise_base + (uint64_t)t * p.noise_tstride + (uint64_t)pb * p.noise_bstride + (uint64_t)pj * A;
#pragma unroll
                  for (int a = 0; a < 8; ++a)
                    if (a < A) {
                      float x = 0.0f;
                      if (valid) {
                        const float mu = tanhf((pd[a] + qcs[(a * 2 + 1) * 128 + row]) + __ldg(m.chain[CH_PI].b3 + a));
                        x = policy_sample(mu, p.min_std, noise_at(p.noise, nbase + a));
                        pi_row[(size_t)t * A + a] = x;
                      }
                      if (!(flags & F_LAST_STEP)) za_store1(za, row, L + a, x, ckb);
                    }
                  if (!(flags & F_LAST_STEP)) { fence_proxy_async_smem(); mbar_arrive(za_full); }
                }
#pragma unroll
                for (int a = 0; a < 8; ++a) pd[a] = 0.0f;
                epi_bar_sync();                  // qcs may be rewritten by the next exchange
              }
            } else if (!to_xs && chunk == 3) {
              // both groups hold a partial reward dot product of this row: combine in group 0.  (A one-way hand-over -- group 1
              // publishes its partial sum and moves on -- was measured: no gain, group 0's three chunks in a row stay critical.)
              sts1(m_dot, dot);
              epi_bar_sync();
              if (grp == 0) {
                const float r = (dot + lds1(x_dot)) + __ldg(ch.b3);
                G = __fadd_rn(G, __fmul_rn(p.disc[t], r));   // G += discount * reward (tdmpc.py:89): mul, then add, as torch does
                rlast = r;
              }
              dot = 0.0f;
            }
            if (tr && (mine || (!to_xs && chunk == 3))) trp[1] = clock64();
          } else {
            if (!mine) continue;
            if (PAIR && epi_only) {
              // CTA pair, rank 1: refill the [z,a] tile with the row rank 0 published for step t + 1 (last step: z_H and pi(z_H))
              mbar_wait(d3_full, d3_phase, 15); d3_phase ^= 1;          // the reward's layer-1 MMAs have read the old tile
              const bool lastf = (flags & F_LAST_STEP) != 0;
              mb_wait(t + 1);
              if (lastf) mb_wait(TDMPC_MAX_H + 1);
              const float* mr = mb_rows + ((size_t)(t + 1) * 128 + row) * m.K1p;
              for (int k0 = 0; k0 < (ckb >= 0 ? m.K1p : m.nkb1 * 64); k0 += 8) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (k0 + i < L + A) ? __ldcg(mr + k0 + i) : 0.0f;
                za_store8(za, row, k0, v, ckb);
              }
              mb_consume(t + 1, lastf ? TDMPC_MAX_H + 1 : -1);
              fence_proxy_async_smem();
              mbar_arrive(za_full);
              continue;
            }
            const bool pub = PAIR;                                     // CTA pair, rank 0: publish what goes into [z,a]
            float* mrow = pub ? mb_rows + ((size_t)((chain == CH_DYN ? t + 1 : p.H)) * 128 + row) * m.K1p : nullptr;
            // the next step's actions of this row are fetched before waiting for the accumulator
            float an[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) an[a] = 0.0f;
            // more than 8 action dimensions (both dims even): the pair path of the z' epilogue below, lane = action pair
            const bool act2 = WIDE && A > 8 && !((A | L) & 1);
            if (chain == CH_DYN && p.mode == MODE_VALUE && !(flags & F_LAST_STEP) && valid && !act2) {
#pragma unroll
              for (int a = 0; a < 8; ++a) if (a < A) an[a] = __ldg(act_row + (size_t)(t + 1) * A + a);
            }
            // policy head: the noise of the first four action dimensions does not depend on the accumulator either (one
            // Philox block when the stream offset is 4-aligned, e.g. A = 4)
            const bool terminal = (flags & F_TERMINAL) != 0;
            const uint64_t nbase = p.noise_base + (terminal ? 0 : (uint64_t)t * p.noise_tstride) + (uint64_t)pb * p.noise_bstride + (uint64_t)pj * A;
            // Any stream offset: element nbase + a sits in Philox block (nbase + a) >> 2, lane (nbase + a) & 3, so the four actions
            // a0..a0+3 take lanes r.. of block nbase/4 + a0/4 and lanes ..r-1 of the next one (r = nbase & 3; A = 38 gives r = 0 / 2
            // on alternate rows).  One new block per group of four, carried over -- not one block + Box-Muller per element.
            const bool nz4 = chain == CH_PI && !p.noise.inj && p.min_std > 0.0f;
            const int nzr = (int)(nbase & 3);
            float zn[4] = {0.f, 0.f, 0.f, 0.f}, zx[4] = {0.f, 0.f, 0.f, 0.f};
            if (nz4 && valid) philox_normal4(p.noise, nbase >> 2, zn);
            if (epi_only) { mbar_wait(d3_full, d3_phase, 15); d3_phase ^= 1; }
            else acc_wait();
            tcgen05_fence_after();
#ifdef TDMPC_TRACE_L3
            if (tr) trp[0] = clock64();
#endif
            const uint32_t l3src = epi_only ? tlane + TM_A : tlane + TM_ACC + b * 128;
            if (chain == CH_DYN) {
              // z' = acc + b3 -> [z', a_{t+1}, 0...] row of the next step.  Columns >= L+A of the row are
              // zero padding written once per tile and never touched again.
              const bool next_act = p.mode == MODE_VALUE && !(flags & F_LAST_STEP);
              for (int k0 = 0; k0 < L; k0 += 32) {
                uint32_t v[32];
                float4 bz[8];
                tmem_ld32(l3src + k0, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) bz[i] = (k0 + 4 * i < L) ? ldg4(ch.b3, (k0 >> 2) + i) : make_float4(0.f, 0.f, 0.f, 0.f);   // b3 is zero padded to a multiple of 4
                tmem_wait_ld_dep(v);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                  const int kk = k0 + c8 * 8;
                  if (kk < L) {
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      const float4 bb = bz[c8 * 2 + (i >> 2)];
                      o[i] = __uint_as_float(v[c8 * 8 + i]) + ((i & 3) == 0 ? bb.x : (i & 3) == 1 ? bb.y : (i & 3) == 2 ? bb.z : bb.w);
                    }
                    if (pub) {
#pragma unroll
                      for (int i = 0; i < 8; ++i) if (kk + i < L) mrow[kk + i] = o[i];
                    }
                    if (kk + 8 <= L) za_store8(za, row, kk, o, ckb);
                    else {
                      // the 8-column group that also holds action columns: z elements only
#pragma unroll
                      for (int i = 0; i < 8; ++i) if (kk + i < L) za_store1(za, row, kk + i, o[i], ckb);
                    }
                  }
                }
              }
#if defined(TDMPC_TRACE_L3) && TDMPC_TRACE_L3 >= 2
              if (tr) trp[0] = clock64();
#endif
              if (next_act && act2) {
                // Dog-run dims (A = 38).  With lane = row every load instruction touches 32 cache lines and every bf16 store
                // recomputes a swizzled address: 14 k cycles per step on one warp per scheduler.  Here lane = action PAIR: one
                // 8-byte load per row fetches the row's next actions as contiguous bytes (sixteen rows in flight), and the
                // bf16x2 store address is a per-lane constant plus a per-row constant (the row's low bits are compile-time).
                const int a2 = 2 * lane, k = L + a2, kb = k >> 6, kk = k & 63;
                const bool on = a2 < A, cmp = kb == ckb;
                const uint32_t lane_part = smem_u32(za) + (uint32_t)kb * TILE_BYTES + (uint32_t)(kk & 7) * 2u + (cmp ? (uint32_t)(kk >> 3) * 128u : 0u);
                const uint32_t xk = cmp ? 0u : (uint32_t)(kk >> 3) << 4, xmask = cmp ? 0u : 0x70u;
                const uint32_t P1 = cmp ? 256u : 1024u, P2 = cmp ? 16u : 128u;       // tc_prims.cuh: k16_offset / sw128_offset
                const uint32_t wpart = lane_part + (uint32_t)(q * 4) * P1;
                // Every load is UNCONDITIONAL (rows past the end of the launch read row 0, lanes past the last pair read pair 0):
                // a `cond ? __ldg(p) : 0` compiles to a predicated load plus a move of its result, and that move makes the warp
                // wait for each load before it issues the next one -- sixteen memory round trips instead of one (seen in the SASS).
                const unsigned long long mine_p = (unsigned long long)(uintptr_t)(valid ? act_row + (size_t)(t + 1) * A : p.actions);
                const int lsel = on ? lane : 0;
#pragma unroll 1
                for (int r0 = 0; r0 < 32; r0 += 16) {
                  float2 av[16];
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const unsigned long long rp = __shfl_sync(0xffffffffu, mine_p, r0 + i);
                    av[i] = __ldg(reinterpret_cast<const float2*>((uintptr_t)rp) + lsel);
                  }
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (on) {
                      const uint32_t addr = wpart + (uint32_t)((r0 + i) >> 3) * P1 + (uint32_t)(i & 7) * P2 + ((xk ^ ((uint32_t)(i & 7) << 4)) & xmask);
                      asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(pack_bf16x2(av[i].x, av[i].y)) : "memory");
                      if (pub) *reinterpret_cast<float2*>(mb_rows + ((size_t)(t + 1) * 128 + q * 32 + r0 + i) * m.K1p + L + a2) = av[i];
                    }
#if defined(TDMPC_TRACE_L3) && TDMPC_TRACE_L3 == 3
                  if (tr && r0 == 0) ktrace[((1 + (grp ^ 1)) * MAX_JOBS + j) * 2] = clock64();
#endif
                }
#if defined(TDMPC_TRACE_L3) && TDMPC_TRACE_L3 == 3
                if (tr) ktrace[((1 + (grp ^ 1)) * MAX_JOBS + j) * 2 + 1] = clock64();
#endif
              } else if (next_act) {
#pragma unroll
                for (int a = 0; a < 8; ++a) if (a < A) { za_store1(za, row, L + a, an[a], ckb); if (pub) mrow[L + a] = an[a]; }
                // action dimensions 8 .. A-1 (dog-run: 38).  (Fetching them before the accumulator wait like an[] was measured:
                // 32 more live registers spill, +6 %.)
#pragma unroll 1
                for (int a0 = 8; WIDE && a0 < A; a0 += 8) {  // odd dims only (the pair path above takes the even ones)
                  float av[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) av[i] = __ldg((valid ? act_row + (size_t)(t + 1) * A : p.actions) + (a0 + i < A ? a0 + i : 0));   // unconditional
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    if (a0 + i < A) { za_store1(za, row, L + a0 + i, valid ? av[i] : 0.0f, ckb); if (pub) mrow[L + a0 + i] = valid ? av[i] : 0.0f; }
                }
              }
            } else {
              // a = clamp(tanh(acc + b3) + clip(eps * std)), written to the action columns of [z,a]
#pragma unroll 1
              for (int a0 = 0; a0 < A; a0 += 4) {
                uint32_t v[4];
                tmem_ld4(tlane + TM_ACC + b * 128 + a0, v);
                if (nz4 && valid && (nzr != 0 || a0 + 4 < A)) philox_normal4(p.noise, (nbase >> 2) + (a0 >> 2) + 1, zx);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int a = a0 + i;
                  if (a < A) {
                    float x = 0.0f;
                    if (valid) {
                      const float mu = tanhf(__uint_as_float(v[i]) + __ldg(ch.b3 + a));
                      // lane nzr + i of the 8 normals {zn, zx} (compile-time i, run-time nzr: selects, no local array)
                      const float w0 = i == 0 ? zn[0] : i == 1 ? zn[1] : i == 2 ? zn[2] : zn[3];
                      const float w1 = i == 0 ? zn[1] : i == 1 ? zn[2] : i == 2 ? zn[3] : zx[0];
                      const float w2 = i == 0 ? zn[2] : i == 1 ? zn[3] : i == 2 ? zx[0] : zx[1];
                      const float w3 = i == 0 ? zn[3] : i == 1 ? zx[0] : i == 2 ? zx[1] : zx[2];
                      const float zsel = nzr == 0 ? w0 : nzr == 1 ? w1 : nzr == 2 ? w2 : w3;
                      x = policy_sample(mu, p.min_std, nz4 ? zsel : noise_at(p.noise, nbase + a));
                      if (!terminal) pi_row[(size_t)t * A + a] = x;
                    }
                    if (!(flags & F_LAST_STEP)) za_store1(za, row, L + a, x, ckb);
                    if (pub && terminal) mrow[L + a] = x;
                  }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) zn[i] = zx[i];
              }
            }
            tcgen05_fence_before();
            if (!epi_only) mbar_arrive(&acc_empty[b]);
            if (!(chain == CH_PI && (flags & F_LAST_STEP))) {
              fence_proxy_async_smem();
              mbar_arrive(za_full);
            }
            // CTA pair: raise the mailbox flag AFTER my own MMA warp has been released (the fence costs ~1 us; rank 1 has slack)
            if (pub && (chain == CH_DYN || terminal)) mb_publish(chain == CH_DYN ? t + 1 : TDMPC_MAX_H + 1);
            if (tr) trp[1] = clock64();
          }
        } else {
          if (tr) trp[0] = clock64();
          // Q head layer: x = acc + b over all 512 columns, LayerNorm, then
          //   J_QL1: tanh -> bf16 -> XS        J_QL2: ELU -> dot with w3 -> q
          // Group g handles chunks g and g+2; statistics are exchanged between the groups.
          const bool first = type == J_QL1;
          // Stage the layer's per-column constants in shared memory: every thread needs all of them, and with 227 KB of
          // the SM given to shared memory the L1 is too small to keep the 8 KB vectors of all chains resident.
          const uint32_t qb = smem_u32(qcs);                                   // bias            (shared-space addresses)
          const uint32_t qg = qb + 2048, qe = qb + 4096, qw = qb + 6144;       // LN gamma, LN beta, head weights w3
          epi_bar_sync();                                                      // the previous layer's readers are done
          {
            const float* srcs[4] = {first ? ch.b1 : ch.b2, ch.ln_g[first ? 0 : 1], ch.ln_b[first ? 0 : 1], ch.w3};
            float4* dstc = reinterpret_cast<float4*>(qcs);
            const int te = (int)threadIdx.x - 128;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int idx = te + k * 256;
              dstc[idx] = ldg4(srcs[idx >> 7], idx & 127);
            }
          }
          epi_bar_sync();
          // Last job of a tile (F_TILE_END): the [z,a] operand was last read by the layer-1 MMAs of this Q head, which both groups
          // have seen complete -- group 0 writes the NEXT tile's rows now, while this layer's MMAs run.
          const bool tile_end = !first && (flags & F_TILE_END);
          const bool early_rel = tile_end && !DRNN;      // (the GRU kernel's next job parks gate products in the h1 region)
          if (tile_end && grp == 0) {
            const int nt = peek_tile();
            if (nt >= 0) { fill_za(nt); prefilled = true; }
          }
          float sum = 0.0f, ssq = 0.0f;
          for (int c = grp; c < 4; c += 2) {
            float4 x0[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x0[i] = lds4(qb + (c * 32 + i) * 16);
            mbar_wait(&q_full[c], q_phase, 8);
            tcgen05_fence_after();
            tmem_stream_aux(tlane + c * 128, 4, x0, [&](int s, float4 (&x)[8]) {
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = lds4(qb + ((c * 4 + s) * 8 + i) * 16);
            },
                            [&](int s, uint32_t (&v)[32], float4 (&x)[8]) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float y0 = __uint_as_float(v[4 * i]) + x[i].x, y1 = __uint_as_float(v[4 * i + 1]) + x[i].y;
                const float y2 = __uint_as_float(v[4 * i + 2]) + x[i].z, y3 = __uint_as_float(v[4 * i + 3]) + x[i].w;
                sum += (y0 + y1) + (y2 + y3);
                ssq = fmaf(y0, y0, ssq); ssq = fmaf(y1, y1, ssq); ssq = fmaf(y2, y2, ssq); ssq = fmaf(y3, y3, ssq);
              }
            });
          }
          q_phase ^= 1;
          sts1(m_sum, sum); sts1(m_ssq, ssq);
          epi_bar_sync();
          sum += lds1(x_sum); ssq += lds1(x_ssq);
          const float mu = sum * (1.0f / 512.0f);
          const float var = fmaxf(ssq * (1.0f / 512.0f) - mu * mu, 0.0f);      // biased variance, as nn.LayerNorm
          const float rstd = 1.0f / sqrtf(var + 1e-5f);
          const float nmr = -mu * rstd;
          float qdot = 0.0f;
          for (int c = grp; c < 4; c += 2) {
            tmem_stream(tlane + c * 128, 4, [&](int s, uint32_t (&v)[32]) {
              float y[32];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int vi = (c * 4 + s) * 8 + i;
                const float4 x = lds4(qb + vi * 16), gg = lds4(qg + vi * 16), bb = lds4(qe + vi * 16);
                y[4 * i] = fmaf(fmaf(__uint_as_float(v[4 * i]) + x.x, rstd, nmr), gg.x, bb.x);          // ((v + b) - mu) * rstd * g + beta
                y[4 * i + 1] = fmaf(fmaf(__uint_as_float(v[4 * i + 1]) + x.y, rstd, nmr), gg.y, bb.y);
                y[4 * i + 2] = fmaf(fmaf(__uint_as_float(v[4 * i + 2]) + x.z, rstd, nmr), gg.z, bb.z);
                y[4 * i + 3] = fmaf(fmaf(__uint_as_float(v[4 * i + 3]) + x.w, rstd, nmr), gg.w, bb.w);
              }
              if (first) {
#pragma unroll
                for (int i = 0; i < 32; ++i) y[i] = tanh_bf16<EXACT>(y[i]);
#pragma unroll
                for (int i = 0; i < 4; ++i) sw_store8(xs, row, c * 128 + s * 32 + i * 8, y + i * 8);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 f = lds4(qw + ((c * 4 + s) * 8 + i) * 16);
                  qdot = fmaf(elu_fast(y[4 * i]), f.x, qdot); qdot = fmaf(elu_fast(y[4 * i + 1]), f.y, qdot);
                  qdot = fmaf(elu_fast(y[4 * i + 2]), f.z, qdot); qdot = fmaf(elu_fast(y[4 * i + 3]), f.w, qdot);
                }
              }
            });
            if (first) { fence_proxy_async_smem(); mbar_arrive(&xs_full[c]); }
            // F_TILE_END: my accumulator buffer (chunk c = grp) goes back as soon as it is read.  The next job is the next
            // tile's first layer 1, which writes buffers 0 / 1 only; the h1 region (chunks 2 - 3) is not written before both
            // groups have passed the barrier below.
            if (early_rel && c == grp) { tcgen05_fence_before(); mbar_arrive(&acc_empty[grp]); }
          }
          if (!first) {
            sts1(m_dot, qdot);
            epi_bar_sync();
            if (grp == 0) {
              const float qv = (qdot + lds1(x_dot)) + __ldg(ch.b3);
              if (PAIR) {
                if (chain == CH_Q1) { mb_q1[row] = qv; mb_publish(TDMPC_MAX_H + 2); }           // rank 0 -> rank 1
                else { mb_wait(TDMPC_MAX_H + 2); q1v = __ldcg(mb_q1 + row); mb_consume(TDMPC_MAX_H + 2, -1); }
              }
              if (chain == CH_Q1) q1v = qv;
              else if (valid) {
                // G += discount * min(Q1, Q2); nan_to_num_   (tdmpc.py:91-92)
                const float val = __fadd_rn(G, __fmul_rn(p.disc[p.H], min_nanprop(q1v, qv)));
                p.values[(size_t)pb * p.Tcap + pj] = nan_to_num0(val);
                if (p.last_reward) p.last_reward[(size_t)pb * p.Tcap + pj] = rlast;
              }
            }
          }
          if (!early_rel) { tcgen05_fence_before(); mbar_arrive(&acc_empty[grp]); }
          if (tr) trp[1] = clock64();
        }
      }
      if (ktrace && blockIdx.x == 0 && threadIdx.x == 128 && it < 128) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
        ktrace[3 * MAX_JOBS * 2 + it] = clock64();
        ktrace[3 * MAX_JOBS * 2 + 128 + it] = ns;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (ktrace && threadIdx.x == 128 && blockIdx.x < 256) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* bs = ktrace + 3 * MAX_JOBS * 2 + 256 + 4 * blockIdx.x;
    bs[0] = (unsigned long long)t_start; bs[1] = (unsigned long long)clock64(); bs[2] = (unsigned long long)tiles_done; bs[3] = smid;
  }
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// Weight packing: fp32 nn.Linear [N][K] -> bf16 tiles [tile_rows x 64], 128B-swizzled, chunk-major.
__global__ void pack_tiles_kernel(const float* __restrict__ W, int N, int K, uint8_t* __restrict__ dst,
                                  int n_chunks, int nkb, int tile_rows) {
  // one thread per 16-byte group (8 consecutive k of one row)
  const long long total = (long long)n_chunks * nkb * tile_rows * 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int g = (int)(i & 7);
  const int r = (int)((i >> 3) % tile_rows);
  const long long ckb = (i >> 3) / tile_rows;
  const int kb = (int)(ckb % nkb), c = (int)(ckb / nkb);
  const int n = c * 128 + r;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kb * 64 + g * 8 + e;
    v[e] = (n < N && k < K) ? W[(size_t)n * K + k] : 0.0f;
  }
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  const size_t tile_off = (size_t)(c * nkb + kb) * tile_rows * 128;
  *reinterpret_cast<uint4*>(dst + tile_off + sw128_offset(r, g * 8)) = u;
}

__global__ void round_bf16_kernel(const float* w, int n, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(__float2bfloat16(w[i]));
}

__global__ void fold_bias_kernel(const float* b, const float* scale, const float* shift, int n, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fmaf(b[i], scale[i], shift[i]);
}

int launch_pack(const float* W, int N, int K, uint8_t* dst, int n_chunks, int nkb, int tile_rows, cudaStream_t st) {
  const long long total = (long long)n_chunks * nkb * tile_rows * 8;
  pack_tiles_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W, N, K, dst, n_chunks, nkb, tile_rows);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

std::vector<uint32_t> build_jobs(int mode, int H, bool drnn = false, int A = 0, int* pair_n0 = nullptr) {
  std::vector<uint32_t> j;
  if (mode == 2) {
    // Value mode over a CTA PAIR (launches of a few tiles: single-environment plans, where one tile's serial chain through H + 2
    // networks is the whole latency).  The reward chain never feeds back into z, so rank 0 runs the dynamics chain, the terminal
    // policy and Q1 and publishes every [z_t | a_t] row through a global-memory mailbox; rank 1 follows one step behind with the
    // reward chain (it owns the discounted return) and Q2, and forms the value.  Rank 1's [z,a] tile is refilled from the mailbox
    // by an epilogue-only job (J_L3 | F_EPI) once the MMA warp's d3_full commit says the reward's layer 1 has read the old one.
    if (drnn) return j;
    auto c12 = [&](int type, int chain, int t, int first_flags, int last_flags = 0) {
      for (int c = 0; c < 4; ++c) j.push_back(tc_job(type, chain, c, t, (c == 0 ? first_flags : 0) | (c == 3 ? last_flags : 0)));
    };
    for (int t = 0; t < H; ++t) {
      c12(J_L1, CH_DYN, t, F_WAIT_ZA);
      c12(J_L2, CH_DYN, t, 0);
      j.push_back(tc_job(J_L3, CH_DYN, 0, t, t == H - 1 ? F_LAST_STEP : 0));
    }
    c12(J_L1, CH_PI, 0, F_WAIT_ZA);
    c12(J_L2, CH_PI, 0, 0);
    j.push_back(tc_job(J_L3, CH_PI, 0, 0, F_TERMINAL));
    j.push_back(tc_job(J_QL1, CH_Q1, 0, 0, F_WAIT_ZA));
    j.push_back(tc_job(J_QL2, CH_Q1, 0, 0, 0));
    if (pair_n0) *pair_n0 = (int)j.size();
    for (int t = 0; t < H; ++t) {
      c12(J_L1, CH_REW, t, F_WAIT_ZA, F_D3);
      c12(J_L2, CH_REW, t, 0);
      j.push_back(tc_job(J_L3, CH_DYN, 0, t, F_EPI | (t == H - 1 ? F_LAST_STEP : 0)));     // refill [z_{t+1} | a_{t+1}] (last: [z_H | pi(z_H)])
    }
    j.push_back(tc_job(J_QL1, CH_Q2, 0, 0, F_WAIT_ZA));
    j.push_back(tc_job(J_QL2, CH_Q2, 0, 0, 0));
    return j;
  }
  auto chain12 = [&](int type, int chain, int t, int first_flags) {
    for (int c = 0; c < 4; ++c) j.push_back(tc_job(type, chain, c, t, c == 0 ? first_flags : 0));
  };
  // layer 2 with chunks 0 and 1 split over k (tc_model.cuh, F_KLO): KLO c0, KLO c1, KHI c0, KHI c1, c2, c3
  const bool split = !drnn && getenv("TDMPC_TC_SPLIT");      // measured slower than whole chunks: off
  auto layer2 = [&](int chain, int t) {
    if (!split) { chain12(J_L2, chain, t, 0); return; }
    j.push_back(tc_job(J_L2, chain, 0, t, F_KLO)); j.push_back(tc_job(J_L2, chain, 1, t, F_KLO));
    j.push_back(tc_job(J_L2, chain, 0, t, F_KHI)); j.push_back(tc_job(J_L2, chain, 1, t, F_KHI));
    j.push_back(tc_job(J_L2, chain, 2, t, 0)); j.push_back(tc_job(J_L2, chain, 3, t, 0));
  };
  if (drnn) {
    // DSSM.next with GRU dynamics (tdmpc_icem_similarity_drnn.py:56-59): gates, h', prior MLP -> z', reward MLP(h').
    // Value mode only: the (short) policy rollout and the single post-plan step run on the fp32 kernel.
    auto gru_prior = [&](int t, bool last) {
      j.push_back(tc_job(J_G, 0, 0, t, F_WAIT_ZA | (t == 0 ? F_WAIT_HS : 0)));
      j.push_back(tc_job(J_G, 0, 1, t, 0));
      j.push_back(tc_job(J_GH, 0, 0, t, 0));
      j.push_back(tc_job(J_GI, 0, 0, t, 0));
      for (int c = 0; c < 4; ++c) j.push_back(tc_job(J_L1, CH_DYN, c, t, F_A_HS | F_BN | (c == 0 ? F_WAIT_HS : 0)));
      j.push_back(tc_job(J_L3, CH_DYN, 0, t, F_A_TMEM | (last ? F_LAST_STEP : 0)));
    };
    if (mode != MODE_VALUE) {
      // policy rollout (tdmpc_icem_similarity_drnn.py:196-200): a_t = pi(z), then the GRU step.  Needs the dot-product
      // policy head (A <= 8); otherwise the list stays empty and the fp32 kernel runs this (short) part.
      if (A > 8) return j;
      for (int t = 0; t < H; ++t) {
        const bool last = t == H - 1;
        chain12(J_L1, CH_PI, t, F_WAIT_ZA);
        for (int c = 0; c < 4; ++c) j.push_back(tc_job(J_L2, CH_PI, c, t, F_PIDOT | (last ? F_LAST_STEP : 0)));
        if (!last) gru_prior(t, false);
      }
      return j;
    }
    for (int t = 0; t < H; ++t) {
      j.push_back(tc_job(J_G, 0, 0, t, F_WAIT_ZA | (t == 0 ? F_WAIT_HS : 0)));
      j.push_back(tc_job(J_G, 0, 1, t, 0));
      j.push_back(tc_job(J_GH, 0, 0, t, 0));
      j.push_back(tc_job(J_GI, 0, 0, t, 0));
      for (int c = 0; c < 4; ++c) j.push_back(tc_job(J_L1, CH_DYN, c, t, F_A_HS | F_BN | (c == 0 ? F_WAIT_HS : 0)));
      j.push_back(tc_job(J_L3, CH_DYN, 0, t, F_A_TMEM | (t == H - 1 ? F_LAST_STEP : 0)));
      for (int c = 0; c < 4; ++c) j.push_back(tc_job(J_L1, CH_REW, c, t, F_A_HS));
      chain12(J_L2, CH_REW, t, 0);
    }
    chain12(J_L1, CH_PI, 0, F_WAIT_ZA);
    chain12(J_L2, CH_PI, 0, 0);
    j.push_back(tc_job(J_L3, CH_PI, 0, 0, F_TERMINAL));
    j.push_back(tc_job(J_QL1, CH_Q1, 0, 0, F_WAIT_ZA));
    j.push_back(tc_job(J_QL2, CH_Q1, 0, 0, 0));
    j.push_back(tc_job(J_QL1, CH_Q2, 0, 0, 0));
    j.push_back(tc_job(J_QL2, CH_Q2, 0, 0, F_TILE_END));
    return j;
  }
  if (mode == MODE_VALUE && !getenv("TDMPC_TC_OVERLAP")) {
    // the plain schedule (default): dynamics and reward alternate on the one TMEM h1 buffer
    for (int t = 0; t < H; ++t) {
      chain12(J_L1, CH_DYN, t, F_WAIT_ZA);
      layer2(CH_DYN, t);
      chain12(J_L1, CH_REW, t, 0);
      j.push_back(tc_job(J_L3, CH_DYN, 0, t, t == H - 1 ? F_LAST_STEP : 0));
      layer2(CH_REW, t);
    }
    chain12(J_L1, CH_PI, 0, F_WAIT_ZA);
    layer2(CH_PI, 0);
    j.push_back(tc_job(J_L3, CH_PI, 0, 0, F_TERMINAL));
    j.push_back(tc_job(J_QL1, CH_Q1, 0, 0, F_WAIT_ZA));
    j.push_back(tc_job(J_QL2, CH_Q1, 0, 0, 0));
    j.push_back(tc_job(J_QL1, CH_Q2, 0, 0, 0));
    j.push_back(tc_job(J_QL2, CH_Q2, 0, 0, F_TILE_END));
  } else if (mode == MODE_VALUE) {
    // Overlapped schedule (TDMPC_TC_OVERLAP=1; measured slower).  The two MLPs of a step share one TMEM h1
    // buffer, which forces dyn and reward to alternate and leaves the tensor pipe idle during every layer-1 epilogue.  Here
    // the reward chain keeps ITS hidden activation in XS (free once the dynamics' layer 3 has read h2), so reward layer 2
    // of step t (SS mode) runs interleaved with layer 1 of the next consumer of z' (dynamics of step t+1, or the terminal
    // policy; its epilogue fills the TMEM h1 buffer).  The dynamics' layer-3 accumulator is parked in the h1 region so that
    // its epilogue (z' -> [z,a]) can wait until the reward's layer-1 MMAs have read the old [z,a].
    chain12(J_L1, CH_DYN, 0, F_WAIT_ZA);
    for (int t = 0; t < H; ++t) {
      const bool last = t == H - 1;
      chain12(J_L2, CH_DYN, t, 0);
      j.push_back(tc_job(J_L3, CH_DYN, 0, t, F_PARK));
      // d3_full: committed behind the last reward layer-1 job -- the parked layer 3 and everything that reads the old
      // [z,a] have then completed; z' -> [z,a] (F_EPI, epilogue group 0) follows
      for (int c = 0; c < 4; ++c) j.push_back(tc_job(J_L1, CH_REW, c, t, F_TO_XS | (c == 0 ? F_BUF0 : 0) | (c == 3 ? F_D3 : 0)));
      j.push_back(tc_job(J_L3, CH_DYN, 0, t, F_EPI | (last ? F_LAST_STEP : 0)));
      // reward layer 2 (R) interleaved with the next layer 1 (N): R0 R1 N0 N1 R2 N2 R3 N3 on alternating accumulators, so
      // no job waits for the drain of its direct predecessor and the last layer-1 epilogue hides under R3's MMAs
      const int nch = last ? CH_PI : CH_DYN, nt = last ? 0 : t + 1;
      auto R = [&](int c) { j.push_back(tc_job(J_L2, CH_REW, c, t, F_A_XS | (c == 0 ? F_BUF0 : 0))); };
      auto N = [&](int c) { j.push_back(tc_job(J_L1, nch, c, nt, c == 0 ? F_WAIT_ZA : 0)); };
      R(0); R(1); N(0); N(1); R(2); N(2); R(3); N(3);
    }
    chain12(J_L2, CH_PI, 0, 0);
    j.push_back(tc_job(J_L3, CH_PI, 0, 0, F_TERMINAL));
    j.push_back(tc_job(J_QL1, CH_Q1, 0, 0, F_WAIT_ZA));
    j.push_back(tc_job(J_QL2, CH_Q1, 0, 0, 0));
    j.push_back(tc_job(J_QL1, CH_Q2, 0, 0, 0));
    j.push_back(tc_job(J_QL2, CH_Q2, 0, 0, 0));
  } else {
    for (int t = 0; t < H; ++t) {
      const bool last = t == H - 1;
      chain12(J_L1, CH_PI, t, F_WAIT_ZA);
      layer2(CH_PI, t);
      j.push_back(tc_job(J_L3, CH_PI, 0, t, last ? F_LAST_STEP : 0));
      if (!last) {                      // the reference also steps z after the last action; its result is unused
        chain12(J_L1, CH_DYN, t, F_WAIT_ZA);
        layer2(CH_DYN, t);
        j.push_back(tc_job(J_L3, CH_DYN, 0, t, 0));
      }
    }
  }
  return j;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
const char* tc_supported(const tdmpc_model_cfg& c) {
  if (c.mlp_dim != 512) return "mlp_dim must be 512";
  if (c.dynamics == TDMPC_DYN_DRNN && c.hidden_dim != 128) return "DRNN: hidden_dim must be 128";
  if (c.latent_dim > 128) return "latent_dim must be <= 128";
  if (c.latent_dim + c.action_dim > 192) return "latent_dim + action_dim must be <= 192";
  return nullptr;
}

static int gi_pad_tiles(int nkb1) { return (4 - (3 * nkb1 + 6) % 4) % 4; }     // GRU tiles per step: 2 (nkb1 + 2) + 2 + nkb1

static int tc_tiles_total(const tdmpc_model_cfg& c, int* nkb1_out, int l3rows[CH_COUNT]) {
  const int K1p = (c.latent_dim + c.action_dim + 15) & ~15;
  const int nkb1 = (K1p + 63) / 64;
  *nkb1_out = nkb1;
  const int n3_dyn = (c.latent_dim + 15) & ~15, n3_pi = (c.action_dim + 15) & ~15;
  for (int i = 0; i < CH_COUNT; ++i) l3rows[i] = 0;
  l3rows[CH_DYN] = n3_dyn <= 64 ? 64 : 128;
  l3rows[CH_PI] = n3_pi <= 64 ? 64 : 128;
  const bool drnn = c.dynamics == TDMPC_DYN_DRNN;
  int tiles = 0;
  for (int ch = 0; ch < CH_COUNT; ++ch) {
    const bool hs_in = drnn && (ch == CH_DYN || ch == CH_REW);         // layer 1 reads the hidden state (K = 128)
    tiles += 4 * (hs_in ? 2 : nkb1);
    if (!(drnn && ch == CH_DYN)) tiles += 32;                           // the prior MLP has no 512 x 512 layer
    if (l3rows[ch]) tiles += l3rows[ch] == 128 ? 8 : 4;
  }
  if (drnn) tiles += 2 * (nkb1 + 2) + 2 + nkb1 + gi_pad_tiles(nkb1);
  return tiles;
}

static size_t jobs_bytes_total() {
  size_t n = 0;
  for (int mode = 0; mode < 3; ++mode)
    for (int H = 1; H <= TDMPC_MAX_H; ++H) n += (build_jobs(mode, H).size() + 3) & ~(size_t)3;
  return n * sizeof(uint32_t);
}

size_t tc_arena_bytes(const tdmpc_model_cfg& c) {
  int nkb1, l3rows[CH_COUNT];
  return (size_t)tc_tiles_total(c, &nkb1, l3rows) * TILE_BYTES + jobs_bytes_total() + (512 + 8 * 512) * sizeof(float) + 1024;
}

int tc_pack(const tdmpc_model_cfg& c, const tdmpc_weights& w, const ModelFp32& m32, void* arena, TcModel* out, cudaStream_t st) {
  TcModel m{};
  const int L = c.latent_dim, A = c.action_dim;
  m.L = L; m.A = A;
  m.K1p = (L + A + 15) & ~15;
  int nkb1;
  const int tiles = tc_tiles_total(c, &nkb1, m.l3rows);
  m.nkb1 = nkb1;
  m.n3p[CH_DYN] = (L + 15) & ~15;
  m.n3p[CH_PI] = (A + 15) & ~15;
  m.tiles = reinterpret_cast<const uint8_t*>(arena);
  uint8_t* base = reinterpret_cast<uint8_t*>(arena);
  if (cudaMemsetAsync(base, 0, (size_t)tiles * TILE_BYTES, st) != cudaSuccess) return -1;

  const bool drnn = c.dynamics == TDMPC_DYN_DRNN;
  const int Hd = 128;
  const float* W1[CH_COUNT] = {drnn ? w.prior_w[0] : w.dyn_w[0], w.rew_w[0], w.pi_w[0], w.q_w[0][0], w.q_w[1][0]};
  const float* W2[CH_COUNT] = {drnn ? nullptr : w.dyn_w[1], w.rew_w[1], w.pi_w[1], w.q_w[0][1], w.q_w[1][1]};
  const float* W3[CH_COUNT] = {drnn ? w.prior_w[1] : w.dyn_w[2], nullptr, w.pi_w[2], nullptr, nullptr};
  const int K1[CH_COUNT] = {drnn ? Hd : L + A, drnn ? Hd : L + A, L, L + A, L + A};
  const int N3[CH_COUNT] = {L, 0, A, 0, 0};
  const LinearT* lin[CH_COUNT] = {m32.dyn, m32.rew, m32.pi, m32.q[0], m32.q[1]};

  int launches = 0;
  uint32_t tile = 0;
  for (int ch = 0; ch < CH_COUNT; ++ch) {
    TcChain& t = m.chain[ch];
    const int nkb = (drnn && (ch == CH_DYN || ch == CH_REW)) ? 2 : nkb1;
    t.l1_tile = tile;
    if (launch_pack(W1[ch], 512, K1[ch], base + (size_t)tile * TILE_BYTES, 4, nkb, 128, st) < 0) return -1;
    tile += 4 * nkb; ++launches;
    t.l2_tile = tile;
    if (W2[ch]) {
      if (launch_pack(W2[ch], 512, 512, base + (size_t)tile * TILE_BYTES, 4, 8, 128, st) < 0) return -1;
      tile += 32; ++launches;
    }
    t.l3_tile = tile;
    if (W3[ch]) {
      if (launch_pack(W3[ch], N3[ch], 512, base + (size_t)tile * TILE_BYTES, 1, 8, m.l3rows[ch], st) < 0) return -1;
      tile += m.l3rows[ch] == 128 ? 8 : 4; ++launches;
    }
    if (drnn && ch == CH_DYN) {
      t.b1 = nullptr; t.b2 = nullptr; t.b3 = m32.prior[1].b; t.w3 = nullptr;      // layer-1 bias lives in bn_shift
    } else {
      t.b1 = lin[ch][0].b; t.b2 = lin[ch][1].b; t.b3 = lin[ch][2].b;
      t.w3 = N3[ch] == 0 ? lin[ch][2].wt : nullptr;      // [512][1] transposed == the weight row itself
    }
    if (ch >= CH_Q1) {
      for (int i = 0; i < 2; ++i) { t.ln_g[i] = m32.q_ln[ch - CH_Q1][i].g; t.ln_b[i] = m32.q_ln[ch - CH_Q1][i].b; }
    }
  }
  m.drnn = drnn ? 1 : 0;
  if (drnn) {
    // gate weights, one 128-row chunk per gate (reset, update, newval), in the order the jobs consume them:
    //   J_G(g): W_ih[g] then W_hh[g];  J_GH: W_hh[newval];  J_GI: W_ih[newval] (+ padding tiles)
    const int KI = L + A;
    for (int g = 0; g < 2; ++g) {
      m.g_tile[g] = tile;
      if (launch_pack(w.gru_w_ih + (size_t)g * Hd * KI, Hd, KI, base + (size_t)tile * TILE_BYTES, 1, nkb1, 128, st) < 0) return -1;
      tile += nkb1;
      if (launch_pack(w.gru_w_hh + (size_t)g * Hd * Hd, Hd, Hd, base + (size_t)tile * TILE_BYTES, 1, 2, 128, st) < 0) return -1;
      tile += 2; launches += 2;
    }
    m.gh_tile = tile;
    if (launch_pack(w.gru_w_hh + (size_t)2 * Hd * Hd, Hd, Hd, base + (size_t)tile * TILE_BYTES, 1, 2, 128, st) < 0) return -1;
    tile += 2;
    m.gi_tile = tile;
    if (launch_pack(w.gru_w_ih + (size_t)2 * Hd * KI, Hd, KI, base + (size_t)tile * TILE_BYTES, 1, nkb1, 128, st) < 0) return -1;
    m.gi_pad = gi_pad_tiles(nkb1);
    tile += nkb1 + m.gi_pad; launches += 2;
    for (int g = 0; g < 3; ++g) { m.gln_g[g] = m32.gru_ln[g].g; m.gln_b[g] = m32.gru_ln[g].b; }
    // eval-mode BatchNorm after the prior's first Linear: y = (acc + b) * scale + shift = acc * scale + (b * scale + shift)
    float* shiftp = reinterpret_cast<float*>(base + (size_t)tiles * TILE_BYTES + jobs_bytes_total());
    fold_bias_kernel<<<2, 256, 0, st>>>(m32.prior[0].b, m32.prior[0].post_scale, m32.prior[0].post_shift, 512, shiftp);
    ++launches;
    m.bn_scale = m32.prior[0].post_scale;
    m.bn_shift = shiftp;
    float* w3q = shiftp + 512;
    if (cudaMemsetAsync(w3q, 0, 8 * 512 * sizeof(float), st) != cudaSuccess) return -1;
    if (A <= 8) { round_bf16_kernel<<<(A * 512 + 255) / 256, 256, 0, st>>>(w.pi_w[2], A * 512, w3q); ++launches; }
    m.pi_w3q = w3q;
  }
  if ((int)tile != tiles) return -1;
  // ticket counter of the dynamic tile scheduler: in the arena's tail, zero between launches (the kernel resets it itself)
  m.tile_counter = reinterpret_cast<unsigned int*>(base + (size_t)tiles * TILE_BYTES + jobs_bytes_total() + (512 + 8 * 512) * sizeof(float));
  if (cudaMemsetAsync(m.tile_counter, 0, 16, st) != cudaSuccess) return -1;
  // job programs for every (mode, horizon)
  uint32_t* jdst = reinterpret_cast<uint32_t*>(base + (size_t)tiles * TILE_BYTES);
  std::vector<uint32_t> all;
  size_t off = 0;
  for (int mode = 0; mode < 3; ++mode)
    for (int H = 1; H <= TDMPC_MAX_H; ++H) {
      std::vector<uint32_t> j = build_jobs(mode, H, drnn, A, &m.pair_n0[H]);
      m.jobs[mode][H] = jdst + off;
      m.njobs[mode][H] = (int)j.size();
      j.resize((j.size() + 3) & ~(size_t)3, 0);
      all.insert(all.end(), j.begin(), j.end());
      off += j.size();
    }
  // pageable-host source: the runtime stages the bytes before returning, so `all` may die here
  if (cudaMemcpyAsync(jdst, all.data(), all.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
  // host copy for the launches (the MMA warp reads its program from the kernel-parameter bank)
  {
    uint32_t* hj = static_cast<uint32_t*>(malloc(all.size() * sizeof(uint32_t) + 4));
    if (!hj) return -1;
    memcpy(hj, all.data(), all.size() * sizeof(uint32_t));
    free(const_cast<uint32_t*>(out->jobs_host));          // re-pack of the same handle
    m.jobs_host = hj;
    for (int mode = 0; mode < 3; ++mode)
      for (int H = 1; H <= TDMPC_MAX_H; ++H) m.jobs_off[mode][H] = (int)(m.jobs[mode][H] - jdst);
  }

  // shared-memory budget -> weight ring depth
  // a last k-block that holds a single k-step is stored unswizzled in 4 KB (tc_prims.cuh: umma_desc_k16); at dog-run dims
  // (K = 138 -> 128 + 16) that is the difference between a 2-slot and a 3-slot weight ring
  m.za_compact = (!drnn && nkb1 >= 2 && m.K1p - 64 * (nkb1 - 1) == 16 && !getenv("TDMPC_TC_NO_COMPACT")) ? 1 : 0;
  m.za_bytes = m.za_compact ? (nkb1 - 1) * TILE_BYTES + K16_BLOCK_BYTES : nkb1 * TILE_BYTES;
  const int fixed = XS_BYTES + m.za_bytes + (XCH_FLOATS + QC_FLOATS) * 4 + MAX_JOBS * 4 + (2 * MAX_STAGES + 26) * 8 + 32 + 1024;
  int stages = (SMEM_LIMIT - fixed) / TILE_BYTES;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages == 5) stages = 4;          // the producers split the ring evenly
  if (stages < 2) return -1;
  m.stages = stages;
  m.smem_bytes = (size_t)fixed + (size_t)stages * TILE_BYTES;
  // The attribute is per function AND per device, not per handle: always raise it to the fixed maximum, so handles with
  // different shapes (or on different GPUs of one process) cannot shrink each other's limit.
  if (cudaFuncSetAttribute(rollout_tc_kernel<false, false, false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<false, false, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<false, true, false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<false, true, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<false, false, true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<false, true, true, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<false, false, false, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<false, true, false, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<false, false, false, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<false, false, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<true, false, false, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<true, false, false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<true, false, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<true, true, false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(rollout_tc_kernel<true, true, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return -1;
  // (the WIDE kernels, A > 8, exist for the plain schedule only: no experiment schedules, no CTA pairs there)
  m.alt_sched = (!drnn && A <= 8 && (getenv("TDMPC_TC_OVERLAP") || getenv("TDMPC_TC_SPLIT"))) ? 1 : 0;
  m.pair_mode = (!drnn && A <= 8 && !m.alt_sched && !getenv("TDMPC_TC_NOPAIR")) ? 1 : 0;
  m.pair_rows = out->pair_rows; m.pair_flags = out->pair_flags;          // kept across re-packs of the same handle
  if (m.pair_mode && !m.pair_rows) {
    const size_t rows = (size_t)PAIR_MAX_TILES * (TDMPC_MAX_H + 1) * 128 * m.K1p + (size_t)PAIR_MAX_TILES * 128;
    if (cudaMalloc(&m.pair_rows, rows * sizeof(float)) != cudaSuccess) return -1;
    if (cudaMalloc(&m.pair_flags, PAIR_MAX_TILES * 32 * sizeof(unsigned int)) != cudaSuccess) return -1;
    if (cudaMemsetAsync(m.pair_flags, 0, PAIR_MAX_TILES * 32 * sizeof(unsigned int), st) != cudaSuccess) return -1;
  }
  m.exact_act = c.precision == TDMPC_PREC_BF16_EXACT ? 1 : 0;
  { const char* e = getenv("TDMPC_TC_DEBUG"); m.debug = e ? atoi(e) : 0; }
  m.producers = m.stages % 2 == 0 ? 2 : (m.stages % 3 == 0 ? 3 : 1);
  if (const char* e = getenv("TDMPC_TC_PRODUCERS")) { const int v = atoi(e); if (v >= 1 && v <= 3 && m.stages % v == 0) m.producers = v; }
  *out = m;
  return launches;
}

void tc_release_host(TcModel* m) {
  if (m) {
    free(const_cast<uint32_t*>(m->jobs_host)); m->jobs_host = nullptr;
    if (m->pair_rows) { cudaFree(m->pair_rows); m->pair_rows = nullptr; }
    if (m->pair_flags) { cudaFree(m->pair_flags); m->pair_flags = nullptr; }
  }
}

static unsigned long long* g_tc_trace = nullptr;
void tc_set_trace(unsigned long long* dev_buf) { g_tc_trace = dev_buf; }
int tc_trace_slots() { return 3 * MAX_JOBS * 2 + 256 + 4 * 256; }      // + end-of-tile stamps of block 0's first 128 tiles + per-block stats

int launch_rollout_tc(const TcModel& m, const RolloutParams& p, int num_sms, cudaStream_t st) {
  if (p.H < 1 || p.H > TDMPC_MAX_H) return -1;
  KParams kp;
  kp.m = m; kp.p = p;
  kp.jobs = m.jobs[p.mode][p.H];
  kp.njobs = m.njobs[p.mode][p.H];
  kp.stages = m.stages;
  const long long total = (long long)p.B * p.T;
  if (total >= (1ll << 31)) return -1;
  kp.num_tiles = (int)((total + TILE_ROWS - 1) / TILE_ROWS);
  kp.trace = g_tc_trace;
  kp.one = 1;
  kp.tile_counter = m.tile_counter;
  kp.debug = m.debug;
  kp.producers = m.producers;
  // CTA pairs for launches of a few tiles (single-environment plans): halves the serial chain of a tile
  // (with the 3-slot weight ring of dog-run dims the pair measured 7 % slower than one CTA: single CTA there)
  kp.pair = (m.pair_mode && m.pair_rows && m.stages >= 4 && p.mode == MODE_VALUE && kp.num_tiles <= PAIR_MAX_TILES && 2 * kp.num_tiles <= num_sms &&
             !g_tc_trace) ? 1 : 0;
  kp.njobs1 = 0; kp.j1off = 0;
  if (kp.pair) {
    kp.jobs = m.jobs[2][p.H];
    kp.njobs = m.pair_n0[p.H];
    kp.njobs1 = m.njobs[2][p.H] - kp.njobs;
    kp.j1off = kp.njobs;
  }
  const int grid = kp.pair ? 2 * kp.num_tiles : (kp.num_tiles < num_sms ? kp.num_tiles : num_sms);
  if (kp.njobs <= 0 || kp.njobs + kp.njobs1 > MAX_JOBS || !m.jobs_host) return -1;
  memcpy(kp.jw, m.jobs_host + m.jobs_off[kp.pair ? 2 : p.mode][p.H], sizeof(uint32_t) * (kp.njobs + kp.njobs1));
  const bool wide = m.A > 8;
#define LAUNCH_TC(...) do { if (wide) rollout_tc_kernel<__VA_ARGS__, true><<<grid, NTHREADS, m.smem_bytes, st>>>(kp); \
                            else rollout_tc_kernel<__VA_ARGS__, false><<<grid, NTHREADS, m.smem_bytes, st>>>(kp); } while (0)
#define LAUNCH_TC_NARROW(...) do { if (wide) return -1; rollout_tc_kernel<__VA_ARGS__, false><<<grid, NTHREADS, m.smem_bytes, st>>>(kp); } while (0)
  if (kp.trace) {
    // the stamping build: default precision mode, plain schedule only (tools/trace_rollout.py)
    if (kp.pair || m.alt_sched || m.exact_act) return -1;
    if (m.drnn && wide) return -1;
    if (m.drnn) LAUNCH_TC_NARROW(true, false, false, false, true);
    else LAUNCH_TC(false, false, false, false, true);
  } else if (m.drnn) {
    if (m.exact_act) LAUNCH_TC(true, true, false, false, false);
    else LAUNCH_TC(true, false, false, false, false);
  } else {
    if (kp.pair) {
      if (m.exact_act) LAUNCH_TC_NARROW(false, true, true, false, false);
      else LAUNCH_TC_NARROW(false, false, true, false, false);
    } else if (m.alt_sched) {
      if (m.exact_act) LAUNCH_TC_NARROW(false, true, false, true, false);
      else LAUNCH_TC_NARROW(false, false, false, true, false);
    } else if (m.exact_act) LAUNCH_TC(false, true, false, false, false);
    else LAUNCH_TC(false, false, false, false, false);
  }
  return 1;
}
