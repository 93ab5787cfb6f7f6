// tcgen05 trajectory rollout: the bf16-operand / fp32-accumulate throughput mode of the planner.
//
// One persistent CTA per SM owns tiles of 128 trajectories (TMEM lane = trajectory) and walks
// each tile through the whole horizon -- TOLD.next (dynamics + reward) per step, then the
// terminal pi / min(Q1,Q2) (algorithm/tdmpc.py:83-92), or pi + dynamics for the policy rollout
// (tdmpc.py:113-118) -- without the latent state ever leaving the SM:
//
//   warp 0   producer: streams the pre-swizzled bf16 weight tiles (16 KB each, laid out in global
//            memory in exactly the order they are consumed) from L2 into a shared-memory ring with
//            bulk-TMA copies (cp.async.bulk ... mbarrier::complete_tx).
//   warp 1   MMA issuer: one thread issues tcgen05.mma (M=128, N=128, K=16) per k-step.
//            layer 1: A = [z,a] tile in shared memory (SS);  layer 2: A = h1 in TENSOR MEMORY (TS),
//            so the 128x512 hidden activation never touches shared memory (the interleaved reward
//            chain keeps its h1 in a second smem buffer, XR, because TMEM is full); layer 3: A = h2 in smem.
//   warps 4-11 epilogue (two warpgroups, one per accumulator buffer): thread = trajectory.
//            Each job's 128 accumulator columns are drained into registers in one go and the buffer is
//            handed back to the MMA warp BEFORE the math (bias, ELU / LayerNorm+tanh, bf16 pack,
//            tcgen05.st / swizzled st.shared, or the fp32 dot product of the N=1 reward / Q heads),
//            so the tensor pipe never waits for MUFU-bound epilogue work.
//
// Job order (build_jobs): MMA-heavy layer-2 jobs of one chain are interleaved with the MUFU-heavy
// layer-1 jobs of the other chain (dynamics L2 | reward L1, then reward L2 | next step's dynamics L1),
// which keeps both the tensor pipe and the epilogue warps busy.  The Q heads need LayerNorm over the
// whole 512-wide row; instead of parking 512 fp32 columns in TMEM (which serialises MMA and epilogue)
// each Q layer is computed TWICE in 128-column chunks: a statistics pass, then a pass that normalises
// (the accumulators are bit-identical both times, the extra MMAs are cheaper than the stall).
//
// TMEM map (512 columns): [0,128) [128,256) accumulator ping-pong, [256,512) h1 as bf16x2 (K=512).
// All three roles walk the same job list (tc_model.cuh), which keeps the mbarrier phases aligned.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_model.cuh"
#include "tc_prims.cuh"

using namespace tc;

namespace {

constexpr int TILE_ROWS = 128;
constexpr int NTHREADS = 384;              // warps 0-3: producer / MMA / 2 idle; warps 4-7, 8-11: epilogue groups 0 and 1
constexpr int REGS_CTRL = 56, REGS_EPI = 224;   // setmaxnreg: frees (168-56)*128 = takes (224-168)*256 registers, exactly
constexpr int XCH_FLOATS = 2 * 3 * 2 * 128;   // epilogue-group exchange: 2 parities x {sum, sum-of-squares, dot} x 2 groups x 128 rows
constexpr int TILE_BYTES = 16384;
constexpr int XS_BYTES = 8 * TILE_BYTES;      // 128 x 512 bf16
constexpr int MAX_JOBS = 17 * TDMPC_MAX_H + 48;
constexpr uint32_t TM_ACC = 0, TM_A = 256;
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
  int cluster;                 // CTAs per cluster sharing one weight stream (TMA multicast): 1, 2 or 4
  int debug;                   // experiments only (TDMPC_TC_DEBUG): bit0 = producer skips the weight copies (timing probe, wrong results)
  unsigned long long* trace;   // optional [3 roles][MAX_JOBS][3] clock64 stamps (job start, after the waits, end) of block 0's third tile (tools/trace_rollout.py)
};

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float elu_fast(float x) {
  const float e = ex2_approx(x * 1.4426950408889634f) - 1.0f;
  return x > 0.0f ? x : e;
}
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = ex2_approx(x * 2.8853900817779268f);       // exp(2x)
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}
__device__ __forceinline__ float4 ldg4(const float* p, int i) { return __ldg(reinterpret_cast<const float4*>(p) + i); }
__device__ __forceinline__ float f4at(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

__device__ __forceinline__ int job_type(uint32_t j) { return j & 3; }
__device__ __forceinline__ int job_epi(uint32_t j) { return (j >> 2) & 15; }
__device__ __forceinline__ int job_chain(uint32_t j) { return (j >> 6) & 7; }
__device__ __forceinline__ int job_chunk(uint32_t j) { return (j >> 9) & 3; }
__device__ __forceinline__ int job_t(uint32_t j) { return (j >> 11) & 31; }
__device__ __forceinline__ int job_flags(uint32_t j) { return (j >> 16) & 255; }

__device__ __forceinline__ int job_tiles(const TcModel& m, uint32_t job) {
  switch (job_type(job)) {
    case J_L1: return m.nkb1;
    case J_L2: return 8;
    default: return m.l3rows[job_chain(job)] == 128 ? 8 : 4;
  }
}
__device__ __forceinline__ uint32_t job_first_tile(const TcModel& m, uint32_t job) {
  const TcChain& c = m.chain[job_chain(job)];
  switch (job_type(job)) {
    case J_L1: return c.l1_tile + job_chunk(job) * m.nkb1;
    case J_L2: return c.l2_tile + job_chunk(job) * 8;
    default: return c.l3_tile;
  }
}

// bf16 store of one element of the [z,a] operand (128B-swizzled k-blocks of 64)
__device__ __forceinline__ void za_store1(uint8_t* za, int row, int k, float v) {
  *reinterpret_cast<__nv_bfloat16*>(za + (k >> 6) * TILE_BYTES + sw128_offset(row, k & 63)) = __float2bfloat16(v);
}
// 8 consecutive k (k0 % 8 == 0) as one 16-byte store
__device__ __forceinline__ void sw_store8(uint8_t* base, int row, int k0, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(base + (k0 >> 6) * TILE_BYTES + sw128_offset(row, k0 & 63)) = u;
}

// Drain this warp's 32 lanes x 128 accumulator columns into registers (4 loads in flight, one wait).
__device__ __forceinline__ void acc_drain(uint32_t taddr, uint32_t (&v)[4][32]) {
  tmem_ld32(taddr, v[0]); tmem_ld32(taddr + 32, v[1]); tmem_ld32(taddr + 64, v[2]); tmem_ld32(taddr + 96, v[3]);
  tmem_wait_ld_dep(v[0]); tmem_wait_ld_dep(v[1]); tmem_wait_ld_dep(v[2]); tmem_wait_ld_dep(v[3]);
}

__global__ void __launch_bounds__(NTHREADS, 1) rollout_tc_kernel(const __grid_constant__ KParams kp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const TcModel& m = kp.m;
  const RolloutParams& p = kp.p;
  const int S = kp.stages;

  // 1024-byte aligned carve-up (the swizzle atoms need it)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xs = smem;                                         // h2 of dynamics / pi: A operand of layer 3
  uint8_t* xr = xs + XS_BYTES;                                // h1 of the interleaved reward chain (only when m.use_xr)
  uint8_t* za = xr + (m.use_xr ? XS_BYTES : 0);               // [z, a] operand
  uint8_t* wring = za + m.nkb1 * TILE_BYTES;
  float* xch = reinterpret_cast<float*>(wring + S * TILE_BYTES);     // epilogue-group exchange
  uint32_t* sjobs = reinterpret_cast<uint32_t*>(xch + XCH_FLOATS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sjobs + ((kp.njobs + 3) & ~3));
  uint64_t* w_full = bars;
  uint64_t* w_empty = w_full + MAX_STAGES;
  uint64_t* acc_full = w_empty + MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* h1_full = acc_empty + 2;     // h1 chunk c written to TMEM (layer-1 epilogue -> TS layer-2 MMAs)
  uint64_t* xr_full = h1_full + 4;       // XR chunk c written to smem (reward layer-1 epilogue -> SS layer-2 MMAs)
  uint64_t* xs_full = xr_full + 4;       // XS chunk c written to smem (layer-2 epilogue -> layer-3 MMAs)
  uint64_t* za_full = xs_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(za_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < kp.njobs; i += NTHREADS) sjobs[i] = kp.jobs[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], kp.cluster); }   // every CTA of the cluster releases a slot
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 256); }
    for (int i = 0; i < 4; ++i) { mbar_init(&h1_full[i], 256); mbar_init(&xr_full[i], 256); mbar_init(&xs_full[i], 256); }
    mbar_init(za_full, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  if (kp.cluster > 1) cluster_sync_all();       // peers' mbarriers are initialised before anything is multicast to them
  tcgen05_fence_after();
  if (*tmem_slot != 0) asm volatile("trap;");   // all 512 columns of the SM are ours: the base is 0, so TMEM addresses are immediates
  constexpr uint32_t tmem = 0;
  // The CTAs of a cluster walk the same job list over C consecutive tiles (one each) in lock-step, because they
  // share one weight stream; tiles past the end are processed as all-invalid rows.
  const int C = kp.cluster;
  const uint32_t crank = C > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << C) - 1u);
  const int n_clusters = (int)gridDim.x / C, cluster_id = (int)blockIdx.x / C;
  const int n_groups = (kp.num_tiles + C - 1) / C;
  const int my_tiles = (n_groups - cluster_id + n_clusters - 1) / n_clusters;

  if (warp < 4) {
    // warps 0-3 form one warpgroup: producer, MMA issuer, two idle warps.  They hand most of their
    // registers to the two epilogue warpgroups.
    setmaxnreg_dec<REGS_CTRL>();
    if (warp != 1) {
      // ============================== producers: weight tiles, L2 -> smem ring ==============================
      // One thread can only issue a bulk copy every ~400 cycles (measured, tools/smem_contention.cu), which would
      // cap the weight stream at ~40 B/clk -- less than the MMAs consume.  So warps 0, 2 and 3 each run a producer
      // thread; producer k of P owns the tiles (and, because P divides the ring depth, the ring slots) k mod P.
      // Every CTA arms its own w_full[stage] for every tile; in a cluster the CTAs take turns issuing the copy,
      // multicast to all of them.  w_empty[stage] collects one arrival per CTA (the MMA warps commit to every
      // CTA of the cluster), so the issuer knows the slot is free everywhere before it overwrites it.
      const int P = kp.producers;
      const int pk = warp == 0 ? 0 : warp - 1;
      if (lane == 0 && pk < P) {
        uint32_t stage = 0, phase = 0, turn = 0, mine = 0;    // mine: tile counter mod P
        for (int it = 0; it < my_tiles; ++it)
          for (int j = 0; j < kp.njobs; ++j) {
            const uint32_t job = sjobs[j];
            const int nt = job_tiles(m, job);
            const uint8_t* src = m.tiles + (size_t)job_first_tile(m, job) * TILE_BYTES;
            for (int i = 0; i < nt; ++i) {
              if (mine == (uint32_t)pk) {
                mbar_wait(&w_empty[stage], phase ^ 1, 1);
                if (kp.debug & 1) mbar_arrive(&w_full[stage]);
                else {
                  mbar_arrive_expect_tx(&w_full[stage], TILE_BYTES);
                  if (C == 1) bulk_g2s(wring + stage * TILE_BYTES, src + (size_t)i * TILE_BYTES, TILE_BYTES, &w_full[stage]);
                  else if (turn == crank)
                    bulk_g2s_multicast(wring + stage * TILE_BYTES, src + (size_t)i * TILE_BYTES, TILE_BYTES, &w_full[stage], cmask);
                }
              }
              if (++mine == (uint32_t)P) mine = 0;
              if (++turn == (uint32_t)C) turn = 0;
              if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; }
            }
          }
      }
    } else if (warp == 1) {
      // ============================== MMA issuer ==============================
      // The whole warp walks the job list converged (waits are warp-wide); one elected lane issues.
      const bool leader = elect_one();      // elected once while converged: the same lane issues every MMA and commit
      uint32_t stage = 0, wphase = 0;
      uint32_t ecnt0 = 0, ecnt1 = 0, nbuf = 0, za_phase = 0;
      // bit c = parity to wait for.  h1 / XR / XS have separate barriers: a shared one would let an
      // epilogue complete a second phase before the MMA warp has waited for the first, and a parity
      // wait cannot tell the two apart.
      uint32_t h1_phase = 0, xr_phase = 0, xs_phase = 0;
      const uint32_t idesc128 = umma_idesc_bf16(128, 128);
      const uint32_t za_addr = smem_u32(za), xs_addr = smem_u32(xs), xr_addr = smem_u32(xr), ring_addr = smem_u32(wring);
      const int ks1_total = m.K1p / 16;

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
      auto tile_release = [&]() {            // commit -> frees the slot (in every CTA of the cluster) when the MMAs that read it retire
        if (leader) {
          if (C == 1) umma_commit(&w_empty[cur_stage]);
          else umma_commit_multicast(&w_empty[cur_stage], cmask);
        }
      };
      auto acc_acquire = [&](int b) {
        if (b == 0) { mbar_wait(&acc_empty[0], (ecnt0 & 1) ^ 1, 3); ++ecnt0; }
        else { mbar_wait(&acc_empty[1], (ecnt1 & 1) ^ 1, 3); ++ecnt1; }
      };
      auto h1_wait = [&](int c) { mbar_wait(&h1_full[c], (h1_phase >> c) & 1, 4); h1_phase ^= 1u << c; };
      auto xr_wait = [&](int c) { mbar_wait(&xr_full[c], (xr_phase >> c) & 1, 9); xr_phase ^= 1u << c; };
      auto xs_wait = [&](int c) { mbar_wait(&xs_full[c], (xs_phase >> c) & 1, 5); xs_phase ^= 1u << c; };
      // one k-block (<= 4 k-steps of 16) with both operands in shared memory
      auto kblock_ss = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, int nks, bool first) {
        const uint32_t alo = umma_desc_lo(a_addr), blo = umma_desc_lo(b_addr);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          if (ks < nks) umma_ss(d, umma_desc_make(alo + 2 * ks), umma_desc_make(blo + 2 * ks), idesc, (first && ks == 0) ? 0u : 1u);
      };

      for (int it = 0; it < my_tiles; ++it) {
        const bool tr = kp.trace && blockIdx.x == 0 && it == 2 && lane == 0;
        for (int j = 0; j < kp.njobs; ++j) {
          const uint32_t job = sjobs[j];
          const int type = job_type(job), chain = job_chain(job), flags = job_flags(job);
          if (tr) kp.trace[(0 * MAX_JOBS + j) * 3] = clock64();
          if (flags & F_WAIT_ZA) { mbar_wait(za_full, za_phase, 6); za_phase ^= 1; }
          const int b = nbuf; nbuf ^= 1;
          acc_acquire(b);
          tcgen05_fence_after();
          if (tr) kp.trace[(0 * MAX_JOBS + j) * 3 + 1] = clock64();
          const uint32_t d = tmem + TM_ACC + b * 128;
          if (type == J_L1) {
            for (int kb = 0; kb < m.nkb1; ++kb) {
              const uint32_t wt = tile_acquire();
              if (leader) kblock_ss(d, za_addr + kb * TILE_BYTES, wt, idesc128, min(4, ks1_total - kb * 4), kb == 0);
              tile_release();
            }
          } else if (type == J_L2) {
            const bool wait_a = (flags & F_WAIT_A) != 0;
            if (flags & F_A_SMEM) {
#pragma unroll
              for (int kb = 0; kb < 8; ++kb) {
                if (wait_a && (kb & 1) == 0) { xr_wait(kb >> 1); tcgen05_fence_after(); }
                const uint32_t wt = tile_acquire();
                if (leader) kblock_ss(d, xr_addr + kb * TILE_BYTES, wt, idesc128, 4, kb == 0);
                tile_release();
              }
            } else {
#pragma unroll
              for (int kb = 0; kb < 8; ++kb) {
                if (wait_a && (kb & 1) == 0) { h1_wait(kb >> 1); tcgen05_fence_after(); }
                const uint32_t wt = tile_acquire();
                if (leader) {
                  const uint32_t blo = umma_desc_lo(wt);
                  const uint32_t a = tmem + TM_A + kb * 32;
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) umma_ts(d, a + ks * 8, umma_desc_make(blo + 2 * ks), idesc128, (kb | ks) ? 1u : 0u);
                }
                tile_release();
              }
            }
          } else {
            const uint32_t idesc = umma_idesc_bf16(128, m.n3p[chain]);
            const bool half = m.l3rows[chain] == 64;
            uint32_t wt = 0;
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
              if ((kb & 1) == 0) { xs_wait(kb >> 1); tcgen05_fence_after(); }
              if (!half || (kb & 1) == 0) wt = tile_acquire();
              if (leader) kblock_ss(d, xs_addr + kb * TILE_BYTES, wt + ((half && (kb & 1)) ? 8192u : 0u), idesc, 4, kb == 0);
              if (!half || (kb & 1)) tile_release();
            }
          }
          if (leader) umma_commit(&acc_full[b]);
          if (tr) kp.trace[(0 * MAX_JOBS + j) * 3 + 2] = clock64();
          __syncwarp();
        }
      }
    }
  } else {
    // ============================== epilogue: two warpgroups, thread = trajectory ==============================
    // Both groups work on EVERY job: group g (warps 4-7: g=0, warps 8-11: g=1) takes accumulator columns
    // [64g, 64g+64) of the 128-column chunk.  Splitting by columns instead of by jobs halves the time an
    // accumulator buffer is held, and 64 columns fit in registers at once: the buffer is handed back to the
    // MMA warp right after the TMEM loads, BEFORE the (MUFU-bound) math.  Two warps per SM sub-partition
    // hide each other's TMEM / L1 / MUFU latencies.  Row-wise reductions (reward dot product, LayerNorm
    // statistics, Q dot product) are combined across the groups through `xch` in shared memory.
    setmaxnreg_inc<REGS_EPI>();
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const int L = m.L, A = m.A;
    const long long total = (long long)p.B * p.T;
    uint32_t fcnt0 = 0, fcnt1 = 0, nbuf = 0, xpar = 0;

    auto acc_wait = [&](int b) {
      if (b == 0) { mbar_wait(&acc_full[0], fcnt0 & 1, 7); ++fcnt0; }
      else { mbar_wait(&acc_full[1], fcnt1 & 1, 7); ++fcnt1; }
    };
    // Q-head constants staged in XS: [head][layer][kind: b, gamma, beta, w3][512] fp32
    auto qconst = [&](int chain, int layer, int kind) -> const float4* {
      return reinterpret_cast<const float4*>(xs) + (((chain - CH_Q1) * 2 + layer) * 4 + kind) * 128;
    };
#define ACCF(g, i) __uint_as_float(v[g][i])

    for (int it = 0; it < my_tiles; ++it) {
      const long long g = ((long long)(cluster_id + it * n_clusters) * C + crank) * TILE_ROWS + row;
      const bool valid = g < total;
      const long long pb = valid ? g / p.T : 0;
      const long long pj = valid ? g - pb * p.T : 0;
      const size_t row_off = (size_t)(pb * p.Tcap + pj);                      // row inside [B, Tcap]
      const uint64_t nrow = (uint64_t)pb * p.noise_bstride + (uint64_t)pj * A;  // this row's offset inside the noise stream
      const float* act_row = p.actions ? p.actions + (row_off * p.H) * A : nullptr;
      float* pi_row = p.pi_actions ? p.pi_actions + (row_off * p.H) * A : nullptr;
      float G = 0.0f, rlast = 0.0f, dot = 0.0f, q1v = 0.0f, sum = 0.0f, ssq = 0.0f, mu = 0.0f, rstd = 1.0f;
      const bool tr = kp.trace && blockIdx.x == 0 && it == 2 && threadIdx.x % 128 == 0;

      // initial [z0, a0, 0...] row: the groups write alternate 8-column pieces
      {
        const float* zr = p.z0 + (size_t)pb * L;
        for (int k0 = grp * 8; k0 < m.nkb1 * 64; k0 += 16) {
          float v8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = k0 + i;
            float x = 0.0f;
            if (valid) {
              if (k < L) x = __ldg(zr + k);
              else if (k < L + A && p.mode == MODE_VALUE) x = __ldg(act_row + (k - L));
            }
            v8[i] = x;
          }
          sw_store8(za, row, k0, v8);
        }
        fence_proxy_async_smem();
        mbar_arrive(za_full);
      }

      for (int j = 0; j < kp.njobs; ++j) {
        const uint32_t job = sjobs[j];
        const int type = job_type(job), epi = job_epi(job), chain = job_chain(job), chunk = job_chunk(job), t = job_t(job), flags = job_flags(job);
        const TcChain& ch = m.chain[chain];
        const int b = nbuf; nbuf ^= 1;
        unsigned long long* trp = kp.trace + ((1 + grp) * MAX_JOBS + j) * 3;
        if (tr) trp[0] = clock64();
        const uint32_t acc = tlane + TM_ACC + b * 128 + grp * 64;      // my 64 columns of the chunk
        const int coff = chunk * 128 + grp * 64;                         // my first column inside the 512-wide layer
        // Per-column constants of my 64 columns: the MLP chains read bias / head weights through L1, the Q heads
        // from the copy staged in shared memory (XS is idle during the Q phase).
        const float* bias = (type == J_L1 ? ch.b1 : ch.b2) + coff;
        acc_wait(b);
        tcgen05_fence_after();
        if (tr) trp[1] = clock64();

        if (epi == E_Z) {
          // z' = acc + b3 -> [z', a_{t+1}, 0...] row of the next step.  Columns >= L+A of the row are zero
          // padding written once per tile and never touched again.  Group g owns row columns [64g, 64g+64);
          // group 0 also owns the pure action columns >= 128 (L <= 128).
          uint32_t v[2][32];
          const int kbase = grp * 64;
          if (kbase < L) {
            tmem_ld32(acc, v[0]);
            if (kbase + 32 < L) tmem_ld32(acc + 32, v[1]);
            tmem_wait_ld_dep(v[0]);
            if (kbase + 32 < L) tmem_wait_ld_dep(v[1]);
          }
          tcgen05_fence_before();
          mbar_arrive(&acc_empty[b]);
          const bool next_act = p.mode == MODE_VALUE && !(flags & F_LAST_STEP);
          const float* a_next = act_row + (size_t)(t + 1) * A;
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const int kk = kbase + c8 * 8;
            if (kk + 8 <= L) {
              // eight z' elements
              const float4 b0 = ldg4(ch.b3, kk >> 2), b1 = ldg4(ch.b3, (kk >> 2) + 1);
              float o[8];
              o[0] = ACCF(c8 >> 2, (c8 & 3) * 8) + b0.x; o[1] = ACCF(c8 >> 2, (c8 & 3) * 8 + 1) + b0.y;
              o[2] = ACCF(c8 >> 2, (c8 & 3) * 8 + 2) + b0.z; o[3] = ACCF(c8 >> 2, (c8 & 3) * 8 + 3) + b0.w;
              o[4] = ACCF(c8 >> 2, (c8 & 3) * 8 + 4) + b1.x; o[5] = ACCF(c8 >> 2, (c8 & 3) * 8 + 5) + b1.y;
              o[6] = ACCF(c8 >> 2, (c8 & 3) * 8 + 6) + b1.z; o[7] = ACCF(c8 >> 2, (c8 & 3) * 8 + 7) + b1.w;
              sw_store8(za, row, kk, o);
            } else if (kk < L + A) {
              // piece with the z / action boundary (or pure action columns): element-wise
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int k = kk + i;
                float x = 0.0f;
                if (k < L) x = ACCF(c8 >> 2, (c8 & 3) * 8 + i) + __ldg(ch.b3 + k);
                else if (k < L + A && next_act && valid) x = __ldg(a_next + (k - L));
                o[i] = x;
              }
              if (next_act) sw_store8(za, row, kk, o);
              else {
                // the action columns must survive (policy rollout / last step): z elements only
#pragma unroll
                for (int i = 0; i < 8; ++i) if (kk + i < L) za_store1(za, row, kk + i, o[i]);
              }
            }
          }
          if (grp == 0 && next_act)
            for (int kk = 128; kk < L + A; kk += 8) {     // action columns beyond the accumulator range
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = (kk + i < L + A && valid) ? __ldg(a_next + (kk + i - L)) : 0.0f;
              sw_store8(za, row, kk, o);
            }
          fence_proxy_async_smem();
          mbar_arrive(za_full);
        } else if (epi == E_PI) {
          // a = clamp(tanh(acc + b3) + clip(eps * std)), written to the action columns of [z,a] (group 0; A <= 64)
          if (grp == 0) {
            const bool terminal = (flags & F_TERMINAL) != 0;
            const uint64_t nbase = p.noise_base + (terminal ? 0 : (uint64_t)t * p.noise_tstride) + nrow;
#pragma unroll 1
            for (int a0 = 0; a0 < A; a0 += 4) {
              uint32_t v4[4];
              tmem_ld4(acc + a0, v4);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int a = a0 + i;
                if (a < A) {
                  float x = 0.0f;
                  if (valid) {
                    const float mu_a = tanhf(__uint_as_float(v4[i]) + __ldg(ch.b3 + a));
                    x = policy_sample(mu_a, p.min_std, noise_at(p.noise, nbase + a));
                    if (!terminal) pi_row[(size_t)t * A + a] = x;
                  }
                  if (!(flags & F_LAST_STEP)) za_store1(za, row, L + a, x);
                }
              }
            }
          }
          tcgen05_fence_before();
          mbar_arrive(&acc_empty[b]);
          if (!(flags & F_LAST_STEP)) {
            fence_proxy_async_smem();
            mbar_arrive(za_full);
          }
          if (flags & F_TERMINAL) {
            // The layer-3 MMAs (last readers of XS) have retired: stage the Q-head constants there.
            const int te = threadIdx.x - 128;               // 0..255
            float4* dst = reinterpret_cast<float4*>(xs);
#pragma unroll 1
            for (int idx = te; idx < 16 * 128; idx += 256) {
              const int arr = idx >> 7, e = idx & 127;
              const TcChain& qc = m.chain[CH_Q1 + (arr >> 3)];
              const int layer = (arr >> 2) & 1, kind = arr & 3;
              const float* src = kind == 0 ? (layer ? qc.b2 : qc.b1) : kind == 1 ? qc.ln_g[layer] : kind == 2 ? qc.ln_b[layer] : qc.w3;
              dst[idx] = ldg4(src, e);
            }
            epi_bar_sync();
          }
        } else {
          // ---- 128-column chunk epilogues: drain my 64 columns, release the accumulator, then do the math ----
          uint32_t v[2][32];
          tmem_ld32(acc, v[0]);
          tmem_ld32(acc + 32, v[1]);
          tmem_wait_ld_dep(v[0]);
          tmem_wait_ld_dep(v[1]);
          tcgen05_fence_before();
          mbar_arrive(&acc_empty[b]);
          if (epi == E_H1) {
            // h1 = ELU(acc + b1) -> bf16x2 -> TMEM A-operand columns
            const uint32_t dst = tlane + TM_A + (coff >> 1);
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {
              uint32_t o[16];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 x = ldg4(bias, gq * 8 + i);
                o[2 * i] = pack_bf16x2(elu_fast(ACCF(gq, 4 * i) + x.x), elu_fast(ACCF(gq, 4 * i + 1) + x.y));
                o[2 * i + 1] = pack_bf16x2(elu_fast(ACCF(gq, 4 * i + 2) + x.z), elu_fast(ACCF(gq, 4 * i + 3) + x.w));
              }
              tmem_st16(dst + gq * 16, o);
            }
            tmem_wait_st();
            tcgen05_fence_before();
            mbar_arrive(&h1_full[chunk]);
          } else if (epi == E_XR || epi == E_XS) {
            // ELU(acc + b) -> bf16 -> 128B-swizzled smem operand (XR: reward h1, XS: dynamics / pi h2)
            uint8_t* dstb = epi == E_XR ? xr : xs;
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 x0 = ldg4(bias, gq * 8 + 2 * i), x1 = ldg4(bias, gq * 8 + 2 * i + 1);
                float h[8];
                h[0] = elu_fast(ACCF(gq, 8 * i) + x0.x); h[1] = elu_fast(ACCF(gq, 8 * i + 1) + x0.y);
                h[2] = elu_fast(ACCF(gq, 8 * i + 2) + x0.z); h[3] = elu_fast(ACCF(gq, 8 * i + 3) + x0.w);
                h[4] = elu_fast(ACCF(gq, 8 * i + 4) + x1.x); h[5] = elu_fast(ACCF(gq, 8 * i + 5) + x1.y);
                h[6] = elu_fast(ACCF(gq, 8 * i + 6) + x1.z); h[7] = elu_fast(ACCF(gq, 8 * i + 7) + x1.w);
                sw_store8(dstb, row, coff + gq * 32 + i * 8, h);
              }
            }
            fence_proxy_async_smem();
            mbar_arrive(epi == E_XR ? &xr_full[chunk] : &xs_full[chunk]);
          } else if (epi == E_DOT) {
            // reward head (N = 1): fp32 dot product with w3 (helper.py:169-176 last Linear)
            const float* w3 = ch.w3 + coff;
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 x = ldg4(bias, gq * 8 + i), w = ldg4(w3, gq * 8 + i);
                dot = fmaf(elu_fast(ACCF(gq, 4 * i) + x.x), w.x, dot);
                dot = fmaf(elu_fast(ACCF(gq, 4 * i + 1) + x.y), w.y, dot);
                dot = fmaf(elu_fast(ACCF(gq, 4 * i + 2) + x.z), w.z, dot);
                dot = fmaf(elu_fast(ACCF(gq, 4 * i + 3) + x.w), w.w, dot);
              }
            }
          } else if (epi == E_STAT) {
            // LayerNorm statistics of x = acc + b over the 512-wide row, accumulated chunk by chunk
            const float4* qb = qconst(chain, type == J_L1 ? 0 : 1, 0) + (coff >> 2);
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 x = qb[gq * 8 + i];
                const float y0 = ACCF(gq, 4 * i) + x.x, y1 = ACCF(gq, 4 * i + 1) + x.y, y2 = ACCF(gq, 4 * i + 2) + x.z, y3 = ACCF(gq, 4 * i + 3) + x.w;
                sum += (y0 + y1) + (y2 + y3);
                ssq = fmaf(y0, y0, ssq); ssq = fmaf(y1, y1, ssq); ssq = fmaf(y2, y2, ssq); ssq = fmaf(y3, y3, ssq);
              }
            }
          } else {
            // E_QN / E_QDOT: the chunk recomputed with the row statistics known: LayerNorm (helper.py:199-200), then
            //   E_QN:   tanh -> bf16x2 -> TMEM h1 chunk          E_QDOT: ELU -> dot with w3
            const int li = type == J_L1 ? 0 : 1;
            const float4* qb = qconst(chain, li, 0) + (coff >> 2);
            const float4* qg = qconst(chain, li, 1) + (coff >> 2);
            const float4* qe = qconst(chain, li, 2) + (coff >> 2);
            const float4* qw = qconst(chain, 1, 3) + (coff >> 2);
            const uint32_t dst = tlane + TM_A + (coff >> 1);
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {
              uint32_t o[16];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 x = qb[gq * 8 + i], gg = qg[gq * 8 + i], bb = qe[gq * 8 + i];
                const float y0 = (ACCF(gq, 4 * i) + x.x - mu) * rstd * gg.x + bb.x;
                const float y1 = (ACCF(gq, 4 * i + 1) + x.y - mu) * rstd * gg.y + bb.y;
                const float y2 = (ACCF(gq, 4 * i + 2) + x.z - mu) * rstd * gg.z + bb.z;
                const float y3 = (ACCF(gq, 4 * i + 3) + x.w - mu) * rstd * gg.w + bb.w;
                if (epi == E_QN) {
                  o[2 * i] = pack_bf16x2(tanh_fast(y0), tanh_fast(y1));
                  o[2 * i + 1] = pack_bf16x2(tanh_fast(y2), tanh_fast(y3));
                } else {
                  const float4 f = qw[gq * 8 + i];
                  dot = fmaf(elu_fast(y0), f.x, dot); dot = fmaf(elu_fast(y1), f.y, dot);
                  dot = fmaf(elu_fast(y2), f.z, dot); dot = fmaf(elu_fast(y3), f.w, dot);
                }
              }
              if (epi == E_QN) tmem_st16(dst + gq * 16, o);
            }
            if (epi == E_QN) {
              tmem_wait_st();
              tcgen05_fence_before();
              mbar_arrive(&h1_full[chunk]);
            }
          }
          if (chunk == 3 && (epi == E_DOT || epi == E_STAT || epi == E_QDOT)) {
            // each group holds a partial row sum: combine through shared memory (double-buffered by parity)
            float* xb = xch + xpar * (XCH_FLOATS / 2);
            xpar ^= 1;
            if (epi == E_STAT) { xb[(0 * 2 + grp) * 128 + row] = sum; xb[(1 * 2 + grp) * 128 + row] = ssq; }
            else xb[(2 * 2 + grp) * 128 + row] = dot;
            epi_bar_sync();
            if (epi == E_STAT) {
              // (group 0's partial) + (group 1's partial) in both groups: identical statistics
              const float s0 = grp == 0 ? sum : xb[(0 * 2 + 0) * 128 + row], s1 = grp == 1 ? sum : xb[(0 * 2 + 1) * 128 + row];
              const float q0 = grp == 0 ? ssq : xb[(1 * 2 + 0) * 128 + row], q1 = grp == 1 ? ssq : xb[(1 * 2 + 1) * 128 + row];
              mu = (s0 + s1) * (1.0f / 512.0f);
              const float var = fmaxf((q0 + q1) * (1.0f / 512.0f) - mu * mu, 0.0f);      // biased variance, as nn.LayerNorm
              rstd = 1.0f / sqrtf(var + 1e-5f);
              sum = 0.0f; ssq = 0.0f;
            } else {
              if (grp == 0) {
                const float r = (dot + xb[(2 * 2 + 1) * 128 + row]) + __ldg(ch.b3);
                if (epi == E_DOT) {
                  G = __fadd_rn(G, __fmul_rn(p.disc[t], r));   // G += discount * reward (tdmpc.py:89): mul, then add, as torch does
                  rlast = r;
                } else if (chain == CH_Q1) {
                  q1v = r;
                } else if (valid) {
                  // G += discount * min(Q1, Q2); nan_to_num_   (tdmpc.py:91-92)
                  const float val = __fadd_rn(G, __fmul_rn(p.disc[p.H], min_nanprop(q1v, r)));
                  p.values[row_off] = nan_to_num0(val);
                  if (p.last_reward) p.last_reward[row_off] = rlast;
                }
              }
              dot = 0.0f;
            }
          }
        }
        if (tr) trp[2] = clock64();
      }
    }
#undef ACCF
  }

  tcgen05_fence_before();
  __syncthreads();
  if (kp.cluster > 1) cluster_sync_all();       // no CTA leaves while a peer may still signal its barriers
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

int launch_pack(const float* W, int N, int K, uint8_t* dst, int n_chunks, int nkb, int tile_rows, cudaStream_t st) {
  const long long total = (long long)n_chunks * nkb * tile_rows * 8;
  pack_tiles_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W, N, K, dst, n_chunks, nkb, tile_rows);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// The job program of one tile.  xr: the reward chain keeps its h1 in the smem buffer XR, which lets its
// layer-1 jobs (MUFU-bound epilogues, 4 MMAs each) hide under the dynamics chain's layer-2 jobs (32 MMAs
// each) and the next step's dynamics layer 1 under the reward chain's layer 2.  Heavy and light jobs are
// issued in pairs (H H L L ...) so consecutive heavy jobs land in different accumulator buffers.
std::vector<uint32_t> build_jobs(int mode, int H, bool xr) {
  std::vector<uint32_t> j;
  auto J = [&](int type, int epi, int chain, int c, int t, int fl) { j.push_back(tc_job(type, epi, chain, c, t, fl)); };
  auto l1 = [&](int chain, int t, int epi, bool wait_za) {
    for (int c = 0; c < 4; ++c) J(J_L1, epi, chain, c, t, (c == 0 && wait_za) ? F_WAIT_ZA : 0);
  };
  auto l2 = [&](int chain, int t, int epi, bool wait_a) {
    for (int c = 0; c < 4; ++c) J(J_L2, epi, chain, c, t, (c == 0 && wait_a) ? F_WAIT_A : 0);
  };
  if (mode == MODE_VALUE) {
    if (xr) {
      l1(CH_DYN, 0, E_H1, true);
      for (int t = 0; t < H; ++t) {
        const bool last = t == H - 1;
        // phase A: dynamics layer 2 (heavy) alternating with reward layer 1 (light)
        for (int c = 0; c < 4; ++c) {
          J(J_L2, E_XS, CH_DYN, c, t, c == 0 ? F_WAIT_A : 0);
          J(J_L1, E_XR, CH_REW, c, t, 0);
        }
        J(J_L3, E_Z, CH_DYN, 0, t, F_WAIT_A | (last ? F_LAST_STEP : 0));
        // phase B: reward layer 2 (heavy) alternating with the next chain's layer 1 (light), shifted by one job so
        // that the z' epilogue has two heavy jobs' worth of time before the first job that reads the new [z,a]
        const int nxt = last ? CH_PI : CH_DYN;
        const int tn = last ? 0 : t + 1;
        J(J_L2, E_DOT, CH_REW, 0, t, F_A_SMEM | F_WAIT_A);
        for (int c = 1; c < 4; ++c) {
          J(J_L2, E_DOT, CH_REW, c, t, F_A_SMEM);
          J(J_L1, E_H1, nxt, c - 1, tn, c == 1 ? F_WAIT_ZA : 0);
        }
        J(J_L1, E_H1, nxt, 3, tn, 0);
      }
    } else {
      for (int t = 0; t < H; ++t) {
        l1(CH_DYN, t, E_H1, true);
        l2(CH_DYN, t, E_XS, true);
        l1(CH_REW, t, E_H1, false);
        J(J_L3, E_Z, CH_DYN, 0, t, F_WAIT_A | (t == H - 1 ? F_LAST_STEP : 0));
        l2(CH_REW, t, E_DOT, true);
      }
      l1(CH_PI, 0, E_H1, true);
    }
    l2(CH_PI, 0, E_XS, true);
    J(J_L3, E_PI, CH_PI, 0, 0, F_WAIT_A | F_TERMINAL);
    for (int qh = CH_Q1; qh <= CH_Q2; ++qh) {
      l1(qh, 0, E_STAT, qh == CH_Q1);      // the terminal policy rewrote the action columns of [z,a]
      l1(qh, 0, E_QN, false);
      l2(qh, 0, E_STAT, true);
      l2(qh, 0, E_QDOT, false);            // same h1 as the statistics pass: nothing new to wait for
    }
  } else {
    for (int t = 0; t < H; ++t) {
      const bool last = t == H - 1;
      l1(CH_PI, t, E_H1, true);
      l2(CH_PI, t, E_XS, true);
      J(J_L3, E_PI, CH_PI, 0, t, F_WAIT_A | (last ? F_LAST_STEP : 0));
      if (!last) {                      // the reference also steps z after the last action; its result is unused
        l1(CH_DYN, t, E_H1, true);
        l2(CH_DYN, t, E_XS, true);
        J(J_L3, E_Z, CH_DYN, 0, t, F_WAIT_A);
      }
    }
  }
  return j;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
const char* tc_supported(const tdmpc_model_cfg& c) {
  if (c.mlp_dim != 512) return "mlp_dim must be 512";
  if (c.latent_dim > 128) return "latent_dim must be <= 128";
  if (c.latent_dim + c.action_dim > 192) return "latent_dim + action_dim must be <= 192";
  return nullptr;
}

static int tc_tiles_total(const tdmpc_model_cfg& c, int* nkb1_out, int l3rows[CH_COUNT]) {
  const int K1p = (c.latent_dim + c.action_dim + 15) & ~15;
  const int nkb1 = (K1p + 63) / 64;
  *nkb1_out = nkb1;
  const int n3_dyn = (c.latent_dim + 15) & ~15, n3_pi = (c.action_dim + 15) & ~15;
  for (int i = 0; i < CH_COUNT; ++i) l3rows[i] = 0;
  l3rows[CH_DYN] = n3_dyn <= 64 ? 64 : 128;
  l3rows[CH_PI] = n3_pi <= 64 ? 64 : 128;
  int tiles = 0;
  for (int ch = 0; ch < CH_COUNT; ++ch) {
    tiles += 4 * nkb1 + 32;
    if (l3rows[ch]) tiles += l3rows[ch] == 128 ? 8 : 4;
  }
  return tiles;
}

static size_t jobs_bytes_total() {
  size_t n = 0;
  for (int mode = 0; mode < 2; ++mode)
    for (int H = 1; H <= TDMPC_MAX_H; ++H) n += (build_jobs(mode, H, true).size() + 3) & ~(size_t)3;   // the two orders have equal length
  return n * sizeof(uint32_t);
}

size_t tc_arena_bytes(const tdmpc_model_cfg& c) {
  int nkb1, l3rows[CH_COUNT];
  return (size_t)tc_tiles_total(c, &nkb1, l3rows) * TILE_BYTES + jobs_bytes_total() + 1024;
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

  const float* W1[CH_COUNT] = {w.dyn_w[0], w.rew_w[0], w.pi_w[0], w.q_w[0][0], w.q_w[1][0]};
  const float* W2[CH_COUNT] = {w.dyn_w[1], w.rew_w[1], w.pi_w[1], w.q_w[0][1], w.q_w[1][1]};
  const float* W3[CH_COUNT] = {w.dyn_w[2], nullptr, w.pi_w[2], nullptr, nullptr};
  const int K1[CH_COUNT] = {L + A, L + A, L, L + A, L + A};
  const int N3[CH_COUNT] = {L, 0, A, 0, 0};
  const LinearT* lin[CH_COUNT] = {m32.dyn, m32.rew, m32.pi, m32.q[0], m32.q[1]};

  int launches = 0;
  uint32_t tile = 0;
  for (int ch = 0; ch < CH_COUNT; ++ch) {
    TcChain& t = m.chain[ch];
    t.l1_tile = tile;
    if (launch_pack(W1[ch], 512, K1[ch], base + (size_t)tile * TILE_BYTES, 4, nkb1, 128, st) < 0) return -1;
    tile += 4 * nkb1; ++launches;
    t.l2_tile = tile;
    if (launch_pack(W2[ch], 512, 512, base + (size_t)tile * TILE_BYTES, 4, 8, 128, st) < 0) return -1;
    tile += 32; ++launches;
    t.l3_tile = tile;
    if (W3[ch]) {
      if (launch_pack(W3[ch], N3[ch], 512, base + (size_t)tile * TILE_BYTES, 1, 8, m.l3rows[ch], st) < 0) return -1;
      tile += m.l3rows[ch] == 128 ? 8 : 4; ++launches;
    }
    t.b1 = lin[ch][0].b; t.b2 = lin[ch][1].b; t.b3 = lin[ch][2].b;
    t.w3 = N3[ch] == 0 ? lin[ch][2].wt : nullptr;      // [512][1] transposed == the weight row itself
    if (ch >= CH_Q1) {
      for (int i = 0; i < 2; ++i) { t.ln_g[i] = m32.q_ln[ch - CH_Q1][i].g; t.ln_b[i] = m32.q_ln[ch - CH_Q1][i].b; }
    }
  }
  // shared-memory budget -> XR buffer (interleaved reward chain) or not, weight ring depth
  auto fixed_bytes = [&](int xr) {
    return XS_BYTES * (1 + xr) + nkb1 * TILE_BYTES + XCH_FLOATS * 4 + MAX_JOBS * 4 + (2 * MAX_STAGES + 32) * 8 + 16 + 1024;
  };
  int use_xr = (SMEM_LIMIT - fixed_bytes(1)) / TILE_BYTES >= 3 ? 1 : 0;
  if (const char* e = getenv("TDMPC_TC_XR")) use_xr = atoi(e) ? ((SMEM_LIMIT - fixed_bytes(1)) / TILE_BYTES >= 2 ? 1 : 0) : 0;   // experiments
  int stages = (SMEM_LIMIT - fixed_bytes(use_xr)) / TILE_BYTES;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages == 5) stages = 4;          // the producers split the ring evenly
  if (stages < 2) return -1;
  m.use_xr = use_xr;
  m.stages = stages;
  m.smem_bytes = (size_t)fixed_bytes(use_xr) + (size_t)stages * TILE_BYTES;
  if (cudaFuncSetAttribute(rollout_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.smem_bytes) != cudaSuccess) return -1;

  // job programs for every (mode, horizon)
  uint32_t* jdst = reinterpret_cast<uint32_t*>(base + (size_t)tiles * TILE_BYTES);
  std::vector<uint32_t> all;
  size_t off = 0;
  for (int mode = 0; mode < 2; ++mode)
    for (int H = 1; H <= TDMPC_MAX_H; ++H) {
      std::vector<uint32_t> j = build_jobs(mode, H, use_xr != 0);
      if ((int)j.size() > MAX_JOBS) return -1;
      m.jobs[mode][H] = jdst + off;
      m.njobs[mode][H] = (int)j.size();
      j.resize((j.size() + 3) & ~(size_t)3, 0);
      all.insert(all.end(), j.begin(), j.end());
      off += j.size();
    }
  // pageable-host source: the runtime stages the bytes before returning, so `all` may die here
  if (cudaMemcpyAsync(jdst, all.data(), all.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;

  *out = m;
  return launches;
}

static unsigned long long* g_tc_trace = nullptr;
void tc_set_trace(unsigned long long* dev_buf) { g_tc_trace = dev_buf; }
int tc_trace_slots() { return 3 * MAX_JOBS * 3; }

int launch_rollout_tc(const TcModel& m, const RolloutParams& p, int num_sms, cudaStream_t st) {
  if (p.H < 1 || p.H > TDMPC_MAX_H) return -1;
  KParams kp;
  kp.m = m; kp.p = p;
  kp.jobs = m.jobs[p.mode][p.H];
  kp.njobs = m.njobs[p.mode][p.H];
  kp.stages = m.stages;
  const long long total = (long long)p.B * p.T;
  kp.num_tiles = (int)((total + TILE_ROWS - 1) / TILE_ROWS);
  kp.trace = g_tc_trace;
  { const char* e = getenv("TDMPC_TC_DEBUG"); kp.debug = e ? atoi(e) : 0; }
  // CTAs per cluster sharing one multicast weight stream (experiments: the L2 -> SM path is not the bottleneck,
  // so the default is 1).
  int C = 1;
  if (const char* e = getenv("TDMPC_TC_CLUSTER")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) C = v; }
  if (kp.num_tiles < C) C = 1;
  kp.cluster = C;
  kp.producers = m.stages % 3 == 0 ? 3 : (m.stages % 2 == 0 ? 2 : 1);
  if (const char* e = getenv("TDMPC_TC_PRODUCERS")) { const int v = atoi(e); if (v >= 1 && v <= 3 && m.stages % v == 0) kp.producers = v; }
  const int groups = (kp.num_tiles + C - 1) / C;
  const int max_clusters = num_sms / C;
  const int grid = (groups < max_clusters ? groups : max_clusters) * C;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = m.smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, rollout_tc_kernel, kp) != cudaSuccess) return -1;
  return 1;
}
