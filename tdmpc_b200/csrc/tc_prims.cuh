// sm_100a primitives used by the tcgen05 rollout kernel: mbarrier, bulk-TMA copies, TMEM
// allocation / load / store, UMMA shared-memory + instruction descriptors and the MMA issue
// wrappers (SS: A,B in shared memory; TS: A in tensor memory).  Inline PTX only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdio>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Explicit shared-space accesses: through a generic pointer the compiler emits LD.E / ST.E, which take the slower generic
// path and its queue (profiles: 'lg' throttle on the Q-head constants).
__device__ __forceinline__ float4 lds4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lds1(uint32_t saddr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr)); return v; }
__device__ __forceinline__ void sts1(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }

// ------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must end in a trap, never in a hung GPU.  The diagnostic printf is a
// debug-build option only (-DTDMPC_DEBUG_HANG): any real function call in the kernel makes ptxas
// apply the smallest setmaxnreg budget to ALL code and spill the epilogue.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
#ifdef TDMPC_DEBUG_HANG
      printf("tdmpc_b200: mbarrier wait timed out (tag %d, smem 0x%x, parity %u, block %d, thread %d)\n", tag, smem_u32(bar),
             parity, (int)blockIdx.x, (int)threadIdx.x);
#endif
      asm volatile("trap;");
    }
  }
}

// ------------------------------------------------------------------------------- fences
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------- bulk TMA (global -> shared)
// One contiguous blob (the weight tiles are pre-swizzled in global memory, so no tensor map is
// needed); completion is signalled on the mbarrier as transaction bytes.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------- TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(NCOLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {         // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS));
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
// tcgen05.wait::ld with the destination registers as in/out operands, so neither nvcc nor ptxas can
// schedule a consumer of v[] above the wait (the loads complete asynchronously).
__device__ __forceinline__ void tmem_wait_ld_dep(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
      : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
        "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
        "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
        "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
      :: "memory");
}
// Streams `nsub` consecutive 32-column groups of this warp's TMEM lanes through f(s, v[32]) with the
// load of group s+1 in flight while group s is processed (two register buffers).
template <typename F>
__device__ __forceinline__ void tmem_stream(uint32_t taddr, int nsub, F&& f) {
  uint32_t va[32], vb[32];
  tmem_ld32(taddr, va);
#pragma unroll 1
  for (int s = 0; s < nsub; s += 2) {
    tmem_wait_ld_dep(va);
    if (s + 1 < nsub) tmem_ld32(taddr + (s + 1) * 32, vb);
    f(s, va);
    if (s + 1 < nsub) {
      tmem_wait_ld_dep(vb);
      if (s + 2 < nsub) tmem_ld32(taddr + (s + 2) * 32, va);
      f(s + 1, vb);
    }
  }
}
// Same, with an auxiliary per-group operand (bias vector slice) prefetched one group ahead as well.
template <typename LA, typename F>
__device__ __forceinline__ void tmem_stream_aux(uint32_t taddr, int nsub, float4 (&xa)[8], LA&& load_aux, F&& f) {
  uint32_t va[32], vb[32];
  float4 xb[8];
  tmem_ld32(taddr, va);                  // xa (group 0) was loaded by the caller before it waited for the accumulator
#pragma unroll 1
  for (int s = 0; s < nsub; s += 2) {
    tmem_wait_ld_dep(va);
    if (s + 1 < nsub) { tmem_ld32(taddr + (s + 1) * 32, vb); load_aux(s + 1, xb); }
    f(s, va, xa);
    if (s + 1 < nsub) {
      tmem_wait_ld_dep(vb);
      if (s + 2 < nsub) { tmem_ld32(taddr + (s + 2) * 32, va); load_aux(s + 2, xa); }
      f(s + 1, vb, xb);
    }
  }
}
// Early-release variant for one 128-column accumulator: all four 32-column groups are pulled into registers first, then
// `release()` hands the accumulator back to the MMA warp (which can start the job after next while this one's activation
// math is still running), then f(s, v, aux) runs on the register copy with the aux operand prefetched one group ahead.
template <typename R, typename LA, typename F>
__device__ __forceinline__ void tmem_drain_aux(uint32_t taddr, float4 (&xa)[8], R&& release, LA&& load_aux, F&& f) {
  uint32_t v0[32], v1[32], v2[32], v3[32];
  float4 xb[8];
  tmem_ld32(taddr, v0);
  tmem_ld32(taddr + 32, v1);
  tmem_ld32(taddr + 64, v2);
  tmem_ld32(taddr + 96, v3);
  tmem_wait_ld_dep(v0); tmem_wait_ld_dep(v1); tmem_wait_ld_dep(v2); tmem_wait_ld_dep(v3);
  release();
  load_aux(1, xb);
  f(0, v0, xa);
  load_aux(2, xa);
  f(1, v1, xb);
  load_aux(3, xb);
  f(2, v2, xa);
  f(3, v3, xb);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
// named barrier 1: the 256 threads of the two epilogue warpgroups
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}

// ------------------------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, K-major operand stored as rows of 64 bf16 (128 B) with the
// 128-byte swizzle: 8-row groups are 1024 B apart (SBO); LBO is unused for swizzled K-major.
// bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr_bytes & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Same descriptor split into a constant high word and an address-dependent low word, so the MMA
// issue loop only does one integer add per operand per k-step (+2 = 32 bytes = UMMA_K bf16).
constexpr uint32_t UMMA_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr_bytes) { return ((saddr_bytes & 0x3FFFF) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t umma_desc_make(uint32_t lo) { return ((uint64_t)UMMA_DESC_HI << 32) | lo; }

// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, both operands K-major.
// [4,6) D fmt=1 (f32) | [7,10) A fmt=1 (bf16) | [10,13) B fmt=1 (bf16) | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T      (single thread issues)
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T      (A: lane = row, 32-bit column = two consecutive bf16 of K)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once every previously issued MMA of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Predicated forms for a CONVERGED issuing warp: every lane executes the statement, the elected lane's predicate sits inside
// the asm.  The operands are then computed in uniform control flow, so ptxas keeps descriptors and TMEM addresses in uniform
// registers (UMOV / UIADD) instead of moving them there with an R2UR per operand inside a divergent `if (leader)` region:
// tools/mma_rate.cu measures 94 cycles per TS MMA (N = 128) for the divergent form and 64 (the pipe's rate) for this one.
__device__ __forceinline__ void umma_ss_p(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t lead) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(lead) : "memory");
}
__device__ __forceinline__ void umma_ts_p(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t lead) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(lead) : "memory");
}
__device__ __forceinline__ void umma_commit_p(uint64_t* bar, uint32_t lead) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)), "r"(lead) : "memory");
}

// ------------------------------------------------------------------------------- layout helpers
// Byte offset of element (row, k) inside one [rows x 64] bf16 K-block stored with the 128 B swizzle.
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

}  // namespace tc
