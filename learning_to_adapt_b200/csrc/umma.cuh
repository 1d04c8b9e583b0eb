// sm_100a building blocks: mbarrier, bulk-copy TMA (cp.async.bulk), tcgen05 alloc / mma / commit / ld, descriptors.
// Hand-written inline PTX; encodings follow the SM100 UMMA descriptor format (K-major, 128-byte swizzle).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace l2a {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
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
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking phase test (mbarrier.test_wait never suspends the thread)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Watchdog: a barrier that does not complete within ~4e9 SM cycles (>2 s) is a protocol bug; trap instead of
// hanging the GPU.
#ifndef L2A_WATCHDOG_CYCLES
#define L2A_WATCHDOG_CYCLES 4000000000ll
#endif
// (No printf by default: a call to a shared non-inlined function from warp-specialised branches makes ptxas allocate every
// branch with the smallest setmaxnreg budget of the kernel.  -DL2A_WATCHDOG_PRINTF restores the message for debugging.)
#ifdef L2A_WATCHDOG_PRINTF
__device__ __noinline__ void watchdog_fail(uint32_t bar_addr, uint32_t parity) {
  printf("l2a_b200 watchdog: mbarrier 0x%x parity %u never completed (block %d thread %d)\n", bar_addr, parity,
         (int)blockIdx.x, (int)threadIdx.x);
  __trap();
}
#else
__device__ __forceinline__ void watchdog_fail(uint32_t, uint32_t) { __trap(); }
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > L2A_WATCHDOG_CYCLES) watchdog_fail(smem_u32(bar), parity);
  }
}
// cluster-scope variants (peer CTA barriers addressed through mapa)
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Remote arrive WITHOUT a cluster-scope release (no MEMBAR.ALL.GPU, ~1.5 k cycles on B200): for arrivals that only certify
// writes to the arriving CTA's OWN shared memory, which a preceding fence.proxy.async.shared::cta (MEMBAR.ALL.CTA +
// FENCE.VIEW.ASYNC.S) has already performed -- shared memory has one copy, so CTA-scope performed is performed for the peer's
// tensor-core reads too -- or data whose arrival is tracked by an mbarrier transaction count.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > L2A_WATCHDOG_CYCLES) watchdog_fail(smem_u32(bar), parity);
  }
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr));
  return v;
}
__device__ __forceinline__ uint32_t ld_dsmem_u32(uint32_t cluster_addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr));
  return v;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// One lane of the (converged) warp; ptxas knows the guarded region runs on a single lane, so tcgen05 instructions in it
// compile to one uniform UTC* instruction instead of a per-lane "waterfall" loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- TMA bulk copy global -> shared (UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------- TMEM
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {      // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (lane_base + i), columns [col, col+16)
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 8 consecutive fp32 columns (thread i: lane lane_base + i, columns [col, col+8)), into r[0..8)
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// 16 lanes x (8 x REP) fp32 columns in the mma-accumulator fragment layout: for column block j (8 columns),
// thread T holds  r[4j+0], r[4j+1] = (lane T/4,     columns 8j + 2(T%4), +1)
//                 r[4j+2], r[4j+3] = (lane T/4 + 8, columns 8j + 2(T%4), +1)        (lanes relative to the address lane)
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// four 8x8 b16 matrices, transposed store: with .trans thread T's register i holds elements (row 2(T%4), col T/4) [low
// half] and (row 2(T%4)+1, col T/4) [high half] of matrix i; lane 8i + r supplies the address of row r of matrix i.
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t smem_addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(smem_addr), "r"(r0), "r"(r1),
               "r"(r2), "r"(r3)
               : "memory");
}
// (hi, lo) bf16x2 split of two floats: hi = {bf16(v1), bf16(v0)} (v0 in the low half), lo likewise for the residuals
__device__ __forceinline__ void split_bf16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v1 - h1), "f"(v0 - h0));
}
// Same split for the pair (a0 + b, a1 + b) after ReLU, with the packed fp32x2 adder (FADD2) for the bias add and the residual
// subtraction: 8 instead of 10 instructions per two elements, bit-identical results (IEEE adds either way).
__device__ __forceinline__ void bias_relu_split_bf16x2(uint32_t a0, uint32_t a1, float bias, uint32_t& hi, uint32_t& lo) {
  uint64_t v, b2, h2, d;
  uint32_t v0, v1;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(a0), "r"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(bias));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(v), "l"(b2));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(v0), "=r"(v1) : "l"(v));
  const float r0 = fmaxf(__uint_as_float(v0), 0.f), r1 = fmaxf(__uint_as_float(v1), 0.f);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(r1), "f"(r0));
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(r0), "f"(r1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(h2) : "r"(hi << 16), "r"(hi & 0xffff0000u));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(v), "l"(h2));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(v0), "=r"(v1) : "l"(d));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(__uint_as_float(v1)), "f"(__uint_as_float(v0)));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B: rows of 128 bytes (64 bf16), 8-row groups of
// 1024 bytes (stride byte offset), start address / LBO / SBO in 16-byte units, version = 1 (sm_100),
// layout type 2 = SWIZZLE_128B at bits [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t variant = 0) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);                 // start address   bits [0,14)
  const uint64_t lbo = (variant & 1u) ? 0ull : 1ull;            // ignored by HW for swizzled K-major
  d |= lbo << 16;                                               // leading byte off bits [16,30)
  d |= (uint64_t)(1024u >> 4) << 32;                            // stride byte off  bits [32,46)
  if (!(variant & 2u)) d |= 1ull << 46;                         // descriptor version (Blackwell)
  d |= 2ull << 61;                                              // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: BF16 x BF16 -> FP32, both operands K-major, M x N tile.
__device__ __host__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the descriptors given as their low 32-bit words (start address / LBO fields); the high word (SBO = 1024 B,
// version 1, SWIZZLE_128B) is the constant kDescHi32.  Lets the issue loop advance along K with one 32-bit add.
constexpr uint32_t kDescHi32 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo32(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void mma_bf16_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi32)
      : "memory");
}
// Same with a collector hint for the A operand: kAKeep = this MMA's A tile stays in the collector buffer for the next MMA
// (.collector::a::fill -> SASS UTCHMMA ...A_KEEP); kAReuse = this MMA takes A from the collector instead of re-reading shared
// memory (.collector::a::lastuse -> A_REUSE).  The caller guarantees that a kAReuse MMA directly follows a kAKeep MMA with
// the identical A descriptor.  An SS-mode MMA at N = 80 is shared-memory-bandwidth bound (4 KB of A + 2.5 KB of B per 40
// tensor cycles), so not re-reading W_hi for the W_hi * x_lo pass removes a fifth of the operand traffic.
enum : int { kANone = 0, kAKeep = 1, kAReuse = 2 };
template <int HINT>
__device__ __forceinline__ void mma_bf16_ss_lo_hint(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  if (HINT == kAKeep) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi32)
        : "memory");
  } else if (HINT == kAReuse) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi32)
        : "memory");
  } else {
    mma_bf16_ss_lo(d_tmem, a_lo, b_lo, idesc, accumulate);
  }
}
// A operand from TENSOR MEMORY (128 lanes x 8 columns per K=16 slice), B from shared memory.
__device__ __forceinline__ void mma_bf16_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi32)
      : "memory");
}
// shared memory -> tensor memory copy of a [128 rows x 256 bit] slice (= one K=16 bf16 A-operand slice), source given by
// the same matrix descriptor an SS-mode MMA would use.  Ordered with the issuing thread's other tcgen05.mma / cp ops.
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t dst_tmem, uint32_t src_lo) {
  asm volatile(
      "{\n\t.reg .b64 ds;\n\t"
      "mov.b64 ds, {%1, %2};\n\t"
      "tcgen05.cp.cta_group::1.128x256b [%0], ds;\n\t}" ::"r"(dst_tmem),
      "r"(src_lo), "r"(kDescHi32)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- CTA pair (cta_group::2): M = 256 across two CTAs of a cluster
// Both CTAs of the pair run the alloc / dealloc (same warp index, same smem slot offset); only the leader (cluster rank 0)
// issues MMAs and commits.  Every tcgen05 instruction of a kernel uses the same cta_group.
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot) {  // whole warp, in BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {     // whole warp, in BOTH CTAs
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256: each CTA's shared memory supplies its 128 rows of A and its N/2 rows of B
// (the descriptors are CTA-local offsets, identical in both CTAs); CTA r's TMEM receives rows [128 r, 128 r + 128) x N.
template <int HINT>
__device__ __forceinline__ void mma2_bf16_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  if (HINT == kAKeep) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16.collector::a::fill [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi32)
        : "memory");
  } else if (HINT == kAReuse) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi32)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi32)
        : "memory");
  }
}
// Same with the high descriptor words of A and B given separately (operands in different canonical layouts, e.g. K-major
// SWIZZLE_128B weights against MN-major unswizzled activations).
template <int HINT>
__device__ __forceinline__ void mma2_bf16_ss_ab(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate) {
  if (HINT == kAKeep) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %6};\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16.collector::a::fill [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(a_hi), "r"(b_hi)
        : "memory");
  } else if (HINT == kAReuse) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %6};\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(a_hi), "r"(b_hi)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %6};\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(a_hi), "r"(b_hi)
        : "memory");
  }
}
// MN-major (M / N contiguous) unswizzled operand of 8x8 core matrices (128 contiguous bytes each): SBO = 128 B between 8-row groups
// along M / N, LBO = `lbo_bytes` between 8-column groups along K.  High word: SBO, descriptor version 1, layout type 0.
constexpr uint32_t kDescHi32Mn = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo32_mn(uint32_t smem_addr, uint32_t lbo_bytes) { return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); }
constexpr uint32_t kIdescAMn = 1u << 15, kIdescBMn = 1u << 16;     // instruction-descriptor bits: A / B operand is MN-major
// all MMAs issued so far by this thread arrive, when complete, on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void mma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// 2-D tiled TMA load into THIS CTA's shared memory whose byte count completes on `mbar_cluster` (a shared::cluster address:
// the leader CTA's barrier for both CTAs of the pair)
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const void* tmap, int32_t c0, int32_t c1, uint32_t mbar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(mbar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// Asynchronous 16-byte store into a cluster peer's shared memory whose completion is counted (16 bytes of complete_tx) on an
// mbarrier of the SAME peer CTA: the sender needs no fence, the receiver sees the data once the barrier's phase completes.
__device__ __forceinline__ void st_async_cluster_v4(uint32_t cluster_addr, uint4 v, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_mbar)
               : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t smem_addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t smem_addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr) : "memory");
  return v;
}
// generic-proxy writes to shared memory of this CTA or of a cluster peer -> visible to the async proxy (tensor core reads).
// (The unqualified fence.proxy.async also covers global memory and compiles to MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC.)
__device__ __forceinline__ void fence_proxy_async_cluster() { asm volatile("fence.proxy.async.shared::cluster;" ::: "memory"); }
// arrive with cluster-scope release on a barrier of THIS CTA (orders this thread's -- and, through a preceding __syncwarp /
// bar.sync, its warp's / CTA's -- stores into the peer's shared memory before the arrival)
__device__ __forceinline__ void mbar_arrive_release_cluster(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_relaxed_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 16-byte relaxed gpu-scope store / load (each 8-byte half is single-copy atomic: the {value, flag} words of the member exchange)
__device__ __forceinline__ void st_relaxed_gpu_v4(uint4* p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_gpu_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// global-memory flags of the member exchange between CTA pairs
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---------------------------------------------------------------- split-bf16 helpers
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits; three MMA passes hi*hi + hi*lo + lo*hi.
__device__ __forceinline__ void split_bf16(float x, uint16_t& hi, uint16_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(l);
}
// byte offset of element (row, col) inside a K-major SWIZZLE_128B tile whose rows are 64 bf16 (128 B)
__device__ __host__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t col) {
  return row * 128u + ((((col >> 3) ^ (row & 7u)) & 7u) << 4) + ((col & 7u) << 1);
}

}  // namespace umma
}  // namespace l2a
