// Diagnostics: one split-bf16 tcgen05 GEMM tile through the exact descriptor / swizzle / TMEM-load code of the rollout
// kernel.  C[128, n] = A[128, k] * B[n, k]^T, fp32 in / out, k % 64 == 0 (k <= 256), n % 16 == 0 (n <= 128).
#pragma once
#include "umma.cuh"

namespace l2a {

__global__ void __launch_bounds__(128, 1) debug_umma_tile_kernel(const float* A, const float* B, float* C, int n, int k,
                                                                 int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nkc = k / 64;
  const int b_chunk = n * 128;
  uint8_t* a_hi = smem;                                  // nkc tiles of 16 KB
  uint8_t* a_lo = a_hi + (size_t)nkc * 16384;
  uint8_t* b_hi = a_lo + (size_t)nkc * 16384;            // nkc chunks of n*128 B (1024-aligned when n % 8 == 0)
  uint8_t* b_lo = b_hi + (size_t)nkc * b_chunk;
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + (size_t)nkc * b_chunk);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int idx = tid; idx < 128 * k; idx += 128) {
    const int r = idx / k, c = idx % k;
    uint16_t hi, lo;
    umma::split_bf16(A[idx], hi, lo);
    const uint32_t off = (uint32_t)(c >> 6) * 16384u + umma::sw128_offset(r, c & 63);
    *reinterpret_cast<uint16_t*>(a_hi + off) = hi;
    *reinterpret_cast<uint16_t*>(a_lo + off) = lo;
  }
  for (int idx = tid; idx < n * k; idx += 128) {
    const int r = idx / k, c = idx % k;
    uint16_t hi, lo;
    umma::split_bf16(B[idx], hi, lo);
    const uint32_t off = (uint32_t)(c >> 6) * b_chunk + umma::sw128_offset(r, c & 63);
    *reinterpret_cast<uint16_t*>(b_hi + off) = hi;
    *reinterpret_cast<uint16_t*>(b_lo + off) = lo;
  }
  if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int passes = (variant & 4) ? 1 : 3;              // variant bit 2: single bf16 pass (hi*hi only)
  if (tid == 0 && (variant & 16)) {
    // A operand staged into tensor memory with tcgen05.cp, then TS-mode MMAs (A from TMEM, B from shared memory)
    const uint32_t idesc = umma::make_idesc_bf16(128, (uint32_t)n);
    const uint32_t a_hi_t = tmem_base + 128u, a_lo_t = tmem_base + 128u + (uint32_t)(nkc * 32);
    for (int kc = 0; kc < nkc; ++kc)
      for (int ks = 0; ks < 4; ++ks) {
        umma::tmem_cp_128x256b(a_hi_t + (uint32_t)(kc * 32 + ks * 8), umma::desc_lo32(umma::smem_u32(a_hi) + kc * 16384 + ks * 32));
        umma::tmem_cp_128x256b(a_lo_t + (uint32_t)(kc * 32 + ks * 8), umma::desc_lo32(umma::smem_u32(a_lo) + kc * 16384 + ks * 32));
      }
    uint32_t acc = 0;
    for (int kc = 0; kc < nkc; ++kc)
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t bh = umma::desc_lo32(umma::smem_u32(b_hi) + kc * b_chunk + ks * 32);
        const uint32_t bl = umma::desc_lo32(umma::smem_u32(b_lo) + kc * b_chunk + ks * 32);
        umma::mma_bf16_ts_lo(tmem_base, a_hi_t + (uint32_t)(kc * 32 + ks * 8), bh, idesc, acc);
        acc = 1;
        if (passes == 3) {
          umma::mma_bf16_ts_lo(tmem_base, a_hi_t + (uint32_t)(kc * 32 + ks * 8), bl, idesc, 1u);
          umma::mma_bf16_ts_lo(tmem_base, a_lo_t + (uint32_t)(kc * 32 + ks * 8), bh, idesc, 1u);
        }
      }
    umma::mma_commit(bar);
  } else if (tid == 0) {
    const uint32_t idesc = umma::make_idesc_bf16(128, (uint32_t)n);
    uint32_t acc = 0;
    for (int kc = 0; kc < nkc; ++kc)
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ah = umma::make_desc_sw128(umma::smem_u32(a_hi) + kc * 16384 + ks * 32, variant & 3);
        const uint64_t al = umma::make_desc_sw128(umma::smem_u32(a_lo) + kc * 16384 + ks * 32, variant & 3);
        const uint64_t bh = umma::make_desc_sw128(umma::smem_u32(b_hi) + kc * b_chunk + ks * 32, variant & 3);
        const uint64_t bl = umma::make_desc_sw128(umma::smem_u32(b_lo) + kc * b_chunk + ks * 32, variant & 3);
        umma::mma_bf16_ss(tmem_base, ah, bh, idesc, acc);
        acc = 1;
        if (passes == 3) {
          umma::mma_bf16_ss(tmem_base, ah, bl, idesc, 1u);
          umma::mma_bf16_ss(tmem_base, al, bh, idesc, 1u);
        }
      }
    umma::mma_commit(bar);
  }
  umma::mbar_wait(bar, 0);
  umma::tc_fence_after();
  if (variant & 8) {
    // accumulator-fragment loads (the rollout epilogue's path): thread T holds lanes T/4, T/4+8 x column pairs
    const int lane = tid & 31;
    for (int half = 0; half < 2; ++half)
      for (int cb = 0; cb < n / 16; ++cb) {
        uint32_t r[8];
        umma::tmem_ld_16x256b_x2(tmem_base + ((uint32_t)(warp * 32 + half * 16) << 16) + (uint32_t)(cb * 16), r);
        umma::tmem_ld_wait();
        for (int j = 0; j < 2; ++j)
          for (int q = 0; q < 2; ++q) {
            const int row = warp * 32 + half * 16 + (lane >> 2) + q * 8;
            const int col = cb * 16 + j * 8 + 2 * (lane & 3);
            C[(size_t)row * n + col] = __uint_as_float(r[4 * j + 2 * q]);
            C[(size_t)row * n + col + 1] = __uint_as_float(r[4 * j + 2 * q + 1]);
          }
      }
  } else {
    for (int c16 = 0; c16 < n / 16; ++c16) {
      uint32_t r[16];
      umma::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c16 * 16), r);
      umma::tmem_ld_wait();
      for (int i = 0; i < 16; ++i) C[(size_t)tid * n + c16 * 16 + i] = __uint_as_float(r[i]);
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) { umma::tc_fence_after(); umma::tmem_dealloc<512>(tmem_base); }
}

}  // namespace l2a

namespace l2a {
// Diagnostics: pure weight-stream pipeline (TMA bulk copies of `tile_bytes` through an `stages`-deep mbarrier ring, the
// consumer only waits and releases, optionally holding each tile for `hold_cycles`).  Reports SM cycles per CTA.
__global__ void __launch_bounds__(128, 1) debug_stream_kernel(const uint8_t* blob, int n_tiles_per_pass, int passes, int stages,
                                                              int tile_bytes, int hold_cycles, int producers, int consumers,
                                                              long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * tile_bytes);
  uint64_t* empty = full + stages;
  const int tid = threadIdx.x;
  blob += (size_t)(blockIdx.x % 5) * (size_t)n_tiles_per_pass * tile_bytes;     // five "members", like the ensemble rollout
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { umma::mbar_init(&full[s], 1); umma::mbar_init(&empty[s], 1); }
    umma::fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  const int total = passes * n_tiles_per_pass;
  // producers: threads 0 and 64 (different warps), consumers: threads 32 and 96; with two of a kind, each takes every other tile
  if ((tid == 0) || (tid == 64 && producers == 2)) {
    const int me = tid / 64;
    int stage = 0, t = 0, turn = 0;
    uint32_t phase = 0;
    for (int i = 0; i < total; ++i) {
      if (turn == me) {
        umma::mbar_wait(&empty[stage], phase ^ 1u);
        umma::mbar_arrive_expect_tx(&full[stage], (uint32_t)tile_bytes);
        umma::bulk_g2s(smem + (size_t)stage * tile_bytes, blob + (size_t)t * tile_bytes, (uint32_t)tile_bytes, &full[stage]);
      }
      if (++turn == producers) turn = 0;
      if (++stage == stages) { stage = 0; phase ^= 1u; }
      if (++t == n_tiles_per_pass) t = 0;
    }
  } else if ((tid == 32) || (tid == 96 && consumers == 2)) {
    const int me = tid / 64;
    int stage = 0, turn = 0;
    uint32_t phase = 0;
    for (int i = 0; i < total; ++i) {
      if (turn == me) {
        umma::mbar_wait(&full[stage], phase);
        if (hold_cycles > 0) { const long long h0 = clock64(); while (clock64() - h0 < hold_cycles) {} }
        umma::mbar_arrive(&empty[stage]);
      }
      if (++turn == consumers) turn = 0;
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  }
  __syncthreads();
  if (tid == 0) cycles_out[blockIdx.x] = clock64() - t0;
}
}  // namespace l2a

namespace l2a {
// Diagnostics: tensor-pipe timing of one "tile pair" (the 12 split-bf16 MMAs of a [128 x 64] weight block against NC
// candidates) in four flavours, operands resident in shared memory (contents irrelevant):
//   mode 0: SS  (A and B from shared memory)                                  -- what rollout_tc_kernel does
//   mode 1: TS  (A pre-copied to TMEM once, only the MMAs are timed)
//   mode 2: CP  (only the 8 tcgen05.cp of the hi and lo tile per pair)
//   mode 3: CP+TS per pair, two TMEM staging slots (cp of pair p+1 issued before the MMAs of pair p)
//   mode 4: SS with the A-collector keep / reuse hints on the two W_hi passes           -- rollout_tc_kernel, hidden layers
//   mode 5-8: the output layer's swapped-role shape (M = 128 candidates, N = 32 / 48 features) with / without the hints
// cycles_out[0] = SM cycles for `iters` pairs.
__device__ uint8_t g_rate_stream_src[65536];        // source of the concurrent TMA stream of modes 9, 10

template <int NC>
__global__ void __launch_bounds__(128, 1) debug_mma_rate_kernel(int mode, int iters, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t rate_smem[];
  uint8_t* a_tiles = rate_smem;                       // 2 pairs x (hi 16 KB + lo 16 KB)
  uint8_t* b_hi = rate_smem + 4 * 16384;              // [NC x 64] chunk
  uint8_t* b_lo = b_hi + NC * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + NC * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (4 * 16384 + 2 * NC * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(rate_smem)[i] = 0x3c003c00u;
  if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = umma::make_idesc_bf16(128, NC);
  // modes 9 / 10 = mode 4 / mode 0 with a concurrent bulk-copy stream into two further 32 KB stages of this CTA's shared memory
  // (what the weight ring of the rollout kernel does beside the MMAs): shows whether the MMA operand reads and the TMA fill
  // writes compete for the shared-memory bandwidth.  cycles_out[2] = 32 KB copies completed while the MMAs ran.
  __shared__ uint64_t s_bars[2];
  __shared__ volatile int s_stop;
  // modes 11-13 = mode 4 with the per-pair handshake instructions of the rollout kernel's issuer on always-satisfied barriers:
  //   11: mbarrier wait + tcgen05.fence::after before the 12 MMAs and tcgen05.commit after them   12: commit only   13: wait + fence only
  // modes 14, 15 = mode 4 fed by a REAL 2 x 32 KB weight ring (producer thread in warp 2 streaming g_rate_stream_src, full / empty
  //   mbarriers, one commit per pair = the rollout kernel's hidden-layer loop in isolation); 15 = W_hi / W_lo halves with their own
  //   barriers and two producer threads (L2A_TC_SPLIT_RING).
  __shared__ uint64_t s_ring[8];                      // full[2], empty[2], full2[2], empty2[2]
  __shared__ uint64_t s_dummy[2];                     // [0]: never arrived (parity 1 is always complete), [1]: commit sink
  const int hs_mode = (mode >= 11 && mode <= 13) ? mode : 0;
  const int ring_mode = (mode == 14 || mode == 15) ? mode : 0;
  if (hs_mode || ring_mode) {
    if (tid == 0) {
      for (int i = 0; i < 8; ++i) umma::mbar_init(&s_ring[i], 1);
      umma::mbar_init(&s_dummy[0], 1);
      umma::mbar_init(&s_dummy[1], 1);
      umma::fence_barrier_init();
    }
    __syncthreads();
    mode = 4;
  }
  if (ring_mode && (warp == 2 || (warp == 3 && ring_mode == 15)) && (tid & 31) == 0) {
    // producer(s): stage s of the ring IS a_tiles + s * 32 KB (the MMA operands themselves)
    const bool second = (warp == 3);
    uint64_t* fullb = second ? &s_ring[4] : &s_ring[0];
    uint64_t* emptyb = second ? &s_ring[6] : &s_ring[2];
    const uint32_t bytes = (ring_mode == 15) ? 16384u : 32768u;
    const uint32_t off = second ? 16384u : 0u;
    uint32_t phase = 0;
    int st = 0;
    for (int it = 0; it < iters; ++it) {
      umma::mbar_wait(&emptyb[st], phase ^ 1u);
      umma::mbar_arrive_expect_tx(&fullb[st], bytes);
      umma::bulk_g2s(a_tiles + st * 32768 + off, g_rate_stream_src + st * 32768 + off, bytes, &fullb[st]);
      if (++st == 2) { st = 0; phase ^= 1u; }
    }
  }
  const bool streaming = (mode == 9 || mode == 10);
  if (streaming) {
    if (tid == 0) { umma::mbar_init(&s_bars[0], 1); umma::mbar_init(&s_bars[1], 1); umma::fence_barrier_init(); s_stop = 0; }
    __syncthreads();
    if (mode == 9) mode = 4; else mode = 0;
  }
  if (streaming && warp == 2 && (tid & 31) == 0) {
    uint8_t* dst = reinterpret_cast<uint8_t*>(bar) + 1024;       // 64 KB behind the probe's own buffers (1024-byte aligned below)
    dst = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dst) + 1023) & ~(uintptr_t)1023);
    long long copies = 0;
    uint32_t ph[2] = {0, 0};
    umma::mbar_arrive_expect_tx(&s_bars[0], 32768);
    umma::bulk_g2s(dst, g_rate_stream_src, 32768, &s_bars[0]);
    umma::mbar_arrive_expect_tx(&s_bars[1], 32768);
    umma::bulk_g2s(dst + 32768, g_rate_stream_src + 32768, 32768, &s_bars[1]);
    int sidx = 0;
    while (!s_stop) {
      umma::mbar_wait(&s_bars[sidx], ph[sidx]);
      ph[sidx] ^= 1u;
      ++copies;
      umma::mbar_arrive_expect_tx(&s_bars[sidx], 32768);
      umma::bulk_g2s(dst + sidx * 32768, g_rate_stream_src + sidx * 32768, 32768, &s_bars[sidx]);
      sidx ^= 1;
    }
    umma::mbar_wait(&s_bars[0], ph[0]);
    umma::mbar_wait(&s_bars[1], ph[1]);
    cycles_out[2] = copies;
  }
  if (warp == 1) {
    const uint32_t a0 = umma::desc_lo32(umma::smem_u32(a_tiles)), bh = umma::desc_lo32(umma::smem_u32(b_hi)), bl = umma::desc_lo32(umma::smem_u32(b_lo));
    const uint32_t stage_step = 32768u >> 4, lo_step = 16384u >> 4;
    const uint32_t t_stage = tmem_base + 384u;        // two staging slots of 64 columns (hi 32 + lo 32)
    long long t0 = 0;
    if (mode == 1 && umma::elect_one()) {
      for (int ks = 0; ks < 4; ++ks) {
        umma::tmem_cp_128x256b(t_stage + ks * 8, a0 + 2 * ks);
        umma::tmem_cp_128x256b(t_stage + 32 + ks * 8, a0 + lo_step + 2 * ks);
      }
    }
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t st = (uint32_t)(it & 1);
      const uint32_t a_hi = a0 + st * stage_step, a_lo = a_hi + lo_step;
      const uint32_t d = tmem_base + (uint32_t)((it & 3) * NC);
      if (hs_mode == 11 || hs_mode == 13) { umma::mbar_wait(&s_dummy[0], 1u); umma::tc_fence_after(); }
      if (ring_mode) { umma::mbar_wait(&s_ring[st], (uint32_t)((it >> 1) & 1)); umma::tc_fence_after(); }
      if (umma::elect_one()) {
        if (mode == 0) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma::mma_bf16_ss_lo(d, a_hi + 2 * ks, bh + 2 * ks, idesc, 1u);
            umma::mma_bf16_ss_lo(d, a_hi + 2 * ks, bl + 2 * ks, idesc, 1u);
            umma::mma_bf16_ss_lo(d, a_lo + 2 * ks, bh + 2 * ks, idesc, 1u);
          }
        } else if (mode == 4 && ring_mode != 15) {
          // SS with A-collector hints: W_hi is read from shared memory once for its two passes
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d, a_hi + 2 * ks, bh + 2 * ks, idesc, 1u);
            umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d, a_hi + 2 * ks, bl + 2 * ks, idesc, 1u);
            umma::mma_bf16_ss_lo(d, a_lo + 2 * ks, bh + 2 * ks, idesc, 1u);
          }
          if (hs_mode == 11 || hs_mode == 12) umma::mma_commit(&s_dummy[1]);
          if (ring_mode == 14) umma::mma_commit(&s_ring[2 + st]);
        } else if (mode == 4) {
          // split ring: the two W_hi passes, release the W_hi half, then (below, after the W_lo wait) the W_lo pass
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d, a_hi + 2 * ks, bh + 2 * ks, idesc, 1u);
            umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d, a_hi + 2 * ks, bl + 2 * ks, idesc, 1u);
          }
          umma::mma_commit(&s_ring[2 + st]);
        } else if (mode >= 5 && mode <= 8) {
          // output-layer shape (roles swapped): M = 128 candidates (A = a 16 KB activation chunk), N = 32 (modes 5, 6) or
          // 48 (modes 7, 8) features; odd modes with the A-collector hints, even modes without
          const uint32_t idesc_o = umma::make_idesc_bf16(128, mode <= 6 ? 32u : 48u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (mode & 1) {
              umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d, a_hi + 2 * ks, bh + 2 * ks, idesc_o, 1u);
              umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d, a_hi + 2 * ks, bl + 2 * ks, idesc_o, 1u);
            } else {
              umma::mma_bf16_ss_lo(d, a_hi + 2 * ks, bh + 2 * ks, idesc_o, 1u);
              umma::mma_bf16_ss_lo(d, a_hi + 2 * ks, bl + 2 * ks, idesc_o, 1u);
            }
            umma::mma_bf16_ss_lo(d, a_lo + 2 * ks, bh + 2 * ks, idesc_o, 1u);
          }
        } else if (mode == 1) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma::mma_bf16_ts_lo(d, t_stage + ks * 8, bh + 2 * ks, idesc, 1u);
            umma::mma_bf16_ts_lo(d, t_stage + ks * 8, bl + 2 * ks, idesc, 1u);
            umma::mma_bf16_ts_lo(d, t_stage + 32 + ks * 8, bh + 2 * ks, idesc, 1u);
          }
        } else {
          const uint32_t ts_cur = t_stage + st * 64u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma::tmem_cp_128x256b(ts_cur + ks * 8, a_hi + 2 * ks);
            umma::tmem_cp_128x256b(ts_cur + 32 + ks * 8, a_lo + 2 * ks);
          }
          if (mode == 3) {
            const uint32_t ts_prev = t_stage + (st ^ 1u) * 64u;     // MMAs of the previous pair (its cp was issued one iteration ago)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma::mma_bf16_ts_lo(d, ts_prev + ks * 8, bh + 2 * ks, idesc, 1u);
              umma::mma_bf16_ts_lo(d, ts_prev + ks * 8, bl + 2 * ks, idesc, 1u);
              umma::mma_bf16_ts_lo(d, ts_prev + 32 + ks * 8, bh + 2 * ks, idesc, 1u);
            }
          }
        }
      }
      __syncwarp();
      if (ring_mode == 15) {
        umma::mbar_wait(&s_ring[4 + st], (uint32_t)((it >> 1) & 1));
        umma::tc_fence_after();
        if (umma::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma::mma_bf16_ss_lo(d, a_lo + 2 * ks, bh + 2 * ks, idesc, 1u);
          umma::mma_commit(&s_ring[6 + st]);
        }
        __syncwarp();
      }
    }
    const long long t_issued = clock64();             // every MMA has been accepted by the tensor pipe's queue
    if (umma::elect_one()) umma::mma_commit(bar);
    __syncwarp();
    umma::mbar_wait(bar, 0);
    if ((tid & 31) == 0) { cycles_out[0] = clock64() - t0; cycles_out[1] = t_issued - t0; s_stop = 1; }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) { umma::tc_fence_after(); umma::tmem_dealloc<512>(tmem_base); }
}
}  // namespace l2a

namespace l2a {
// Diagnostics for the next step of K1: a CTA PAIR issuing tcgen05.mma.cta_group::2 (M = 256 across the two
// CTAs, each CTA's shared memory supplies its 128 weight rows and its NC candidate rows of B; N = 2*NC), and the DSMEM bulk
// copy the pair's epilogue all-to-all would use.
//   mode 0: `iters` tile pairs (12 split-bf16 MMAs each, with the A-collector hints) -> cycles_out[0] = total, [1] = issue time
//   mode 1: `iters` bulk copies of `copy_bytes` from this CTA's shared memory into the peer's (cp.async.bulk.shared::cluster,
//           complete_tx on the peer's mbarrier), both directions at once -> cycles_out[2 + rank] = cycles seen by the receiver
// Launch with a cluster of 2.  Operand contents are irrelevant (timing only).
template <int NC>
__global__ void __launch_bounds__(128, 1) debug_pair_kernel(int mode, int iters, int copy_bytes, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t pair_smem[];
  uint8_t* a_tiles = pair_smem;                       // 2 stages x (hi 16 KB + lo 16 KB): this CTA's 128 rows of the 256-row tile
  uint8_t* b_hi = pair_smem + 4 * 16384;              // [NC x 64] chunk: this CTA's half of the N = 2*NC columns
  uint8_t* b_lo = b_hi + NC * 128;
  uint8_t* xfer_src = b_lo + NC * 128;                // 16 KB source / 16 KB destination of the DSMEM copy test
  uint8_t* xfer_dst = xfer_src + 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(xfer_dst + 16384);   // [0] MMA completion, [1] incoming copy
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = umma::cluster_ctarank();
  for (int i = tid; i < (4 * 16384 + 2 * NC * 128 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(pair_smem)[i] = 0x3c003c00u;
  if (tid == 0) { umma::mbar_init(&bar[0], 1); umma::mbar_init(&bar[1], 1); umma::fence_barrier_init(); }
  umma::cluster_sync_all();
  if (warp == 0) {                                    // both CTAs of the pair take part in the 2-CTA allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  umma::cluster_sync_all();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (mode == 0) {
    if (rank == 0 && warp == 1) {                     // only the leader CTA issues; descriptors are CTA-local offsets valid in both
      constexpr uint32_t idesc = umma::make_idesc_bf16(256, 2 * NC);
      const uint32_t a0 = umma::desc_lo32(umma::smem_u32(a_tiles)), bh = umma::desc_lo32(umma::smem_u32(b_hi)), bl = umma::desc_lo32(umma::smem_u32(b_lo));
      const uint32_t stage_step = 32768u >> 4, lo_step = 16384u >> 4;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t a_hi = a0 + (uint32_t)(it & 1) * stage_step, a_lo = a_hi + lo_step;
        const uint32_t d = tmem_base + (uint32_t)((it & 1) * 2 * NC);
        if (umma::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\tsetp.ne.b32 p, 1, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::f16.collector::a::fill [%0], da, db, %3, p;\n\t}" ::"r"(d), "r"(a_hi + 2 * ks), "r"(bh + 2 * ks), "r"(idesc), "r"(umma::kDescHi32) : "memory");
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\tsetp.ne.b32 p, 1, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse [%0], da, db, %3, p;\n\t}" ::"r"(d), "r"(a_hi + 2 * ks), "r"(bl + 2 * ks), "r"(idesc), "r"(umma::kDescHi32) : "memory");
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\tsetp.ne.b32 p, 1, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d), "r"(a_lo + 2 * ks), "r"(bh + 2 * ks), "r"(idesc), "r"(umma::kDescHi32) : "memory");
          }
        }
        __syncwarp();
      }
      const long long t_issued = clock64();
      if (umma::elect_one())
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(umma::smem_u32(&bar[0])), "h"((uint16_t)1) : "memory");
      __syncwarp();
      umma::mbar_wait(&bar[0], 0);
      if ((tid & 31) == 0) { cycles_out[0] = clock64() - t0; cycles_out[1] = t_issued - t0; }
    }
  } else if (mode == 2 || mode == 3) {
    // DSMEM generic stores from registers (the pair kernel's epilogue): `copy_bytes` / 16 warps-wide 16-byte-per-lane
    // st.shared::cluster instructions per warp, 4 warps, into the peer's 32 KB xfer region (b-tile area reused), then one
    // cluster-scope release fence.  mode 2: a warp instruction covers 512 contiguous bytes; mode 3: four 128-byte pieces 1152
    // bytes apart (what the MN-major epilogue issues).  cycles_out[2 + rank] = cycles until the fence has completed.
    const uint32_t peer = rank ^ 1u;
    const uint32_t base = umma::map_to_cta(umma::smem_u32(pair_smem), peer);        // the peer's 64 KB weight-ring area
    const int lane = tid & 31;
    const uint4 v = make_uint4(tid, 1, 2, 3);
    umma::cluster_sync_all();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t slot = (uint32_t)((it * 4 + warp) & 31);                          // 32 slots of 2 KB
      uint32_t off;
      if (mode == 2) off = slot * 2048u + (uint32_t)lane * 16u;
      else off = (slot & 7u) * 128u + (uint32_t)(lane >> 3) * 1152u + (uint32_t)(lane & 7) * 16u + (slot >> 3) * 4608u;
      umma::st_cluster_v4(base + off, v);
    }
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
    const long long t1 = clock64();
    if (tid == 0) cycles_out[2 + rank] = t1 - t0;
    umma::cluster_sync_all();
  } else {
    // each CTA pushes `iters` copies into the peer's xfer_dst; the receiver waits on its own barrier
    if (warp == 1 && (tid & 31) == 0) {
      const uint32_t peer = rank ^ 1u;
      const uint32_t dst = umma::map_to_cta(umma::smem_u32(xfer_dst), peer), rbar = umma::map_to_cta(umma::smem_u32(&bar[1]), peer);
      for (int it = 0; it < iters; ++it)
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "r"(umma::smem_u32(xfer_src)), "r"((uint32_t)copy_bytes), "r"(rbar) : "memory");
    }
    if (warp == 2 && (tid & 31) == 0) {
      const long long t0 = clock64();
      umma::mbar_arrive_expect_tx(&bar[1], (uint32_t)copy_bytes * (uint32_t)iters);
      umma::mbar_wait(&bar[1], 0);
      cycles_out[2 + rank] = clock64() - t0;
    }
  }
  umma::tc_fence_before();
  umma::cluster_sync_all();
  if (warp == 0) {
    umma::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}
}  // namespace l2a
