// ReBAL on tensor cores: persistent fused H-step rollout through a single-layer LSTM dynamics model (tcgen05, sm_100a).
// Same building blocks and conventions as rollout_tc.cuh (transposed products, split-bf16 x3, 32 KB weight-pair ring, whole-warp
// MMA issuer, 16x256b fragments + stmatrix epilogue); what differs is the layer graph:
//   z[4Hs, cand] = Wk^T[4Hs, 64 + Hs] * [x ; h][64 + Hs, cand]        gates i | j | f | o  (TF LSTMCell column order)
//   c' = sigmoid(f + 1) * c + sigmoid(i) * tanh(j) ;  h' = sigmoid(o) * tanh(c')     (the four gates of a unit sit in the SAME TMEM
//        lane of four different accumulator blocks, so one thread combines them; c stays in shared memory as fp32 [unit][cand])
//   y[D, cand] = Wo^T[D, Hs] * h'[Hs, cand]   -> delta -> reward / state update -> next x
// Replaces policies/rnn_mpc_controller.py:112-134 + dynamics/rnn_dynamics.py:233-252 (cell: dynamics/core/utils.py:193-198).
#pragma once
#include "common.cuh"
#include "rollout_tc.cuh"
#include "umma.cuh"

namespace l2a {

constexpr int kRnnTcThreads = 192;
constexpr int kRnnTcStages = 2;
constexpr int kRnnTcMaxHs = 256;

struct RnnTcPlan {
  int hs, nmbg, nkc, nko, nks0, pairs_per_set;
};

inline bool rnn_tc_make_plan(int D, int A, int hs, RnnTcPlan* p) {
  if (hs != 128 && hs != 256) return false;
  if (D > kTcMaxObs || A > kTcMaxAct || tc_obs_pad(D) + A > 64 || D < 3) return false;
  p->hs = hs;
  p->nmbg = 4 * hs / 128;
  p->nkc = 1 + hs / 64;
  p->nko = hs / 64;
  p->nks0 = (tc_obs_pad(D) + A + 15) / 16;
  p->pairs_per_set = p->nmbg * p->nkc + p->nko;
  return true;
}

// grid.x = pairs; builds the (hi, lo) tile pair in consumption order: gates (mb-outer, kc-inner), then the output layer.
struct RnnPrepArgs {
  RnnTcPlan plan;
  int D, A;
  const float* wk;   // [(D+A+Hs), 4Hs]
  const float* wo;   // [Hs, D]
  uint8_t* blob;
};

__global__ void __launch_bounds__(256) rnn_tc_prep_kernel(const RnnPrepArgs a) {
  const RnnTcPlan& pl = a.plan;
  const int pr = blockIdx.x, hs = pl.hs, IN = a.D + a.A;
  const bool gates = pr < pl.nmbg * pl.nkc;
  const int mb = gates ? pr / pl.nkc : 0;
  const int kc = gates ? pr % pl.nkc : (pr - pl.nmbg * pl.nkc);
  uint8_t* tile_hi = a.blob + (size_t)pr * 2 * kTcTileBytes;
  uint8_t* tile_lo = tile_hi + kTcTileBytes;
  for (int item = threadIdx.x; item < 128 * 8; item += blockDim.x) {
    const int r = item & 127, ch = item >> 7;
    uint16_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = ch * 8 + i;
      float w = 0.f;
      if (gates) {
        const int f = mb * 128 + r;                                       // gate column (< 4 Hs always)
        int krow = (kc == 0) ? tc_in0_of_col(a.D, a.A, c) : IN + (kc - 1) * 64 + c;
        if (krow >= 0) w = a.wk[(size_t)krow * (4 * hs) + f];
      } else if (r < a.D) {
        w = a.wo[(size_t)(kc * 64 + c) * a.D + r];
      }
      umma::split_bf16(w, hi[i], lo[i]);
    }
    const uint32_t off = umma::sw128_offset(r, ch * 8);
    *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16), hi[4] | (hi[5] << 16), hi[6] | (hi[7] << 16));
    *reinterpret_cast<uint4*>(tile_lo + off) = make_uint4(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16), lo[4] | (lo[5] << 16), lo[6] | (lo[7] << 16));
  }
}

struct RnnTcArgs {
  RnnTcPlan plan;
  int D, A;
  NormDev norm;
  const float* bk;            // [4Hs]
  const float* bo;            // [D]
  const uint8_t* blob;
  const float* obs0;          // [m, D]
  const float* c0;            // [m, Hs]
  const float* h0;
  const float* actions;
  long long act_stride_t, act_stride_row;
  const float* discount_pow;
  int n_candidates, n_envs, horizon, reward_kind;
  float dt;
  int groups_per_env;
  float* returns;
  ReduceArgs red;
};

template <int NC>
struct RnnTcSmem {
  static constexpr int kChunkBytes = NC * 128;
  static constexpr int kMaxChunks = 1 + kRnnTcMaxHs / 64;
  static constexpr int kCP = NC + 4;                                    // c-state row pitch (floats)
  static constexpr int kNCP = NC + 1;
  static constexpr size_t act_bytes = (size_t)2 * kMaxChunks * kChunkBytes;
  static constexpr size_t stage_off = act_bytes;
  static constexpr size_t cst_off = stage_off + (size_t)kRnnTcStages * 2 * kTcTileBytes;
  static size_t total(int D, int A, int hs) {
    return cst_off + sizeof(float) * ((size_t)hs * kCP + (size_t)D * kNCP + 4 * (size_t)D + 2 * (size_t)A) + 64 + 16 * sizeof(uint64_t) + 64;
  }
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

template <int NC>
__global__ void __launch_bounds__(kRnnTcThreads, 1) rollout_rnn_tc_kernel(const RnnTcArgs a) {
  using S = RnnTcSmem<NC>;
  constexpr int kChunkBytes = S::kChunkBytes;
  constexpr int CP = S::kCP, NCP = S::kNCP;
  constexpr uint32_t kIdesc = umma::make_idesc_bf16(128, NC);
  constexpr int kStageBytes = 2 * kTcTileBytes;
  static_assert(NC % 16 == 0 && 8 * NC <= 512, "eight gate accumulator blocks must fit the 512 TMEM columns");

  extern __shared__ __align__(1024) uint8_t rnn_tc_smem[];
  uint8_t* const smem = rnn_tc_smem;
  if ((umma::smem_u32(smem) & 1023u) != 0) __trap();
  const RnnTcPlan& pl = a.plan;
  const int D = a.D, A = a.A, hs = pl.hs, H = a.horizon;
  uint8_t* act_hi = smem;
  uint8_t* act_lo = smem + (size_t)S::kMaxChunks * kChunkBytes;
  uint8_t* stages = smem + S::stage_off;
  float* cst = reinterpret_cast<float*>(smem + S::cst_off);            // [hs][CP]
  float* dbuf = cst + (size_t)hs * CP;                                 // [D][NCP]
  float* n_obs_mean = dbuf + (size_t)D * NCP;
  float* n_obs_rden = n_obs_mean + D;
  float* n_dmean = n_obs_rden + D;
  float* n_dscale = n_dmean + D;
  float* n_act_mean = n_dscale + D;
  float* n_act_rden = n_act_mean + A;
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(n_act_rden + A) + 15) & ~(uintptr_t)15);
  uint64_t* full = bars;                  // [2]
  uint64_t* empty = bars + 2;             // [2]
  uint64_t* gates_full = bars + 4;
  uint64_t* out_full = bars + 5;
  uint64_t* x_ready = bars + 6;           // 128 arrivals: chunk 0 (x) written -- and, by program order, the h chunks before it
  uint64_t* h_ready = bars + 7;           // 128 arrivals: the new h is in the activation chunks
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* red_v = reinterpret_cast<float*>(tmem_slot + 2);
  int* red_i = reinterpret_cast<int*>(red_v + 4);
  int* s_flag = red_i + 4;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int env = blockIdx.x / a.groups_per_env, group = blockIdx.x % a.groups_per_env;
  const int c0i = group * NC;
  const int nvalid = min(NC, a.n_candidates - c0i);

  if (tid == 0) {
    for (int s = 0; s < kRnnTcStages; ++s) { umma::mbar_init(&full[s], 1); umma::mbar_init(&empty[s], 1); }
    umma::mbar_init(gates_full, 1);
    umma::mbar_init(out_full, 1);
    umma::mbar_init(x_ready, 128);
    umma::mbar_init(h_ready, 128);
    umma::fence_barrier_init();
  }
  if (warp == 5) umma::tmem_alloc<512>(tmem_slot);
  for (int i = tid; i < D; i += kRnnTcThreads) {
    n_obs_mean[i] = a.norm.obs_mean[i];
    n_obs_rden[i] = 1.0f / a.norm.obs_den[i];
    n_dmean[i] = a.norm.delta_mean[i];
    n_dscale[i] = a.norm.delta_scale[i];
  }
  for (int i = tid; i < A; i += kRnnTcThreads) { n_act_mean[i] = a.norm.act_mean[i]; n_act_rden[i] = 1.0f / a.norm.act_den[i]; }
  // initial cell / hidden state of this env, repeated for every candidate (repeat_hidden, rnn_mpc_controller.py:165-187)
  for (int idx = tid; idx < hs * NC; idx += kRnnTcThreads) {
    const int u = idx / NC, n = idx % NC;
    cst[u * CP + n] = a.c0[(size_t)env * hs + u];
    uint16_t hi, lo;
    umma::split_bf16(a.h0[(size_t)env * hs + u], hi, lo);
    const uint32_t off = (uint32_t)(1 + (u >> 6)) * kChunkBytes + umma::sw128_offset((uint32_t)n, (uint32_t)(u & 63));
    *reinterpret_cast<uint16_t*>(act_hi + off) = hi;
    *reinterpret_cast<uint16_t*>(act_lo + off) = lo;
  }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < H; ++t)
        for (int pr = 0; pr < pl.pairs_per_set; ++pr) {
          umma::mbar_wait(&empty[stage], phase ^ 1u);
          umma::mbar_arrive_expect_tx(&full[stage], kStageBytes);
          umma::bulk_g2s(stages + (size_t)stage * kStageBytes, a.blob + (size_t)pr * kStageBytes, kStageBytes, &full[stage]);
          if (++stage == kRnnTcStages) { stage = 0; phase ^= 1u; }
        }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ================================================================ MMA issuer (whole warp walks the loops)
    int stage = 0;
    uint32_t phase = 0, xr_phase = 0, hr_phase = 0;
    const uint32_t hi_lo32 = umma::desc_lo32(umma::smem_u32(act_hi)), lo_lo32 = umma::desc_lo32(umma::smem_u32(act_lo));
    const uint32_t st_lo32 = umma::desc_lo32(umma::smem_u32(stages));
    constexpr uint32_t kChunkStep = (uint32_t)kChunkBytes >> 4, kStageStep = (uint32_t)kStageBytes >> 4, kLoStep = (uint32_t)kTcTileBytes >> 4;
    auto tile_pair = [&](uint32_t d_tmem, int kc, bool first, int nks) {
      const uint32_t bh = hi_lo32 + (uint32_t)kc * kChunkStep, bl = lo_lo32 + (uint32_t)kc * kChunkStep;
      umma::mbar_wait(&full[stage], phase);
      umma::tc_fence_after();
      const uint32_t a_hi = st_lo32 + (uint32_t)stage * kStageStep, a_lo = a_hi + kLoStep;
      if (umma::elect_one()) {
        for (int ks = 0; ks < nks; ++ks) {
          // W_hi is fetched from shared memory once for its two passes (A-collector keep / reuse, see umma.cuh)
          umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d_tmem, a_hi + 2 * ks, bh + 2 * ks, kIdesc, (first && ks == 0) ? 0u : 1u);
          umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d_tmem, a_hi + 2 * ks, bl + 2 * ks, kIdesc, 1u);
          umma::mma_bf16_ss_lo(d_tmem, a_lo + 2 * ks, bh + 2 * ks, kIdesc, 1u);
        }
        umma::mma_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == kRnnTcStages) { stage = 0; phase ^= 1u; }
    };
    for (int t = 0; t < H; ++t) {
      umma::mbar_wait(x_ready, xr_phase);
      xr_phase ^= 1u;
      umma::tc_fence_after();
      for (int mb = 0; mb < pl.nmbg; ++mb)
        for (int kc = 0; kc < pl.nkc; ++kc) tile_pair(tmem_base + (uint32_t)(mb * NC), kc, kc == 0, kc == 0 ? pl.nks0 : 4);
      if (umma::elect_one()) umma::mma_commit(gates_full);
      __syncwarp();
      umma::mbar_wait(h_ready, hr_phase);
      hr_phase ^= 1u;
      umma::tc_fence_after();
      for (int kc = 1; kc < pl.nkc; ++kc) tile_pair(tmem_base, kc, kc == 1, 4);
      if (umma::elect_one()) umma::mma_commit(out_full);
      __syncwarp();
    }
  } else {
    // ================================================================ epilogue + env step (warps 0-3)
    const int n = tid;
    const bool has_cand = n < NC, valid = n < nvalid;
    const long long row = (long long)env * a.n_candidates + c0i + (valid ? n : 0);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t act_hi_addr = umma::smem_u32(act_hi), act_lo_addr = umma::smem_u32(act_lo);
    uint32_t gf_phase = 0, of_phase = 0;
    float ret = 0.f, asq = 0.f;
    float a_cur[kTcMaxAct];
    float st[kTcMaxObs];
#pragma unroll
    for (int k = 0; k < kTcMaxObs; ++k) st[k] = (k < D) ? __ldg(a.obs0 + (size_t)env * D + k) : 0.f;
    const int d8 = tc_obs_pad(D);
    const int nub = hs / 128;

    auto load_actions = [&](int t) {
      const float* src = a.actions + (long long)t * a.act_stride_t + row * a.act_stride_row;
#pragma unroll
      for (int j = 0; j < kTcMaxAct; ++j) a_cur[j] = (j < A && valid && has_cand) ? __ldg(src + j) : 0.f;
    };
    auto write_x = [&]() {
      if (has_cand) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kTcMaxAct; ++j) s = fmaf(a_cur[j], a_cur[j], s);
        asq = s;
        auto store_group = [&](int g, const float (&v)[8]) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) umma::split_bf16x2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
          const uint32_t off = umma::sw128_offset((uint32_t)n, (uint32_t)g * 8u);
          *reinterpret_cast<uint4*>(act_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(act_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        };
#pragma unroll
        for (int g = 0; g < kTcMaxObs / 8; ++g) {
          if (g * 8 < d8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int k = g * 8 + i;
              v[i] = (k < D) ? (st[k] - n_obs_mean[k]) * n_obs_rden[k] : 0.f;
            }
            store_group(g, v);
          }
        }
#pragma unroll
        for (int ga = 0; ga < kTcMaxAct / 8; ++ga) {
          if (ga * 8 < A) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int j = ga * 8 + i;
              v[i] = (j < A) ? (a_cur[j] - n_act_mean[j]) * n_act_rden[j] : 0.f;
            }
            store_group(d8 / 8 + ga, v);
          }
        }
        {
          const int used = d8 / 8 + (A + 7) / 8, need = pl.nks0 * 2;
          const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int g = used; g < need; ++g) store_group(g, z);
        }
      }
      umma::fence_proxy_async_smem();
      umma::tc_fence_before();
      umma::mbar_arrive(x_ready);
    };

    load_actions(0);
    write_x();

    const int cand_l = (lane & 7) + ((lane >> 4) & 1) * 8;
    const int fsel = ((lane >> 3) & 1) * 8;
    for (int t = 0; t < H; ++t) {
      const float disc_t = __ldg(a.discount_pow + t);       // discount**t, fetched a whole step before its use
      // ---------------- LSTM cell: the four gate accumulators of a unit meet in this thread
      umma::mbar_wait(gates_full, gf_phase);
      gf_phase ^= 1u;
      umma::tc_fence_after();
      if (t + 1 < H) load_actions(t + 1);
      for (int ub = 0; ub < nub; ++ub) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int ubase = ub * 128 + warp * 32 + half * 16;            // 16 units handled by this (warp, half)
          const int u_a = ubase + (lane >> 2), u_b = u_a + 8;
          float bias[4][2];
#pragma unroll
          for (int g = 0; g < 4; ++g) { bias[g][0] = __ldg(a.bk + g * hs + u_a); bias[g][1] = __ldg(a.bk + g * hs + u_b); }
          const uint32_t lane_sel = (uint32_t)(warp * 32 + half * 16) << 16;
          const uint32_t chunk_off = (uint32_t)(1 + (ubase >> 6)) * kChunkBytes;
          const uint32_t fcol = (uint32_t)((ubase & 63) + fsel) >> 3;
#pragma unroll 1
          for (int cb = 0; cb < NC / 16; ++cb) {
            uint32_t z[4][8];
#pragma unroll
            for (int g = 0; g < 4; ++g)
              umma::tmem_ld_16x256b_x2(tmem_base + lane_sel + (uint32_t)((g * nub + ub) * NC + cb * 16), z[g]);
            umma::tmem_ld_wait();
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int w = q & 1;                                       // 0: unit u_a, 1: unit u_b
              const int unit = w ? u_b : u_a;
              const int cand = cb * 16 + (q >> 1) * 8 + 2 * (lane & 3);
              float2 c = *reinterpret_cast<const float2*>(cst + (size_t)unit * CP + cand);
              float hv[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float zi = __uint_as_float(z[0][2 * q + e]) + bias[0][w];
                const float zj = __uint_as_float(z[1][2 * q + e]) + bias[1][w];
                const float zf = __uint_as_float(z[2][2 * q + e]) + bias[2][w];
                const float zo = __uint_as_float(z[3][2 * q + e]) + bias[3][w];
                const float c_old = e ? c.y : c.x;
                const float c_new = fast_sigmoid(zf + 1.0f) * c_old + fast_sigmoid(zi) * fast_tanh(zj);   // LSTMCell, forget_bias 1
                if (e) c.y = c_new; else c.x = c_new;
                hv[e] = fast_sigmoid(zo) * fast_tanh(c_new);
              }
              *reinterpret_cast<float2*>(cst + (size_t)unit * CP + cand) = c;
              umma::split_bf16x2(hv[0], hv[1], hi[q], lo[q]);
            }
            const uint32_t cand_row = (uint32_t)(cb * 16 + cand_l);
            const uint32_t off = chunk_off + cand_row * 128u + (((fcol ^ cand_row) & 7u) << 4);
            umma::stmatrix_x4_trans(act_hi_addr + off, hi[0], hi[1], hi[2], hi[3]);
            umma::stmatrix_x4_trans(act_lo_addr + off, lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      umma::fence_proxy_async_smem();
      umma::tc_fence_before();
      umma::mbar_arrive(h_ready);
      // ---------------- output layer: y -> denormalised delta
      umma::mbar_wait(out_full, of_phase);
      of_phase ^= 1u;
      umma::tc_fence_after();
      if (warp * 32 < D) {
        const int f = tid;
        const bool frow = f < D;
        const float bias = frow ? __ldg(a.bo + f) : 0.f;
        const float sc = frow ? n_dscale[f] : 0.f, mu = frow ? n_dmean[f] : 0.f;
        uint32_t r[NC / 16][16];
#pragma unroll
        for (int c16 = 0; c16 < NC / 16; ++c16) umma::tmem_ld_32x32b_x16(tmem_base + lane_base + (uint32_t)(c16 * 16), r[c16]);
        umma::tmem_ld_wait();
        if (frow) {
#pragma unroll
          for (int c16 = 0; c16 < NC / 16; ++c16)
#pragma unroll
            for (int i = 0; i < 16; ++i) dbuf[f * NCP + c16 * 16 + i] = (__uint_as_float(r[c16][i]) + bias) * sc + mu;   // rnn_dynamics.py:244
        }
      }
      umma::tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // ---------------- env step
      if (has_cand) {
        float dx = 0.f, nx0 = 0.f, nx1 = 0.f, nx2 = 0.f;
#pragma unroll
        for (int k = 0; k < kTcMaxObs; ++k) {
          if (k < D) {
            const float d = dbuf[k * NCP + n];
            const float s_new = st[k] + d;
            st[k] = s_new;
            if (k == D - 3) { dx = d; nx0 = s_new; }
            if (k == D - 2) nx1 = s_new;
            if (k == D - 1) nx2 = s_new;
          }
        }
        ret = fmaf(disc_t, reward_value(a.reward_kind, 0.f, a.dt, asq, dx, nx0, nx1, nx2), ret);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");       // dbuf is rewritten by the next step's output epilogue
      if (t + 1 < H) write_x();
    }

    float v = -__int_as_float(0x7f800000);
    int idx = 0x7fffffff;
    if (valid && has_cand) { v = ret; idx = c0i + n; }
    if (a.returns && valid && has_cand) a.returns[(size_t)env * a.n_candidates + c0i + n] = ret;
    warp_argmax(v, idx);
    if (lane == 0) { red_v[warp] = v; red_i[warp] = idx; }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 0) {
      v = (lane < 4) ? red_v[lane] : -__int_as_float(0x7f800000);
      idx = (lane < 4) ? red_i[lane] : 0x7fffffff;
      warp_argmax(v, idx);
      if (lane == 0) { red_v[0] = v; red_i[0] = idx; }
    }
  }

  umma::tc_fence_before();
  __syncthreads();
  publish_and_reduce(a.red, env, group, red_v[0], red_i[0], tid, s_flag);
  if (warp == 5) {
    umma::tc_fence_after();
    umma::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace l2a
