// K1, CTA-pair variant: the fused H-step rollout of rollout_tc.cuh on tcgen05.mma.cta_group::2 (sm_100a).
//
// Why a pair: at N = 80 candidates a single-CTA SS-mode MMA is shared-memory-bandwidth bound (4 KB of A + 2.5 KB of B per
// 40 tensor cycles) and a 32 KB weight stage carries only ~560 cycles of MMA time against a ~900-cycle refill.  With
// cta_group::2 the MMA is M = 256 output features x N = 2*NC candidates across two CTAs of a cluster:
//   * A = W^T tile [256 x 64]: each CTA streams only ITS 128 rows (half the weight bytes per CTA),
//   * B = activations [2*NC x 64]: each CTA holds only ITS NC candidates (resident in shared memory for the whole rollout),
//   * D: CTA r's TMEM holds features [128 r, 128 r + 128) of the M-block for all 2*NC candidates,
// so one 32 KB stage per CTA feeds 12 MMAs of 128 x 2NC per SM = 961 cycles at NC = 80 (measured, profiles/r02_pair_probe.txt:
// the tensor pipe's own rate), operand reads drop to ~81 B/clk per SM, and the ring handshakes are amortised over twice the
// MMA time.
// The price is an all-to-all in the hidden-layer epilogue: CTA r produces features of ALL 2*NC candidates, but the next
// layer's B operand of the peer's candidates lives in the peer's shared memory.  The hidden activations are therefore kept
// MN-MAJOR (candidates contiguous, unswizzled 8x8 core matrices): an epilogue thread owns one output feature (its TMEM lane),
// holds 8 consecutive candidates per 32x32b TMEM load and writes them as ONE 16-byte store, into its own CTA's buffer or, for
// the peer's candidates, into the peer's with st.async (DSMEM, the bytes counted on the peer's `in_ready` mbarrier) -- no
// transposition, no staging, no fence on the sending side.  A warp reports to the LEADER's mbarriers (only the leader CTA issues
// MMAs) once its own rows are in place and its CTA's incoming bytes have landed; the peer's arrivals are relaxed remote arrives:
// no cluster-scope release (MEMBAR.ALL.GPU, 1.2-1.8 k cycles on B200) anywhere on the layer-to-layer path.
// (First version: K-major activations transposed through a per-warp stmatrix bounce buffer: 5.5 k cycles per M-block epilogue.)
// The output layer keeps the swapped roles of rollout_tc.cuh (M = candidates: each CTA's own 128 activation rows are its half
// of A; the [out_n x 64] weight tiles are B, split by N between the CTAs: x_hi * [W_hi ; W_lo] has W_hi in the leader and
// W_lo in the peer), so every candidate's deltas land in the TMEM lanes of its own CTA's env-step thread.
// Weight stages arrive through a 2-D tensor-map TMA (cp.async.bulk.tensor.cta_group::2) so that the peer's copies complete
// on the leader's full barrier without a forwarding thread.
// Ensemble mode: E pairs per candidate tile (one member each); the members' deltas are exchanged through an L2-resident
// scratch with one release/acquire flag per (member, rank) and step, instead of the cluster barrier of rollout_tc.cuh (the
// cluster dimension is taken by the pair).  The E pairs of a tile are adjacent in launch order.
//
// Replaces policies/mpc_controller.py:116-129 + dynamics/mlp_dynamics.py:204-222 / meta_mlp_dynamics.py:296-306.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "umma.cuh"
#include "rollout_tc.cuh"

namespace l2a {

// timing experiments only: 1 = every epilogue store goes to the CTA's own shared memory (garbage results); 3 = cluster-scope
// proxy fence by every lane in the hidden epilogue (see there)
#ifndef L2A_TC2_EXPERIMENT
#define L2A_TC2_EXPERIMENT 0
#endif
// warpgroups per CTA: 3 = epilogue/env + (producer, issuer, helpers) + (helpers, 2nd producer); 4 adds a second helper group
// (warps 12-15), every non-env warp then runs at 88 registers
// member exchange of the ensemble mode: 0 (default) = rows + one release flag per (member, rank); 1 = flag-in-data ("LL": every
// 16-byte unit carries two {value, step} words, readers poll the data itself: no fence, no flag round trip, no CTA barrier).
// Measured (B200, headline): LL 0.819 ms vs 0.602 ms -- 128 threads x 40 polled 16-byte gpu-scope loads per CTA make the
// exchange 27 k cycles per step instead of 7 k (profiles/r02_negative_results.md); kept for the record, parity-green.
#ifndef L2A_TC2_LL
#define L2A_TC2_LL 0
#endif
// epilogue all-to-all: 1 (default) = st.async into the peer with the bytes counted on the peer's `in_ready` mbarrier (no fence on the
// sending side; every warp waits for its own CTA's incoming bytes, then arrives on the leader WITHOUT a cluster-scope release);
// 0 = plain st.shared::cluster + one mbarrier.arrive.release.cluster per warp (MEMBAR.ALL.GPU: 1.2-1.8 k cycles per M-block)
#ifndef L2A_TC2_STASYNC
#define L2A_TC2_STASYNC 1
#endif
// hidden epilogue arithmetic with the packed fp32x2 adder (1) or scalar adds (0); results are bit-identical
#ifndef L2A_TC2_FADD2
#define L2A_TC2_FADD2 1
#endif
#ifndef L2A_TC2_WGS
#define L2A_TC2_WGS 3
#endif
constexpr int kTc2Threads = 128 * L2A_TC2_WGS;
constexpr int kTc2Parts = L2A_TC2_WGS - 1;          // warps that share a TMEM lane quadrant in the hidden epilogue
#define L2A_TC2_DEC_REGS() asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(L2A_TC2_WGS == 4 ? 88 : 128))
constexpr int kTc2StageBytes = 32768;     // per CTA and ring stage: W_hi tile + W_lo tile of this CTA's 128 rows (or output-layer tiles)
constexpr int kTc2MaxStages = 4;
constexpr int kTc2XChunk = kTcMaxChunks - 1;   // activation chunk that holds the layer-0 input (not written by the M-block-0 epilogue)

// Stage enumeration of one weight set's pair blob, consumed in exactly this order.  A stage is 2 x 32 KB: the leader's
// half then the peer's half.  Hidden layer l: for mb (256-feature M-blocks): for kc: one stage; a layer-0 whose padded input is
// <= 32 wide and which has two M-blocks is PACKED into one stage (M-block 0 in K columns [0,32), M-block 1 in [32,64)).
// Output layer: per 64-wide K chunk a region of out_kc_bytes: Y [out_n x 64] (leader: W_hi, peer: W_lo) then X [out_n/2 x 64]
// (W_hi rows [0, out_n/2) in the leader, [out_n/2, out_n) in the peer); out_kcs chunks per stage.
struct Tc2Plan {
  int n_layers;
  int nmb[kMaxLayers];
  int nkc[kMaxLayers];
  int nks_last[kMaxLayers];
  int stage_off[kMaxLayers];
  int l0_packed;
  int hidden_stages;
  int out_n, out_kc_bytes, out_kcs, out_stages;
  int stages_per_set;
  long long set_bytes;
};

__host__ __device__ inline int tc2_layer_stages(const Tc2Plan& p, int l) {
  return (l == 0 && p.l0_packed) ? 1 : p.nmb[l] * p.nkc[l];
}

inline bool tc2_make_plan(const MlpDims& md, Tc2Plan* p) {
  memset(p, 0, sizeof(*p));
  if (md.act_dim > kTcMaxAct || md.obs_dim > kTcMaxObs || md.obs_dim < 3) return false;
  if (tc_obs_pad(md.obs_dim) + md.act_dim > 64) return false;
  if (md.n_layers < 2) return false;
  p->n_layers = md.n_layers;
  int st = 0;
  for (int l = 0; l < md.n_layers; ++l) {
    const int din = md.dims[l], dout = md.dims[l + 1];
    if (l + 1 < md.n_layers) {
      if (dout % 256 != 0 || dout > 64 * kTcMaxChunks) return false;
      p->nmb[l] = dout / 256;
    }
    if (l > 0 && din % 64 != 0) return false;
    const int din_eff = (l == 0) ? tc_obs_pad(md.obs_dim) + md.act_dim : din;
    p->nkc[l] = (din_eff + 63) / 64;
    const int rem = din_eff - (p->nkc[l] - 1) * 64;
    p->nks_last[l] = (rem + 15) / 16;
    p->stage_off[l] = st;
    if (l == 0) p->l0_packed = (p->nkc[0] == 1 && p->nks_last[0] <= 2 && p->nmb[0] == 2) ? 1 : 0;
    if (l + 1 < md.n_layers) st += tc2_layer_stages(*p, l);
  }
  p->hidden_stages = st;
  p->out_n = (md.obs_dim + 15) / 16 * 16;
  p->out_kc_bytes = p->out_n * 128 + (p->out_n / 2) * 128;
  p->out_kcs = kTc2StageBytes / p->out_kc_bytes;
  p->out_stages = (p->nkc[md.n_layers - 1] + p->out_kcs - 1) / p->out_kcs;
  p->stages_per_set = p->hidden_stages + p->out_stages;
  p->set_bytes = (long long)p->stages_per_set * 2 * kTc2StageBytes;
  return true;
}

struct Prep2Args {
  MlpDims dims;
  Tc2Plan plan;
  const float* params;
  uint8_t* blobs;
  int first_set;
};

// fp32 [in, out] kernels -> the pair blob (bf16 hi / lo, K-major, SWIZZLE_128B pre-applied, zero padded).
// grid.x = hidden stages, then one block per K chunk of the output layer; grid.y = set; grid.z = CTA rank of the pair.
__global__ void __launch_bounds__(256) tc2_prep_kernel(const Prep2Args a) {
  const Tc2Plan& p = a.plan;
  const int set = a.first_set + blockIdx.y, rank = blockIdx.z;
  uint8_t* const set_base = a.blobs + (size_t)set * p.set_bytes;
  auto pack8 = [](const uint16_t (&v)[8]) {
    return make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
  };
  if ((int)blockIdx.x >= p.hidden_stages) {
    const int l = p.n_layers - 1;
    const int kc = (int)blockIdx.x - p.hidden_stages;
    const int din = a.dims.dims[l], dout = a.dims.dims[l + 1];
    const float* W = a.params + (size_t)set * a.dims.set_stride + a.dims.w_off[l];
    uint8_t* cb = set_base + ((size_t)(p.hidden_stages + kc / p.out_kcs) * 2 + rank) * kTc2StageBytes + (size_t)(kc % p.out_kcs) * p.out_kc_bytes;
    uint8_t* Y = cb;
    uint8_t* X = cb + p.out_n * 128;
    const int hn = p.out_n / 2;
    for (int item = threadIdx.x; item < p.out_n * 8; item += blockDim.x) {
      const int f = item % p.out_n, ch = item / p.out_n;
      uint16_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kc * 64 + ch * 8 + i;
        const float w = (f < dout && k < din) ? W[(size_t)k * dout + f] : 0.f;
        umma::split_bf16(w, hi[i], lo[i]);
      }
      *reinterpret_cast<uint4*>(Y + umma::sw128_offset(f, ch * 8)) = (rank == 0) ? pack8(hi) : pack8(lo);
    }
    for (int item = threadIdx.x; item < hn * 8; item += blockDim.x) {
      const int fl = item % hn, ch = item / hn;
      const int f = rank * hn + fl;
      uint16_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kc * 64 + ch * 8 + i;
        const float w = (f < dout && k < din) ? W[(size_t)k * dout + f] : 0.f;
        umma::split_bf16(w, hi[i], lo[i]);
      }
      *reinterpret_cast<uint4*>(X + umma::sw128_offset(fl, ch * 8)) = pack8(hi);
    }
    return;
  }
  int l = 0, rem = blockIdx.x;
  while (l + 2 < p.n_layers && rem >= tc2_layer_stages(p, l)) { rem -= tc2_layer_stages(p, l); ++l; }
  const int din = a.dims.dims[l], dout = a.dims.dims[l + 1];
  const float* W = a.params + (size_t)set * a.dims.set_stride + a.dims.w_off[l];
  uint8_t* tile_hi = set_base + ((size_t)blockIdx.x * 2 + rank) * kTc2StageBytes;
  uint8_t* tile_lo = tile_hi + kTcTileBytes;
  const bool packed = (l == 0 && p.l0_packed);
  int mb = 0, kc = 0;
  if (!packed) {
    if (l == 0) { mb = rem / p.nkc[l]; kc = rem % p.nkc[l]; }
    else {
      // hidden layers behind layer 0: K halves outermost -- chunks [0,4) for every M-block, then chunks [4, nkc) for every M-block
      // (the first half needs only M-block 0 of the previous layer's epilogue)
      const int h0 = p.nkc[l] < 4 ? p.nkc[l] : 4, h1 = p.nkc[l] - h0;
      if (rem < p.nmb[l] * h0) { mb = rem / h0; kc = rem % h0; }
      else { rem -= p.nmb[l] * h0; mb = rem / h1; kc = h0 + rem % h1; }
    }
  }
  for (int item = threadIdx.x; item < 128 * 8; item += blockDim.x) {
    const int r = item & 127, ch = item >> 7;
    const int f = (packed ? (ch >> 2) : mb) * 256 + rank * 128 + r;
    uint16_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int k = packed ? ((ch & 3) * 8 + i) : (kc * 64 + ch * 8 + i);
      if (l == 0) k = tc_in0_of_col(a.dims.obs_dim, a.dims.act_dim, k);
      const float w = (f < dout && k >= 0 && k < din) ? W[(size_t)k * dout + f] : 0.f;
      umma::split_bf16(w, hi[i], lo[i]);
    }
    const uint32_t off = umma::sw128_offset(r, ch * 8);
    *reinterpret_cast<uint4*>(tile_hi + off) = pack8(hi);
    *reinterpret_cast<uint4*>(tile_lo + off) = pack8(lo);
  }
}

struct Tc2Args {
  MlpDims dims;
  Tc2Plan plan;
  NormDev norm;
  const float* params;
  const float* obs0;
  const float* actions;
  long long act_stride_t, act_stride_row;
  const float* discount_pow;
  int n_candidates, n_envs, horizon;
  int set_mode, first_set, n_sets;
  int reward_kind;
  float dt;
  int groups_per_env;             // candidate tiles of 2*NC per env
  float* returns;
  ReduceArgs red;                 // tiles_per_env = 2 * groups_per_env (one partial per CTA)
  float* xch;                     // member exchange rows: [tile][2 (step parity)][E][2 (rank)][DMAX/4][NC] float4 (L2-resident)
  unsigned int* flags;            // [tile][2 (rank)][E] step counters, zeroed before the launch
  long long* timeline;
};

template <int NC>
struct Tc2Smem {
  static constexpr int kChunkBytes = NC * 128;
  static constexpr size_t act_bytes = (size_t)2 * kTcMaxChunks * kChunkBytes;
  static constexpr size_t stage_off = act_bytes;
  static constexpr size_t kMisc = 2048;
  static constexpr int kFit = (int)((232448 - (long long)act_bytes - (long long)kMisc) / kTc2StageBytes);
  static constexpr int kStages = kFit > kTc2MaxStages ? kTc2MaxStages : kFit;
  static_assert(kStages >= 2, "no room for the weight ring");
  static constexpr size_t misc_off = stage_off + (size_t)kStages * kTc2StageBytes;
  static constexpr size_t total = misc_off + kMisc;
  static_assert(kChunkBytes % 1024 == 0, "activation chunks must stay 1024-byte aligned (SWIZZLE_128B atoms): NC % 8 == 0");
};

template <int NC, int DMAX>
__global__ void __launch_bounds__(kTc2Threads, 1) rollout_tc2_kernel(const Tc2Args a, const __grid_constant__ CUtensorMap wmap) {
  using S = Tc2Smem<NC>;
  constexpr int kChunkBytes = S::kChunkBytes;
  constexpr int kStages = S::kStages;
  // W_hi / W_lo halves of a stage with their own barriers and producer threads only where the ring is 2 deep (NC >= 72): a deeper
  // ring already covers the refill latency, and one handshake per stage instead of two is worth ~100 cycles per stage there
  constexpr bool kSplit = (kStages == 2);
  constexpr int N2 = 2 * NC;                               // MMA N = candidates of the pair
  constexpr uint32_t kIdesc = umma::make_idesc_bf16(256, N2);
  static_assert(N2 % 16 == 0 && N2 <= 256 && 3 * N2 <= 512, "UMMA N constraint (M = 256) / three accumulator slots in 512 TMEM columns");

  extern __shared__ __align__(1024) uint8_t tc2_smem[];
  uint8_t* const smem = tc2_smem;
  if ((umma::smem_u32(smem) & 1023u) != 0) __trap();
  const MlpDims& md = a.dims;
  const Tc2Plan& plan = a.plan;
  const int D = md.obs_dim, A = md.act_dim, L = md.n_layers, H = a.horizon;

  uint8_t* act_hi = smem;
  uint8_t* act_lo = smem + (size_t)kTcMaxChunks * kChunkBytes;
  uint8_t* stages = smem + S::stage_off;
  float* n_obs_mean = reinterpret_cast<float*>(smem + S::misc_off);
  float* n_obs_den = n_obs_mean + DMAX;
  float* n_dmean = n_obs_den + DMAX;
  float* n_dscale = n_dmean + DMAX;
  float* n_bias_out = n_dscale + DMAX;
  float* n_act_mean = n_bias_out + DMAX;
  float* n_act_den = n_act_mean + kTcMaxAct;
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(n_act_den + kTcMaxAct) + 15) & ~(uintptr_t)15);
  uint64_t* full_hi = bars;                        // [kStages]  leader only: both CTAs' W_hi halves have landed
  uint64_t* full_lo = bars + kTc2MaxStages;        // [kStages]  leader only
  uint64_t* empty_hi = bars + 2 * kTc2MaxStages;   // [kStages]  both CTAs (multicast commit)
  uint64_t* empty_lo = bars + 3 * kTc2MaxStages;   // [kStages]
  uint64_t* layer_full = bars + 4 * kTc2MaxStages; // both CTAs
  uint64_t* early = layer_full + 1;                // both CTAs: M-block 0 of a hidden layer may be drained
  uint64_t* act_ready = layer_full + 2;            // [2] leader only: 512 arrivals (256 epilogue / helper threads of each CTA)
  uint64_t* x_ready = layer_full + 4;              // leader only: one arrival per env-step warp of both CTAs
  uint64_t* in_ready = layer_full + 5;             // [2] both CTAs: the peer's st.async bytes of M-block mb have landed here (tx count)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(layer_full + 7);
  float* red_v = reinterpret_cast<float*>(tmem_slot + 2);
  int* red_i = reinterpret_cast<int*>(red_v + 4);
  int* s_flag = red_i + 4;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = umma::cluster_ctarank();           // 0 = leader (issues every MMA of the pair)
  const bool ensemble = (a.set_mode == L2A_SETS_ENSEMBLE_MEAN) && a.n_sets > 1;
  const int csize = ensemble ? a.n_sets : 1;
  const int pair_id = (int)(blockIdx.x >> 1);
  const int member = pair_id % csize;
  const int tile_id = pair_id / csize;
  const int env = tile_id / a.groups_per_env;
  const int group = tile_id % a.groups_per_env;
  const int c0 = group * N2 + (int)rank * NC;              // first candidate of this CTA
  const int nvalid = min(NC, a.n_candidates - c0);         // may be <= 0 for the peer of the last tile
  int set = a.first_set;
  if (a.set_mode == L2A_SETS_PER_ENV) set += env;
  if (ensemble) set += member;
  const float* P = a.params + (size_t)set * md.set_stride;

  // ------------------------------------------------------------------ setup
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      umma::mbar_init(&full_hi[s], 1); umma::mbar_init(&full_lo[s], 1);
      umma::mbar_init(&empty_hi[s], 1); umma::mbar_init(&empty_lo[s], 1);
    }
    umma::mbar_init(layer_full, 1);
    umma::mbar_init(early, 1);
    umma::mbar_init(&act_ready[0], 8 * kTc2Parts);   // one arrival per epilogue / helper warp of both CTAs
    umma::mbar_init(&act_ready[1], 8 * kTc2Parts);
    umma::mbar_init(x_ready, 8);                 // one arrival per env-step warp of both CTAs
    umma::mbar_init(&in_ready[0], 1);            // one expect_tx arrival per use + the peer's bytes
    umma::mbar_init(&in_ready[1], 1);
    umma::fence_barrier_init();
  }
  for (int i = tid; i < DMAX; i += kTc2Threads) {
    const bool in = i < D;
    n_obs_mean[i] = in ? a.norm.obs_mean[i] : 0.f;
    n_obs_den[i] = in ? 1.0f / a.norm.obs_den[i] : 0.f;
    n_dmean[i] = in ? a.norm.delta_mean[i] : 0.f;
    n_dscale[i] = in ? a.norm.delta_scale[i] : 0.f;
    n_bias_out[i] = in ? P[md.b_off[L - 1] + i] : 0.f;
  }
  for (int i = tid; i < kTcMaxAct; i += kTc2Threads) {
    n_act_mean[i] = (i < A) ? a.norm.act_mean[i] : 0.f;
    n_act_den[i] = (i < A) ? 1.0f / a.norm.act_den[i] : 0.f;
  }
  __syncthreads();
  umma::cluster_sync_all();                      // both CTAs' barriers exist before any remote arrive / multicast commit
  if (warp == 5) umma::tmem_alloc2<512>(tmem_slot);
  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync_all();
  umma::tc_fence_after();
  const uint32_t tmem_base = umma::ld_dsmem_u32(umma::map_to_cta(umma::smem_u32(tmem_slot), 0));   // the pair's allocation (leader's slot)

  // (Waits on barriers the peer arrives on use the plain CTA-scope try_wait: everything they order lives in shared memory;
  // the .acquire.cluster form adds a CCTL.IVALL -- an L1 invalidation -- to every successful wait of the MMA issuer.)
  // leader-side addresses of the barriers the peer arrives on
  const uint32_t act_ready_leader = umma::map_to_cta(umma::smem_u32(&act_ready[0]), 0);
  const uint32_t x_ready_leader = umma::map_to_cta(umma::smem_u32(x_ready), 0);

  // Hidden-layer epilogue of this warp's TMEM lane quadrant wq.  The hidden activations are stored MN-MAJOR (candidates
  // contiguous, no swizzle): element (candidate n, feature f) at (f/8) * kSlabBytes + (n/8) * 128 + (f%8) * 16 + (n%8) * 2, i.e.
  // 8x8 core matrices of 128 contiguous bytes, SBO = 128 B between candidate groups, LBO = kSlabBytes between feature groups.
  // With 32x32b TMEM loads a thread IS one output feature (TMEM lane) and holds 8 consecutive candidates per load: bias + ReLU
  // -> bf16 hi / lo -> ONE 16-byte store per part, straight from registers into the owning CTA's buffer (st.shared for its own
  // candidate groups, st.shared::cluster for the peer's) -- no transposition, no staging.  A warp's 32 lanes cover 4 feature
  // groups x 8 rows: every quarter-warp stores 128 contiguous bytes.
  // The kTc2Parts warps of a quadrant (epilogue warp + helpers) split the candidate groups: part p takes the p-th share of the
  // CTA's own groups and the (kTc2Parts-1-p)-th share of the peer's, so every warp mixes local and DSMEM stores.
  constexpr int kSlabBytes = NC * 16;                       // 8 features x NC candidates x 2 B
  constexpr int kGr = NC / 8;                               // 8-candidate groups per CTA
  auto hidden_epilogue = [&](auto part_tag, int t, int l, int slot_a, int wq, bool stamps, uint32_t& lf_phase, uint32_t& early_phase, uint32_t& in_phase) {
    constexpr int PART = decltype(part_tag)::value, RPART = kTc2Parts - 1 - PART;
    const uint32_t act_hi_addr = umma::smem_u32(act_hi), act_lo_addr = umma::smem_u32(act_lo);
    const uint32_t peer_hi = umma::map_to_cta(act_hi_addr, rank ^ 1u), peer_lo = umma::map_to_cta(act_lo_addr, rank ^ 1u);
    for (int mb = 0; mb < plan.nmb[l]; ++mb) {
      if (mb == 0) umma::mbar_wait(early, early_phase);
      else umma::mbar_wait(layer_full, lf_phase);
      umma::tc_fence_after();
      if (stamps) L2A_STAMP(32 + 4 * l + (mb == 0 ? 0 : 3));
      const int slot = (mb == 0) ? slot_a : (slot_a + 1) % 3;
      const int gfeat = mb * 256 + (int)rank * 128 + wq * 32 + lane;           // layer output feature of this thread's TMEM lane
      const float bias = __ldg(P + md.b_off[l] + gfeat);
      const uint32_t t_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(slot * N2);
      const uint32_t row_off = (uint32_t)(gfeat >> 3) * (uint32_t)kSlabBytes + (uint32_t)(gfeat & 7) * 16u;
      const uint32_t peer_in = umma::map_to_cta(umma::smem_u32(&in_ready[mb]), rank ^ 1u);
      if (L2A_TC2_STASYNC && PART == 0 && wq == 0 && lane == 0)       // this CTA expects the peer's half of the M-block: 128 features x NC candidates x (hi, lo)
        umma::mbar_arrive_expect_tx(&in_ready[mb], (uint32_t)(NC * 512));
      auto drain = [&](auto g0_tag, auto g1_tag) {
        constexpr int G0 = decltype(g0_tag)::value, G1 = decltype(g1_tag)::value, CNT = G1 - G0;
        if constexpr (CNT > 0) {
          uint32_t r[CNT][8];
#pragma unroll
          for (int i = 0; i < CNT; ++i) umma::tmem_ld_32x32b_x8(t_addr + (uint32_t)((G0 + i) * 8), r[i]);
          umma::tmem_ld_wait();
          L2A_TIMELINE(if (stamps && l == 1 && mb == 0 && a.timeline && blockIdx.x == 0 && t == 1 && lane == 0) a.timeline[100 + (G0 >= kGr ? 3 : 0)] = clock64());
#pragma unroll
          for (int i = 0; i < CNT; ++i) {
            const int g = G0 + i;
            uint4 uh, ul;
            {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
#if L2A_TC2_FADD2
                umma::bias_relu_split_bf16x2(r[i][2 * q], r[i][2 * q + 1], bias, hi[q], lo[q]);  // core/utils.py:119-126 (ReLU dense)
#else
                const float v0 = fmaxf(__uint_as_float(r[i][2 * q]) + bias, 0.f);          // core/utils.py:119-126 (ReLU dense)
                const float v1 = fmaxf(__uint_as_float(r[i][2 * q + 1]) + bias, 0.f);
                umma::split_bf16x2(v0, v1, hi[q], lo[q]);
#endif
              }
              uh = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              ul = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            const uint32_t owner = (g >= kGr) ? 1u : 0u;
            const uint32_t off = row_off + (uint32_t)(g - (g >= kGr ? kGr : 0)) * 128u;
            if (owner == rank || L2A_TC2_EXPERIMENT == 1) {
              umma::st_shared_v4(act_hi_addr + off, uh);
              umma::st_shared_v4(act_lo_addr + off, ul);
            } else if (L2A_TC2_STASYNC) {
              umma::st_async_cluster_v4(peer_hi + off, uh, peer_in);
              umma::st_async_cluster_v4(peer_lo + off, ul, peer_in);
            } else {
              umma::st_cluster_v4(peer_hi + off, uh);
              umma::st_cluster_v4(peer_lo + off, ul);
            }
          }
        }
      };
      // the peer's groups first: their DSMEM stores are in flight while the own groups are converted
      drain(IntTag<kGr + (RPART * kGr) / kTc2Parts>{}, IntTag<kGr + ((RPART + 1) * kGr) / kTc2Parts>{});
      L2A_TIMELINE(if (stamps && l == 1 && mb == 0 && a.timeline && blockIdx.x == 0 && t == 1 && lane == 0) a.timeline[104] = clock64());
      drain(IntTag<(PART * kGr) / kTc2Parts>{}, IntTag<((PART + 1) * kGr) / kTc2Parts>{});
      L2A_TIMELINE(if (stamps && l == 1 && mb == 0 && a.timeline && blockIdx.x == 0 && t == 1 && lane == 0) a.timeline[101] = clock64());
      // every lane: its rows in this CTA's own shared memory -> async proxy (CTA-scope proxy fence).  The rows sent to the peer are
      // st.async stores counted on the peer's in_ready barrier (L2A_TC2_STASYNC); the earlier version used plain st.shared::cluster
      // and ONE cluster-scope release per warp in front of the arrival (SASS: MEMBAR.ALL.GPU, 1.2-1.8 k cycles per M-block on the
      // critical path of the next layer: headline 0.596 vs 0.573 ms, cfg1 0.118 vs 0.110 ms).
      if (L2A_TC2_EXPERIMENT == 3) umma::fence_proxy_async_cluster(); else umma::fence_proxy_async_smem();
      umma::tc_fence_before();
      __syncwarp();
      L2A_TIMELINE(if (stamps && l == 1 && mb == 0 && a.timeline && blockIdx.x == 0 && t == 1 && lane == 0) a.timeline[105] = clock64());
#if L2A_TC2_STASYNC
      // this warp's rows are in place in its own CTA (proxy fence above) and on their way into the peer (counted there); once the
      // peer's rows have landed HERE the warp reports to the leader -- no cluster-scope release anywhere on this path
      umma::mbar_wait(&in_ready[mb], (in_phase >> mb) & 1u);
      if (lane == 0) {
        if (rank == 0) umma::mbar_arrive(&act_ready[mb]);
        else umma::mbar_arrive_remote_relaxed(act_ready_leader + (uint32_t)mb * 8u);
      }
#else
      if (lane == 0) {
        if (rank == 0) umma::mbar_arrive_release_cluster(&act_ready[mb]);
        else umma::mbar_arrive_remote(act_ready_leader + (uint32_t)mb * 8u);
      }
#endif
      if (stamps) L2A_STAMP(32 + 4 * l + (mb == 0 ? 1 : 2));
    }
    if (plan.nmb[l] == 1) umma::mbar_wait(layer_full, lf_phase);        // keeps the layer_full phase in step
    early_phase ^= 1u;
    lf_phase ^= 1u;
    in_phase ^= (plan.nmb[l] == 1) ? 1u : 3u;
  };

  if (warp >= 6) {
    // ================================================================ epilogue helpers (warps 6-9 [, 12-15]), second producer (warp 10)
    L2A_TC2_DEC_REGS();
    if (warp < 10 || warp >= 12) {
      const int wq = warp & 3;
      uint32_t lf_phase = 0, early_phase = 0, in_phase = 0;
      int slot_a = 0;
      for (int t = 0; t < H; ++t) {
        for (int l = 0; l + 1 < L; ++l) {
          if (kTc2Parts == 3 && warp >= 12) hidden_epilogue(IntTag<kTc2Parts - 1>{}, t, l, slot_a, wq, false, lf_phase, early_phase, in_phase);
          else hidden_epilogue(IntTag<1>{}, t, l, slot_a, wq, false, lf_phase, early_phase, in_phase);
          slot_a = (slot_a + 2) % 3;
        }
        umma::mbar_wait(layer_full, lf_phase);                           // the output layer's completion
        lf_phase ^= 1u;
        slot_a = (slot_a + 2) % 3;
      }
    } else if (kSplit && warp == 10 && lane == 0) {
      // W_lo halves of every stage (both CTAs stream their own rows; the bytes complete on the leader's barrier)
      const uint32_t full_leader = umma::map_to_cta(umma::smem_u32(&full_lo[0]), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < H; ++t) {
        for (int st = 0; st < plan.stages_per_set; ++st) {
          umma::mbar_wait(&empty_lo[stage], phase ^ 1u);
          if (rank == 0) umma::mbar_arrive_expect_tx(&full_lo[stage], 2u * kTcTileBytes);
          const int row = ((set * plan.stages_per_set + st) * 2 + (int)rank) * 256 + 128;
          umma::tma2_load_2d(stages + (size_t)stage * kTc2StageBytes + kTcTileBytes, &wmap, 0, row, full_leader + (uint32_t)stage * 8u);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 4) {
    // ================================================================ TMA producer: W_hi halves
    L2A_TC2_DEC_REGS();
    if (lane == 0) {
      const uint32_t full_leader = umma::map_to_cta(umma::smem_u32(&full_hi[0]), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < H; ++t) {
        for (int st = 0; st < plan.stages_per_set; ++st) {
          umma::mbar_wait(&empty_hi[stage], phase ^ 1u);
          if (rank == 0) umma::mbar_arrive_expect_tx(&full_hi[stage], (kSplit ? 2u : 4u) * kTcTileBytes);
          const int row = ((set * plan.stages_per_set + st) * 2 + (int)rank) * 256;
          umma::tma2_load_2d(stages + (size_t)stage * kTc2StageBytes, &wmap, 0, row, full_leader + (uint32_t)stage * 8u);
          if (!kSplit)
            umma::tma2_load_2d(stages + (size_t)stage * kTc2StageBytes + kTcTileBytes, &wmap, 0, row + 128, full_leader + (uint32_t)stage * 8u);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    L2A_TC2_DEC_REGS();
    // ================================================================ MMA issuer (leader CTA only; the whole warp walks the loops)
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0, act_phase = 0, xr_phase = 0;
      const uint32_t hi_lo32 = umma::desc_lo32(umma::smem_u32(act_hi)), lo_lo32 = umma::desc_lo32(umma::smem_u32(act_lo));
      const uint32_t st_lo32 = umma::desc_lo32(umma::smem_u32(stages));
      constexpr uint32_t kChunkStep = (uint32_t)kChunkBytes >> 4, kStageStep = (uint32_t)kTc2StageBytes >> 4, kLoStep = (uint32_t)kTcTileBytes >> 4;
      auto advance = [&]() { if (++stage == kStages) { stage = 0; phase ^= 1u; } };
      // hidden activations as MN-major B operand (see the epilogue): LBO = one 8-feature slab, a 16-wide k-step = two slabs
      const uint32_t hi_mn32 = umma::desc_lo32_mn(umma::smem_u32(act_hi), (uint32_t)(NC * 16)), lo_mn32 = umma::desc_lo32_mn(umma::smem_u32(act_lo), (uint32_t)(NC * 16));
      constexpr uint32_t kKsMn = (uint32_t)(2 * NC * 16) >> 4;         // k-step advance of an MN-major activation descriptor
      constexpr uint32_t kIdescMn = kIdesc | umma::kIdescBMn;
      // one stage = the three split-bf16 passes of one [256 x 64] weight block against activation chunk `ch` of all 2*NC candidates;
      // MN = the chunk is a hidden activation (MN-major), else the K-major SWIZZLE_128B layer-0 input
      auto tile_pair = [&](auto mn_tag, uint32_t d_tmem, int ch, bool first, int nks) {
        constexpr bool MN = decltype(mn_tag)::value != 0;
        constexpr uint32_t kb = MN ? kKsMn : 2u, b_hi32 = MN ? umma::kDescHi32Mn : umma::kDescHi32, idesc = MN ? kIdescMn : kIdesc;
        const uint32_t bh = (MN ? hi_mn32 : hi_lo32) + (uint32_t)ch * kChunkStep, bl = (MN ? lo_mn32 : lo_lo32) + (uint32_t)ch * kChunkStep;
        const uint32_t a_hi = st_lo32 + (uint32_t)stage * kStageStep, a_lo = a_hi + kLoStep;
        umma::mbar_wait(&full_hi[stage], phase);
        umma::tc_fence_after();
        if constexpr (!kSplit) {
          // whole stage behind one barrier: per k-step W_hi * x_hi (keep), W_hi * x_lo (reuse), W_lo * x_hi; one release
          if (umma::elect_one()) {
            if (nks == 4) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma::mma2_bf16_ss_ab<umma::kAKeep>(d_tmem, a_hi + 2 * ks, umma::kDescHi32, bh + kb * ks, b_hi32, idesc, (first && ks == 0) ? 0u : 1u);
                umma::mma2_bf16_ss_ab<umma::kAReuse>(d_tmem, a_hi + 2 * ks, umma::kDescHi32, bl + kb * ks, b_hi32, idesc, 1u);
                umma::mma2_bf16_ss_ab<umma::kANone>(d_tmem, a_lo + 2 * ks, umma::kDescHi32, bh + kb * ks, b_hi32, idesc, 1u);
              }
            } else {
              for (int ks = 0; ks < nks; ++ks) {
                umma::mma2_bf16_ss_ab<umma::kAKeep>(d_tmem, a_hi + 2 * ks, umma::kDescHi32, bh + kb * ks, b_hi32, idesc, (first && ks == 0) ? 0u : 1u);
                umma::mma2_bf16_ss_ab<umma::kAReuse>(d_tmem, a_hi + 2 * ks, umma::kDescHi32, bl + kb * ks, b_hi32, idesc, 1u);
                umma::mma2_bf16_ss_ab<umma::kANone>(d_tmem, a_lo + 2 * ks, umma::kDescHi32, bh + kb * ks, b_hi32, idesc, 1u);
              }
            }
            umma::mma2_commit_both(&empty_hi[stage]);
          }
          __syncwarp();
          advance();
          return;
        }
        if (umma::elect_one()) {
          if (nks == 4) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma::mma2_bf16_ss_ab<umma::kAKeep>(d_tmem, a_hi + 2 * ks, umma::kDescHi32, bh + kb * ks, b_hi32, idesc, (first && ks == 0) ? 0u : 1u);
              umma::mma2_bf16_ss_ab<umma::kAReuse>(d_tmem, a_hi + 2 * ks, umma::kDescHi32, bl + kb * ks, b_hi32, idesc, 1u);
            }
          } else {
            for (int ks = 0; ks < nks; ++ks) {
              umma::mma2_bf16_ss_ab<umma::kAKeep>(d_tmem, a_hi + 2 * ks, umma::kDescHi32, bh + kb * ks, b_hi32, idesc, (first && ks == 0) ? 0u : 1u);
              umma::mma2_bf16_ss_ab<umma::kAReuse>(d_tmem, a_hi + 2 * ks, umma::kDescHi32, bl + kb * ks, b_hi32, idesc, 1u);
            }
          }
          umma::mma2_commit_both(&empty_hi[stage]);          // the W_hi halves of both CTAs may be refilled
        }
        __syncwarp();
        umma::mbar_wait(&full_lo[stage], phase);
        umma::tc_fence_after();
        if (umma::elect_one()) {
          if (nks == 4) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma::mma2_bf16_ss_ab<umma::kANone>(d_tmem, a_lo + 2 * ks, umma::kDescHi32, bh + kb * ks, b_hi32, idesc, 1u);
          } else {
            for (int ks = 0; ks < nks; ++ks) umma::mma2_bf16_ss_ab<umma::kANone>(d_tmem, a_lo + 2 * ks, umma::kDescHi32, bh + kb * ks, b_hi32, idesc, 1u);
          }
          umma::mma2_commit_both(&empty_lo[stage]);
        }
        __syncwarp();
        advance();
      };
      // packed layer 0: one stage holds both M-blocks (K columns [0,32) and [32,64) of the tile), k-steps 0 .. nks-1 of the input chunk
      auto packed_stage = [&](uint32_t d0, uint32_t d1, int nks) {
        const uint32_t a_hi = st_lo32 + (uint32_t)stage * kStageStep, a_lo = a_hi + kLoStep;
        const uint32_t xh = hi_lo32 + (uint32_t)kTc2XChunk * kChunkStep, xl = lo_lo32 + (uint32_t)kTc2XChunk * kChunkStep;
        umma::mbar_wait(&full_hi[stage], phase);
        if (kSplit) umma::mbar_wait(&full_lo[stage], phase);
        umma::tc_fence_after();
        if (umma::elect_one()) {
#pragma unroll
          for (int mbsel = 0; mbsel < 2; ++mbsel) {
            const uint32_t d = mbsel ? d1 : d0;
            for (int ks = 0; ks < nks; ++ks) {
              const uint32_t ka = (uint32_t)(2 * (2 * mbsel + ks)), kb = (uint32_t)(2 * ks);
              umma::mma2_bf16_ss_lo<umma::kAKeep>(d, a_hi + ka, xh + kb, kIdesc, ks == 0 ? 0u : 1u);
              umma::mma2_bf16_ss_lo<umma::kAReuse>(d, a_hi + ka, xl + kb, kIdesc, 1u);
              umma::mma2_bf16_ss_lo<umma::kANone>(d, a_lo + ka, xh + kb, kIdesc, 1u);
            }
          }
          umma::mma2_commit_both(&empty_hi[stage]);
          if (kSplit) umma::mma2_commit_both(&empty_lo[stage]);
        }
        __syncwarp();
        advance();
      };
      auto commit = [&](uint64_t* bar) {
        if (umma::elect_one()) umma::mma2_commit_both(bar);
        __syncwarp();
      };
      // Accumulator slots: three of N2 TMEM columns.  Layer i: M-block 0 -> slot a_i, M-block 1 -> a_i + 1; a_{i+1} = a_i + 2
      // (mod 3): the slot layer i does not touch, so M-block 0 of layer i+1 starts while layer i's epilogue drains, and its
      // M-block 1 reuses a_i, drained before layer i+1's inputs were complete.
      int slot_a = 0;
      const uint32_t idesc_out = umma::make_idesc_bf16(256, (uint32_t)plan.out_n) | umma::kIdescAMn;        // A = MN-major activations
      const uint32_t idesc_out2 = umma::make_idesc_bf16(256, (uint32_t)(2 * plan.out_n)) | umma::kIdescAMn;
      const uint32_t out_kc_step = (uint32_t)plan.out_kc_bytes >> 4, out_x_step = (uint32_t)(plan.out_n * 128) >> 4;
      for (int t = 0; t < H; ++t) {
        for (int l = 0; l + 1 < L; ++l) {
          const int nmb = plan.nmb[l], nkc = plan.nkc[l], nks_last = plan.nks_last[l];
          const uint32_t dA = tmem_base + (uint32_t)(slot_a * N2), dB = tmem_base + (uint32_t)(((slot_a + 1) % 3) * N2);
          L2A_STAMP(4 * l + 0);
          L2A_TIMELINE(if (a.timeline && blockIdx.x == 0 && t == 2 && l == 0 && lane == 0) a.timeline[80] = clock64());
          if (l == 0) {
            umma::mbar_wait(x_ready, xr_phase);
            xr_phase ^= 1u;
            umma::tc_fence_after();
            L2A_STAMP(4 * l + 3);
            if (plan.l0_packed) {
              packed_stage(dA, dB, nks_last);
              commit(early);
            } else {
              tile_pair(IntTag<0>{}, dA, kTc2XChunk, true, nks_last);
              commit(early);                                  // M-block 0's epilogue writes chunks 0-3, the input sits in kTc2XChunk
              if (nmb > 1) tile_pair(IntTag<0>{}, dB, kTc2XChunk, true, nks_last);
            }
          } else {
            const int nsrc = plan.nmb[l - 1];
            // K halves outermost: event ev = M-block ev of the previous layer's epilogue published chunks [4 ev, 4 ev + 4); both
            // M-blocks of this layer consume them before the next event is needed, so the second half of the stages has the whole
            // epilogue of the previous layer's M-block 1 (and the first half that of M-block 0) to hide behind.
            // M-block 0 may be drained once its own accumulation is complete AND nothing reads chunks 0-3 any more (its epilogue
            // overwrites them in place): after its part of the last event when there are two events, else with the whole layer.
            for (int ev = 0; ev < nsrc; ++ev) {
              umma::mbar_wait(&act_ready[ev], (act_phase >> ev) & 1u);
              act_phase ^= (1u << ev);
              umma::tc_fence_after();
              if (ev == 0) L2A_STAMP(4 * l + 3);
              const int kc_end = min(nkc, 4 * ev + 4);
              for (int mb = 0; mb < nmb; ++mb) {
                for (int kc = 4 * ev; kc < kc_end; ++kc) tile_pair(IntTag<1>{}, mb ? dB : dA, kc, kc == 0, (kc == nkc - 1) ? nks_last : 4);
                if (mb == 0 && ev == nsrc - 1 && nsrc > 1) { commit(early); L2A_STAMP(4 * l + 1); }
              }
            }
            if (nsrc == 1) commit(early);
          }
          commit(layer_full);
          L2A_STAMP(4 * l + 2);
          slot_a = (slot_a + 2) % 3;
        }
        // ---- output layer, roles swapped: D[cand, feat] (+)= X[cand, 64-chunk] * W[64-chunk, feat]; A = each CTA's own
        // activation rows, B = the [out_n x 64] weight tiles split by N between the CTAs (see the header)
        {
          const int l = L - 1;
          const int nkc = plan.nkc[l], nsrc = plan.nmb[l - 1];
          const uint32_t d_out = tmem_base + (uint32_t)(slot_a * N2);
          int j = 0;
          L2A_STAMP(4 * l + 0);
          for (int ev = 0; ev < nsrc; ++ev) {
            umma::mbar_wait(&act_ready[ev], (act_phase >> ev) & 1u);
            act_phase ^= (1u << ev);
            umma::tc_fence_after();
            if (ev == 0) L2A_STAMP(4 * l + 3);
            const int kc_end = min(nkc, 4 * ev + 4);
            for (int kc = 4 * ev; kc < kc_end; ++kc) {
              if (j == 0) {
                umma::mbar_wait(&full_hi[stage], phase);
                if (kSplit) umma::mbar_wait(&full_lo[stage], phase);
                umma::tc_fence_after();
              }
              const uint32_t xh = hi_mn32 + (uint32_t)kc * kChunkStep, xl = lo_mn32 + (uint32_t)kc * kChunkStep;
              const uint32_t wy = st_lo32 + (uint32_t)stage * kStageStep + (uint32_t)j * out_kc_step, wx = wy + out_x_step;
              const bool last_in_stage = (j + 1 == plan.out_kcs) || (kc == nkc - 1);
              if (umma::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  umma::mma2_bf16_ss_ab<umma::kANone>(d_out, xh + kKsMn * ks, umma::kDescHi32Mn, wy + 2 * ks, umma::kDescHi32, idesc_out2,
                                                      (kc == 0 && ks == 0) ? 0u : 1u);                                               // x_hi * [W_hi ; W_lo]
                  umma::mma2_bf16_ss_ab<umma::kANone>(d_out, xl + kKsMn * ks, umma::kDescHi32Mn, wx + 2 * ks, umma::kDescHi32, idesc_out, 1u);   // x_lo * W_hi
                }
                if (last_in_stage) {
                  umma::mma2_commit_both(&empty_hi[stage]);
                  if (kSplit) umma::mma2_commit_both(&empty_lo[stage]);
                }
              }
              __syncwarp();
              if (last_in_stage) { j = 0; advance(); } else ++j;
            }
          }
          commit(layer_full);
          L2A_STAMP(4 * l + 2);
          slot_a = (slot_a + 2) % 3;
        }
      }
    }
  } else {
    // ================================================================ epilogue + env step (warps 0-3, 128 threads)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    const int n = tid;
    const bool has_cand = n < NC;
    const bool valid = n < nvalid;
    const long long row = (long long)env * a.n_candidates + c0 + (valid ? n : 0);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t lf_phase = 0, early_phase = 0, in_phase = 0;
    int slot_a = 0;
    float ret = 0.f, asq = 0.f;
    constexpr int AMAX = (DMAX <= 24) ? 8 : kTcMaxAct;
    float a_cur[AMAX];
    float st[DMAX];
#pragma unroll
    for (int k = 0; k < DMAX; ++k) st[k] = (k < D) ? __ldg(a.obs0 + (size_t)env * D + k) : 0.f;
    const int d8 = tc_obs_pad(D);

    auto load_actions = [&](int t) {
      const float* src = a.actions + (long long)t * a.act_stride_t + row * a.act_stride_row;
#pragma unroll
      for (int j = 0; j < AMAX; ++j) a_cur[j] = (j < A && valid && has_cand) ? __ldg(src + j) : 0.f;
    };
    auto write_x = [&]() {
      if (has_cand) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < AMAX; ++j) s = fmaf(a_cur[j], a_cur[j], s);
        asq = s;
        auto store_group = [&](int g, const float (&v)[8]) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) umma::split_bf16x2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
          const uint32_t off = (uint32_t)kTc2XChunk * (uint32_t)kChunkBytes + umma::sw128_offset((uint32_t)n, (uint32_t)g * 8u);
          *reinterpret_cast<uint4*>(act_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(act_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        };
#pragma unroll
        for (int g = 0; g < DMAX / 8; ++g) {
          if (g * 8 < d8) {
            float mu[8], rd[8], v[8];
            *reinterpret_cast<float4*>(&mu[0]) = *reinterpret_cast<const float4*>(n_obs_mean + g * 8);
            *reinterpret_cast<float4*>(&mu[4]) = *reinterpret_cast<const float4*>(n_obs_mean + g * 8 + 4);
            *reinterpret_cast<float4*>(&rd[0]) = *reinterpret_cast<const float4*>(n_obs_den + g * 8);
            *reinterpret_cast<float4*>(&rd[4]) = *reinterpret_cast<const float4*>(n_obs_den + g * 8 + 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float x = (st[g * 8 + i] - mu[i]) * rd[i];                              // mlp_dynamics.py:265-266
              v[i] = (g * 8 + i < D) ? x : 0.f;
            }
            store_group(g, v);
          }
        }
#pragma unroll
        for (int ga = 0; ga < AMAX / 8; ++ga) {
          if (ga * 8 < A) {
            float mu[8], rd[8], v[8];
            *reinterpret_cast<float4*>(&mu[0]) = *reinterpret_cast<const float4*>(n_act_mean + ga * 8);
            *reinterpret_cast<float4*>(&mu[4]) = *reinterpret_cast<const float4*>(n_act_mean + ga * 8 + 4);
            *reinterpret_cast<float4*>(&rd[0]) = *reinterpret_cast<const float4*>(n_act_den + ga * 8);
            *reinterpret_cast<float4*>(&rd[4]) = *reinterpret_cast<const float4*>(n_act_den + ga * 8 + 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float x = (a_cur[ga * 8 + i] - mu[i]) * rd[i];
              v[i] = (ga * 8 + i < A) ? x : 0.f;
            }
            store_group(d8 / 8 + ga, v);
          }
        }
        {
          const int used = d8 / 8 + (A + 7) / 8, need = plan.nks_last[0] * 2;
          const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int g = used; g < need; ++g) store_group(g, z);
        }
      }
      umma::fence_proxy_async_smem();              // (this CTA's own shared memory only)
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) umma::mbar_arrive(x_ready);
        else umma::mbar_arrive_remote_relaxed(x_ready_leader);     // own shared memory only, performed by the CTA-scope proxy fence above
      }
    };

    load_actions(0);
    write_x();

    for (int t = 0; t < H; ++t) {
      const float disc_t = __ldg(a.discount_pow + t);
      for (int l = 0; l + 1 < L; ++l) {
        hidden_epilogue(IntTag<0>{}, t, l, slot_a, warp, warp == 0, lf_phase, early_phase, in_phase);
        slot_a = (slot_a + 2) % 3;
      }
      if (t + 1 < H) load_actions(t + 1);
      umma::mbar_wait(layer_full, lf_phase);
      lf_phase ^= 1u;
      umma::tc_fence_after();
      if (warp == 0) L2A_STAMP(60);
      float dl[DMAX];
      {
        uint32_t r[DMAX], r2[DMAX];
        const uint32_t t_addr = tmem_base + lane_base + (uint32_t)(slot_a * N2);
#pragma unroll
        for (int g = 0; g < DMAX / 8; ++g) umma::tmem_ld_32x32b_x8(t_addr + (uint32_t)(g * 8), &r[g * 8]);
#pragma unroll
        for (int g = 0; g < DMAX / 8; ++g) umma::tmem_ld_32x32b_x8(t_addr + (uint32_t)(plan.out_n + g * 8), &r2[g * 8]);
        umma::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < DMAX / 4; ++q) {
          const float4 b = *reinterpret_cast<const float4*>(n_bias_out + 4 * q);
          const float4 sc = *reinterpret_cast<const float4*>(n_dscale + 4 * q);
          const float4 mu = *reinterpret_cast<const float4*>(n_dmean + 4 * q);
          const float y0 = ((__uint_as_float(r[4 * q]) + __uint_as_float(r2[4 * q])) + b.x) * sc.x + mu.x;        // mlp_dynamics.py:269-270
          const float y1 = ((__uint_as_float(r[4 * q + 1]) + __uint_as_float(r2[4 * q + 1])) + b.y) * sc.y + mu.y;
          const float y2 = ((__uint_as_float(r[4 * q + 2]) + __uint_as_float(r2[4 * q + 2])) + b.z) * sc.z + mu.z;
          const float y3 = ((__uint_as_float(r[4 * q + 3]) + __uint_as_float(r2[4 * q + 3])) + b.w) * sc.w + mu.w;
          dl[4 * q] = (4 * q < D) ? y0 : 0.f;
          dl[4 * q + 1] = (4 * q + 1 < D) ? y1 : 0.f;
          dl[4 * q + 2] = (4 * q + 2 < D) ? y2 : 0.f;
          dl[4 * q + 3] = (4 * q + 3 < D) ? y3 : 0.f;
        }
      }
      slot_a = (slot_a + 2) % 3;
      umma::tc_fence_before();
      if (warp == 0) L2A_STAMP(61);
#if L2A_TC2_LL
      if (ensemble) {
        // Member exchange through L2, flag in data: a row is DMAX/2 16-byte units {d[2q], step+1, d[2q+1], step+1}, laid out
        // [member][rank][unit][candidate] (a warp's store of one unit is 512 contiguous bytes).  Every thread publishes its own row with
        // relaxed gpu-scope 16-byte stores and polls the same row of the other members until both step words of every unit match
        // (each 8-byte {value, step} half is single-copy atomic, so a matching word carries its value): one L2 round trip after
        // the slowest member's stores, no fence, no barrier.  Members are averaged in member order 0..E-1 (bit-identical on every
        // member).  Double-buffered by step parity; the host zeroes the scratch before every launch.
        constexpr int DU = DMAX / 2;
        const int du = (D + 1) >> 1;
        constexpr size_t kBlkU = (size_t)NC * DU;
        uint4* const blk0 = reinterpret_cast<uint4*>(a.xch) + ((size_t)(tile_id * 2 + (t & 1)) * csize) * 2 * kBlkU;
        const uint32_t flag = (uint32_t)(t + 1);
        if (has_cand) {
          uint4* mine = blk0 + (size_t)(member * 2 + (int)rank) * kBlkU + n;
#pragma unroll
          for (int q = 0; q < DU; ++q)
            if (q < du) umma::st_relaxed_gpu_v4(mine + q * NC, make_uint4(__float_as_uint(dl[2 * q]), flag, __float_as_uint(dl[2 * q + 1]), flag));
        }
        if (warp == 0) L2A_STAMP(65);
        L2A_TIMELINE(if (a.timeline && (blockIdx.x & 1) == 0 && blockIdx.x < 10 && t == 1 && tid == 0) a.timeline[86 + blockIdx.x] = (long long)umma::globaltimer_ns());
        if (has_cand) {
          constexpr int EB = 2, QB = (DMAX <= 24) ? DU : 6;               // members x units polled together (register budget)
          float acc[DMAX];
#pragma unroll
          for (int k = 0; k < DMAX; ++k) acc[k] = 0.f;
          const uint4* rows = blk0 + (size_t)rank * kBlkU + n;              // member e: + e * 2 * kBlkU
#pragma unroll
          for (int qb = 0; qb < DU; qb += QB) {
            if (qb < du) {
              for (int e0 = 0; e0 < csize; e0 += EB) {
                uint4 u[EB][QB];
                const long long w0 = clock64();
                bool ok;
                do {
                  ok = true;
#pragma unroll
                  for (int j = 0; j < EB; ++j)
#pragma unroll
                    for (int q = 0; q < QB; ++q)
                      if (qb + q < DU && e0 + j < csize && e0 + j != member && qb + q < du) {
                        u[j][q] = umma::ld_relaxed_gpu_v4(rows + (size_t)(e0 + j) * 2 * kBlkU + (qb + q) * NC);
                        ok = ok && (u[j][q].y == flag) && (u[j][q].w == flag);
                      }
                  if (!ok && clock64() - w0 > L2A_WATCHDOG_CYCLES) __trap();
                } while (!ok);
#pragma unroll
                for (int j = 0; j < EB; ++j)
                  if (e0 + j < csize) {
                    const bool own = (e0 + j == member);
#pragma unroll
                    for (int q = 0; q < QB; ++q)
                      if (qb + q < DU) {
                        const int k = 2 * (qb + q);
                        acc[k] += own ? dl[k] : __uint_as_float(u[j][q].x);
                        acc[k + 1] += own ? dl[k + 1] : __uint_as_float(u[j][q].z);
                      }
                  }
              }
            }
          }
          const float inv_e = 1.0f / (float)csize;
#pragma unroll
          for (int k = 0; k < DMAX; ++k) dl[k] = (k < 2 * du) ? acc[k] * inv_e : dl[k];
        }
        if (warp == 0) L2A_STAMP(67);
      }
#else
      if (ensemble) {
        // Member exchange through L2 (see the header): rows stored -> gpu-scope fence -> CTA barrier -> one release flag per
        // (member, rank) -> the E-1 other flags acquired -> CTA barrier -> the other members' rows read with ld.global.cg and
        // averaged in member order 0..E-1 (bit-identical on every member).  Double-buffered by step parity.
        constexpr int DQ = DMAX / 4;
        const int dq = (D + 3) >> 2;
        constexpr size_t kBlk = (size_t)NC * DQ;                                   // float4 units of one (member, rank) block
        float4* const blk0 = reinterpret_cast<float4*>(a.xch) + ((size_t)(tile_id * 2 + (t & 1)) * csize) * 2 * kBlk;
        if (has_cand) {
          float4* mine = blk0 + (size_t)(member * 2 + (int)rank) * kBlk + n;
#pragma unroll
          for (int q = 0; q < DQ; ++q)
            if (q < dq) mine[q * NC] = make_float4(dl[4 * q], dl[4 * q + 1], dl[4 * q + 2], dl[4 * q + 3]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 0) L2A_STAMP(65);
        unsigned int* const flags = a.flags + ((size_t)tile_id * 2 + rank) * csize;
        if (tid == 0) {                              // one gpu-scope release for the CTA's rows (cumulative through the barrier)
          umma::fence_acq_rel_gpu();
          umma::st_relaxed_gpu(flags + member, (unsigned int)(t + 1));
        }
        if (warp == 0) L2A_STAMP(66);
        L2A_TIMELINE(if (a.timeline && (blockIdx.x & 1) == 0 && blockIdx.x < 10 && t == 1 && tid == 0) a.timeline[86 + blockIdx.x] = (long long)umma::globaltimer_ns());
        if (tid < csize && tid != member) {
          const long long w0 = clock64();
          while (umma::ld_acquire_gpu(flags + tid) < (unsigned int)(t + 1)) {
            if (clock64() - w0 > L2A_WATCHDOG_CYCLES) __trap();
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 0) L2A_STAMP(67);
        L2A_TIMELINE(if (a.timeline && (blockIdx.x & 1) == 0 && blockIdx.x < 10 && t == 1 && tid == 0) a.timeline[87 + blockIdx.x] = (long long)umma::globaltimer_ns());
        if (has_cand) {
          const float inv_e = 1.0f / (float)csize;
          const float4* rows = blk0 + (size_t)rank * kBlk + n;                     // member e: + e * 2 * kBlk
          auto mean_rows = [&](auto em_tag, auto qb_tag) {
            constexpr int EM = decltype(em_tag)::value, QB = decltype(qb_tag)::value;
#pragma unroll
            for (int qb = 0; qb < DQ; qb += QB) {
              if (qb < dq) {
                float4 v[EM][QB];
#pragma unroll
                for (int e = 0; e < EM; ++e)
#pragma unroll
                  for (int q = 0; q < QB; ++q)
                    if (qb + q < DQ && e < csize && e != member && qb + q < dq) v[e][q] = __ldcg(rows + (size_t)e * 2 * kBlk + (qb + q) * NC);
#pragma unroll
                for (int q = 0; q < QB; ++q) {
                  if (qb + q >= DQ) continue;
                  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                  const int k0 = 4 * (qb + q);
#pragma unroll
                  for (int e = 0; e < EM; ++e)
                    if (e < csize) {
                      const bool own = (e == member);
                      acc.x += own ? dl[k0] : v[e][q].x;
                      acc.y += own ? dl[k0 + 1] : v[e][q].y;
                      acc.z += own ? dl[k0 + 2] : v[e][q].z;
                      acc.w += own ? dl[k0 + 3] : v[e][q].w;
                    }
                  if (qb + q < dq) { dl[k0] = acc.x * inv_e; dl[k0 + 1] = acc.y * inv_e; dl[k0 + 2] = acc.z * inv_e; dl[k0 + 3] = acc.w * inv_e; }
                }
              }
            }
          };
          if (csize <= 5) mean_rows(IntTag<5>{}, IntTag<(DMAX <= 24) ? 5 : 2>{});
          else mean_rows(IntTag<8>{}, IntTag<1>{});
        }
      }
#endif
      if (warp == 0) L2A_STAMP(62);
      // ---------------- env step: (mean) delta -> reward -> state update -> next normalised input
      if (has_cand) {
        float dx = 0.f, nx0 = 0.f, nx1 = 0.f, nx2 = 0.f;
#pragma unroll
        for (int k = 0; k < DMAX; ++k) {
          const float d = dl[k];
          const float s_new = st[k] + d;                        // mlp_dynamics.py:220
          st[k] = s_new;
          dx = (k == D - 3) ? d : dx;
          nx0 = (k == D - 3) ? s_new : nx0;
          nx1 = (k == D - 2) ? s_new : nx1;
          nx2 = (k == D - 1) ? s_new : nx2;
        }
        const float rew = reward_value(a.reward_kind, 0.f, a.dt, asq, dx, nx0, nx1, nx2);
        ret = fmaf(disc_t, rew, ret);                        // mpc_controller.py:126
      }
      if (warp == 0) L2A_STAMP(63);
      if (t + 1 < H) write_x();
      if (warp == 0) L2A_STAMP(64);
    }

    float v = -__int_as_float(0x7f800000);
    int idx = 0x7fffffff;
    if (valid && has_cand) { v = ret; idx = c0 + n; }
    if (member == 0 && a.returns && valid && has_cand) a.returns[(size_t)env * a.n_candidates + c0 + n] = ret;
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

  // ------------------------------------------------------------------ teardown
  umma::tc_fence_before();
  __syncthreads();
  if (member == 0) publish_and_reduce(a.red, env, group * 2 + (int)rank, red_v[0], red_i[0], tid, s_flag);
  umma::cluster_sync_all();                      // the peer's shared memory / TMEM stay alive until both CTAs are done
  if (warp == 5) {
    umma::tc_fence_after();
    umma::tmem_dealloc2<512>(tmem_base);
  }
}

}  // namespace l2a
