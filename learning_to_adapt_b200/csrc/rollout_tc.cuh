// K1, tensor-core variant: persistent fused H-step rollout on tcgen05 (sm_100a).
//
// Orientation: the dense layers are computed TRANSPOSED, Y^T[out, cand] = W^T[out, in] * X^T[in, cand], so that
//   * the MMA "M" dimension (128 TMEM lanes) is the layer's output features  -> bias / ReLU are per-thread scalars,
//   * the MMA "N" dimension is the CTA's NC candidates                       -> NC need only be a multiple of 16,
//   * A = W^T tiles [128 out x 64 in] stream from L2 by TMA bulk copies of pre-swizzled (hi, lo) tile pairs (32 KB, one
//     full/empty mbarrier handshake per pair = per 12 MMAs),
//   * B = activations [NC cand x K] live in shared memory for the whole rollout (K-major, 128B swizzle), updated in place,
//   * D = fp32 accumulators in TMEM: six 128 x NC slots, rotated so the epilogue of layer l overlaps the first MMAs of
//     layer l+1 (K-outer "phase A" on the activation chunks as they are published).
// Precision: split-bf16.  Every fp32 operand x is carried as hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits) and
// each product is three MMA passes  W_hi*x_hi + W_hi*x_lo + W_lo*x_hi  accumulated in fp32 (measured return error
// vs the fp32/fp64 reference: ~5e-6 relative).  A single bf16 pass misses the 1e-4 parity bar.
//
// Warp roles (384 threads = three warpgroups): warps 0-3 epilogue + "env step" (16x256b TMEM fragments -> bias/ReLU/split ->
// stmatrix; the candidate's state in registers, reward, normalisation, argmax), warp 4 TMA producer (one lane), warp 5 MMA
// issuer (the whole warp walks the loops so descriptors stay in uniform registers; an elect.sync lane issues) + TMEM owner,
// warps 6-9 epilogue helpers: warp w shares TMEM lane quadrant w % 4 with epilogue warp w % 4 and converts the upper
// candidate blocks of every hidden-layer M-block (one warp per scheduler cannot hide its own ALU latency: the epilogue of
// an M-block took 1.35 k cycles, and the output layer, which needs the whole previous epilogue, was bound by it).
// Registers: launched at 168 per thread; warpgroups 1 and 2 release down to 128 (setmaxnreg.dec: 2 x 128 x 40 registers back
// to the CTA pool), warpgroup 0 grows to 240 (needs 128 x 72).
// The OUTPUT layer runs with the roles swapped, D[cand, feat] = X[cand, in] * W[in, feat]: the resident activation chunks
// are the A operand (M = 128 TMEM lanes = candidates; rows >= NC read whatever follows the chunk and land in unused lanes),
// the weights are small [out_n x 64] B tiles (out_n = obs dim padded to 16).  The MMA N drops from NC to out_n, the weight
// stream of the layer from 256 KB to 64 KB per step, and every candidate's D deltas arrive in the registers of its own
// env-step thread with one tcgen05.ld -- no transposition through shared memory.
// Ensemble mode (BASELINE "ensemble=E"): a thread-block cluster of E CTAs, one member each, same candidates; every step
// each candidate thread publishes its member's denormalised delta row to an L2-resident scratch, the cluster meets on an
// mbarrier, and the thread reads back its row of the other E-1 members and averages in member order, so all E CTAs carry
// bit-identical states.
//
// Replaces policies/mpc_controller.py:116-129 + dynamics/mlp_dynamics.py:204-222 / meta_mlp_dynamics.py:296-306.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace l2a {

constexpr int kTcTileBytes = 16384;     // one [128 x 64] bf16 weight tile (hi or lo part)
constexpr int kTcMaxStages = 4;         // ring depth is per NC: as many 32 KB (hi, lo) pair stages as shared memory allows (2 at NC=80)
constexpr int kTcMaxChunks = 8;         // activation width <= 512
constexpr int kTcThreads = 384;         // three warpgroups: 0 = epilogue + env step, 1 = producer / MMA issuer / helpers, 2 = helpers
constexpr int kTcMaxAct = 16;           // action dim limit of this variant
constexpr int kTcMaxObs = 48;           // obs dim limit of this variant (candidate state lives in registers)

// Schedule options of the rollout kernel (compile-time; see DESIGN.md "K1 schedule"):
//  L2A_TC_EARLY_EPI  the epilogue of M-blocks 0 and 1 of a hidden layer starts while the layer's phase B is still running, as
//                    soon as phase B has consumed the activation chunks that epilogue overwrites in place (chunks 0,1 / 2,3);
//                    the layer-0 input then lives in the LAST chunk so that layer 0 can do the same.
//  L2A_TC_PIPE_WAIT  the MMA issuer tests the next ring stage's full barrier (non-blocking) in the middle of the current
//                    stage's MMA burst, so a stage that has already landed costs no handshake bubble.
#ifndef L2A_TC_EARLY_EPI
#define L2A_TC_EARLY_EPI 1
#endif
#ifndef L2A_TC_PIPE_WAIT
#define L2A_TC_PIPE_WAIT 0
#endif
//  L2A_TC_SPLIT_RING the W_hi and W_lo halves of a ring stage travel separately: two producer threads (warps 4 and 10), one 16 KB
//                    bulk copy each, own full / empty barriers.  The issuer runs the 8 W_hi MMAs of a pair first and releases
//                    the W_hi half right after them, so its refill is requested 4 MMAs earlier and has to cover only 16 KB of
//                    transfer before the next use: the 2-deep ring's refill latency (~860 cycles for 32 KB against 563 of MMA
//                    time per pair) drops below the MMA time of the stages in between.
#ifndef L2A_TC_SPLIT_RING
#define L2A_TC_SPLIT_RING 1
#endif
// timing experiments only (results are garbage): 1 = the issuer never waits for the weight ring; 2 = the producers signal the
// stages full without copying anything
#ifndef L2A_TC_EXPERIMENT
#define L2A_TC_EXPERIMENT 0
#endif
constexpr int kTcXChunk = L2A_TC_EARLY_EPI ? (kTcMaxChunks - 1) : 0;   // activation chunk that holds the layer-0 input

// Tile enumeration of one weight set's blob; consumed in exactly this order by the kernel.  Within a layer the M-blocks
// are split into phase A (the first min(2, nmb) blocks) and phase B (the rest); each phase is K-OUTER:
//   A: for kc: for mb in A: (hi, lo)      B: for kc: for mb in B: (hi, lo)
// so that phase A of layer l+1 can start on the first activation chunks while layer l's epilogue is still writing the
// later ones (see the schedule notes in rollout_tc_kernel).
struct TcPlan {
  int n_layers;
  int nmb[kMaxLayers];
  int nkc[kMaxLayers];
  int nks_last[kMaxLayers];
  int tile_off[kMaxLayers];
  int l0_packed;         // layer 0 uses <= 32 of a tile's 64 K columns: two M-blocks share one tile pair (K halves)
  int hidden_pairs;      // 32 KB ring stages (tile pairs) of the hidden layers
  int out_n;             // output layer: MMA N = obs dim padded to a multiple of 16
  int out_kcs;           // output layer: K chunks ([out_n x 64] hi + lo tiles) packed into one 32 KB ring stage
  int out_stages;        // output layer: ring stages
  int stages_per_set;    // hidden_pairs + out_stages
  long long set_bytes;   // stages_per_set * 32 KB
};

__host__ __device__ inline int tc_layer_pairs(const TcPlan& p, int l) {
  return (l == 0 && p.l0_packed) ? p.nmb[0] / 2 : p.nmb[l] * p.nkc[l];
}
__host__ __device__ inline int tc_tile_index(const TcPlan& p, int l, int mb, int kc, int part) {
  if (l == 0 && p.l0_packed) return p.tile_off[0] + (mb / 2) * 2 + part;     // M-blocks 2j, 2j+1 = K halves of pair j
  const int nmb = p.nmb[l], nkc = p.nkc[l];
  const int nA = nmb < 2 ? nmb : 2, nB = nmb - nA;
  if (mb < nA) return p.tile_off[l] + ((kc * nA + mb) * 2 + part);
  return p.tile_off[l] + nkc * nA * 2 + ((kc * nB + (mb - nA)) * 2 + part);
}

// Layer-0 input feature order of the tensor-core path: [obs (D) | zero pad to a multiple of 8 | act (A) | zero pad]; the
// padding lets the env step write whole 16-byte groups of state features and of action features with static register
// indices.  in0_of_col maps a layer-0 K column to the reference's input feature (or -1 for padding).
__host__ __device__ inline int tc_obs_pad(int D) { return (D + 7) & ~7; }
__host__ __device__ inline int tc_in0_of_col(int D, int A, int c) {
  const int d8 = tc_obs_pad(D);
  if (c < d8) return c < D ? c : -1;
  return (c - d8 < A) ? D + (c - d8) : -1;
}

inline bool tc_make_plan(const MlpDims& md, TcPlan* p) {
  if (md.act_dim > kTcMaxAct || md.obs_dim > kTcMaxObs || md.obs_dim < 3) return false;
  if (tc_obs_pad(md.obs_dim) + md.act_dim > 64) return false;
  p->n_layers = md.n_layers;
  if (md.n_layers < 2) return false;
  int off = 0;
  for (int l = 0; l < md.n_layers; ++l) {
    const int din = md.dims[l], dout = md.dims[l + 1];
    if (l + 1 < md.n_layers && (dout % 128 != 0 || dout > 64 * kTcMaxChunks)) return false;
    if ((dout + 127) / 128 > 4) return false;
    if (l > 0 && din % 64 != 0) return false;
    if (din > 64 * kTcMaxChunks) return false;
    p->nmb[l] = (dout + 127) / 128;
    const int din_eff = (l == 0) ? tc_obs_pad(md.obs_dim) + md.act_dim : din;   // padded layer-0 layout
    p->nkc[l] = (din_eff + 63) / 64;
    const int rem = din_eff - (p->nkc[l] - 1) * 64;
    p->nks_last[l] = (rem + 15) / 16;
    p->tile_off[l] = off;
    if (l == 0) p->l0_packed = (p->nkc[0] == 1 && p->nks_last[0] <= 2 && p->nmb[0] % 2 == 0) ? 1 : 0;
    if (l + 1 < md.n_layers) off += tc_layer_pairs(*p, l) * 2;
  }
  p->hidden_pairs = off / 2;
  // output layer (roles swapped): per K chunk one [out_n x 64] hi tile followed by the lo tile, out_kcs chunks per stage
  p->out_n = (md.obs_dim + 15) / 16 * 16;
  p->out_kcs = (2 * kTcTileBytes) / (2 * p->out_n * 128);
  p->out_stages = (p->nkc[md.n_layers - 1] + p->out_kcs - 1) / p->out_kcs;
  p->stages_per_set = p->hidden_pairs + p->out_stages;
  p->set_bytes = (long long)p->stages_per_set * 2 * kTcTileBytes;
  return true;
}

// fp32 [in, out] kernels -> bf16 hi/lo tiles, transposed to [out, in] (K-major) and pre-swizzled (SWIZZLE_128B),
// zero padded.  grid.x = (layer, mb, kc) triples of the hidden layers, then one block per K chunk of the output layer;
// grid.y = set.
struct PrepArgs {
  MlpDims dims;
  TcPlan plan;
  const float* params;
  uint8_t* blobs;
  int first_set;
};

__global__ void __launch_bounds__(256) tc_prep_kernel(const PrepArgs a) {
  const int set = a.first_set + blockIdx.y;
  if ((int)blockIdx.x >= a.plan.hidden_pairs) {
    // output layer: B-operand tile [out_n features x 64 inputs] of K chunk kc, hi part then lo part
    const int l = a.plan.n_layers - 1;
    const int kc = (int)blockIdx.x - a.plan.hidden_pairs;
    const int din = a.dims.dims[l], dout = a.dims.dims[l + 1];
    const float* W = a.params + (size_t)set * a.dims.set_stride + a.dims.w_off[l];
    const int part_bytes = a.plan.out_n * 128;
    uint8_t* tile_hi = a.blobs + (size_t)set * a.plan.set_bytes +
                       (size_t)(a.plan.hidden_pairs + kc / a.plan.out_kcs) * (2 * kTcTileBytes) +
                       (size_t)(kc % a.plan.out_kcs) * 2 * part_bytes;
    uint8_t* tile_lo = tile_hi + part_bytes;
    for (int item = threadIdx.x; item < a.plan.out_n * 8; item += blockDim.x) {
      const int f = item % a.plan.out_n, ch = item / a.plan.out_n;
      uint16_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kc * 64 + ch * 8 + i;
        const float w = (f < dout && k < din) ? W[(size_t)k * dout + f] : 0.f;
        umma::split_bf16(w, hi[i], lo[i]);
      }
      const uint32_t off = umma::sw128_offset(f, ch * 8);
      *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16), hi[4] | (hi[5] << 16), hi[6] | (hi[7] << 16));
      *reinterpret_cast<uint4*>(tile_lo + off) = make_uint4(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16), lo[4] | (lo[5] << 16), lo[6] | (lo[7] << 16));
    }
    return;
  }
  int l = 0, rem = blockIdx.x;
  while (l + 2 < a.plan.n_layers && rem >= tc_layer_pairs(a.plan, l)) { rem -= tc_layer_pairs(a.plan, l); ++l; }
  if (l == 0 && a.plan.l0_packed) {
    // packed layer 0: tile pair `rem` holds M-block 2*rem in K columns [0, 32) and M-block 2*rem + 1 in [32, 64)
    const int din = a.dims.dims[0], dout = a.dims.dims[1];
    const float* W = a.params + (size_t)set * a.dims.set_stride + a.dims.w_off[0];
    uint8_t* tile_hi = a.blobs + (size_t)set * a.plan.set_bytes + (size_t)tc_tile_index(a.plan, 0, 2 * rem, 0, 0) * kTcTileBytes;
    uint8_t* tile_lo = tile_hi + kTcTileBytes;
    for (int item = threadIdx.x; item < 128 * 8; item += blockDim.x) {
      const int r = item & 127, ch = item >> 7;
      const int f = (2 * rem + (ch >> 2)) * 128 + r;
      uint16_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = tc_in0_of_col(a.dims.obs_dim, a.dims.act_dim, (ch & 3) * 8 + i);
        const float w = (f < dout && k >= 0 && k < din) ? W[(size_t)k * dout + f] : 0.f;
        umma::split_bf16(w, hi[i], lo[i]);
      }
      const uint32_t off = umma::sw128_offset(r, ch * 8);
      *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16), hi[4] | (hi[5] << 16), hi[6] | (hi[7] << 16));
      *reinterpret_cast<uint4*>(tile_lo + off) = make_uint4(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16), lo[4] | (lo[5] << 16), lo[6] | (lo[7] << 16));
    }
    return;
  }
  const int mb = rem / a.plan.nkc[l], kc = rem % a.plan.nkc[l];
  const int din = a.dims.dims[l], dout = a.dims.dims[l + 1];
  const float* W = a.params + (size_t)set * a.dims.set_stride + a.dims.w_off[l];
  uint8_t* tile_hi = a.blobs + (size_t)set * a.plan.set_bytes + (size_t)tc_tile_index(a.plan, l, mb, kc, 0) * kTcTileBytes;
  uint8_t* tile_lo = tile_hi + kTcTileBytes;
  for (int item = threadIdx.x; item < 128 * 8; item += blockDim.x) {
    const int r = item & 127, ch = item >> 7;
    const int f = mb * 128 + r;
    uint16_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int k = kc * 64 + ch * 8 + i;
      if (l == 0) k = tc_in0_of_col(a.dims.obs_dim, a.dims.act_dim, k);
      const float w = (f < dout && k >= 0 && k < din) ? W[(size_t)k * dout + f] : 0.f;
      umma::split_bf16(w, hi[i], lo[i]);
    }
    const uint32_t off = umma::sw128_offset(r, ch * 8);
    *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16), hi[4] | (hi[5] << 16), hi[6] | (hi[7] << 16));
    *reinterpret_cast<uint4*>(tile_lo + off) = make_uint4(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16), lo[4] | (lo[5] << 16), lo[6] | (lo[7] << 16));
  }
}

struct TcArgs {
  MlpDims dims;
  TcPlan plan;
  NormDev norm;
  const float* params;
  const uint8_t* blobs;
  const float* obs0;
  const float* actions;
  long long act_stride_t, act_stride_row;
  const float* discount_pow;
  int n_candidates, n_envs, horizon;
  int set_mode, first_set, n_sets;
  int reward_kind;
  float dt;
  int groups_per_env;
  float* returns;
  ReduceArgs red;
  float* xch;                   // ensemble exchange scratch in global memory: [clusters][2][E][NC][DMAX] floats (L2-resident)
  long long* timeline;          // diagnostics: clock64 stamps of CTA 0 at step 1 (null = off)
};

template <int NC>
struct TcSmem {
  static constexpr int kChunkBytes = NC * 128;
  static constexpr size_t act_bytes = (size_t)2 * kTcMaxChunks * kChunkBytes;
  static constexpr size_t stage_off = act_bytes;
  // ring stage = one (hi, lo) tile pair = 32 KB = the three split-bf16 passes of a [128 x 64] weight block: one bulk copy,
  // one full/empty handshake per 12 MMAs (the per-handshake mbarrier latency of the issuing threads, ~300 cycles, is what
  // bounds a 16 KB-granular ring -- scripts/stream_probe.py)
  static constexpr int kStageBytes = 2 * kTcTileBytes;
  static constexpr int kFit = (int)((232448 - (long long)act_bytes - 2048) / kStageBytes);
  static constexpr int kStages = kFit > kTcMaxStages ? kTcMaxStages : kFit;
  static_assert(kStages >= 2, "no room for the weight-tile ring");
  static constexpr size_t misc_off = stage_off + (size_t)kStages * kStageBytes;
  static size_t total(int D, int A) {
    (void)D; (void)A;
    size_t misc = sizeof(float) * (5 * (size_t)kTcMaxObs + 2 * (size_t)kTcMaxAct) + 64 /*pad*/ + 40 * sizeof(uint64_t) + 64;
    return misc_off + misc;       // the dynamic shared window is 1024-byte aligned (checked in the kernel)
  }
};

template <int V> struct IntTag { static constexpr int value = V; };

#ifdef L2A_DEBUG_KERNELS
#define L2A_STAMP(slot) do { if (a.timeline && blockIdx.x == 0 && t == 1 && (threadIdx.x & 31) == 0) a.timeline[(slot)] = clock64(); } while (0)
#define L2A_TIMELINE(expr) do { expr; } while (0)
#else
#define L2A_STAMP(slot) do { } while (0)
#define L2A_TIMELINE(expr) do { } while (0)
#endif

// DMAX: compile-time bound of the observation dimension (24 with act_dim <= 8, or 48): sizes the register-resident candidate state and the
// unrolled env-step code (a 48-wide instance costs HalfCheetah's D = 20 twice the instruction-cache footprint).
template <int NC, int DMAX>
__global__ void __launch_bounds__(kTcThreads, 1) rollout_tc_kernel(const TcArgs a) {
  using S = TcSmem<NC>;
  constexpr int kChunkBytes = S::kChunkBytes;
  constexpr int kTcStages = S::kStages;
  // split hi / lo ring only where the ring is 2 deep (NC = 80): deeper rings (NC <= 64) already cover the refill latency and
  // measured slower with the extra handshakes (cfg1 +7 %, cfg2 +6 %; headline -1.4 %)
  constexpr bool kSplit = (L2A_TC_SPLIT_RING != 0) && (S::kStages == 2);
  constexpr uint32_t kIdesc = umma::make_idesc_bf16(128, NC);
  static_assert(NC % 16 == 0 && NC >= 16 && 6 * NC <= 512, "UMMA N constraint / six accumulator slots must fit the 512 TMEM columns");

  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  if ((umma::smem_u32(smem) & 1023u) != 0) __trap();     // SWIZZLE_128B atoms need 1024-byte alignment
  const MlpDims& md = a.dims;
  const TcPlan& plan = a.plan;
  const int D = md.obs_dim, A = md.act_dim, L = md.n_layers, H = a.horizon;

  uint8_t* act_hi = smem;
  uint8_t* act_lo = smem + (size_t)kTcMaxChunks * kChunkBytes;
  uint8_t* stages = smem + S::stage_off;
  // per-feature constants, zero padded to DMAX / kTcMaxAct so the env step reads them as 16-byte vectors without bounds tests
  float* n_obs_mean = reinterpret_cast<float*>(smem + S::misc_off);
  float* n_obs_den = n_obs_mean + DMAX;
  float* n_dmean = n_obs_den + DMAX;
  float* n_dscale = n_dmean + DMAX;
  float* n_bias_out = n_dscale + DMAX;        // output-layer bias of this CTA's weight set
  float* n_act_mean = n_bias_out + DMAX;
  float* n_act_den = n_act_mean + kTcMaxAct;
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(n_act_den + kTcMaxAct) + 15) & ~(uintptr_t)15);
  uint64_t* full = bars;                      // [kTcStages]
  uint64_t* empty = bars + kTcStages;         // [kTcStages]
  uint64_t* full2 = bars + 2 * kTcStages;     // [kTcStages]  W_lo halves (L2A_TC_SPLIT_RING)
  uint64_t* empty2 = bars + 3 * kTcStages;    // [kTcStages]
  uint64_t* layer_full = bars + 4 * kTcStages;
  uint64_t* act_ready = layer_full + 1;       // [4]: one barrier per readiness event (source M-block) of a layer's input, so the
                                              // MMA issuer can lag several events behind without mbarrier parity aliasing
  uint64_t* peer_ready = layer_full + 5;
  uint64_t* x_ready = layer_full + 6;         // the layer-0 input of the next step is written (the 128 env-step threads)
  uint64_t* early = layer_full + 7;           // [2]: M-block 0 / 1 of a hidden layer may be drained (L2A_TC_EARLY_EPI)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(layer_full + 9);
  float* red_v = reinterpret_cast<float*>(tmem_slot + 2);   // [4]
  int* red_i = reinterpret_cast<int*>(red_v + 4);           // [4]
  int* s_flag = red_i + 4;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool ensemble = (a.set_mode == L2A_SETS_ENSEMBLE_MEAN) && a.n_sets > 1;
  const int csize = ensemble ? a.n_sets : 1;
  const int crank = ensemble ? (int)umma::cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / csize;
  const int env = cluster_id / a.groups_per_env;
  const int group = cluster_id % a.groups_per_env;
  const int c0 = group * NC;
  const int nvalid = min(NC, a.n_candidates - c0);
  int set = a.first_set;
  if (a.set_mode == L2A_SETS_PER_ENV) set += env;
  if (ensemble) set += crank;
  const float* P = a.params + (size_t)set * md.set_stride;
  const uint8_t* blob = a.blobs + (size_t)set * plan.set_bytes;

  // ------------------------------------------------------------------ setup
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      umma::mbar_init(&full[s], 1); umma::mbar_init(&empty[s], 1);
      umma::mbar_init(&full2[s], 1); umma::mbar_init(&empty2[s], 1);
    }
    umma::mbar_init(layer_full, 1);
    for (int j = 0; j < 4; ++j) umma::mbar_init(&act_ready[j], 256);   // epilogue warps + helper warps
    umma::mbar_init(x_ready, 128);
    umma::mbar_init(&early[0], 1);
    umma::mbar_init(&early[1], 1);
    umma::mbar_init(peer_ready, csize);
    umma::fence_barrier_init();
  }
  if (warp == 5) umma::tmem_alloc<512>(tmem_slot);
  for (int i = tid; i < DMAX; i += kTcThreads) {
    const bool in = i < D;
    n_obs_mean[i] = in ? a.norm.obs_mean[i] : 0.f;
    n_obs_den[i] = in ? 1.0f / a.norm.obs_den[i] : 0.f;   // reciprocal: x_n = (x - mean) * (1 / (std + 1e-10))
    n_dmean[i] = in ? a.norm.delta_mean[i] : 0.f;
    n_dscale[i] = in ? a.norm.delta_scale[i] : 0.f;
    n_bias_out[i] = in ? P[md.b_off[L - 1] + i] : 0.f;
  }
  for (int i = tid; i < kTcMaxAct; i += kTcThreads) {
    n_act_mean[i] = (i < A) ? a.norm.act_mean[i] : 0.f;
    n_act_den[i] = (i < A) ? 1.0f / a.norm.act_den[i] : 0.f;
  }
  umma::tc_fence_before();
  __syncthreads();
  if (ensemble) umma::cluster_sync_all();      // peers' barriers are initialised before any remote arrive
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Hidden-layer epilogue of one layer for the candidate blocks [CB0, CB1) (16 candidates each) of this warp's TMEM lane
  // quadrant wq: accumulator fragments (16x256b TMEM loads: thread T holds features T/4, T/4+8 x candidate pairs) -> bias +
  // ReLU -> bf16 hi/lo split -> the next layer's K-major B operand with transposed 8x8 stmatrix stores (16-byte rows of 8
  // features per candidate), in place; one readiness event per M-block (chunks 2mb, 2mb+1 of the next layer's input).
  constexpr int kCbAll = NC / 16, kCbMain = (kCbAll + 1) / 2;          // the epilogue warps take the lower blocks, the helpers the rest
  static_assert(kCbMain < kCbAll, "the helper warps take part in every act_ready barrier: NC must be >= 32");
  auto hidden_epilogue = [&](auto cb0_tag, auto cb1_tag, int t, int l, int pair_a, int wq, bool stamps, uint32_t& lf_phase,
                             uint32_t& early_phase) {
    constexpr int CB0 = decltype(cb0_tag)::value, CB1 = decltype(cb1_tag)::value;
    const uint32_t act_hi_addr = umma::smem_u32(act_hi), act_lo_addr = umma::smem_u32(act_lo);
    const int cand_l = (lane & 7) + ((lane >> 4) & 1) * 8;          // stmatrix row owned by this lane (within 16 candidates)
    const int fsel = ((lane >> 3) & 1) * 8;                         // ... of the feature-group matrix 0 / +8
#if !L2A_TC_EARLY_EPI
    umma::mbar_wait(layer_full, lf_phase);
    umma::tc_fence_after();
    if (stamps) L2A_STAMP(32 + 4 * l + 0);
#endif
    for (int mb = 0; mb < plan.nmb[l]; ++mb) {
#if L2A_TC_EARLY_EPI
      // M-blocks 0, 1 (phase A accumulators): drained as soon as the issuer says phase B no longer reads the chunks they
      // overwrite; M-blocks 2, 3 after the whole layer
      if (mb < 2) {
        umma::mbar_wait(&early[mb], (early_phase >> mb) & 1u);
        umma::tc_fence_after();
        if (stamps && mb == 0) L2A_STAMP(32 + 4 * l + 0);
      } else if (mb == 2) {
        umma::mbar_wait(layer_full, lf_phase);
        umma::tc_fence_after();
      }
#endif
      const int slot = (mb < 2) ? (2 * pair_a + mb) : (2 * ((pair_a + 1) % 3) + (mb - 2));
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int fbase = mb * 128 + wq * 32 + half * 16;           // 16 features handled by this (warp, half)
        const float bias_a = __ldg(P + md.b_off[l] + fbase + (lane >> 2));
        const float bias_b = __ldg(P + md.b_off[l] + fbase + (lane >> 2) + 8);
        const uint32_t t_addr = tmem_base + ((uint32_t)(wq * 32 + half * 16) << 16) + (uint32_t)(slot * NC);
        const uint32_t chunk_off = (uint32_t)(fbase >> 6) * kChunkBytes;
        const uint32_t fcol = (uint32_t)((fbase & 63) + fsel) >> 3;  // 16-byte column of this lane's matrix rows
        uint32_t r[CB1 - CB0][8];
#pragma unroll
        for (int cb = CB0; cb < CB1; ++cb) umma::tmem_ld_16x256b_x2(t_addr + (uint32_t)(cb * 16), r[cb - CB0]);
        umma::tmem_ld_wait();                                        // one wait for the whole block
#pragma unroll
        for (int cb = CB0; cb < CB1; ++cb) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float b = (q & 1) ? bias_b : bias_a;
            const float v0 = fmaxf(__uint_as_float(r[cb - CB0][2 * q]) + b, 0.f);      // core/utils.py:119-126 (ReLU dense)
            const float v1 = fmaxf(__uint_as_float(r[cb - CB0][2 * q + 1]) + b, 0.f);
            umma::split_bf16x2(v0, v1, hi[q], lo[q]);
          }
          const uint32_t cand = (uint32_t)(cb * 16 + cand_l);
          const uint32_t off = chunk_off + cand * 128u + (((fcol ^ cand) & 7u) << 4);
          umma::stmatrix_x4_trans(act_hi_addr + off, hi[0], hi[1], hi[2], hi[3]);
          umma::stmatrix_x4_trans(act_lo_addr + off, lo[0], lo[1], lo[2], lo[3]);
        }
      }
      umma::fence_proxy_async_smem();
      umma::tc_fence_before();
      umma::mbar_arrive(&act_ready[mb]);
      if (stamps && mb == 0) L2A_STAMP(32 + 4 * l + 1);
    }
#if L2A_TC_EARLY_EPI
    if (plan.nmb[l] <= 2) umma::mbar_wait(layer_full, lf_phase);      // keeps the layer_full phase in step
    early_phase ^= 3u;
#endif
    lf_phase ^= 1u;
  };

  // Register re-partitioning between the warpgroups (see the header): the env-step warps need ~240, everything else < 128.
  // Each setmaxnreg sits at the top of its role's branch so that ptxas allocates the branch with that budget.
  if (warp >= 6) {
    // ================================================================ epilogue helpers (warps 6-9; 10, 11 idle)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 128;");
    if (warp < 10) {
      const int wq = warp & 3;
      uint32_t lf_phase = 0, early_phase = 0;
      int pair_a = 0;
      for (int t = 0; t < H; ++t) {
        for (int l = 0; l + 1 < L; ++l) {
          hidden_epilogue(IntTag<kCbMain>{}, IntTag<kCbAll>{}, t, l, pair_a, wq, false, lf_phase, early_phase);
          pair_a = (pair_a + 2) % 3;
        }
        umma::mbar_wait(layer_full, lf_phase);                       // the output layer's completion: keeps the phase in step
        lf_phase ^= 1u;
        pair_a = (pair_a + 2) % 3;
      }
    }
    else if (kSplit && warp == 10 && lane == 0) {
      // ============================================================== second TMA producer: the W_lo half of every ring stage
      int stage = 0;
      uint32_t phase = 0;
      const int out_part2 = 2 * plan.out_n * 128;
      const int out_nkc = plan.nkc[L - 1];
      for (int t = 0; t < H; ++t) {
        for (int pr = 0; pr < plan.stages_per_set; ++pr) {
          uint32_t bytes = S::kStageBytes;
          if (pr >= plan.hidden_pairs) bytes = (uint32_t)(min(plan.out_kcs, out_nkc - (pr - plan.hidden_pairs) * plan.out_kcs) * out_part2);
          const uint32_t bytes_l = bytes > (uint32_t)kTcTileBytes ? bytes - (uint32_t)kTcTileBytes : 0u;
          umma::mbar_wait(&empty2[stage], phase ^ 1u);
          if (bytes_l && L2A_TC_EXPERIMENT != 2) {
            umma::mbar_arrive_expect_tx(&full2[stage], bytes_l);
            umma::bulk_g2s(stages + (size_t)stage * S::kStageBytes + kTcTileBytes, blob + (size_t)pr * S::kStageBytes + kTcTileBytes, bytes_l, &full2[stage]);
          } else {
            umma::mbar_arrive(&full2[stage]);
          }
          if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 4) {
    // ================================================================ TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 128;");
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int out_part2 = 2 * plan.out_n * 128;               // one K chunk of the output layer: hi + lo tile
      const int out_nkc = plan.nkc[L - 1];
      for (int t = 0; t < H; ++t) {
        for (int pr = 0; pr < plan.stages_per_set; ++pr) {
          uint32_t bytes = S::kStageBytes;
          if (pr >= plan.hidden_pairs) bytes = (uint32_t)(min(plan.out_kcs, out_nkc - (pr - plan.hidden_pairs) * plan.out_kcs) * out_part2);
          if (kSplit && bytes > (uint32_t)kTcTileBytes) bytes = (uint32_t)kTcTileBytes;       // the W_hi half; warp 10 streams the rest
          umma::mbar_wait(&empty[stage], phase ^ 1u);
          if (L2A_TC_EXPERIMENT == 2) { umma::mbar_arrive(&full[stage]); if (++stage == kTcStages) { stage = 0; phase ^= 1u; } continue; }
          umma::mbar_arrive_expect_tx(&full[stage], bytes);
          umma::bulk_g2s(stages + (size_t)stage * S::kStageBytes, blob + (size_t)pr * S::kStageBytes, bytes, &full[stage]);
          if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 128;");
    // ================================================================ MMA issuer
    // The WHOLE warp walks the loops (so every descriptor is computed warp-uniformly and lands in uniform registers);
    // only the elected lane issues tcgen05.mma / tcgen05.commit.
    {
      int stage = 0;
      uint32_t phase = 0, act_phase = 0, xr_phase = 0;
      bool ready = false;        // full[stage] was already seen complete for `phase` by a mid-burst probe (L2A_TC_PIPE_WAIT)
      const uint32_t hi_lo32 = umma::desc_lo32(umma::smem_u32(act_hi)), lo_lo32 = umma::desc_lo32(umma::smem_u32(act_lo));
      const uint32_t st_lo32 = umma::desc_lo32(umma::smem_u32(stages));
      constexpr uint32_t kChunkStep = (uint32_t)kChunkBytes >> 4, kStageStep = (uint32_t)S::kStageBytes >> 4, kLoStep = (uint32_t)kTcTileBytes >> 4;
      // ring protocol of the consumer side: acquire (wait unless a probe already saw the stage full) -> MMAs -> commit -> advance
      auto acquire = [&]() {
        if (!ready && L2A_TC_EXPERIMENT != 1) umma::mbar_wait(&full[stage], phase);
        if (kSplit && L2A_TC_EXPERIMENT != 1) umma::mbar_wait(&full2[stage], phase);
        umma::tc_fence_after();
      };
      // whole stage consumed (elected lane only)
      auto release = [&]() {
        umma::mma_commit(&empty[stage]);
        if (kSplit) umma::mma_commit(&empty2[stage]);
      };
      auto probe_next = [&]() {
#if L2A_TC_PIPE_WAIT
        const int ns = (stage + 1 == kTcStages) ? 0 : stage + 1;
        ready = umma::mbar_test_wait(&full[ns], (ns == 0) ? (phase ^ 1u) : phase);
#endif
      };
      auto advance = [&](bool probed) {
        if (!probed) ready = false;
        if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
      };
      // one (hi tile, lo tile) pair = the three split-bf16 passes of one [128 x 64] weight block against activation chunk `ch`
      auto tile_pair = [&](uint32_t d_tmem, int ch, bool first, bool full_k, int nks_last) {
        const uint32_t bh = hi_lo32 + (uint32_t)ch * kChunkStep, bl = lo_lo32 + (uint32_t)ch * kChunkStep;
        if constexpr (kSplit) {
          const uint32_t a_hi = st_lo32 + (uint32_t)stage * kStageStep, a_lo = a_hi + kLoStep;
          const int nks = full_k ? 4 : nks_last;
          // the two W_hi passes of every k-step (W_hi slice read from shared memory once: A-collector keep / reuse) ...
          if (L2A_TC_EXPERIMENT != 1) umma::mbar_wait(&full[stage], phase);
          umma::tc_fence_after();
          if (umma::elect_one()) {
            if (full_k) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d_tmem, a_hi + 2 * ks, bh + 2 * ks, kIdesc, (first && ks == 0) ? 0u : 1u);
                umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d_tmem, a_hi + 2 * ks, bl + 2 * ks, kIdesc, 1u);
              }
            } else {
              for (int ks = 0; ks < nks; ++ks) {
                umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d_tmem, a_hi + 2 * ks, bh + 2 * ks, kIdesc, (first && ks == 0) ? 0u : 1u);
                umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d_tmem, a_hi + 2 * ks, bl + 2 * ks, kIdesc, 1u);
              }
            }
            umma::mma_commit(&empty[stage]);                 // the W_hi half may be refilled
          }
          __syncwarp();
          // ... then the W_lo * x_hi pass
          if (L2A_TC_EXPERIMENT != 1) umma::mbar_wait(&full2[stage], phase);
          umma::tc_fence_after();
          if (umma::elect_one()) {
            if (full_k) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma::mma_bf16_ss_lo(d_tmem, a_lo + 2 * ks, bh + 2 * ks, kIdesc, 1u);
            } else {
              for (int ks = 0; ks < nks; ++ks) umma::mma_bf16_ss_lo(d_tmem, a_lo + 2 * ks, bh + 2 * ks, kIdesc, 1u);
            }
            umma::mma_commit(&empty2[stage]);
          }
          __syncwarp();
          advance(false);
          return;
        }
        acquire();
        const uint32_t a_hi = st_lo32 + (uint32_t)stage * kStageStep, a_lo = a_hi + kLoStep;
        // per 16-wide k-step: W_hi * x_hi, W_hi * x_lo (the W_hi slice is fetched from shared memory once for the two:
        // A-collector keep / reuse), W_lo * x_hi
        if (full_k) {
          if (umma::elect_one()) {
            umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d_tmem, a_hi, bh, kIdesc, first ? 0u : 1u);
            umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d_tmem, a_hi, bl, kIdesc, 1u);
            umma::mma_bf16_ss_lo(d_tmem, a_lo, bh, kIdesc, 1u);
            umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d_tmem, a_hi + 2, bh + 2, kIdesc, 1u);
            umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d_tmem, a_hi + 2, bl + 2, kIdesc, 1u);
            umma::mma_bf16_ss_lo(d_tmem, a_lo + 2, bh + 2, kIdesc, 1u);
          }
          __syncwarp();
          probe_next();            // its latency hides behind the MMAs already queued on the tensor pipe
          if (umma::elect_one()) {
#pragma unroll
            for (int ks = 2; ks < 4; ++ks) {
              umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d_tmem, a_hi + 2 * ks, bh + 2 * ks, kIdesc, 1u);
              umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d_tmem, a_hi + 2 * ks, bl + 2 * ks, kIdesc, 1u);
              umma::mma_bf16_ss_lo(d_tmem, a_lo + 2 * ks, bh + 2 * ks, kIdesc, 1u);
            }
            umma::mma_commit(&empty[stage]);
          }
          __syncwarp();
          advance(true);
        } else {
          if (umma::elect_one()) {
            for (int ks = 0; ks < nks_last; ++ks) {
              umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d_tmem, a_hi + 2 * ks, bh + 2 * ks, kIdesc, (first && ks == 0) ? 0u : 1u);
              umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d_tmem, a_hi + 2 * ks, bl + 2 * ks, kIdesc, 1u);
              umma::mma_bf16_ss_lo(d_tmem, a_lo + 2 * ks, bh + 2 * ks, kIdesc, 1u);
            }
            umma::mma_commit(&empty[stage]);
          }
          __syncwarp();
          advance(false);
        }
      };
      // Packed layer 0 (its input is <= 32 wide): one stage holds TWO M-blocks, W[mb0] in K columns [0, 32) of the tile and
      // W[mb1] in [32, 64); both multiply the input chunk, k-steps 0 .. nks-1, into the two accumulators of a slot pair.
      // Halves the layer's weight stream and its stage handshakes.
      auto packed_pair = [&](uint32_t d_tmem0, int nks) {
        acquire();
        const uint32_t a_hi = st_lo32 + (uint32_t)stage * kStageStep, a_lo = a_hi + kLoStep;
        const uint32_t xh = hi_lo32 + (uint32_t)kTcXChunk * kChunkStep, xl = lo_lo32 + (uint32_t)kTcXChunk * kChunkStep;
        if (umma::elect_one()) {
#pragma unroll
          for (int mbsel = 0; mbsel < 2; ++mbsel) {
            const uint32_t d = d_tmem0 + (uint32_t)(mbsel * NC);
            for (int ks = 0; ks < nks; ++ks) {
              const uint32_t ka = (uint32_t)(2 * (2 * mbsel + ks)), kb = (uint32_t)(2 * ks);
              umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d, a_hi + ka, xh + kb, kIdesc, ks == 0 ? 0u : 1u);
              umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d, a_hi + ka, xl + kb, kIdesc, 1u);
              umma::mma_bf16_ss_lo(d, a_lo + ka, xh + kb, kIdesc, 1u);
            }
          }
          release();
        }
        __syncwarp();
        advance(false);
      };
      // "M-blocks 0 / 1 of this layer may be drained": all MMAs issued so far (phase A entirely, phase B up to the chunks those
      // epilogues overwrite) are complete when the commit arrives
      auto commit_early = [&](int which) {
#if L2A_TC_EARLY_EPI
        if (umma::elect_one()) umma::mma_commit(&early[which]);
        __syncwarp();
#else
        (void)which;
#endif
      };
      // Accumulator slots: 3 pairs x 2 slots x NC TMEM columns.  Layer i accumulates its M-blocks {0,1} in pair a_i and
      // {2,3} in pair b_i = a_i + 1; a_{i+1} = a_i + 2 (mod 3) is the pair layer i does not touch, so phase A of layer i+1
      // can run while layer i's epilogue is still draining a_i / b_i, and b_{i+1} = a_i is drained by the time phase B starts.
      int pair_a = 0;
      const uint32_t idesc_out = umma::make_idesc_bf16(128, (uint32_t)plan.out_n);
      const uint32_t idesc_out2 = umma::make_idesc_bf16(128, (uint32_t)(2 * plan.out_n));
      const bool out_stacked = (2 * plan.out_n <= 2 * NC);             // x_hi * [W_hi; W_lo] as ONE N = 2*out_n MMA (see below)
      const uint32_t out_part = (uint32_t)(plan.out_n * 128) >> 4;     // one [out_n x 64] tile in descriptor units
      for (int t = 0; t < H; ++t) {
        for (int l = 0; l + 1 < L; ++l) {
          const int nmb = plan.nmb[l], nkc = plan.nkc[l], nks_last = plan.nks_last[l];
          const int nA = nmb < 2 ? nmb : 2;
          const int pair_b = (pair_a + 1) % 3;
          const int nsrc = (l == 0) ? 1 : plan.nmb[l - 1];          // readiness events of this layer's input
          const int cpe = (l == 0) ? nkc : 2;                       // activation chunks published per event
          const int ch0 = (l == 0) ? kTcXChunk : 0;                 // layer 0 reads its single input chunk from kTcXChunk
          L2A_STAMP(4 * l + 0);
          L2A_TIMELINE(if (a.timeline && blockIdx.x == 0 && t == 2 && l == 0 && lane == 0) a.timeline[80] = clock64());   // step length
          // phase A: K-outer over the chunks as the previous layer's epilogue publishes them
          for (int ev = 0; ev < nsrc; ++ev) {
            if (l == 0) {
              umma::mbar_wait(x_ready, xr_phase);
              xr_phase ^= 1u;
            } else {
              umma::mbar_wait(&act_ready[ev], (act_phase >> ev) & 1u);
              act_phase ^= (1u << ev);
            }
            umma::tc_fence_after();
            if (ev == 0) L2A_STAMP(4 * l + 3);
            if (l == 0 && plan.l0_packed) {
              packed_pair(tmem_base + (uint32_t)(2 * pair_a * NC), nks_last);       // M-blocks 0 and 1 from one stage
              continue;
            }
            const int kc_end = (ev == nsrc - 1) ? nkc : min(nkc, (ev + 1) * cpe);
            for (int kc = ev * cpe; kc < kc_end; ++kc) {
              const bool full_k = (kc != nkc - 1) || (nks_last == 4);
              for (int mb = 0; mb < nA; ++mb)
                tile_pair(tmem_base + (uint32_t)((2 * pair_a + mb) * NC), ch0 + kc, kc == 0, full_k, nks_last);
            }
          }
          L2A_STAMP(4 * l + 1);
          // phase B: every chunk is there and the previous layer's accumulators (pair_b) are drained.
          // Early drain of M-blocks 0 / 1: layer 0 keeps its input in chunk kTcXChunk, which those epilogues do not write ->
          // right after phase A; wider inputs -> once phase B is past chunks {0,1} / {2,3}; no phase B -> now.
          const bool early_now = (l == 0) || (nmb <= nA);
          if (early_now) { commit_early(0); commit_early(1); }
          if (l == 0 && plan.l0_packed) {
            if (nmb > 2) packed_pair(tmem_base + (uint32_t)(2 * pair_b * NC), nks_last);   // M-blocks 2 and 3
          } else
          for (int kc = 0; kc < nkc; ++kc) {
            const bool full_k = (kc != nkc - 1) || (nks_last == 4);
            for (int mb = nA; mb < nmb; ++mb)
              tile_pair(tmem_base + (uint32_t)((2 * pair_b + (mb - nA)) * NC), ch0 + kc, kc == 0, full_k, nks_last);
            if (!early_now) {
              if (kc == min(nkc - 1, 1)) commit_early(0);
              if (kc == min(nkc - 1, 3)) commit_early(1);
            }
          }
          if (umma::elect_one()) umma::mma_commit(layer_full);
          __syncwarp();
          L2A_STAMP(4 * l + 2);
          pair_a = (pair_a + 2) % 3;
        }
        // ---- output layer, roles swapped: D[cand, feat] (+)= X[cand, 64-chunk] * W[64-chunk, feat]; A = the resident
        // activation chunk (x_hi is kept in the collector for its two passes), B = [out_n x 64] weight tiles, out_kcs K
        // chunks per ring stage.  The accumulator is the first out_n columns of pair a.
        {
          const int l = L - 1;
          const int nkc = plan.nkc[l], nsrc = plan.nmb[l - 1];
          const uint32_t d_out = tmem_base + (uint32_t)(2 * pair_a * NC);
          int j = 0;                                                // K chunk within the current ring stage
          L2A_STAMP(4 * l + 0);
          for (int ev = 0; ev < nsrc; ++ev) {
            umma::mbar_wait(&act_ready[ev], (act_phase >> ev) & 1u);
            act_phase ^= (1u << ev);
            umma::tc_fence_after();
            if (ev == 0) L2A_STAMP(4 * l + 3);
            const int kc_end = (ev == nsrc - 1) ? nkc : min(nkc, (ev + 1) * 2);
            for (int kc = ev * 2; kc < kc_end; ++kc) {
              if (j == 0) acquire();
              const uint32_t xh = hi_lo32 + (uint32_t)kc * kChunkStep, xl = lo_lo32 + (uint32_t)kc * kChunkStep;
              const uint32_t wh = st_lo32 + (uint32_t)stage * kStageStep + (uint32_t)j * 2u * out_part, wl = wh + out_part;
              const bool last_in_stage = (j + 1 == plan.out_kcs) || (kc == nkc - 1);
              if (umma::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint32_t acc0 = (kc == 0 && ks == 0) ? 0u : 1u;
                  if (out_stacked) {
                    // the W_hi and W_lo tiles of a chunk are adjacent = one [2*out_n x 64] B operand: columns [0, out_n) of the
                    // accumulator take x_hi * W_hi (+ x_lo * W_hi below), columns [out_n, 2*out_n) take x_hi * W_lo; the epilogue
                    // adds the halves.  Two MMAs per k-step instead of three (a small-N MMA costs its 4 KB A read, not N).
                    umma::mma_bf16_ss_lo(d_out, xh + 2 * ks, wh + 2 * ks, idesc_out2, acc0);
                  } else {
                    umma::mma_bf16_ss_lo_hint<umma::kAKeep>(d_out, xh + 2 * ks, wh + 2 * ks, idesc_out, acc0);    // x_hi * W_hi
                    umma::mma_bf16_ss_lo_hint<umma::kAReuse>(d_out, xh + 2 * ks, wl + 2 * ks, idesc_out, 1u);     // x_hi * W_lo
                  }
                  umma::mma_bf16_ss_lo(d_out, xl + 2 * ks, wh + 2 * ks, idesc_out, 1u);                             // x_lo * W_hi
                }
                if (last_in_stage) release();
              }
              __syncwarp();
              if (last_in_stage) {
                j = 0;
                advance(false);
              } else {
                ++j;
              }
            }
          }
          if (umma::elect_one()) umma::mma_commit(layer_full);
          __syncwarp();
          L2A_STAMP(4 * l + 2);
          pair_a = (pair_a + 2) % 3;
        }
      }
    }
  } else {
    // ================================================================ epilogue + env step (warps 0-3, 128 threads)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    const int n = tid;                                  // candidate owned in the env phase
    const bool has_cand = n < NC;
    const bool valid = n < nvalid;
    const long long row = (long long)env * a.n_candidates + c0 + (valid ? n : 0);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t lf_phase = 0, pr_phase = 0, early_phase = 0;
    int pair_a = 0;                                     // same accumulator-pair rotation as the MMA issuer
    float ret = 0.f, asq = 0.f;
    constexpr int AMAX = (DMAX <= 24) ? 8 : kTcMaxAct;    // the small instance serves act_dim <= 8 (see launch dispatch)
    float a_cur[AMAX];
    float st[DMAX];                                      // this candidate's state, float32, in registers for the whole rollout
#pragma unroll
    for (int k = 0; k < DMAX; ++k) st[k] = (k < D) ? __ldg(a.obs0 + (size_t)env * D + k) : 0.f;
    const int d8 = tc_obs_pad(D);

    auto load_actions = [&](int t) {
      const float* src = a.actions + (long long)t * a.act_stride_t + row * a.act_stride_row;
#pragma unroll
      for (int j = 0; j < AMAX; ++j) a_cur[j] = (j < A && valid && has_cand) ? __ldg(src + j) : 0.f;
    };
    // normalised network input of the candidate for the step whose actions are in a_cur: features [state | action | 0]
    int t_stamp = -1;
    auto write_x = [&]() {
      if (has_cand) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < AMAX; ++j) s = fmaf(a_cur[j], a_cur[j], s);
        asq = s;
        // layer-0 input layout: [obs | pad8 | act | pad]  (tc_in0_of_col); groups of 8 features = one 16-byte store
        auto store_group = [&](int g, const float (&v)[8]) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) umma::split_bf16x2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
          const uint32_t off = (uint32_t)kTcXChunk * (uint32_t)kChunkBytes + umma::sw128_offset((uint32_t)n, (uint32_t)g * 8u);
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
        // zero the rest of the K range the MMAs read (nks_last 16-wide k-steps of the single chunk)
        {
          const int used = d8 / 8 + (A + 7) / 8, need = plan.nks_last[0] * 2;
          const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int g = used; g < need; ++g) store_group(g, z);
        }
      }
      L2A_TIMELINE(if (warp == 0 && t_stamp == 1 && a.timeline && blockIdx.x == 0 && lane == 0) a.timeline[70] = clock64());
      umma::fence_proxy_async_smem();
      L2A_TIMELINE(if (warp == 0 && t_stamp == 1 && a.timeline && blockIdx.x == 0 && lane == 0) a.timeline[71] = clock64());
      umma::tc_fence_before();
      umma::mbar_arrive(x_ready);
    };

    load_actions(0);
    write_x();

    for (int t = 0; t < H; ++t) {
      const float disc_t = __ldg(a.discount_pow + t);       // discount**t, fetched a whole step before its use
      // ---------------- hidden layers: TMEM -> bias + ReLU -> split -> next layer's B operand (in place)
      for (int l = 0; l + 1 < L; ++l) {
        hidden_epilogue(IntTag<0>{}, IntTag<kCbMain>{}, t, l, pair_a, warp, warp == 0, lf_phase, early_phase);
        if (warp == 0) L2A_STAMP(32 + 4 * l + 2);
        pair_a = (pair_a + 2) % 3;
      }
      // ---------------- output layer: this candidate's D outputs (TMEM lane = candidate) -> denormalised delta, in registers
      if (t + 1 < H) load_actions(t + 1);                 // prefetch the next step's actions (HBM) under the MMA wait
      umma::mbar_wait(layer_full, lf_phase);
      lf_phase ^= 1u;
      umma::tc_fence_after();
      if (warp == 0) L2A_STAMP(60);
      float dl[DMAX];
      {
        uint32_t r[DMAX];
        const uint32_t t_addr = tmem_base + lane_base + (uint32_t)(2 * pair_a * NC);
#pragma unroll
        for (int g = 0; g < DMAX / 8; ++g) umma::tmem_ld_32x32b_x8(t_addr + (uint32_t)(g * 8), &r[g * 8]);
        if (2 * plan.out_n <= 2 * NC) {                     // stacked output MMAs: add the x_hi * W_lo half (columns out_n ..)
          uint32_t r2[DMAX];
#pragma unroll
          for (int g = 0; g < DMAX / 8; ++g) umma::tmem_ld_32x32b_x8(t_addr + (uint32_t)(plan.out_n + g * 8), &r2[g * 8]);
          umma::tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < DMAX; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) + __uint_as_float(r2[k]));
        }
        umma::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < DMAX / 4; ++q) {
          const float4 b = *reinterpret_cast<const float4*>(n_bias_out + 4 * q);
          const float4 sc = *reinterpret_cast<const float4*>(n_dscale + 4 * q);
          const float4 mu = *reinterpret_cast<const float4*>(n_dmean + 4 * q);
          const float y0 = (__uint_as_float(r[4 * q]) + b.x) * sc.x + mu.x;                 // mlp_dynamics.py:269-270
          const float y1 = (__uint_as_float(r[4 * q + 1]) + b.y) * sc.y + mu.y;
          const float y2 = (__uint_as_float(r[4 * q + 2]) + b.z) * sc.z + mu.z;
          const float y3 = (__uint_as_float(r[4 * q + 3]) + b.w) * sc.w + mu.w;
          dl[4 * q] = (4 * q < D) ? y0 : 0.f;             // (columns >= D hold padding / stale accumulator data)
          dl[4 * q + 1] = (4 * q + 1 < D) ? y1 : 0.f;
          dl[4 * q + 2] = (4 * q + 2 < D) ? y2 : 0.f;
          dl[4 * q + 3] = (4 * q + 3 < D) ? y3 : 0.f;
        }
      }
      pair_a = (pair_a + 2) % 3;
      umma::tc_fence_before();
      if (warp == 0) L2A_STAMP(61);
      L2A_TIMELINE(if (a.timeline && blockIdx.x < 5 && t == 1 && tid == 0) a.timeline[86 + 2 * blockIdx.x] = clock64());       // member skew
      if (ensemble) {
        // Exchange of the E members' deltas through L2: every candidate thread publishes its own row (DMAX floats, 16-byte
        // stores), the cluster meets on the peers' mbarriers (release / acquire at cluster scope), then the thread reads
        // the same row of the other E-1 members (ld.global.cg) and averages in member order 0..E-1 -- identical on every
        // member.  The scratch is double-buffered by step parity: a peer can only be one exchange behind.
        // Scratch layout [member][16-byte unit q][candidate]: a warp's accesses to one unit are 512 contiguous bytes.
        constexpr int DQ = DMAX / 4;                         // 16-byte units per candidate row
        const int dq = (D + 3) >> 2;                         // ... that carry data
        float4* const blk0 = reinterpret_cast<float4*>(a.xch) + (size_t)(cluster_id * 2 + (t & 1)) * csize * (size_t)(NC * DQ);
        if (has_cand) {
          float4* mine = blk0 + (size_t)crank * (NC * DQ) + n;
#pragma unroll
          for (int q = 0; q < DQ; ++q)
            if (q < dq) mine[q * NC] = make_float4(dl[4 * q], dl[4 * q + 1], dl[4 * q + 2], dl[4 * q + 3]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 0) L2A_STAMP(65);
        if (tid < csize) umma::mbar_arrive_remote(umma::map_to_cta(umma::smem_u32(peer_ready), (uint32_t)tid));
        if (warp == 0) L2A_STAMP(66);
        umma::mbar_wait_cluster(peer_ready, pr_phase);
        pr_phase ^= 1u;
        if (warp == 0) L2A_STAMP(67);
        if (has_cand) {
          const float inv_e = 1.0f / (float)csize;
          const float4* rows = blk0 + n;
          // All rows of one batch are in flight together; a batch is bounded by the register budget (EM members x QB units x 4).
          // Up to five members (BASELINE's ensemble) and a small obs dim: the whole exchange is ONE batch = one L2 round trip.
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
                    if (qb + q < DQ && e < csize && e != crank && qb + q < dq) v[e][q] = __ldcg(rows + (size_t)e * (NC * DQ) + (qb + q) * NC);
#pragma unroll
                for (int q = 0; q < QB; ++q) {
                  if (qb + q >= DQ) continue;
                  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                  const int k0 = 4 * (qb + q);
#pragma unroll
                  for (int e = 0; e < EM; ++e)
                    if (e < csize) {
                      const bool own = (e == crank);
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
          else mean_rows(IntTag<8>{}, IntTag<1>{});          // 6-8 members: one unit per batch (register budget)
        }
      }
      if (warp == 0) L2A_STAMP(62);
      L2A_TIMELINE(if (a.timeline && blockIdx.x < 5 && t == 1 && tid == 0) a.timeline[87 + 2 * blockIdx.x] = clock64());
      // ---------------- env step: (mean) delta -> reward -> state update -> next normalised input
      if (has_cand) {
        float dx = 0.f, nx0 = 0.f, nx1 = 0.f, nx2 = 0.f;
#pragma unroll
        for (int k = 0; k < DMAX; ++k) {
          const float d = dl[k];                                // ensemble: already the member mean; 0 for k >= D
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
      t_stamp = t;
      if (t + 1 < H) write_x();
      if (warp == 0) L2A_STAMP(64);
    }

    // ---------------- per-CTA argmax; one CTA per cluster publishes
    float v = -__int_as_float(0x7f800000);
    int idx = 0x7fffffff;
    if (valid && has_cand) { v = ret; idx = c0 + n; }
    if (crank == 0 && a.returns && valid && has_cand) a.returns[(size_t)env * a.n_candidates + c0 + n] = ret;
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
  if (crank == 0) publish_and_reduce(a.red, env, group, red_v[0], red_i[0], tid, s_flag);
  if (ensemble) umma::cluster_sync_all();      // nobody exits while a peer may still read its shared memory
  if (warp == 5) {
    umma::tc_fence_after();
    umma::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace l2a
