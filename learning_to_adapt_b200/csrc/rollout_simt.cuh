// K1 / K4, fp32 SIMT variant: fused H-step rollout (or a single predict step) for ANY MLP shape.
// One CTA owns a tile of RT candidate rows of one env and walks the whole horizon on chip:
//   normalise -> dense stack (fp32 FFMA, weights streamed from L2, activations in shared memory as [feature][row])
//   -> denormalise -> (ensemble mean) -> delta add -> reward -> discounted accumulate -> per-CTA argmax.
// Thread j of the CTA owns output feature(s) j, j+T, ... for all RT rows (RT accumulators in registers), so the
// weight reads are coalesced across the CTA and the activation reads are shared-memory broadcasts.
// Replaces policies/mpc_controller.py:116-129 + dynamics/mlp_dynamics.py:204-222 (+ meta_mlp_dynamics.py:296-306).
#pragma once
#include "common.cuh"

namespace l2a {

constexpr int kSimtRT = 32;        // rows per CTA
constexpr int kSimtThreads = 256;

struct SimtArgs {
  MlpDims dims;
  NormDev norm;
  const float* params;             // base of all weight sets
  const float* obs;                // rollout: [m, D] (one per env); predict: [n, D] (one per row)
  const float* actions;
  long long act_stride_t, act_stride_row;
  const float* discount_pow;       // [H]
  int rows_per_group;              // N (rollout) / rows per weight-set chunk (predict)
  int n_groups;                    // m
  int horizon;
  int set_mode, first_set, n_sets;
  int reward_kind;
  float dt;
  float* returns;                  // [m, N] or null
  float* delta_out;                // predict only
  float* next_out;                 // predict only
  ReduceArgs red;
};

// one dense layer for the CTA's RT rows: out[j][r] = act(b[j] + sum_k in[k][r] * W[k][j])
template <int RT>
__device__ __forceinline__ void simt_dense(const float* __restrict__ W, const float* __restrict__ b, int din, int dout,
                                           const float* in, float* out, float* scratch, bool relu, int tid, int nthreads) {
  // k-split for narrow layers so the whole CTA stays busy (deterministic reduction through `scratch`)
  int ksplit = 1;
  if (dout * 2 <= nthreads) {
    ksplit = nthreads / dout;
    if (ksplit > 16) ksplit = 16;
    if (ksplit > din) ksplit = din;
  }
  if (ksplit == 1) {
    for (int j = tid; j < dout; j += nthreads) {
      float acc[RT];
      const float bj = b[j];
#pragma unroll
      for (int r = 0; r < RT; ++r) acc[r] = bj;
      for (int k = 0; k < din; ++k) {
        const float w = __ldg(&W[(size_t)k * dout + j]);
        const float4* hv = reinterpret_cast<const float4*>(in + (size_t)k * RT);
#pragma unroll
        for (int q = 0; q < RT / 4; ++q) {
          const float4 h = hv[q];
          acc[4 * q + 0] = fmaf(h.x, w, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(h.y, w, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(h.z, w, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(h.w, w, acc[4 * q + 3]);
        }
      }
      float4* ov = reinterpret_cast<float4*>(out + (size_t)j * RT);
#pragma unroll
      for (int q = 0; q < RT / 4; ++q) {
        float4 o = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        ov[q] = o;
      }
    }
    __syncthreads();
    return;
  }
  const int kchunk = (din + ksplit - 1) / ksplit;
  if (tid < dout * ksplit) {
    const int j = tid % dout, ks = tid / dout;
    const int k0 = ks * kchunk, k1 = min(din, k0 + kchunk);
    float acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = 0.f;
    for (int k = k0; k < k1; ++k) {
      const float w = __ldg(&W[(size_t)k * dout + j]);
      const float4* hv = reinterpret_cast<const float4*>(in + (size_t)k * RT);
#pragma unroll
      for (int q = 0; q < RT / 4; ++q) {
        const float4 h = hv[q];
        acc[4 * q + 0] = fmaf(h.x, w, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(h.y, w, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(h.z, w, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(h.w, w, acc[4 * q + 3]);
      }
    }
    float4* pv = reinterpret_cast<float4*>(scratch + ((size_t)ks * dout + j) * RT);
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) pv[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
  }
  __syncthreads();
  for (int idx = tid; idx < dout * RT; idx += nthreads) {
    const int j = idx / RT, r = idx % RT;
    float s = b[j];
    for (int ks = 0; ks < ksplit; ++ks) s += scratch[((size_t)ks * dout + j) * RT + r];
    out[(size_t)j * RT + r] = relu ? fmaxf(s, 0.f) : s;
  }
  __syncthreads();
}

template <bool PREDICT>
__global__ void __launch_bounds__(kSimtThreads, 1) rollout_simt_kernel(const SimtArgs a) {
  constexpr int RT = kSimtRT;
  extern __shared__ __align__(16) float smem[];
  const MlpDims& md = a.dims;
  const int D = md.obs_dim, A = md.act_dim, W = md.max_width;
  float* xin = smem;                       // [D+A][RT]
  float* act0 = xin + (size_t)(D + A) * RT;  // [W][RT]
  float* act1 = act0 + (size_t)W * RT;     // [W][RT]
  float* state = act1 + (size_t)W * RT;    // [D][RT]
  float* dsum = state + (size_t)D * RT;    // [D][RT]
  float* asq = dsum + (size_t)D * RT;      // [RT]
  float* ret = asq + RT;                   // [RT]
  float* scratch = ret + RT;               // [kSimtThreads][RT] k-split partials of narrow layers
  __shared__ int s_flag;

  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int tiles_per_group = (a.rows_per_group + RT - 1) / RT;
  const int group = blockIdx.x / tiles_per_group;
  const int tile = blockIdx.x % tiles_per_group;
  const int c0 = tile * RT;                               // first candidate of the tile within its group
  const int nvalid = min(RT, a.rows_per_group - c0);
  const long long row0 = (long long)group * a.rows_per_group + c0;   // global row of local row 0

  for (int idx = tid; idx < D * RT; idx += nthreads) {
    const int k = idx / RT, r = idx % RT;
    float v = 0.f;
    if (r < nvalid) v = PREDICT ? a.obs[(row0 + r) * D + k] : a.obs[(long long)group * D + k];
    state[idx] = v;
  }
  if (tid < RT) ret[tid] = 0.f;
  __syncthreads();

  int set0 = a.first_set, nset = 1;
  if (a.set_mode == L2A_SETS_PER_ENV) set0 = a.first_set + group;
  if (a.set_mode == L2A_SETS_ENSEMBLE_MEAN) nset = a.n_sets;
  const float inv_nset = 1.0f / (float)nset;

  for (int t = 0; t < a.horizon; ++t) {
    const float* act_t = a.actions + (long long)t * a.act_stride_t;
    // normalised network input, mlp_dynamics.py:242-251
    for (int idx = tid; idx < (D + A) * RT; idx += nthreads) {
      const int k = idx / RT, r = idx % RT;
      float v = 0.f;
      if (r < nvalid) {
        if (k < D) v = (state[idx] - a.norm.obs_mean[k]) / a.norm.obs_den[k];
        else {
          const float av = act_t[(row0 + r) * a.act_stride_row + (k - D)];
          v = (av - a.norm.act_mean[k - D]) / a.norm.act_den[k - D];
        }
      }
      xin[idx] = v;
    }
    if (tid < RT) {
      float s = 0.f;
      if (tid < nvalid)
        for (int j = 0; j < A; ++j) { const float av = act_t[(row0 + tid) * a.act_stride_row + j]; s = fmaf(av, av, s); }
      asq[tid] = s;
    }
    for (int idx = tid; idx < D * RT; idx += nthreads) dsum[idx] = 0.f;
    __syncthreads();

    for (int e = 0; e < nset; ++e) {
      const float* P = a.params + (size_t)(set0 + e) * md.set_stride;
      const float* in = xin;
      float* out = act0;
      for (int l = 0; l < md.n_layers; ++l) {
        const int din = md.dims[l], dout = md.dims[l + 1];
        simt_dense<RT>(P + md.w_off[l], P + md.b_off[l], din, dout, in, out, scratch, l < md.n_layers - 1, tid, nthreads);
        in = out;
        out = (out == act0) ? act1 : act0;
      }
      // denormalise (mlp_dynamics.py:269-270) and accumulate over the sets
      for (int idx = tid; idx < D * RT; idx += nthreads) {
        const int k = idx / RT;
        dsum[idx] += in[idx] * a.norm.delta_scale[k] + a.norm.delta_mean[k];
      }
      __syncthreads();
    }

    if (PREDICT) {
      for (int idx = tid; idx < D * RT; idx += nthreads) {
        const int k = idx / RT, r = idx % RT;
        if (r < nvalid) {
          const float d = dsum[idx] * inv_nset;
          if (a.delta_out) a.delta_out[(row0 + r) * D + k] = d;
          if (a.next_out) a.next_out[(row0 + r) * D + k] = state[idx] + d;
        }
      }
    } else {
      if (nset > 1) {
        for (int idx = tid; idx < D * RT; idx += nthreads) dsum[idx] *= inv_nset;
        __syncthreads();
      }
      if (tid < RT) {
        const int r = tid;
        const float dx = dsum[(D - 3) * RT + r];
        const float n0 = state[(D - 3) * RT + r] + dx;
        const float n1 = state[(D - 2) * RT + r] + dsum[(D - 2) * RT + r];
        const float n2 = state[(D - 1) * RT + r] + dsum[(D - 1) * RT + r];
        const float rew = reward_value(a.reward_kind, 0.f, a.dt, asq[r], dx, n0, n1, n2);
        ret[r] = fmaf(a.discount_pow[t], rew, ret[r]);          // mpc_controller.py:126
      }
      __syncthreads();
      for (int idx = tid; idx < D * RT; idx += nthreads) state[idx] += dsum[idx];   // :127
      __syncthreads();
    }
  }

  if (!PREDICT) {
    if (a.returns && tid < nvalid) a.returns[(long long)group * a.rows_per_group + c0 + tid] = ret[tid];
    float v = -__int_as_float(0x7f800000);
    int idx = 0x7fffffff;
    if (tid < 32) {
      if (tid < nvalid) { v = ret[tid]; idx = c0 + tid; }
      warp_argmax(v, idx);
    }
    publish_and_reduce(a.red, group, tile, v, idx, tid, &s_flag);
  }
}

inline size_t simt_smem_bytes(const MlpDims& md) {
  const size_t RT = kSimtRT;
  return sizeof(float) * ((size_t)(md.obs_dim + md.act_dim) * RT + 2 * (size_t)md.max_width * RT +
                          2 * (size_t)md.obs_dim * RT + 2 * RT + (size_t)kSimtThreads * RT);
}

}  // namespace l2a
