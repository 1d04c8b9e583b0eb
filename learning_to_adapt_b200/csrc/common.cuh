// Shared device-side structures of libl2a_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/l2a_b200.h"

namespace l2a {

constexpr int kMaxLayers = L2A_MAX_LAYERS;

// Dense stack geometry + per-set fp32 storage layout (one contiguous block per weight set):
//   W_l at w_off[l] ([in_l, out_l] row-major, the reference's "kernel" layout), b_l at b_off[l].
struct MlpDims {
  int n_layers;                 // dense layers incl. output
  int dims[kMaxLayers + 1];     // dims[0] = D + A, dims[n_layers] = D
  int w_off[kMaxLayers];
  int b_off[kMaxLayers];
  int set_stride;               // floats per weight set
  int obs_dim, act_dim;
  int max_width;                // max over dims[]
};

struct NormDev {
  const float* obs_mean;
  const float* obs_den;         // std + 1e-10
  const float* act_mean;
  const float* act_den;
  const float* delta_mean;
  const float* delta_scale;     // std_delta + 1e-10
};

// np.argmax semantics: first maximum wins; a NaN beats every number (first NaN wins).
__device__ __forceinline__ bool better(float a, int ia, float b, int ib) {
  const bool na = (a != a), nb = (b != b);
  if (na != nb) return na;
  if (na) return ia < ib;
  return (a > b) || (a == b && ia < ib);
}

__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, i, off);
    if (better(ov, oi, v, i)) { v = ov; i = oi; }
  }
}

// Reward closed forms on (asq = sum_j a_j^2, d = denormalised delta, next = state + d); D = obs_dim.
//   half_cheetah: (next[D-3]-obs[D-3])/dt - 0.05*asq      envs/half_cheetah_env.py:58-65
//   ant         : (next[D-3]-obs[D-3])/dt + 0.05          envs/ant_env.py:56-66
//   arm_7dof    : -||next[D-3:]||_2 - 0.005*asq           envs/arm_7dof_env.py:91-99
// (next - obs) is the delta itself: the float64 reference loses nothing in that subtraction, so using d
// directly is the closest fp32 restatement.
__device__ __forceinline__ float reward_value(int kind, float inv_dt_num, float dt, float asq, float d_x,
                                              float n0, float n1, float n2) {
  (void)inv_dt_num;
  if (kind == L2A_REWARD_HALF_CHEETAH) return d_x / dt - 0.05f * asq;
  if (kind == L2A_REWARD_ANT) return d_x / dt + 0.05f;
  return -sqrtf(n0 * n0 + n1 * n1 + n2 * n2) - 0.005f * asq;
}

// Final per-env reduction shared by both rollout kernels: every CTA publishes its (best return, candidate index)
// and the last CTA of an env to arrive scans the partials in tile order (deterministic, lowest index wins ties).
struct ReduceArgs {
  float* part_ret;              // [m, tiles_per_env]
  int* part_idx;                // [m, tiles_per_env]
  unsigned int* counters;       // [m], zero on entry, reset by the last CTA
  float* best_ret;              // [m]
  int* best_idx;                // [m]
  float* best_act;              // [m, A]
  const float* actions;
  long long act_stride_row;
  int act_dim;
  int n_candidates;
  int tiles_per_env;
};

__device__ __forceinline__ void publish_and_reduce(const ReduceArgs& ra, int env, int tile, float v, int idx,
                                                    int tid, int* s_flag) {
  if (tid == 0) {
    ra.part_ret[env * ra.tiles_per_env + tile] = v;
    ra.part_idx[env * ra.tiles_per_env + tile] = idx;
    __threadfence();
    const unsigned int prev = atomicAdd(&ra.counters[env], 1u);
    *s_flag = (prev == (unsigned int)(ra.tiles_per_env - 1));
  }
  __syncthreads();
  if (*s_flag && tid < 32) {
    __threadfence();
    float bv = 0.f;
    int bi = 0x7fffffff;
    bool have = false;
    for (int t = tid; t < ra.tiles_per_env; t += 32) {
      const float pv = __ldcg(&ra.part_ret[env * ra.tiles_per_env + t]);
      const int pi = __ldcg(&ra.part_idx[env * ra.tiles_per_env + t]);
      if (pi < 0) continue;
      if (!have || better(pv, pi, bv, bi)) { bv = pv; bi = pi; have = true; }
    }
    if (!have) { bv = -__int_as_float(0x7f800000); bi = 0x7fffffff; }
    // lanes without a candidate carry (-inf, INT_MAX): any real candidate beats them
    warp_argmax(bv, bi);
    if (tid == 0) {
      ra.best_ret[env] = bv;
      ra.best_idx[env] = bi;
      ra.counters[env] = 0u;
    }
    const long long row = (long long)env * ra.n_candidates + bi;
    for (int j = tid; j < ra.act_dim; j += 32) ra.best_act[env * ra.act_dim + j] = ra.actions[row * ra.act_stride_row + j];
  }
}

}  // namespace l2a
