// Device-resident adaptation window: the last M+1 (observation, action) pairs of every env's
// running path, kept as a ring in HBM so that GrBAL's per-step adapt windows
//   obs[-M-1:-1], act[-M-1:-1], obs[-M:]            (samplers/sampler.py:82-90)
// and their float64 normalisation (dynamics/meta_mlp_dynamics.py:334-339 -> mlp_dynamics.py:242-251, 265-267) are formed on
// the device, stream-ordered in front of K2, instead of being sliced out of Python lists, stacked and uploaded every env step.
// Arithmetic is IEEE float64 (sub, sub, div -- no contraction possible) rounded once to float32, i.e. bit-identical to the
// numpy expressions the reference feeds its float32 placeholders with.
#pragma once
#include <cstdint>

namespace l2a {

struct WindowDev {
  double* obs;        // [n_envs][cap][D]
  double* act;        // [n_envs][cap][A]
  int* count;         // [n_envs] pairs appended since the path started (not saturating)
  const double* norm; // obs_mean[D] obs_std[D] act_mean[A] act_std[A] delta_mean[D] delta_std[D]
  int n_envs, cap, D, A;
};

// running_paths[idx]["observations"].append(observation); ["actions"].append(action)   (sampler.py:109-110)
// act_stride: doubles between consecutive envs' action rows (A for a packed [n_envs, A] array)
__global__ void window_push_kernel(WindowDev w, const double* __restrict__ obs, const double* __restrict__ act, int act_stride) {
  const int env = blockIdx.x;
  const int slot = w.count[env] % w.cap;
  for (int c = threadIdx.x; c < w.D + w.A; c += blockDim.x) {
    if (c < w.D) w.obs[((size_t)env * w.cap + slot) * w.D + c] = obs[(size_t)env * w.D + c];
    else w.act[((size_t)env * w.cap + slot) * w.A + (c - w.D)] = act[(size_t)env * act_stride + (c - w.D)];
  }
  __syncthreads();
  if (threadIdx.x == 0) w.count[env] += 1;
}

// running_paths[idx] = _get_empty_running_paths_dict()   (sampler.py:128); env < 0: every env
__global__ void window_reset_kernel(WindowDev w, int env) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w.n_envs && (env < 0 || env == i)) w.count[i] = 0;
}

// x[k, j, :] = [ (obs_j - mu_o)/(sd_o + 1e-10) | (act_j - mu_a)/(sd_a + 1e-10) ],
// target[k, j, :] = ((obs_{j+1} - obs_j) - mu_d)/(sd_d + 1e-10)   for the M pairs j = L-M-1 .. L-2 of env k's path (length L).
__global__ void window_gather_kernel(WindowDev w, int M, float* __restrict__ x, float* __restrict__ target) {
  const int env = blockIdx.x;
  const int L = w.count[env];
  const int D = w.D, A = w.A;
  const double* mu_o = w.norm;
  const double* sd_o = w.norm + D;
  const double* mu_a = w.norm + 2 * D;
  const double* sd_a = w.norm + 2 * D + A;
  const double* mu_d = w.norm + 2 * D + 2 * A;
  const double* sd_d = w.norm + 3 * D + 2 * A;
  const int per_row = 2 * D + A;
  for (int i = threadIdx.x; i < M * per_row; i += blockDim.x) {
    const int j = i / per_row, c = i % per_row;
    const int slot = (L - M - 1 + j) % w.cap;
    const double* o = w.obs + ((size_t)env * w.cap + slot) * D;
    if (c < D) {
      x[((size_t)env * M + j) * (D + A) + c] = (float)__ddiv_rn(__dsub_rn(o[c], mu_o[c]), __dadd_rn(sd_o[c], 1e-10));
    } else if (c < D + A) {
      const int a = c - D;
      const double v = w.act[((size_t)env * w.cap + slot) * A + a];
      x[((size_t)env * M + j) * (D + A) + c] = (float)__ddiv_rn(__dsub_rn(v, mu_a[a]), __dadd_rn(sd_a[a], 1e-10));
    } else {
      const int d = c - D - A;
      const double* on = w.obs + ((size_t)env * w.cap + (slot + 1) % w.cap) * D;
      const double delta = __dsub_rn(on[d], o[d]);
      target[((size_t)env * M + j) * D + d] = (float)__ddiv_rn(__dsub_rn(delta, mu_d[d]), __dadd_rn(sd_d[d], 1e-10));
    }
  }
}

}  // namespace l2a
