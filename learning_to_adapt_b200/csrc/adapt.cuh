// K2: GrBAL one-step inner adaptation (dynamics/meta_mlp_dynamics.py:321-345, _adapt_sym 409-421, graph 96-120).
//   theta'_k = theta - lr * d/dtheta mean_{M x D}((target_k - f_theta(x_k))^2)        for each task k < K.
// Two phases:
//   adapt_fwd_bwd_kernel : one CTA per task; forward over the M context rows keeping every activation, then the
//                          backward chain of layer-output gradients g_l.  Tiny (M <= 32 rows), latency-bound.
//   adapt_update_kernel  : grid over (task, layer, tiles); theta'[i][j] = theta[i][j] - lr * sum_r h_l[r][i] g_l[r][j].
//                          Pure streaming: reads theta once, writes theta' once -> HBM-bound (4 B in + 4 B out / param).
#pragma once
#include "common.cuh"

namespace l2a {

constexpr int kAdaptMaxM = 32;
constexpr int kAdaptThreads = 256;

struct AdaptArgs {
  MlpDims dims;
  const float* params;      // all sets
  int src_set, dst_first_set;
  const float* x;           // [K, M, D+A] normalised
  const float* target;      // [K, M, D] normalised delta
  int K, M;
  float lr;
  float* acts;              // workspace [K][sum_l dims[l]][M]   (layer inputs h_l, feature-major)
  float* grads;             // workspace [K][sum_l dims[l+1]][M] (layer-output gradients g_l)
  int act_off[kMaxLayers + 1];
  int grad_off[kMaxLayers + 1];
  float* params_out;        // == params (sets dst_first_set + k)
};

// out[j][r] = act(b[j] + sum_i in[i][r] W[i][j]) for r < M (feature-major activations in global/L2)
__global__ void __launch_bounds__(kAdaptThreads, 1) adapt_fwd_bwd_kernel(const AdaptArgs a) {
  const MlpDims& md = a.dims;
  const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, M = a.M;
  const float* P = a.params + (size_t)a.src_set * md.set_stride;
  float* A0 = a.acts + (size_t)k * (size_t)a.act_off[md.n_layers];
  float* G0 = a.grads + (size_t)k * (size_t)a.grad_off[md.n_layers];
  const int din0 = md.dims[0];
  // h_0 = x (transpose to feature-major)
  for (int idx = tid; idx < din0 * M; idx += nt) {
    const int i = idx / M, r = idx % M;
    A0[a.act_off[0] + idx] = a.x[((size_t)k * M + r) * din0 + i];
  }
  __syncthreads();
  // forward: keep every layer input; the network output goes to the slot after the last input
  for (int l = 0; l < md.n_layers; ++l) {
    const int din = md.dims[l], dout = md.dims[l + 1];
    const float* W = P + md.w_off[l];
    const float* b = P + md.b_off[l];
    const float* in = A0 + a.act_off[l];
    float* out = (l + 1 < md.n_layers) ? (A0 + a.act_off[l + 1]) : (G0 + a.grad_off[l]);   // y lands in g_L slot
    for (int j = tid; j < dout; j += nt) {
      float acc[kAdaptMaxM];
#pragma unroll
      for (int r = 0; r < kAdaptMaxM; ++r) acc[r] = b[j];
      for (int i = 0; i < din; ++i) {
        const float w = __ldg(&W[(size_t)i * dout + j]);
#pragma unroll
        for (int r = 0; r < kAdaptMaxM; ++r)
          if (r < M) acc[r] = fmaf(in[(size_t)i * M + r], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < kAdaptMaxM; ++r)
        if (r < M) out[(size_t)j * M + r] = (l + 1 < md.n_layers) ? fmaxf(acc[r], 0.f) : acc[r];
    }
    __syncthreads();
  }
  // dL/dy = 2/(M*D) * (y - target)                                  (meta_mlp_dynamics.py:118)
  {
    const int L = md.n_layers - 1, dout = md.dims[md.n_layers];
    float* g = G0 + a.grad_off[L];
    const float scale = 2.0f / (float)(M * dout);
    for (int idx = tid; idx < dout * M; idx += nt) {
      const int j = idx / M, r = idx % M;
      g[idx] = scale * (g[idx] - a.target[((size_t)k * M + r) * dout + j]);
    }
    __syncthreads();
  }
  // backward chain: g_{l-1}[i][r] = relu'(h_l[i][r]) * sum_j g_l[j][r] W_l[i][j]
  for (int l = md.n_layers - 1; l >= 1; --l) {
    const int din = md.dims[l], dout = md.dims[l + 1];
    const float* W = P + md.w_off[l];
    const float* g = G0 + a.grad_off[l];
    const float* h = A0 + a.act_off[l];
    float* gp = G0 + a.grad_off[l - 1];
    for (int i = tid; i < din; i += nt) {
      float acc[kAdaptMaxM];
#pragma unroll
      for (int r = 0; r < kAdaptMaxM; ++r) acc[r] = 0.f;
      const float* wrow = W + (size_t)i * dout;
      for (int j = 0; j < dout; ++j) {
        const float w = __ldg(&wrow[j]);
#pragma unroll
        for (int r = 0; r < kAdaptMaxM; ++r)
          if (r < M) acc[r] = fmaf(g[(size_t)j * M + r], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < kAdaptMaxM; ++r)
        if (r < M) gp[(size_t)i * M + r] = (h[(size_t)i * M + r] > 0.f) ? acc[r] : 0.f;
    }
    __syncthreads();
  }
}

// grid.x = tiles over the flattened [in_l * out_l] (+ out_l bias) of one layer, grid.y = layer, grid.z = task
__global__ void __launch_bounds__(256) adapt_update_kernel(const AdaptArgs a) {
  const MlpDims& md = a.dims;
  const int l = blockIdx.y, k = blockIdx.z, M = a.M;
  const int din = md.dims[l], dout = md.dims[l + 1];
  const float* P = a.params + (size_t)a.src_set * md.set_stride;
  float* Q = a.params_out + (size_t)(a.dst_first_set + k) * md.set_stride;
  const float* h = a.acts + (size_t)k * (size_t)a.act_off[md.n_layers] + a.act_off[l];
  const float* g = a.grads + (size_t)k * (size_t)a.grad_off[md.n_layers] + a.grad_off[l];
  const int total = din * dout + dout;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    if (idx < din * dout) {
      const int i = idx / dout, j = idx % dout;
      float s = 0.f;
      for (int r = 0; r < M; ++r) s = fmaf(h[(size_t)i * M + r], g[(size_t)j * M + r], s);
      Q[md.w_off[l] + idx] = P[md.w_off[l] + idx] - a.lr * s;                 // _adapt_sym :416-417
    } else {
      const int j = idx - din * dout;
      float s = 0.f;
      for (int r = 0; r < M; ++r) s += g[(size_t)j * M + r];
      Q[md.b_off[l] + j] = P[md.b_off[l] + j] - a.lr * s;
    }
  }
}

}  // namespace l2a
