// K2: GrBAL one-step inner adaptation (dynamics/meta_mlp_dynamics.py:321-345, _adapt_sym 409-421, graph 96-120).
//   theta'_k = theta - lr * d/dtheta mean_{M x D}((target_k - f_theta(x_k))^2)        for each task k < K.
// Two kernels:
//   adapt_fwd_bwd_kernel : a thread-block CLUSTER of 16 (K <= 8 tasks) or 8 CTAs per task walks the 2L-1 dependent stages (forward
//                          layers, loss gradient, backward chain).  Every stage is a skinny product over the M <= 32 context
//                          rows: each CTA computes a slice of the stage's output features from the full input (32 KB, re-read
//                          from L2 into shared memory), publishes it to the global workspace and the cluster barrier
//                          (release / acquire) hands it to the next stage.  Latency-bound: ~3 us per stage.
//   adapt_update_kernel  : grid over (task, layer, tiles); theta'[i][j] = theta[i][j] - lr * sum_r h_l[i][r] g_l[r][j].
//                          Pure streaming: reads theta once, writes theta' once -> HBM-bound (4 B in + 4 B out / param).
// Workspace layouts: layer inputs h_l as [feature][M] (broadcast reads in the forward product), layer-output gradients g_l as
// [M][feature] (coalesced reads in the backward product and in the update).
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace l2a {

constexpr int kAdaptMaxM = 32;
constexpr int kAdaptThreads = 256;
constexpr int kAdaptCluster = 8;          // CTAs per task: 16 (the non-portable maximum, one cluster per GPC) while all tasks' clusters are
constexpr int kAdaptClusterMax = 16;      // resident at once (K <= 8 on B200's 8 GPCs), else 8

struct AdaptArgs {
  MlpDims dims;
  const float* params;      // all sets
  int src_set, dst_first_set;
  const float* x;           // [K, M, D+A] normalised
  const float* target;      // [K, M, D] normalised delta
  int K, M;
  float lr;
  float* acts;              // workspace [K][sum_l dims[l]][M]      layer inputs h_l, [feature][M]
  float* grads;             // workspace [K][M][sum_l dims[l+1]]    layer-output gradients g_l, each [M][dims[l+1]]
  int act_off[kMaxLayers + 1];    // float offset of h_l inside a task's block (act_off[n_layers] = block size)
  int grad_off[kMaxLayers + 1];   // float offset of g_l inside a task's block
  float* params_out;        // == params (sets dst_first_set + k)
  int csize;                // cluster size of adapt_fwd_bwd_kernel (CTAs per task)
};

template <int MR>
__global__ void __launch_bounds__(kAdaptThreads, 1) adapt_fwd_bwd_kernel(const AdaptArgs a) {
  extern __shared__ __align__(16) float adapt_smem[];
  const MlpDims& md = a.dims;
  const int M = a.M, W = md.max_width;
  float* s_in = adapt_smem;                       // stage input: h_l as [feature][MR] or g_l as [MR][feature]
  float* s_part = s_in + (size_t)W * MR;          // k-split partial sums [4 * kAdaptThreads][MR]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int crank = (int)umma::cluster_ctarank();
  const int k = blockIdx.x / a.csize;             // task
  const float* P = a.params + (size_t)a.src_set * md.set_stride;
  float* A0 = a.acts + (size_t)k * (size_t)a.act_off[md.n_layers];
  float* G0 = a.grads + (size_t)k * (size_t)a.grad_off[md.n_layers];
  const int L = md.n_layers;

  // h_0 = x, transposed to [feature][M] (every CTA writes the same values of its own slice; rank 0 publishes)
  if (crank == 0) {
    const int din0 = md.dims[0];
    for (int idx = tid; idx < din0 * M; idx += kAdaptThreads) {
      const int i = idx / M, r = idx % M;
      A0[a.act_off[0] + idx] = a.x[((size_t)k * M + r) * din0 + i];
    }
  }
  umma::cluster_sync_all();

  // ------------------------------------------------------------------ forward
  for (int l = 0; l < L; ++l) {
    const int din = md.dims[l], dout = md.dims[l + 1];
    const float* Wl = P + md.w_off[l];
    const float* bl = P + md.b_off[l];
    const float* h_in = A0 + a.act_off[l];
    if (M == MR) {                                 // contiguous [din][M] block: 16-byte copies, all loads in flight
      const int n4 = din * M / 4;
      const float4* src = reinterpret_cast<const float4*>(h_in);
      float4* dst = reinterpret_cast<float4*>(s_in);
#pragma unroll 8
      for (int idx = tid; idx < n4; idx += kAdaptThreads) dst[idx] = __ldcg(src + idx);
    } else {
      for (int idx = tid; idx < din * M; idx += kAdaptThreads) {
        const int i = idx / M, r = idx % M;
        s_in[i * MR + r] = __ldcg(h_in + idx);
      }
    }
    __syncthreads();
    const int fs = (dout + a.csize - 1) / a.csize;                      // output features per CTA
    const int j0 = crank * fs, nf = max(0, min(fs, dout - j0));
    // thread = (group of 4 adjacent output features, k-slice): one 16-byte weight load feeds 4 x MR FMAs, 8 loads in flight
    const bool vec4 = (dout % 4 == 0) && (fs % 4 == 0);
    const int fw = vec4 ? 4 : 1;
    const int ngrp = (nf + fw - 1) / fw;
    int ksplit = ngrp > 0 ? kAdaptThreads / ngrp : 1;
    if (ksplit > 32) ksplit = 32;
    if (ksplit > din) ksplit = din;
    if (ksplit < 1) ksplit = 1;
    const int kchunk = (din + ksplit - 1) / ksplit;
    if (tid < ngrp * ksplit) {
      const int fg = tid % ngrp, ks = tid / ngrp, j = j0 + fg * fw;
      const int i0 = ks * kchunk, i1 = min(din, i0 + kchunk);
      if (vec4) {
        float acc[4][MR];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int r = 0; r < MR; ++r) acc[c][r] = 0.f;
#pragma unroll 8
        for (int i = i0; i < i1; ++i) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(&Wl[(size_t)i * dout + j]));
          const float4* hv = reinterpret_cast<const float4*>(s_in + (size_t)i * MR);
#pragma unroll
          for (int q = 0; q < MR / 4; ++q) {
            const float4 h = hv[q];
            const float hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              acc[0][4 * q + u] = fmaf(hh[u], w.x, acc[0][4 * q + u]);
              acc[1][4 * q + u] = fmaf(hh[u], w.y, acc[1][4 * q + u]);
              acc[2][4 * q + u] = fmaf(hh[u], w.z, acc[2][4 * q + u]);
              acc[3][4 * q + u] = fmaf(hh[u], w.w, acc[3][4 * q + u]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int r = 0; r < MR; ++r) s_part[((size_t)ks * nf + fg * 4 + c) * MR + r] = acc[c][r];
      } else {
        float acc[MR];
#pragma unroll
        for (int r = 0; r < MR; ++r) acc[r] = 0.f;
#pragma unroll 8
        for (int i = i0; i < i1; ++i) {
          const float w = __ldg(&Wl[(size_t)i * dout + j]);
#pragma unroll
          for (int r = 0; r < MR; ++r) acc[r] = fmaf(s_in[(size_t)i * MR + r], w, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < MR; ++r) s_part[((size_t)ks * nf + fg) * MR + r] = acc[r];
      }
    }
    __syncthreads();
    for (int idx = tid; idx < nf * M; idx += kAdaptThreads) {
      const int f = idx / M, r = idx % M, j = j0 + f;
      float s = __ldg(&bl[j]);
      for (int ks = 0; ks < ksplit; ++ks) s += s_part[((size_t)ks * nf + f) * MR + r];        // fixed order: deterministic
      if (l + 1 < L) {
        A0[a.act_off[l + 1] + j * M + r] = fmaxf(s, 0.f);                                     // ReLU dense
      } else {
        // dL/dy = 2 / (M * D) * (y - target)                                  (meta_mlp_dynamics.py:118)
        G0[a.grad_off[l] + r * dout + j] = (2.0f / (float)(M * dout)) * (s - a.target[((size_t)k * M + r) * dout + j]);
      }
    }
    umma::cluster_sync_all();
  }

  // ------------------------------------------------------------------ backward chain
  //   g_{l-1}[r][i] = relu'(h_l[i][r]) * sum_j g_l[r][j] W_l[i][j]       (one warp per input feature i: W row read coalesced)
  for (int l = L - 1; l >= 1; --l) {
    const int din = md.dims[l], dout = md.dims[l + 1];
    const float* Wl = P + md.w_off[l];
    const float* g_in = G0 + a.grad_off[l];
    const float* h_l = A0 + a.act_off[l];
    float* g_out = G0 + a.grad_off[l - 1];
    if (dout % 4 == 0 && W % 4 == 0) {
      const int d4 = dout / 4;
#pragma unroll 8
      for (int idx = tid; idx < M * d4; idx += kAdaptThreads) {
        const int r = idx / d4, q = idx % d4;
        reinterpret_cast<float4*>(s_in + (size_t)r * W)[q] = __ldcg(reinterpret_cast<const float4*>(g_in + (size_t)r * dout) + q);
      }
    } else {
      for (int idx = tid; idx < M * dout; idx += kAdaptThreads) {
        const int r = idx / dout, j = idx % dout;
        s_in[r * W + j] = __ldcg(g_in + idx);
      }
    }
    __syncthreads();
    const int is = (din + a.csize - 1) / a.csize;
    const int i0 = crank * is, ni = max(0, min(is, din - i0));
    for (int ii = warp; ii < ni; ii += kAdaptThreads / 32) {
      const int i = i0 + ii;
      const float* wrow = Wl + (size_t)i * dout;
      float acc[MR];
#pragma unroll
      for (int r = 0; r < MR; ++r) acc[r] = 0.f;
#pragma unroll 8
      for (int j = lane; j < dout; j += 32) {
        const float w = __ldg(&wrow[j]);
#pragma unroll
        for (int r = 0; r < MR; ++r) acc[r] = fmaf(s_in[r * W + j], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < MR; ++r) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], off);
      }
      float mine = 0.f;
#pragma unroll
      for (int r = 0; r < MR; ++r) mine = (lane == r) ? acc[r] : mine;
      if (lane < M) g_out[lane * din + i] = (__ldcg(&h_l[i * M + lane]) > 0.f) ? mine : 0.f;
    }
    umma::cluster_sync_all();
  }
}

// theta'[i][j] = theta[i][j] - lr * sum_r h[i][r] g[r][j]   (+ the bias row: b'[j] = b[j] - lr * sum_r g[r][j])      _adapt_sym :416-417
// grid.x = tiles of kUpdTI input rows x kUpdTJ output columns (row tiles fastest; one extra row tile carries the bias), grid.y = layer,
// grid.z = task.  A CTA stages its h rows ([r][i], transposed) and g columns in shared memory once; a thread owns 4 rows x 4
// adjacent columns: per context row r one 16-byte g load + one 16-byte h load feed 16 FMAs, theta is read and theta' written as
// 16-byte vectors (HBM / L2 streaming: 4 B in + 4 B out per parameter).  Sum over r in ascending order, as the first version did.
constexpr int kUpdTI = 32, kUpdTJ = 128;
__global__ void __launch_bounds__(256) adapt_update_kernel(const AdaptArgs a) {
  __shared__ __align__(16) float h_s[kAdaptMaxM][kUpdTI];
  __shared__ __align__(16) float g_s[kAdaptMaxM][kUpdTJ];
  const MlpDims& md = a.dims;
  const int l = blockIdx.y, k = blockIdx.z, M = a.M;
  const int din = md.dims[l], dout = md.dims[l + 1];
  const int tiles_j = (dout + kUpdTJ - 1) / kUpdTJ, tiles_i = (din + kUpdTI - 1) / kUpdTI;
  if ((int)blockIdx.x >= tiles_j * (tiles_i + 1)) return;
  const int tj = blockIdx.x / (tiles_i + 1), ti = blockIdx.x % (tiles_i + 1);
  const int i0 = ti * kUpdTI, j0 = tj * kUpdTJ;
  const float* P = a.params + (size_t)a.src_set * md.set_stride;
  float* Q = a.params_out + (size_t)(a.dst_first_set + k) * md.set_stride;
  const float* h = a.acts + (size_t)k * (size_t)a.act_off[md.n_layers] + a.act_off[l];       // [din][M]
  const float* g = a.grads + (size_t)k * (size_t)a.grad_off[md.n_layers] + a.grad_off[l];    // [M][dout]
  const int tid = threadIdx.x;
  for (int idx = tid; idx < M * kUpdTJ; idx += 256) {
    const int r = idx / kUpdTJ, j = idx % kUpdTJ;
    g_s[r][j] = (j0 + j < dout) ? __ldcg(&g[(size_t)r * dout + j0 + j]) : 0.f;
  }
  if (ti < tiles_i)
    for (int idx = tid; idx < kUpdTI * M; idx += 256) {
      const int i = idx / M, r = idx % M;
      h_s[r][i] = (i0 + i < din) ? __ldcg(&h[(size_t)(i0 + i) * M + r]) : 0.f;
    }
  __syncthreads();
  const int jq = (tid & 31) * 4, iq = (tid >> 5) * 4;          // 32 column groups x 8 row groups
  if (ti == tiles_i) {                                         // the bias row of this column tile
    if (tid < kUpdTJ && j0 + tid < dout) {
      float s = 0.f;
      for (int r = 0; r < M; ++r) s += g_s[r][tid];
      Q[md.b_off[l] + j0 + tid] = P[md.b_off[l] + j0 + tid] - a.lr * s;
    }
    return;
  }
  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
  for (int r = 0; r < M; ++r) {
    const float4 gv = *reinterpret_cast<const float4*>(&g_s[r][jq]);
    const float4 hv = *reinterpret_cast<const float4*>(&h_s[r][iq]);
    const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc[u][0] = fmaf(hh[u], gv.x, acc[u][0]);
      acc[u][1] = fmaf(hh[u], gv.y, acc[u][1]);
      acc[u][2] = fmaf(hh[u], gv.z, acc[u][2]);
      acc[u][3] = fmaf(hh[u], gv.w, acc[u][3]);
    }
  }
  const bool vec = (dout % 4 == 0) && (j0 + jq + 3 < dout) && (md.w_off[l] % 4 == 0);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + iq + u;
    if (i >= din) continue;
    const size_t off = (size_t)md.w_off[l] + (size_t)i * dout + j0 + jq;
    if (vec) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(P + off));
      *reinterpret_cast<float4*>(Q + off) = make_float4(p.x - a.lr * acc[u][0], p.y - a.lr * acc[u][1], p.z - a.lr * acc[u][2], p.w - a.lr * acc[u][3]);
    } else {
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (j0 + jq + v < dout) Q[off + v] = P[off + v] - a.lr * acc[u][v];
    }
  }
}

inline size_t adapt_smem_bytes(const MlpDims& md, int mr) {
  return sizeof(float) * ((size_t)md.max_width * mr + (size_t)4 * kAdaptThreads * mr);
}

}  // namespace l2a
