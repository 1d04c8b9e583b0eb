// ReBAL: fused H-step rollout through a single-layer LSTM dynamics model, fp32 SIMT.
// Replaces policies/rnn_mpc_controller.py:112-134 (+ repeat_hidden :165-187) and dynamics/rnn_dynamics.py:233-252; the cell is
// TF 1.13 tf.nn.rnn_cell.LSTMCell (dynamics/core/utils.py:193-198):
//   z = [x, h] @ kernel + bias ; i, j, f, o = split(z, 4) ; c' = sigmoid(f + 1) * c + sigmoid(i) * tanh(j) ; h' = sigmoid(o) * tanh(c')
// followed by the dense output layer (core/utils.py:227-234).  One CTA owns RT candidate rows and keeps observation, cell
// and hidden state in shared memory for the whole horizon; the two dense products reuse simt_dense (rollout_simt.cuh).
#pragma once
#include "common.cuh"
#include "rollout_simt.cuh"

namespace l2a {

constexpr int kRnnRT = 16;
constexpr int kRnnThreads = 256;

struct RnnDims {
  int obs_dim, act_dim, hidden;
  int wk_off, bk_off, wo_off, bo_off;     // float offsets inside the parameter block
  int total;
};

struct RnnArgs {
  RnnDims dims;
  NormDev norm;
  const float* params;
  const float* obs;                // rollout: [m, D]; predict: [n, D]
  const float* c0;                 // rollout: [m, Hs]; predict: [n, Hs]
  const float* h0;
  const float* actions;
  long long act_stride_t, act_stride_row;
  const float* discount_pow;
  int rows_per_group, n_groups, horizon;
  int reward_kind;
  float dt;
  float* returns;
  float* delta_out;                // predict only
  float* c_out;                    // predict only [n, Hs]
  float* h_out;
  ReduceArgs red;
};

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <bool PREDICT>
__global__ void __launch_bounds__(kRnnThreads, 1) rollout_rnn_simt_kernel(const RnnArgs a) {
  constexpr int RT = kRnnRT;
  extern __shared__ __align__(16) float rnn_smem[];
  const RnnDims& rd = a.dims;
  const int D = rd.obs_dim, A = rd.act_dim, Hs = rd.hidden, IN = D + A;
  float* xin = rnn_smem;                          // [(IN + Hs)][RT]: normalised (obs, act) then h_prev
  float* gates = xin + (size_t)(IN + Hs) * RT;    // [4 Hs][RT]
  float* cbuf = gates + (size_t)4 * Hs * RT;      // [Hs][RT]
  float* state = cbuf + (size_t)Hs * RT;          // [D][RT]
  float* ybuf = state + (size_t)D * RT;           // [D][RT]
  float* asq = ybuf + (size_t)D * RT;             // [RT]
  float* ret = asq + RT;                          // [RT]
  float* scratch = ret + RT;                      // [kRnnThreads][RT]
  __shared__ int s_flag;

  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int tiles_per_group = (a.rows_per_group + RT - 1) / RT;
  const int group = blockIdx.x / tiles_per_group, tile = blockIdx.x % tiles_per_group;
  const int c0i = tile * RT;
  const int nvalid = min(RT, a.rows_per_group - c0i);
  const long long row0 = (long long)group * a.rows_per_group + c0i;
  const float* P = a.params;

  for (int idx = tid; idx < D * RT; idx += nthreads) {
    const int k = idx / RT, r = idx % RT;
    state[idx] = (r < nvalid) ? (PREDICT ? a.obs[(row0 + r) * D + k] : a.obs[(long long)group * D + k]) : 0.f;
  }
  for (int idx = tid; idx < Hs * RT; idx += nthreads) {
    const int u = idx / RT, r = idx % RT;
    const long long src = PREDICT ? (row0 + r) : (long long)group;       // repeat_hidden: every candidate of env g starts from its state
    cbuf[idx] = (r < nvalid) ? a.c0[src * Hs + u] : 0.f;
    xin[(size_t)(IN + u) * RT + r] = (r < nvalid) ? a.h0[src * Hs + u] : 0.f;
  }
  if (tid < RT) ret[tid] = 0.f;
  __syncthreads();

  for (int t = 0; t < a.horizon; ++t) {
    const float* act_t = a.actions + (long long)t * a.act_stride_t;
    for (int idx = tid; idx < IN * RT; idx += nthreads) {
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
    __syncthreads();
    // z = [x, h] @ kernel + bias
    simt_dense<RT>(P + rd.wk_off, P + rd.bk_off, IN + Hs, 4 * Hs, xin, gates, scratch, false, tid, nthreads);
    // LSTMCell update (forget_bias = 1.0)
    for (int idx = tid; idx < Hs * RT; idx += nthreads) {
      const float zi = gates[idx], zj = gates[(size_t)Hs * RT + idx], zf = gates[(size_t)2 * Hs * RT + idx], zo = gates[(size_t)3 * Hs * RT + idx];
      const float c_new = sigmoid_acc(zf + 1.0f) * cbuf[idx] + sigmoid_acc(zi) * tanhf(zj);
      cbuf[idx] = c_new;
      xin[(size_t)IN * RT + idx] = sigmoid_acc(zo) * tanhf(c_new);
    }
    __syncthreads();
    // y = h' @ W_out + b_out
    simt_dense<RT>(P + rd.wo_off, P + rd.bo_off, Hs, D, xin + (size_t)IN * RT, ybuf, scratch, false, tid, nthreads);
    for (int idx = tid; idx < D * RT; idx += nthreads) {
      const int k = idx / RT;
      ybuf[idx] = ybuf[idx] * a.norm.delta_scale[k] + a.norm.delta_mean[k];          // rnn_dynamics.py:244
    }
    __syncthreads();
    if (PREDICT) {
      for (int idx = tid; idx < D * RT; idx += nthreads) {
        const int k = idx / RT, r = idx % RT;
        if (r < nvalid) a.delta_out[(row0 + r) * D + k] = ybuf[idx];
      }
      for (int idx = tid; idx < Hs * RT; idx += nthreads) {
        const int u = idx / RT, r = idx % RT;
        if (r < nvalid) {
          a.c_out[(row0 + r) * Hs + u] = cbuf[idx];
          a.h_out[(row0 + r) * Hs + u] = xin[(size_t)IN * RT + idx];
        }
      }
    } else {
      if (tid < RT) {
        const int r = tid;
        const float dx = ybuf[(D - 3) * RT + r];
        const float n0 = state[(D - 3) * RT + r] + dx;
        const float n1 = state[(D - 2) * RT + r] + ybuf[(D - 2) * RT + r];
        const float n2 = state[(D - 1) * RT + r] + ybuf[(D - 1) * RT + r];
        ret[r] = fmaf(a.discount_pow[t], reward_value(a.reward_kind, 0.f, a.dt, asq[r], dx, n0, n1, n2), ret[r]);
      }
      __syncthreads();
      for (int idx = tid; idx < D * RT; idx += nthreads) state[idx] += ybuf[idx];
      __syncthreads();
    }
  }
  if (!PREDICT) {
    if (a.returns && tid < nvalid) a.returns[(long long)group * a.rows_per_group + c0i + tid] = ret[tid];
    float v = -__int_as_float(0x7f800000);
    int idx = 0x7fffffff;
    if (tid < 32) {
      if (tid < nvalid) { v = ret[tid]; idx = c0i + tid; }
      warp_argmax(v, idx);
    }
    publish_and_reduce(a.red, group, tile, v, idx, tid, &s_flag);
  }
}

inline size_t rnn_smem_bytes(const RnnDims& rd) {
  const size_t RT = kRnnRT;
  return sizeof(float) * ((size_t)(rd.obs_dim + rd.act_dim + rd.hidden) * RT + (size_t)4 * rd.hidden * RT + (size_t)rd.hidden * RT +
                          2 * (size_t)rd.obs_dim * RT + 2 * RT + (size_t)kRnnThreads * RT);
}

}  // namespace l2a
