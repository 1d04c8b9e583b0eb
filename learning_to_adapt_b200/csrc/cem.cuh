// K1c: CEM sampling and elite refit around the rollout kernel (policies/mpc_controller.py:84-104).
// The planner state (mean, std) is float64 like the reference's numpy arrays; samples are rounded to fp32 for the
// rollout (the TF feed does the same cast).
#pragma once
#include "common.cuh"

namespace l2a {

// a = mean + z * std (:86); clipped copy (:87).  Layout [n, m, H*A] exactly as numpy's (n, m, hA) array, so that the
// reference's reshape to (n*m, H, A) (:88) is the same buffer with act_stride_row = H*A, act_stride_t = A.
__global__ void cem_sample_kernel(const float* __restrict__ z, const double* __restrict__ mean, const double* __restrict__ std_,
                                  const float* __restrict__ clip_low, const float* __restrict__ clip_high, int n, int m, int ha,
                                  float* __restrict__ samples, float* __restrict__ clipped) {
  const long long total = (long long)n * m * ha;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % ha);
    const int e = (int)((i / ha) % m);
    const double a = mean[e * ha + j] + (double)z[i] * std_[e * ha + j];
    samples[i] = (float)a;
    clipped[i] = (float)fmin(fmax(a, (double)clip_low[j]), (double)clip_high[j]);
  }
}

// The same with float64 draws (the host-buffer planning call keeps numpy's float64 normals): additionally writes the float64
// first action of every row, first64[row][A] with row = flat index / (H*A) of the (n*m, H, A) view -- the value the reference
// returns for the winning row (mpc_controller.py:94, 106: cand_a from the UNclipped float64 samples).
__global__ void cem_sample64_kernel(const double* __restrict__ z, const double* __restrict__ mean, const double* __restrict__ std_,
                                    const float* __restrict__ clip_low, const float* __restrict__ clip_high, int n, int m, int ha, int A,
                                    float* __restrict__ samples, double* __restrict__ clipped, double* __restrict__ first64) {
  const long long total = (long long)n * m * ha;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % ha);
    const int e = (int)((i / ha) % m);
    const double a = __dadd_rn(mean[e * ha + j], __dmul_rn(z[i], std_[e * ha + j]));       // mean + z * std (:86), no FMA
    samples[i] = (float)a;
    clipped[i] = fmin(fmax(a, (double)clip_low[j]), (double)clip_high[j]);      // np.clip of the float64 samples (:87)
    if (j < A) first64[(i / ha) * A + j] = a;
  }
}

// mean = 0, std = 1 (:79-80) at the start of every planning call
__global__ void cem_init_kernel(double* __restrict__ mean, double* __restrict__ std_, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) { mean[i] = 0.0; std_[i] = 1.0; }
}

// rank[e][c] = position of candidate c in the descending-return order of env e (= np.argsort(-returns) inverse).
// Ties: the lower index ranks first.  Counting rank: candidate c is preceded by every j with a larger return (or an equal one
// and a lower index).  grid = (ceil(n/256), m, kRankSplit): the j range is split over blockIdx.z, each part counted against a
// shared-memory tile of the returns and added atomically -- `rank` must be ZERO on entry (the launcher clears it).
constexpr int kRankSplit = 4;
constexpr int kRankTile = 1024;
__global__ void __launch_bounds__(256) cem_rank_kernel(const float* __restrict__ returns, int n, int* __restrict__ rank) {
  __shared__ float tile[kRankTile];
  const int e = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const float* r = returns + (size_t)e * n;
  const float rc = (c < n) ? r[c] : 0.f;
  const int per = (n + kRankSplit - 1) / kRankSplit;
  const int j_begin = blockIdx.z * per, j_end = min(n, j_begin + per);
  int cnt = 0;
  for (int j0 = j_begin; j0 < j_end; j0 += kRankTile) {
    const int lim = min(kRankTile, j_end - j0);
    for (int i = threadIdx.x; i < lim; i += blockDim.x) tile[i] = r[j0 + i];
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < lim; ++i) {
      const float rj = tile[i];
      cnt += (rj > rc) || (rj == rc && (j0 + i) < c);
    }
    __syncthreads();
  }
  if (c < n && cnt) atomicAdd(&rank[(size_t)e * n + c], cnt);
}

// Elite statistics + refit.  One CTA per action dimension j (< H*A).
//  compat != 0 (reference, :101): mask[p][e] = argsort(-returns[e])[p] < k, i.e. the elite ROWS are the rank positions of
//    candidates 0..k-1: rows {rank[e][c] : c < k}; pooled over envs (:102).
//  compat == 0: rows {c : rank[e][c] < k} (true top-k).
// mean' = alpha*mean + (1-alpha)*mean(elites) (:103), std' = std(elites) (ddof 0, :104), broadcast to all envs.
template <typename T>
__global__ void cem_refit_kernel(const int* __restrict__ rank, const T* __restrict__ clipped, int n, int m, int ha, int k,
                                 double alpha, int compat, double* __restrict__ mean, double* __restrict__ std_) {
  const int j = blockIdx.x;
  __shared__ double s_sum[256];
  __shared__ double s_mean;
  const int tid = threadIdx.x;
  const int limit = compat ? k : n;
  double acc = 0.0;
  for (int idx = tid; idx < m * limit; idx += blockDim.x) {
    const int e = idx / limit, c = idx % limit;
    const int rk = rank[(size_t)e * n + c];
    int rowi = -1;
    if (compat) rowi = rk; else if (rk < k) rowi = c;
    if (rowi >= 0) acc += (double)clipped[((size_t)rowi * m + e) * ha + j];
  }
  s_sum[tid] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (tid < s) s_sum[tid] += s_sum[tid + s]; __syncthreads(); }
  if (tid == 0) s_mean = s_sum[0] / (double)(m * k);
  __syncthreads();
  const double mu = s_mean;
  acc = 0.0;
  for (int idx = tid; idx < m * limit; idx += blockDim.x) {
    const int e = idx / limit, c = idx % limit;
    const int rk = rank[(size_t)e * n + c];
    int rowi = -1;
    if (compat) rowi = rk; else if (rk < k) rowi = c;
    if (rowi >= 0) { const double d = (double)clipped[((size_t)rowi * m + e) * ha + j] - mu; acc += d * d; }
  }
  s_sum[tid] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (tid < s) s_sum[tid] += s_sum[tid + s]; __syncthreads(); }
  if (tid == 0) {
    const double sd = sqrt(s_sum[0] / (double)(m * k));
    for (int e = 0; e < m; ++e) {
      mean[e * ha + j] = mean[e * ha + j] * alpha + (1.0 - alpha) * mu;
      std_[e * ha + j] = sd;
    }
  }
}

}  // namespace l2a
