// K3: candidate-shard glue around the one NCCL all-gather of a multi-GPU planning call (no reference counterpart).
//   pack   : per env (best return, global candidate index as two exact 16-bit halves, first action) -> [m, 3 + A] fp32
//   select : gathered [G, m, 3 + A] -> winner per env with np.argmax semantics over the concatenated candidates
//            (maximum return; NaN beats numbers; ties -> lowest global index).
#pragma once
#include "common.cuh"

namespace l2a {

__global__ void shard_pack_kernel(const float* __restrict__ best_ret, const int* __restrict__ best_idx, const float* __restrict__ best_act,
                                  long long idx_offset, int m, int A, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const long long g = (long long)best_idx[e] + idx_offset;
  float* o = out + (size_t)e * (3 + A);
  o[0] = best_ret[e];
  o[1] = (float)(g >> 16);
  o[2] = (float)(g & 0xFFFF);
  for (int j = 0; j < A; ++j) o[3 + j] = best_act[e * A + j];
}

__global__ void shard_select_kernel(const float* __restrict__ gathered, int G, int m, int A, float* __restrict__ best_ret,
                                    long long* __restrict__ best_idx, float* __restrict__ best_act) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int row = 3 + A;
  int win = 0;
  float bv = 0.f;
  long long bi = 0;
  for (int g = 0; g < G; ++g) {
    const float* p = gathered + ((size_t)g * m + e) * row;
    const float v = p[0];
    const long long idx = ((long long)p[1] << 16) | (long long)p[2];
    bool take = (g == 0);
    if (!take) {
      const bool nv = (v != v), nb = (bv != bv);
      if (nv != nb) take = nv;
      else if (nv) take = idx < bi;
      else take = (v > bv) || (v == bv && idx < bi);
    }
    if (take) { win = g; bv = v; bi = idx; }
  }
  best_ret[e] = bv;
  best_idx[e] = bi;
  const float* p = gathered + ((size_t)win * m + e) * row;
  for (int j = 0; j < A; ++j) best_act[e * A + j] = p[3 + j];
}

}  // namespace l2a
