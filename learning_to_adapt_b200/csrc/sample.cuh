// Candidate sampling on the device: U[low, high) action sequences from Philox4x32-10 (Salmon et al., SC'11), the
// throughput-mode replacement of MPCController.get_random_action (policies/mpc_controller.py:67-69, 114).  Counter-based:
// element block i of planning call c is Philox(key = seed, counter = (i, c)), so a captured CUDA graph can be replayed with
// the call index read from device memory and every replay draws fresh candidates.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace l2a {

__device__ __forceinline__ void philox4x32_10(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
    const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n2 = hi0 ^ ctr[3] ^ k1;
    ctr[0] = n0; ctr[1] = lo1; ctr[2] = n2; ctr[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// out[i] = low[i % A] + (high[i % A] - low[i % A]) * u_i, u_i in [0, 1) with 24 random bits; four values per Philox block.
// call_index: device pointer to the 64-bit planning-call counter (two 32-bit words), or NULL to use call_value.
__global__ void __launch_bounds__(256) sample_uniform_kernel(const float* __restrict__ low, const float* __restrict__ high,
                                                             float* __restrict__ out, long long total, int A, uint64_t seed,
                                                             const uint32_t* __restrict__ call_index, uint64_t call_value) {
  const long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // Philox block = 4 consecutive elements
  const long long i0 = blk * 4;
  if (i0 >= total) return;
  const uint32_t c_lo = call_index ? call_index[0] : (uint32_t)call_value;
  const uint32_t c_hi = call_index ? call_index[1] : (uint32_t)(call_value >> 32);
  uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), c_lo, c_hi};
  philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
  float v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = (int)((i0 + q) % A);
    const float u = (float)(ctr[q] >> 8) * (1.0f / 16777216.0f);
    const float lo = __ldg(low + j), hi = __ldg(high + j);
    v[q] = fmaf(hi - lo, u, lo);
  }
  if (i0 + 4 <= total) {
    *reinterpret_cast<float4*>(out + i0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    for (int q = 0; q < 4 && i0 + q < total; ++q) out[i0 + q] = v[q];
  }
}

// Standard normals from the same Philox stream (throughput-mode replacement of np.random.normal in the CEM planner,
// policies/mpc_controller.py:85): Box-Muller on two 24-bit uniforms per output pair, u1 in (0, 1], u2 in [0, 1):
//   r = sqrt(-2 ln u1),  z0 = r cos(2 pi u2),  z1 = r sin(2 pi u2)      (float32 arithmetic, stored as float64).
// One Philox block (4 words) -> 4 normals: pairs (w0, w1) and (w2, w3).  call_index: device pointer to the 64-bit call counter;
// `stream_id` separates the CEM iterations of one call.
__global__ void __launch_bounds__(256) sample_normal_kernel(double* __restrict__ out, long long total, uint64_t seed,
                                                            const uint32_t* __restrict__ call_index, uint32_t stream_id) {
  const long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i0 = blk * 4;
  if (i0 >= total) return;
  uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32) ^ (stream_id << 16), call_index[0], call_index[1]};
  philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x5EEDu);
  float z[4];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float u1 = (float)((ctr[2 * q] >> 8) + 1u) * (1.0f / 16777216.0f);
    const float u2 = (float)(ctr[2 * q + 1] >> 8) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincosf(6.283185307179586f * u2, &sn, &cs);
    z[2 * q] = r * cs;
    z[2 * q + 1] = r * sn;
  }
  for (int q = 0; q < 4 && i0 + q < total; ++q) out[i0 + q] = (double)z[q];
}

}  // namespace l2a
