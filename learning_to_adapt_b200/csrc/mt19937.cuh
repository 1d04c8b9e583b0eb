// numpy's global MT19937 stream, regenerated bit-exactly on the device: the PARITY-mode replacement of
// MPCController.get_random_action (policies/mpc_controller.py:67-69, 114: np.random.uniform(low, high, (H*N*m, A))) and of the
// CEM draw np.random.normal(size=(n, m, H*A)) (:85).  The host uploads np.random.get_state() (624 key words + position
// [+ cached gaussian]), the device produces exactly the words the host generator would have produced, and the advanced
// state goes back so that np.random.set_state() leaves the process-wide stream where the reference would have left it.
//
// Algorithm restated from the published MT19937 (Matsumoto & Nishimura 1998) and numpy's legacy distributions
// (numpy/random/src/mt19937/mt19937.h: mt19937_next / mt19937_next_double; src/distributions + legacy-distributions.c:
// random_uniform = lower + range * next_double, legacy_gauss = polar Box-Muller with a cached second value); numpy itself
// (installed on the test box) is the oracle: tests compare the device draws with np.random bit for bit.
//
// Three kernels:
//   mt19937_raw_kernel      ONE CTA walks the linear recurrence x[k+624] = x[k+397] ^ twist(x[k], x[k+1]).  Thread k < 227
//                           owns elements k, k+227, k+454 of every 624-word block: new[k+227] and new[k+454] depend only on the
//                           thread's own previous results (lag 227), so a block costs one dependent 3-element chain and one
//                           CTA barrier (element 623, which needs two other chains' results, is recomputed by a ninth warp).  Writes the RAW (untempered) blocks to global memory: the state at any position of
//                           the stream is then just a 624-word window of that buffer.
//   mt19937_uniform_kernel  massively parallel: tempering, 53-bit doubles, low + range * d in float64 (separate mul / add
//                           roundings like the C code), float32 candidate tensor + the float64 copy of time step 0.
//   mt19937_gauss_kernel / _count / _scan / _scatter   legacy polar method: every attempt (2 doubles) in parallel, accepted
//                           attempts compacted in stream order by a three-launch prefix sum over the accept flags.
//   mt19937_state_out_kernel  final (key, pos[, has_gauss, cached]) after exactly the consumed number of words.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace l2a {

constexpr int kMtN = 624, kMtM = 397;

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
  const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
// mt19937_next_double: a = next >> 5, b = next >> 6, (a * 2^26 + b) / 2^53   (exact in float64)
__device__ __forceinline__ double mt_double(uint32_t w0, uint32_t w1) {
  const uint32_t a = mt_temper(w0) >> 5, b = mt_temper(w1) >> 6;
  return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}

// state_in: [624] key words, pos_in[0] = position in [0, 624] (624 = "refill before the next draw").
// raw_out: blocks of 624 raw words; block 0 = the incoming key, block b >= 1 = b-th refill.  Word w (0-based) of the stream the
// host generator would deliver from here is temper(raw_out[pos + w]).  Generates ceil((pos + n_words) / 624) blocks.
constexpr int kMtRawThreads = 288;      // warps 0-7: the 227 element chains; warp 8 (thread 256): element 623 on its own
__global__ void __launch_bounds__(kMtRawThreads, 1) mt19937_raw_kernel(const uint32_t* __restrict__ state_in, const int* __restrict__ pos_in,
                                                                       long long n_words, uint32_t* __restrict__ raw_out) {
  __shared__ uint32_t buf[2][kMtN + 1];
  const int k = threadIdx.x;
  for (int i = k; i < kMtN; i += blockDim.x) {
    const uint32_t v = state_in[i];
    buf[0][i] = v;
    raw_out[i] = v;
  }
  __syncthreads();
  const long long total = (long long)pos_in[0] + n_words;
  const long long nblocks = (total + kMtN - 1) / kMtN;               // block 0 included
  const int kc = (k + 454 < kMtN - 1) ? k + 454 : kMtN - 2;          // clamped index of the chain's third element (k < 169)
  int cur = 0;
  uint32_t* out = raw_out;
  for (long long b = 1; b < nblocks; ++b) {
    const uint32_t* o = buf[cur];
    uint32_t* nw = buf[cur ^ 1];
    out += kMtN;
    if (k < 227) {
      // every input of the chain is an OLD value except the thread's own previous result: all loads first, one dependent
      // chain of three twists, then the stores
      const uint32_t a0 = o[k], a1 = o[k + 1], a2 = o[k + kMtM];
      const uint32_t b0 = o[k + 227], b1 = o[k + 228];
      const uint32_t c0 = o[kc], c1 = o[kc + 1];
      const uint32_t n0 = mt_twist(a0, a1, a2);                      // element k      : old[k], old[k+1], old[k+397]
      const uint32_t n1 = mt_twist(b0, b1, n0);                      // element k+227  : old[k+227], old[k+228], new[k]
      const uint32_t n2 = mt_twist(c0, c1, n1);                      // element k+454  : old[k+454], old[k+455], new[k+227]
      nw[k] = n0;
      nw[k + 227] = n1;
      if (k < 169) nw[k + 454] = n2;
      out[k] = n0;
      out[k + 227] = n1;
      if (k < 169) out[k + 454] = n2;
    } else if (k == 256) {
      // element 623: old[623], NEW[0], NEW[396]; both recomputed here from old values so that no thread waits for another:
      // new[0] = f(old[0], old[1], old[397]);  new[396] = f(old[396], old[397], new[169]);  new[169] = f(old[169], old[170], old[566])
      const uint32_t new0 = mt_twist(o[0], o[1], o[kMtM]);
      const uint32_t new169 = mt_twist(o[169], o[170], o[169 + kMtM]);
      const uint32_t new396 = mt_twist(o[396], o[397], new169);
      const uint32_t n623 = mt_twist(o[623], new0, new396);
      nw[623] = n623;
      out[623] = n623;
    }
    __syncthreads();
    cur ^= 1;
  }
}

// Candidate tensor of random shooting: element e of the flattened [H, rows, A] tensor = low[j] + range[j] * d_e, j = e % A,
// d_e from stream words 2e, 2e+1.  actions: float32 (the TF feed cast); act64_t0: the float64 values of time step 0
// ([rows, A]; what the reference returns as the chosen action, mpc_controller.py:118,129).
__global__ void __launch_bounds__(256) mt19937_uniform_kernel(const uint32_t* __restrict__ raw, const int* __restrict__ pos_in,
                                                              const double* __restrict__ low, const double* __restrict__ range,
                                                              long long total, long long t0_elems, int A,
                                                              float* __restrict__ actions, double* __restrict__ act64_t0) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const uint32_t* w = raw + pos_in[0] + 2 * e;
  const double d = mt_double(w[0], w[1]);
  const int j = (int)(e % A);
  const double v = __dadd_rn(low[j], __dmul_rn(range[j], d));          // lower + range * next_double, no fused multiply-add
  actions[e] = __double2float_rn(v);
  if (e < t0_elems) act64_t0[e] = v;
}

// legacy_gauss attempts.  Attempt i uses stream doubles 2i, 2i+1 (words 4i .. 4i+3):
//   x1 = 2 d0 - 1, x2 = 2 d1 - 1, r2 = x1^2 + x2^2; accepted iff 0 < r2 < 1; f = sqrt(-2 log(r2) / r2);
//   the generator returns f * x2 first and caches f * x1 for the next call.
// flags[i] = accepted; vals[2i] = f * x2, vals[2i + 1] = f * x1.
__global__ void __launch_bounds__(256) mt19937_gauss_kernel(const uint32_t* __restrict__ raw, const int* __restrict__ pos_in,
                                                            long long n_attempts, uint8_t* __restrict__ flags, double* __restrict__ vals) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_attempts) return;
  const uint32_t* w = raw + pos_in[0] + 4 * i;
  const double x1 = __dadd_rn(__dmul_rn(2.0, mt_double(w[0], w[1])), -1.0);
  const double x2 = __dadd_rn(__dmul_rn(2.0, mt_double(w[2], w[3])), -1.0);
  const double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
  const bool ok = (r2 < 1.0) && (r2 != 0.0);
  flags[i] = ok ? 1 : 0;
  if (ok) {
    const double f = sqrt(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
    vals[2 * i] = __dmul_rn(f, x2);
    vals[2 * i + 1] = __dmul_rn(f, x1);
  }
}

// Compaction of the accepted attempts in stream order: the j-th accepted attempt supplies normals 2j (f * x2) and 2j + 1
// (f * x1).  Three small launches: per-CTA accept counts -> exclusive scan of the counts (one CTA) -> scatter.
// `carry` (has_gauss on entry) shifts the outputs by one: z[0] = the cached value.
// meta_out[0] = attempts consumed (-1: the budget was too small), [1] = has_gauss after the draw; cached_out = its value.
constexpr int kGaussChunk = 2048;                            // attempts per CTA (256 threads x 8)

__global__ void __launch_bounds__(256) mt19937_gauss_count_kernel(const uint8_t* __restrict__ flags, long long n_attempts, int* __restrict__ counts) {
  __shared__ int s_warp[8];
  const long long base = (long long)blockIdx.x * kGaussChunk + threadIdx.x * 8;
  int c = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) c += (base + q < n_attempts) ? (int)flags[base + q] : 0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s_warp[w];
    counts[blockIdx.x] = t;
  }
}

// one CTA: counts[i] -> exclusive prefix (in place); n_chunks <= a few thousand
__global__ void __launch_bounds__(1024) mt19937_gauss_scan_kernel(int* __restrict__ counts, int n_chunks) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int start = 0; start < n_chunks; start += 1024) {
    const int i = start + tid;
    const int v = (i < n_chunks) ? counts[i] : 0;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int ws = s_warp[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, ws, off);
        if (lane >= off) ws += u;
      }
      s_warp[lane] = ws;
    }
    __syncthreads();
    if (i < n_chunks) counts[i] = s_base + (warp ? s_warp[warp - 1] : 0) + incl - v;
    __syncthreads();
    if (tid == 0) s_base += s_warp[31];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) mt19937_gauss_scatter_kernel(const uint8_t* __restrict__ flags, const double* __restrict__ vals,
                                                                    const int* __restrict__ offsets, long long n_attempts, long long n_normals,
                                                                    const int* __restrict__ gauss_in, const double* __restrict__ cached_in,
                                                                    double* __restrict__ z_out, int* __restrict__ meta_out,
                                                                    double* __restrict__ cached_out) {
  __shared__ int s_warp[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int carry = gauss_in[0] ? 1 : 0;
  const long long need = n_normals - carry;                  // normals to draw from fresh attempts
  const long long pairs_needed = (need + 1) / 2;
  if (blockIdx.x == 0 && tid == 0) {
    if (carry && n_normals > 0) z_out[0] = cached_in[0];
    if (pairs_needed == 0) {                                 // nothing drawn: the cached value is consumed or kept
      meta_out[0] = 0;
      meta_out[1] = (n_normals > 0) ? 0 : carry;
      cached_out[0] = (n_normals > 0) ? 0.0 : cached_in[0];
    }
  }
  const long long base = (long long)blockIdx.x * kGaussChunk + tid * 8;
  int f[8], mine = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) { f[q] = (base + q < n_attempts) ? (int)flags[base + q] : 0; mine += f[q]; }
  int incl = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += u;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += s_warp[w];
  long long rank = (long long)offsets[blockIdx.x] + before + incl - mine;      // accepted attempts before this thread's first
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (f[q] && rank < pairs_needed) {
      const long long i = base + q;
      const long long o = carry + 2 * rank;
      z_out[o] = vals[2 * i];
      if (o + 1 < n_normals) z_out[o + 1] = vals[2 * i + 1];
      if (rank == pairs_needed - 1) {
        meta_out[0] = (int)(i + 1);                          // attempts consumed
        const int leftover = (need & 1) ? 1 : 0;             // an odd draw leaves f * x1 cached
        meta_out[1] = leftover;
        cached_out[0] = leftover ? vals[2 * i + 1] : 0.0;
      }
    }
    rank += f[q];
  }
}

// The generator state after `words` stream words (words_dev != NULL: 4 * words_dev[0] attempts consumed, CEM; else words_fixed):
// key = the raw block holding the position, pos in [0, 624] with numpy's convention (624 = block exhausted, not yet refilled).
__global__ void mt19937_state_out_kernel(const uint32_t* __restrict__ raw, const int* __restrict__ pos_in, long long words_fixed,
                                         const int* __restrict__ attempts_dev, uint32_t* __restrict__ key_out, int* __restrict__ pos_out) {
  const long long words = attempts_dev ? 4ll * (long long)attempts_dev[0] : words_fixed;
  const long long total = (long long)pos_in[0] + words;
  long long blk = total / kMtN;
  int pos = (int)(total % kMtN);
  if (pos == 0 && total > 0 && words > 0) { blk -= 1; pos = kMtN; }
  if (words == 0) { blk = 0; pos = pos_in[0]; }
  for (int i = threadIdx.x; i < kMtN; i += blockDim.x) key_out[i] = raw[blk * kMtN + i];
  if (threadIdx.x == 0) pos_out[0] = pos;
}

}  // namespace l2a
