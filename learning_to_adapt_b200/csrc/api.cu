// libl2a_b200.so -- C ABI of the B200-native MPC planning engine (see include/l2a_b200.h).
// Host side: argument validation, workspace management, kernel selection and launches.  sm_100a only.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "adapt.cuh"
#include "cem.cuh"
#include "common.cuh"
#ifdef L2A_DEBUG_KERNELS
#include "debug_tile.cuh"
#endif
#include "mt19937.cuh"
#include "rollout_simt.cuh"
#include "rollout_rnn_simt.cuh"
#include "rollout_rnn_tc.cuh"
#include "rollout_tc.cuh"
#include "rollout_tc2.cuh"
#include "sample.cuh"
#include "shard.cuh"
#include "window.cuh"

using namespace l2a;

// AUTO kernel choice: 1 = prefer the CTA-pair tcgen05 rollout where the shape allows it (L2A_TC_PAIR=0/1 overrides at run time)
#ifndef L2A_TC_PAIR_DEFAULT
#define L2A_TC_PAIR_DEFAULT 1
#endif

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) return fail(L2A_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                              \
  } while (0)

struct l2a_ctx {
  float* xch = nullptr;            // ensemble exchange scratch (global / L2), grown on demand
  size_t xch_cap = 0;
  unsigned int* xflags = nullptr;  // CTA-pair rollout: per (tile, rank, member) step counters of the member exchange
  size_t xflags_cap = 0;
  long long* timeline = nullptr;   // optional diagnostics buffer (l2a_debug_set_timeline)
  int device = 0;
  int num_sms = 0;
  int max_smem_optin = 0;
  long long launches = 0;
  long long ws_epoch = 0;          // bumped whenever a workspace pointer a captured graph may hold changes
  // per-env reduction workspace
  float* part_ret = nullptr;
  int* part_idx = nullptr;
  unsigned int* counters = nullptr;
  size_t part_cap = 0, counter_cap = 0;
  // adapt workspace
  float* adapt_acts = nullptr;
  float* adapt_grads = nullptr;
  size_t adapt_acts_cap = 0, adapt_grads_cap = 0;
  int adapt_max_clusters16[2] = {-1, -1};   // cudaOccupancyMaxActiveClusters of adapt_fwd_bwd_kernel<16 / 32> at cluster size 16
};

struct l2a_model {
  l2a_mlp_desc desc;
  MlpDims dims;
  TcPlan plan;
  bool tc_ok = false;
  float* params = nullptr;     // [n_sets][set_stride]
  uint8_t* blobs = nullptr;    // [n_sets][plan.set_bytes]
  Tc2Plan plan2;               // CTA-pair (cta_group::2) rollout: its own stage order and blob
  bool tc2_ok = false;
  uint8_t* blobs2 = nullptr;   // [n_sets][plan2.set_bytes]
  CUtensorMap wmap2;           // blobs2 as a 2-D tensor of 128-byte rows, box = one 16 KB tile
  std::vector<uint8_t> stale1; // per set: the single-CTA blob is behind the fp32 parameters (re-tiled on demand, see launch_prep)
  float* norm = nullptr;       // obs_mean[D] obs_den[D] act_mean[A] act_den[A] delta_mean[D] delta_scale[D]
  bool norm_set = false;
  NormDev norm_dev() const {
    const int D = dims.obs_dim, A = dims.act_dim;
    NormDev n;
    n.obs_mean = norm;
    n.obs_den = norm + D;
    n.act_mean = norm + 2 * D;
    n.act_den = norm + 2 * D + A;
    n.delta_mean = norm + 2 * D + 2 * A;
    n.delta_scale = norm + 3 * D + 2 * A;
    return n;
  }
};

extern "C" const char* l2a_last_error(void) { return g_err; }
extern "C" int l2a_version(void) { return 100; }

extern "C" int l2a_ctx_create(int device, l2a_ctx** out) {
  if (!out) return fail(L2A_ERR_INVALID, "out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(L2A_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(L2A_ERR_INVALID, "device %d out of range [0,%d)", device, count);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(L2A_ERR_UNSUPPORTED, "device %d is sm_%d%d; libl2a_b200 is built for sm_100a (B200) only", device, prop.major,
                prop.minor);
  l2a_ctx* c = new (std::nothrow) l2a_ctx();
  if (!c) return fail(L2A_ERR_INVALID, "out of host memory");
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  *out = c;
  return L2A_OK;
}

extern "C" int l2a_ctx_destroy(l2a_ctx* c) {
  if (!c) return L2A_OK;
  cudaSetDevice(c->device);
  cudaFree(c->part_ret);
  cudaFree(c->part_idx);
  cudaFree(c->counters);
  cudaFree(c->adapt_acts);
  cudaFree(c->adapt_grads);
  cudaFree(c->xch);
  cudaFree(c->xflags);
  delete c;
  return L2A_OK;
}

extern "C" int64_t l2a_ctx_launch_count(const l2a_ctx* c) { return c ? c->launches : 0; }

// dense-stack geometry + fp32 storage offsets of one weight set from the C descriptor (host only)
static void fill_dims(const l2a_mlp_desc* d, MlpDims* out) {
  MlpDims& md = *out;
  memset(&md, 0, sizeof(md));
  md.n_layers = d->n_hidden + 1;
  md.obs_dim = d->obs_dim;
  md.act_dim = d->act_dim;
  md.dims[0] = d->obs_dim + d->act_dim;
  for (int i = 0; i < d->n_hidden; ++i) md.dims[i + 1] = d->hidden[i];
  md.dims[md.n_layers] = d->obs_dim;
  int off = 0, maxw = 0;
  for (int l = 0; l < md.n_layers; ++l) {
    md.w_off[l] = off;
    off += md.dims[l] * md.dims[l + 1];
    md.b_off[l] = off;
    off += md.dims[l + 1];
    off = (off + 3) & ~3;                        // keep every kernel 16-byte aligned
  }
  for (int l = 0; l <= md.n_layers; ++l) maxw = std::max(maxw, md.dims[l]);
  md.set_stride = off;
  md.max_width = maxw;
}

// Host-only query of the tensor-core tiling a model of this shape would get (no device needed): out[0] = 1 if the tcgen05
// rollout supports the shape, then hidden_pairs, out_n, out_kcs, out_stages, stages_per_set, set_bytes (low, high 32 bits).
extern "C" int l2a_tc_plan_query(const l2a_mlp_desc* d, int32_t* out8) {
  if (!d || !out8) return fail(L2A_ERR_INVALID, "NULL argument");
  if (d->n_hidden < 1 || d->n_hidden > kMaxLayers - 1) return fail(L2A_ERR_INVALID, "n_hidden %d not in [1,%d]", d->n_hidden, kMaxLayers - 1);
  for (int i = 0; i < d->n_hidden; ++i)
    if (d->hidden[i] < 1) return fail(L2A_ERR_INVALID, "hidden[%d] = %d", i, d->hidden[i]);
  if (d->obs_dim < 3 || d->act_dim < 1) return fail(L2A_ERR_INVALID, "bad obs_dim/act_dim");
  MlpDims md;
  fill_dims(d, &md);
  TcPlan plan;
  memset(&plan, 0, sizeof(plan));
  const bool ok = tc_make_plan(md, &plan);
  memset(out8, 0, 8 * sizeof(int32_t));
  out8[0] = ok ? 1 : 0;
  if (ok) {
    out8[1] = plan.hidden_pairs; out8[2] = plan.out_n; out8[3] = plan.out_kcs; out8[4] = plan.out_stages;
    out8[5] = plan.stages_per_set; out8[6] = (int32_t)(plan.set_bytes & 0xFFFFFFFFll); out8[7] = (int32_t)(plan.set_bytes >> 32);
  }
  return L2A_OK;
}

// The pair blob as a 2-D tensor map: rows of 128 bytes (one swizzled K-major row of a tile), box = 128 rows = one 16 KB tile;
// no swizzle in the map itself (the tiles are stored pre-swizzled).  cuTensorMapEncodeTiled comes from the driver through the
// runtime's entry-point query, so the library does not link libcuda.
typedef CUresult (*l2a_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool encode_weight_map(CUtensorMap* map, void* base, size_t bytes) {
  static l2a_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      cudaGetLastError();
      return false;
    }
    fn = (l2a_encode_tiled_fn)p;
  }
  const cuuint64_t gdim[2] = {128, (cuuint64_t)(bytes / 128)};
  const cuuint64_t gstride[1] = {128};
  const cuuint32_t box[2] = {128, 128};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int pick_nc2(int num_sms, int n_cand, int n_envs, int csize);

extern "C" int l2a_tc2_plan_query(const l2a_mlp_desc* d, int32_t n_candidates, int32_t n_envs, int32_t n_members, int32_t num_sms, int32_t* out12) {
  if (!d || !out12) return fail(L2A_ERR_INVALID, "NULL argument");
  if (d->n_hidden < 1 || d->n_hidden > kMaxLayers - 1) return fail(L2A_ERR_INVALID, "n_hidden %d not in [1,%d]", d->n_hidden, kMaxLayers - 1);
  for (int i = 0; i < d->n_hidden; ++i)
    if (d->hidden[i] < 1) return fail(L2A_ERR_INVALID, "hidden[%d] = %d", i, d->hidden[i]);
  if (d->obs_dim < 3 || d->act_dim < 1 || n_candidates < 1 || n_envs < 1 || n_members < 1 || num_sms < 2) return fail(L2A_ERR_INVALID, "bad argument");
  MlpDims md;
  fill_dims(d, &md);
  Tc2Plan plan;
  const bool ok = tc2_make_plan(md, &plan) && n_members <= 8;
  memset(out12, 0, 12 * sizeof(int32_t));
  out12[0] = ok ? 1 : 0;
  if (ok) {
    out12[1] = plan.hidden_stages; out12[2] = plan.l0_packed; out12[3] = plan.out_n; out12[4] = plan.out_kcs; out12[5] = plan.out_stages;
    out12[6] = plan.stages_per_set; out12[7] = (int32_t)(plan.set_bytes & 0xFFFFFFFFll); out12[8] = (int32_t)(plan.set_bytes >> 32);
    const int nc = pick_nc2(num_sms, n_candidates, n_envs, n_members);
    const int groups = (n_candidates + 2 * nc - 1) / (2 * nc);
    out12[9] = nc; out12[10] = groups; out12[11] = n_envs * groups * n_members * 2;
  }
  return L2A_OK;
}

// --------------------------------------------------------------------------------------------- model
extern "C" int l2a_model_create(l2a_ctx* c, const l2a_mlp_desc* d, l2a_model** out) {
  if (!c || !d || !out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (d->n_hidden < 1 || d->n_hidden > kMaxLayers - 1) return fail(L2A_ERR_INVALID, "n_hidden %d not in [1,%d]", d->n_hidden, kMaxLayers - 1);
  if (d->obs_dim < 3 || d->act_dim < 1 || d->n_sets < 1) return fail(L2A_ERR_INVALID, "bad obs_dim/act_dim/n_sets");
  for (int i = 0; i < d->n_hidden; ++i)
    if (d->hidden[i] < 1) return fail(L2A_ERR_INVALID, "hidden[%d] = %d", i, d->hidden[i]);
  CUDA_TRY(cudaSetDevice(c->device));
  l2a_model* m = new (std::nothrow) l2a_model();
  if (!m) return fail(L2A_ERR_INVALID, "out of host memory");
  m->desc = *d;
  MlpDims& md = m->dims;
  fill_dims(d, &md);
  m->tc_ok = tc_make_plan(md, &m->plan);
  m->stale1.assign((size_t)d->n_sets, 0);
  const size_t pbytes = (size_t)d->n_sets * md.set_stride * sizeof(float);
  if (cudaMalloc(&m->params, pbytes) != cudaSuccess) { delete m; return fail(L2A_ERR_CUDA, "cudaMalloc(params, %zu)", pbytes); }
  cudaMemset(m->params, 0, pbytes);
  if (m->tc_ok) {
    const size_t bbytes = (size_t)d->n_sets * (size_t)m->plan.set_bytes;
    if (cudaMalloc(&m->blobs, bbytes) != cudaSuccess) { cudaFree(m->params); delete m; return fail(L2A_ERR_CUDA, "cudaMalloc(blobs, %zu)", bbytes); }
    cudaMemset(m->blobs, 0, bbytes);
  }
  m->tc2_ok = tc2_make_plan(md, &m->plan2);
  if (m->tc2_ok) {
    const size_t bbytes = (size_t)d->n_sets * (size_t)m->plan2.set_bytes;
    if (cudaMalloc(&m->blobs2, bbytes) != cudaSuccess) { cudaFree(m->params); cudaFree(m->blobs); delete m; return fail(L2A_ERR_CUDA, "cudaMalloc(blobs2, %zu)", bbytes); }
    cudaMemset(m->blobs2, 0, bbytes);
    if (!encode_weight_map(&m->wmap2, m->blobs2, bbytes)) {          // no driver entry point: the pair kernel is simply not offered
      cudaFree(m->blobs2);
      m->blobs2 = nullptr;
      m->tc2_ok = false;
    }
  }
  const size_t nbytes = sizeof(float) * (size_t)(4 * d->obs_dim + 2 * d->act_dim);
  if (cudaMalloc(&m->norm, nbytes) != cudaSuccess) { cudaFree(m->params); cudaFree(m->blobs); cudaFree(m->blobs2); delete m; return fail(L2A_ERR_CUDA, "cudaMalloc(norm)"); }
  *out = m;
  return L2A_OK;
}

extern "C" int l2a_model_destroy(l2a_ctx* c, l2a_model* m) {
  if (!m) return L2A_OK;
  if (c) cudaSetDevice(c->device);
  cudaFree(m->params);
  cudaFree(m->blobs);
  cudaFree(m->blobs2);
  cudaFree(m->norm);
  delete m;
  return L2A_OK;
}

static bool prefer_pair() {
  static const int v = [] { const char* e = getenv("L2A_TC_PAIR"); return e ? atoi(e) : L2A_TC_PAIR_DEFAULT; }();
  return v != 0;
}

static int launch_prep1(l2a_ctx* c, l2a_model* m, int first_set, int n_sets, cudaStream_t st) {
  PrepArgs pa;
  pa.dims = m->dims;
  pa.plan = m->plan;
  pa.params = m->params;
  pa.blobs = m->blobs;
  pa.first_set = first_set;
  dim3 grid(m->plan.hidden_pairs + m->plan.nkc[m->plan.n_layers - 1], n_sets);
  tc_prep_kernel<<<grid, 256, 0, st>>>(pa);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  for (int s = first_set; s < first_set + n_sets; ++s) m->stale1[s] = 0;
  return L2A_OK;
}

// Re-tile weight sets after their fp32 parameters changed (set_params, adapt, fit).  The blob of the kernel AUTO would choose
// is rebuilt here, on the caller's stream; when that is the CTA-pair blob the single-CTA blob is only marked stale and rebuilt
// by the first rollout that explicitly asks for that kernel (GrBAL adapts every env step: one re-tiling pass, not two).
static int launch_prep(l2a_ctx* c, l2a_model* m, int first_set, int n_sets, cudaStream_t st) {
  if (m->tc2_ok) {
    Prep2Args pa;
    pa.dims = m->dims;
    pa.plan = m->plan2;
    pa.params = m->params;
    pa.blobs = m->blobs2;
    pa.first_set = first_set;
    dim3 grid(m->plan2.hidden_stages + m->plan2.nkc[m->plan2.n_layers - 1], n_sets, 2);
    tc2_prep_kernel<<<grid, 256, 0, st>>>(pa);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  if (!m->tc_ok) return L2A_OK;
  if (m->tc2_ok && prefer_pair()) {
    for (int s = first_set; s < first_set + n_sets; ++s) m->stale1[s] = 1;
    return L2A_OK;
  }
  return launch_prep1(c, m, first_set, n_sets, st);
}

extern "C" int l2a_model_set_params(l2a_ctx* c, l2a_model* m, int set, const float* const* W, const float* const* b, void* stream) {
  if (!c || !m || !W || !b) return fail(L2A_ERR_INVALID, "NULL argument");
  if (set < 0 || set >= m->desc.n_sets) return fail(L2A_ERR_INVALID, "set %d out of range [0,%d)", set, m->desc.n_sets);
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const MlpDims& md = m->dims;
  float* base = m->params + (size_t)set * md.set_stride;
  for (int l = 0; l < md.n_layers; ++l) {
    if (!W[l] || !b[l]) return fail(L2A_ERR_INVALID, "W[%d] or b[%d] is NULL", l, l);
    CUDA_TRY(cudaMemcpyAsync(base + md.w_off[l], W[l], sizeof(float) * (size_t)md.dims[l] * md.dims[l + 1], cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(base + md.b_off[l], b[l], sizeof(float) * (size_t)md.dims[l + 1], cudaMemcpyDeviceToDevice, st));
  }
  return launch_prep(c, m, set, 1, st);
}

extern "C" int l2a_model_get_params(l2a_ctx* c, l2a_model* m, int set, float* const* W, float* const* b, void* stream) {
  if (!c || !m || !W || !b) return fail(L2A_ERR_INVALID, "NULL argument");
  if (set < 0 || set >= m->desc.n_sets) return fail(L2A_ERR_INVALID, "set %d out of range [0,%d)", set, m->desc.n_sets);
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const MlpDims& md = m->dims;
  const float* base = m->params + (size_t)set * md.set_stride;
  for (int l = 0; l < md.n_layers; ++l) {
    CUDA_TRY(cudaMemcpyAsync(W[l], base + md.w_off[l], sizeof(float) * (size_t)md.dims[l] * md.dims[l + 1], cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b[l], base + md.b_off[l], sizeof(float) * (size_t)md.dims[l + 1], cudaMemcpyDeviceToDevice, st));
  }
  return L2A_OK;
}

// Direct access to the resident fp32 parameters of weight sets (training on the device without a host round trip, SURVEY.md
// 8(f) f2): the block of set `set` ([W_0 | b_0 | pad | W_1 | ...], floats_per_set floats, layer l at w_off / b_off as in
// fill_dims).  After writing, l2a_model_refresh re-tiles the sets for the tensor-core rollout.
extern "C" int l2a_model_param_block(l2a_ctx* c, l2a_model* m, int set, float** ptr_out, int64_t* floats_per_set, int32_t* w_off, int32_t* b_off) {
  if (!c || !m || !ptr_out || !floats_per_set) return fail(L2A_ERR_INVALID, "NULL argument");
  if (set < 0 || set >= m->desc.n_sets) return fail(L2A_ERR_INVALID, "set %d out of range [0,%d)", set, m->desc.n_sets);
  *ptr_out = m->params + (size_t)set * m->dims.set_stride;
  *floats_per_set = m->dims.set_stride;
  for (int l = 0; l < m->dims.n_layers; ++l) {
    if (w_off) w_off[l] = m->dims.w_off[l];
    if (b_off) b_off[l] = m->dims.b_off[l];
  }
  return L2A_OK;
}

extern "C" int l2a_model_refresh(l2a_ctx* c, l2a_model* m, int first_set, int n_sets, void* stream) {
  if (!c || !m) return fail(L2A_ERR_INVALID, "NULL argument");
  if (first_set < 0 || n_sets < 1 || first_set + n_sets > m->desc.n_sets) return fail(L2A_ERR_INVALID, "sets [%d,%d) out of range", first_set, first_set + n_sets);
  CUDA_TRY(cudaSetDevice(c->device));
  return launch_prep(c, m, first_set, n_sets, (cudaStream_t)stream);
}

extern "C" int l2a_model_set_normalization(l2a_ctx* c, l2a_model* m, const float* obs_mean, const float* obs_den,
                                           const float* act_mean, const float* act_den, const float* delta_mean,
                                           const float* delta_scale, void* stream) {
  if (!c || !m || !obs_mean || !obs_den || !act_mean || !act_den || !delta_mean || !delta_scale)
    return fail(L2A_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int D = m->dims.obs_dim, A = m->dims.act_dim;
  NormDev n = m->norm_dev();
  CUDA_TRY(cudaMemcpyAsync((void*)n.obs_mean, obs_mean, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync((void*)n.obs_den, obs_den, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync((void*)n.act_mean, act_mean, sizeof(float) * A, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync((void*)n.act_den, act_den, sizeof(float) * A, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync((void*)n.delta_mean, delta_mean, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync((void*)n.delta_scale, delta_scale, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
  m->norm_set = true;
  return L2A_OK;
}

// --------------------------------------------------------------------------------------------- workspace
static int ensure_reduce_ws(l2a_ctx* c, size_t n_part, size_t n_env, cudaStream_t st) {
  if (n_part > c->part_cap) {
    cudaFree(c->part_ret);
    cudaFree(c->part_idx);
    c->part_ret = nullptr;
    c->part_idx = nullptr;
    c->part_cap = 0;
    const size_t cap = std::max<size_t>(n_part * 2, 4096);
    CUDA_TRY(cudaMalloc(&c->part_ret, cap * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->part_idx, cap * sizeof(int)));
    c->part_cap = cap;
    c->ws_epoch++;
  }
  if (n_env > c->counter_cap) {
    cudaFree(c->counters);
    c->counters = nullptr;
    c->counter_cap = 0;
    const size_t cap = std::max<size_t>(n_env * 2, 1024);
    CUDA_TRY(cudaMalloc(&c->counters, cap * sizeof(unsigned int)));
    CUDA_TRY(cudaMemsetAsync(c->counters, 0, cap * sizeof(unsigned int), st));
    c->counter_cap = cap;
    c->ws_epoch++;
  }
  return L2A_OK;
}

// --------------------------------------------------------------------------------------------- rollout
template <int NC, int DMAX>
static int launch_tc(l2a_ctx* c, const TcArgs& ta, int csize, cudaStream_t st) {
  const size_t smem = TcSmem<NC>::total(ta.dims.obs_dim, ta.dims.act_dim);
  if ((int)smem > c->max_smem_optin) return fail(L2A_ERR_UNSUPPORTED, "tcgen05 rollout needs %zu B shared memory (> %d)", smem, c->max_smem_optin);
  CUDA_TRY(cudaFuncSetAttribute(rollout_tc_kernel<NC, DMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(ta.n_envs * ta.groups_per_env * csize));
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, rollout_tc_kernel<NC, DMAX>, ta));
  c->launches++;
  return L2A_OK;
}

template <int NC, int DMAX>
static int launch_tc2(l2a_ctx* c, const l2a_model* m, const Tc2Args& ta, int csize, cudaStream_t st) {
  const size_t smem = Tc2Smem<NC>::total;
  if ((int)smem > c->max_smem_optin) return fail(L2A_ERR_UNSUPPORTED, "CTA-pair rollout needs %zu B shared memory (> %d)", smem, c->max_smem_optin);
  CUDA_TRY(cudaFuncSetAttribute(rollout_tc2_kernel<NC, DMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(ta.n_envs * ta.groups_per_env * csize * 2));
  cfg.blockDim = dim3(kTc2Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, rollout_tc2_kernel<NC, DMAX>, ta, m->wmap2));
  c->launches++;
  return L2A_OK;
}

// candidates per CTA of the pair kernel (a tile is 2 * nc candidates): fewest waves of CTA pairs, then the smaller tile
static int pick_nc2(int num_sms, int n_cand, int n_envs, int csize) {
  static const int opts[4] = {80, 72, 48, 32};
  if (const char* ov = getenv("L2A_TC2_NC")) {                  // tuning / experiments only
    const int v = atoi(ov);
    if (v == 80 || v == 72 || v == 48 || v == 32) return v;
  }
  int best = 72;
  double best_cost = 1e300;
  // candidate tiles resident at once: 1 CTA / SM, and the csize member pairs of a tile only make progress together (they
  // meet every horizon step), so a wave holds whole tiles
  const int slots = std::max(1, (num_sms / 2) / csize);
  for (int i = 0; i < 4; ++i) {
    const int nc = opts[i];
    const long long tiles = (long long)n_envs * ((n_cand + 2 * nc - 1) / (2 * nc));
    const long long waves = (tiles + slots - 1) / slots;
    const double cost = (double)waves * (48.0 + nc);
    if (cost < best_cost) { best_cost = cost; best = nc; }
  }
  return best;
}

static int pick_nc(const l2a_ctx* c, int n_cand, int n_envs, int csize) {
  static const int opts[4] = {80, 64, 48, 32};
  if (const char* ov = getenv("L2A_TC_NC")) {                   // tuning / experiments only
    const int v = atoi(ov);
    if (v == 80 || v == 64 || v == 48 || v == 32) return v;
  }
  int best = 80;
  double best_cost = 1e300;
  const int slots = std::max(1, c->num_sms / csize);           // clusters resident at once (1 CTA / SM)
  for (int i = 0; i < 4; ++i) {
    const int nc = opts[i];
    const long long clusters = (long long)n_envs * ((n_cand + nc - 1) / nc);
    const long long waves = (clusters + slots - 1) / slots;
    const double cost = (double)waves * (48.0 + nc);           // fixed per-step weight-stream cost + N-proportional MMA time
    if (cost < best_cost) { best_cost = cost; best = nc; }
  }
  return best;
}

extern "C" int l2a_rollout(l2a_ctx* c, l2a_model* m, const l2a_rollout_params* p, const float* obs0, const float* actions,
                           const float* discount_pow, float* returns, float* best_ret, int32_t* best_idx, float* best_act,
                           void* stream) {
  if (!c || !m || !p || !obs0 || !actions || !discount_pow || !best_ret || !best_idx || !best_act)
    return fail(L2A_ERR_INVALID, "NULL argument");
  if (!m->norm_set) return fail(L2A_ERR_INVALID, "normalization not set (l2a_model_set_normalization)");
  if (p->n_candidates < 1 || p->n_envs < 1 || p->horizon < 1) return fail(L2A_ERR_INVALID, "n_candidates/n_envs/horizon must be >= 1");
  if (p->reward_kind < 0 || p->reward_kind > 2) return fail(L2A_ERR_INVALID, "reward_kind %d", p->reward_kind);
  if (!(p->dt > 0.f)) return fail(L2A_ERR_INVALID, "dt must be > 0");
  int last_set = p->first_set;
  if (p->set_mode == L2A_SETS_PER_ENV) last_set = p->first_set + p->n_envs - 1;
  else if (p->set_mode == L2A_SETS_ENSEMBLE_MEAN) {
    if (p->n_sets < 1) return fail(L2A_ERR_INVALID, "n_sets must be >= 1");
    last_set = p->first_set + p->n_sets - 1;
  } else if (p->set_mode != L2A_SETS_SHARED) return fail(L2A_ERR_INVALID, "set_mode %d", p->set_mode);
  if (p->first_set < 0 || last_set >= m->desc.n_sets)
    return fail(L2A_ERR_INVALID, "weight sets [%d,%d] out of range [0,%d)", p->first_set, last_set, m->desc.n_sets);
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;

  int kernel = p->kernel;
  const int csize = (p->set_mode == L2A_SETS_ENSEMBLE_MEAN) ? p->n_sets : 1;
  const bool tc_possible = m->tc_ok && csize <= 8;
  const bool tc2_possible = m->tc2_ok && csize <= 8;
  if (kernel == L2A_KERNEL_AUTO) {
    kernel = (tc2_possible && prefer_pair()) ? L2A_KERNEL_TCGEN05_PAIR : tc_possible ? L2A_KERNEL_TCGEN05 : L2A_KERNEL_SIMT;
  }
  if (kernel == L2A_KERNEL_TCGEN05_PAIR && !tc2_possible)
    return fail(L2A_ERR_UNSUPPORTED, "the CTA-pair tcgen05 rollout needs hidden widths that are multiples of 256 (<= 512), obs_dim <= 48, "
                                     "act_dim <= 16 and <= 8 ensemble members");
  if (kernel == L2A_KERNEL_TCGEN05 && !tc_possible)
    return fail(L2A_ERR_UNSUPPORTED, "tcgen05 rollout needs hidden widths that are multiples of 128 (<= 512), obs_dim <= 128, "
                                     "act_dim <= 16 and <= 8 ensemble members");

  ReduceArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.best_ret = best_ret;
  ra.best_idx = best_idx;
  ra.best_act = best_act;
  ra.actions = actions;
  ra.act_stride_row = p->act_stride_row;
  ra.act_dim = m->dims.act_dim;
  ra.n_candidates = p->n_candidates;

  if (kernel == L2A_KERNEL_SIMT) {
    const int tiles = (p->n_candidates + kSimtRT - 1) / kSimtRT;
    int rc = ensure_reduce_ws(c, (size_t)tiles * p->n_envs, p->n_envs, st);
    if (rc) return rc;
    ra.part_ret = c->part_ret;
    ra.part_idx = c->part_idx;
    ra.counters = c->counters;
    ra.tiles_per_env = tiles;
    SimtArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.dims = m->dims;
    sa.norm = m->norm_dev();
    sa.params = m->params;
    sa.obs = obs0;
    sa.actions = actions;
    sa.act_stride_t = p->act_stride_t;
    sa.act_stride_row = p->act_stride_row;
    sa.discount_pow = discount_pow;
    sa.rows_per_group = p->n_candidates;
    sa.n_groups = p->n_envs;
    sa.horizon = p->horizon;
    sa.set_mode = p->set_mode;
    sa.first_set = p->first_set;
    sa.n_sets = p->n_sets;
    sa.reward_kind = p->reward_kind;
    sa.dt = p->dt;
    sa.returns = returns;
    sa.red = ra;
    const size_t smem = simt_smem_bytes(m->dims);
    if ((int)smem > c->max_smem_optin)
      return fail(L2A_ERR_UNSUPPORTED, "SIMT rollout needs %zu B shared memory (> %d): layer width %d too large", smem, c->max_smem_optin, m->dims.max_width);
    CUDA_TRY(cudaFuncSetAttribute(rollout_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rollout_simt_kernel<false><<<(unsigned)(tiles * p->n_envs), kSimtThreads, smem, st>>>(sa);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return L2A_OK;
  }

  if (kernel == L2A_KERNEL_TCGEN05_PAIR) {
    const int nc = pick_nc2(c->num_sms, p->n_candidates, p->n_envs, csize);
    const int groups = (p->n_candidates + 2 * nc - 1) / (2 * nc);
    int rc = ensure_reduce_ws(c, (size_t)groups * 2 * p->n_envs, p->n_envs, st);
    if (rc) return rc;
    ra.part_ret = c->part_ret;
    ra.part_idx = c->part_idx;
    ra.counters = c->counters;
    ra.tiles_per_env = groups * 2;
    Tc2Args ta;
    memset(&ta, 0, sizeof(ta));
    ta.dims = m->dims;
    ta.plan = m->plan2;
    ta.norm = m->norm_dev();
    ta.params = m->params;
    ta.obs0 = obs0;
    ta.actions = actions;
    ta.act_stride_t = p->act_stride_t;
    ta.act_stride_row = p->act_stride_row;
    ta.discount_pow = discount_pow;
    ta.n_candidates = p->n_candidates;
    ta.n_envs = p->n_envs;
    ta.horizon = p->horizon;
    ta.set_mode = p->set_mode;
    ta.first_set = p->first_set;
    ta.n_sets = p->n_sets;
    ta.reward_kind = p->reward_kind;
    ta.dt = p->dt;
    ta.groups_per_env = groups;
    ta.returns = returns;
    ta.red = ra;
    ta.timeline = c->timeline;
    const bool small = m->dims.obs_dim <= 24 && m->dims.act_dim <= 8;
    if (csize > 1) {
      const size_t tiles = (size_t)p->n_envs * groups;
      // floats: [tile][parity][E][rank][NC][DMAX]; the flag-in-data exchange stores a step word beside every value (2x)
      const size_t need = tiles * 2 * csize * 2 * (size_t)nc * (small ? 24 : 48) * (L2A_TC2_LL ? 2 : 1);
      if (need > c->xch_cap) {
        cudaFree(c->xch);
        c->xch = nullptr;
        c->xch_cap = 0;
        CUDA_TRY(cudaMalloc(&c->xch, need * sizeof(float)));
        c->xch_cap = need;
        c->ws_epoch++;
      }
      const size_t nflags = tiles * 2 * csize;
      if (nflags > c->xflags_cap) {
        cudaFree(c->xflags);
        c->xflags = nullptr;
        c->xflags_cap = 0;
        CUDA_TRY(cudaMalloc(&c->xflags, nflags * 2 * sizeof(unsigned int)));
        c->xflags_cap = nflags * 2;
        c->ws_epoch++;
      }
      CUDA_TRY(cudaMemsetAsync(c->xflags, 0, nflags * sizeof(unsigned int), st));        // step counters start at 0 every launch
      if (L2A_TC2_LL) CUDA_TRY(cudaMemsetAsync(c->xch, 0, need * sizeof(float), st));      // ... and so do the step words in the rows
      ta.xch = c->xch;
      ta.flags = c->xflags;
    }
    if (small) {
      switch (nc) {
        case 80: return launch_tc2<80, 24>(c, m, ta, csize, st);
        case 72: return launch_tc2<72, 24>(c, m, ta, csize, st);
        case 48: return launch_tc2<48, 24>(c, m, ta, csize, st);
        default: return launch_tc2<32, 24>(c, m, ta, csize, st);
      }
    }
    switch (nc) {
      case 80: return launch_tc2<80, 48>(c, m, ta, csize, st);
      case 72: return launch_tc2<72, 48>(c, m, ta, csize, st);
      case 48: return launch_tc2<48, 48>(c, m, ta, csize, st);
      default: return launch_tc2<32, 48>(c, m, ta, csize, st);
    }
  }

  for (int s = p->first_set; s <= last_set; ++s)
    if (m->stale1[s]) {                                  // (see launch_prep) contiguous runs of stale sets, re-tiled on this stream
      int e = s;
      while (e + 1 <= last_set && m->stale1[e + 1]) ++e;
      int rc1 = launch_prep1(c, m, s, e - s + 1, st);
      if (rc1) return rc1;
      s = e;
    }
  const int nc = pick_nc(c, p->n_candidates, p->n_envs, csize);
  const int groups = (p->n_candidates + nc - 1) / nc;
  int rc = ensure_reduce_ws(c, (size_t)groups * p->n_envs, p->n_envs, st);
  if (rc) return rc;
  ra.part_ret = c->part_ret;
  ra.part_idx = c->part_idx;
  ra.counters = c->counters;
  ra.tiles_per_env = groups;
  TcArgs ta;
  memset(&ta, 0, sizeof(ta));
  ta.dims = m->dims;
  ta.plan = m->plan;
  ta.norm = m->norm_dev();
  ta.params = m->params;
  ta.blobs = m->blobs;
  ta.obs0 = obs0;
  ta.actions = actions;
  ta.act_stride_t = p->act_stride_t;
  ta.act_stride_row = p->act_stride_row;
  ta.discount_pow = discount_pow;
  ta.n_candidates = p->n_candidates;
  ta.n_envs = p->n_envs;
  ta.horizon = p->horizon;
  ta.set_mode = p->set_mode;
  ta.first_set = p->first_set;
  ta.n_sets = p->n_sets;
  ta.reward_kind = p->reward_kind;
  ta.dt = p->dt;
  ta.groups_per_env = groups;
  ta.returns = returns;
  ta.red = ra;
  ta.timeline = c->timeline;
  if (csize > 1) {
    const bool small = m->dims.obs_dim <= 24 && m->dims.act_dim <= 8;
    const size_t blk = (size_t)nc * (small ? 24 : 48);                                // floats per member block: [NC][DMAX]
    const size_t need = (size_t)p->n_envs * groups * 2 * csize * blk;
    if (need > c->xch_cap) {
      cudaFree(c->xch);
      c->xch = nullptr;
      c->xch_cap = 0;
      CUDA_TRY(cudaMalloc(&c->xch, need * sizeof(float)));
      c->xch_cap = need;
      c->ws_epoch++;
    }
    ta.xch = c->xch;
  }
  if (m->dims.obs_dim <= 24 && m->dims.act_dim <= 8) {
    switch (nc) {
      case 80: return launch_tc<80, 24>(c, ta, csize, st);
      case 64: return launch_tc<64, 24>(c, ta, csize, st);
      case 48: return launch_tc<48, 24>(c, ta, csize, st);
      default: return launch_tc<32, 24>(c, ta, csize, st);
    }
  }
  switch (nc) {
    case 80: return launch_tc<80, 48>(c, ta, csize, st);
    case 64: return launch_tc<64, 48>(c, ta, csize, st);
    case 48: return launch_tc<48, 48>(c, ta, csize, st);
    default: return launch_tc<32, 48>(c, ta, csize, st);
  }
}


// --------------------------------------------------------------------------------------------- host-buffer planning call
// One planning call with HOST buffers on both sides (what MPCController.get_actions is to its caller,
// policies/mpc_controller.py:59-65 + 108-129): H2D(obs, call index, [numpy generator state]) -> candidate sampling (Philox, or
// numpy's MT19937 stream regenerated on the device) -> K1 -> [candidate-shard exchange with the peer GPUs] -> D2H(result,
// [advanced generator state]).  The stream operations are captured once into a CUDA graph and replayed (one cudaGraphLaunch
// per call); everything that changes from call to call travels in the pinned input block.
struct l2a_window;
static int adapt_impl(l2a_ctx* c, l2a_model* m, const float* x, const float* target, int K, int M, float inner_lr, int src_set,
                      int dst_first_set, void* stream);
static int window_gather_own(l2a_ctx* c, l2a_window* w, const float** x_out, const float** target_out, int* K_out, int* M_out, cudaStream_t st);
static int window_push_strided(l2a_ctx* c, l2a_window* w, const double* obs, const double* act, int act_stride, bool count_on_host, cudaStream_t st);
static bool window_matches(const l2a_window* w, int n_envs, int D, int A);
static void window_count_push(l2a_window* w);

struct l2a_plan {
  l2a_model* model = nullptr;
  l2a_rollout_params p;          // p.n_candidates = this rank's candidates per env
  l2a_plan_opts o;
  int D = 0, A = 0, rec = 0;     // rec = doubles per result record: (return, global index, action[A])
  uint64_t calls = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr;
  // CEM with numpy's normal stream: the generator chain (raw words -> polar attempts -> compaction -> state) of iteration i+1 runs on
  // a second stream beside K1 of iteration i (it needs the generator state after iteration i's draw, not its returns)
  cudaStream_t stream_mt = nullptr;
  std::vector<cudaEvent_t> ev_mt;    // fork, then per iteration: normals ready, normals consumed; last: chain done
  uint8_t* in_host = nullptr;    // pinned input block (layout: in_* offsets)
  uint8_t* in_dev = nullptr;
  uint8_t* out_host = nullptr;   // pinned output block (layout: out_* offsets)
  uint8_t* out_dev = nullptr;
  size_t in_bytes = 0, out_bytes = 0;
  size_t in_call = 0, in_key = 0, in_pos = 0;                               // byte offsets inside the input block
  size_t out_ret = 0, out_idx = 0, out_act = 0, out_rec = 0, out_final = 0, out_key = 0, out_pos = 0;
  float* actions = nullptr;      // [H, m*N, A]
  double* act64_t0 = nullptr;    // MT19937: float64 candidates of time step 0, [m*N, A]
  uint32_t* mt_raw = nullptr;    // MT19937: raw generator blocks
  long long mt_words = 0;
  float* consts = nullptr;       // low [A], high [A], discount_pow [H], then (CEM) clip_low [H*A], clip_high [H*A]
  double* consts64 = nullptr;    // low [A], high - low [A]
  // CEM planner state (o.planner == L2A_PLANNER_CEM); `actions` then holds the samples [N, m, H*A]
  double* z64 = nullptr;         // [N, m, H*A] standard normals
  double* clipped = nullptr;     // [N, m, H*A] float64 like the reference's a_stacked
  float* returns = nullptr;      // [m, N]
  int32_t* rank = nullptr;       // [m, N]
  double* mean = nullptr;        // [m, H*A], std right behind it
  double* first64 = nullptr;     // [N*m, A] float64 first actions of the rows
  uint8_t* gflags = nullptr;     // MT19937 gauss attempts: accept flags, values
  double* gvals = nullptr;
  int* gcounts = nullptr;        // per-chunk accept counts / offsets
  long long attempts = 0;        // attempt budget per iteration
  uint8_t* mt_scratch = nullptr; // two generator-state blocks (MtStateBlock) + per-iteration meta int[2 * iters]
  size_t out_mean = 0, out_meta = 0;
  // GrBAL: adaptation window attached to the plan (l2a_plan_attach_window)
  l2a_window* window = nullptr;
  float win_lr = 0.f;
  int win_src_set = 0, win_dst_first_set = 1;
  size_t in_obs64 = 0;           // float64 copy of the observations (the window keeps float64 pairs)
  int cur_flags = 0;             // L2A_PLAN_* flags of the call being enqueued
  cudaGraphExec_t exec_flags[4] = {nullptr, nullptr, nullptr, nullptr};   // one captured graph per flag combination
  // candidate shard (o.shard_world > 1): exchange buffer of THIS rank (peers write into it) and the peers' buffers
  uint8_t* xbuf = nullptr;
  size_t xbuf_bytes = 0, xarea_bytes = 0;
  uint8_t* res_scratch = nullptr;   // l2a_plan_exchange_resident: this rank's records + its sequence counter
  void** peers_dev = nullptr;    // device array [world] of peer xbuf pointers
  bool peers_attached = false;
  long long calls_flags[4] = {0, 0, 0, 0};
  int launches_flags[4] = {0, 0, 0, 0};
  long long graph_epoch = -1;
  bool use_graph = true;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// numpy generator state as it travels in the pinned blocks / between CEM iterations on the device
struct MtStateBlock {
  uint32_t key[kMtN];
  int32_t pos;
  int32_t has_gauss;
  double cached;
};
static_assert(sizeof(MtStateBlock) == 2512, "MtStateBlock layout");

// ---- candidate-shard exchange over peer memory (NVLink): every rank writes its per-env record into every peer's buffer,
// raises a sequence flag there (release at system scope), waits for all ranks' flags in its own buffer and selects the winner
// with np.argmax semantics over the concatenated candidates (max return, NaN wins, ties -> lowest global index).  One CTA.
// Buffer layout per rank: flags uint64[world] (padded to 256 B) | data [2 parities][world][m][rec] doubles.
struct ShardXArgs {
  void* const* peers;            // [world] device pointers to the ranks' exchange buffers (own included)
  int rank, world, m, rec;
  const uint32_t* call_index;    // 64-bit call counter in the input block (low word first), or NULL:
  unsigned long long* own_seq;   // ... a device counter this kernel advances itself (l2a_plan_exchange_resident)
  const double* mine;            // [m][rec] this rank's records
  double* final_rec;             // [m][rec] winner per env
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(128) shard_exchange_kernel(const ShardXArgs a) {
  const int tid = threadIdx.x;
  __shared__ unsigned long long s_seq;
  if (tid == 0) s_seq = a.call_index ? (((unsigned long long)a.call_index[0] | ((unsigned long long)a.call_index[1] << 32)) + 1ull)
                                     : ++(*a.own_seq);
  __syncthreads();
  const unsigned long long seq = s_seq;
  const int par = (int)(seq & 1ull);
  const size_t flag_bytes = ((size_t)a.world * 8 + 255) / 256 * 256;
  const size_t blk = (size_t)a.m * a.rec;                                   // doubles per (parity, rank) block
  for (int g = 0; g < a.world; ++g) {
    double* dst = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(a.peers[g]) + flag_bytes) + ((size_t)par * a.world + a.rank) * blk;
    for (size_t i = tid; i < blk; i += blockDim.x) dst[i] = a.mine[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < a.world) st_release_sys_u64(reinterpret_cast<unsigned long long*>(a.peers[tid]) + a.rank, seq);
  if (tid < a.world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(a.peers[a.rank]) + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(f) < seq) {
      if (clock64() - t0 > 200000000000ll) __trap();                        // ~100 s: a peer never made this call
    }
  }
  __syncthreads();
  const double* data = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(a.peers[a.rank]) + flag_bytes) + (size_t)par * a.world * blk;
  for (int e = tid; e < a.m; e += blockDim.x) {
    int win = 0;
    double bv = 0.0, bi = 0.0;
    for (int g = 0; g < a.world; ++g) {
      const double* r = data + ((size_t)g * a.m + e) * a.rec;
      const double v = __ldcv(r), idx = __ldcv(r + 1);
      bool take = (g == 0);
      if (!take) {
        const bool nv = (v != v), nb = (bv != bv);
        if (nv != nb) take = nv;
        else if (nv) take = idx < bi;
        else take = (v > bv) || (v == bv && idx < bi);
      }
      if (take) { win = g; bv = v; bi = idx; }
    }
    const double* r = data + ((size_t)win * a.m + e) * a.rec;
    for (int j = 0; j < a.rec; ++j) a.final_rec[(size_t)e * a.rec + j] = __ldcv(r + j);
  }
}

// per-env result record of this rank: (return, global candidate index, chosen first action in float64)
__global__ void plan_record_kernel(const float* __restrict__ best_ret, const int* __restrict__ best_idx, const float* __restrict__ best_act,
                                   const double* __restrict__ act64_t0, long long idx_offset, int n_loc, int m, int A, double* __restrict__ rec_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  double* r = rec_out + (size_t)e * (2 + A);
  const int bi = best_idx[e];
  r[0] = (double)best_ret[e];
  r[1] = (double)((long long)bi + idx_offset);
  for (int j = 0; j < A; ++j)
    r[2 + j] = act64_t0 ? act64_t0[((size_t)e * n_loc + bi) * A + j] : (double)best_act[e * A + j];
}

// MT19937 uniform draw restricted to this rank's slice of the reference tensor [H, m, N_total, A] -> [H, m * n_loc, A]
__global__ void __launch_bounds__(256) mt19937_uniform_slice_kernel(const uint32_t* __restrict__ raw, const int* __restrict__ pos_in,
                                                                    const double* __restrict__ low, const double* __restrict__ range,
                                                                    long long total_loc, int n_total, int lo, int n_loc, int m, int A,
                                                                    float* __restrict__ actions, double* __restrict__ act64_t0) {
  const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // destination element
  if (d >= total_loc) return;
  const int j = (int)(d % A);
  const long long row = d / A;                                               // t * (m * n_loc) + env * n_loc + c
  const int c = (int)(row % n_loc);
  const long long te = row / n_loc;                                          // t * m + env
  const long long e = ((te * n_total) + lo + c) * A + j;                     // element of the reference's full tensor
  const uint32_t* w = raw + pos_in[0] + 2 * e;
  const double v = __dadd_rn(low[j], __dmul_rn(range[j], mt_double(w[0], w[1])));
  actions[d] = __double2float_rn(v);
  if (te < m) act64_t0[d] = v;
}

extern "C" int l2a_plan_destroy(l2a_ctx* c, l2a_plan* pl) {
  if (!pl) return L2A_OK;
  if (c) cudaSetDevice(c->device);
  if (pl->stream) cudaStreamSynchronize(pl->stream);
  for (int f = 0; f < 4; ++f)
    if (pl->exec_flags[f]) cudaGraphExecDestroy(pl->exec_flags[f]);
  if (pl->ev_in) cudaEventDestroy(pl->ev_in);
  for (cudaEvent_t ev : pl->ev_mt) cudaEventDestroy(ev);
  if (pl->stream_mt) cudaStreamDestroy(pl->stream_mt);
  if (pl->stream) cudaStreamDestroy(pl->stream);
  cudaFreeHost(pl->in_host);
  cudaFreeHost(pl->out_host);
  cudaFree(pl->in_dev);
  cudaFree(pl->out_dev);
  cudaFree(pl->actions);
  cudaFree(pl->act64_t0);
  cudaFree(pl->mt_raw);
  cudaFree(pl->consts);
  cudaFree(pl->consts64);
  cudaFree(pl->xbuf);
  cudaFree(pl->res_scratch);
  cudaFree(pl->peers_dev);
  cudaFree(pl->z64);
  cudaFree(pl->clipped);
  cudaFree(pl->returns);
  cudaFree(pl->rank);
  cudaFree(pl->mean);
  cudaFree(pl->first64);
  cudaFree(pl->gflags);
  cudaFree(pl->gvals);
  cudaFree(pl->gcounts);
  cudaFree(pl->mt_scratch);
  delete pl;
  return L2A_OK;
}

extern "C" int l2a_plan_create_ex(l2a_ctx* c, l2a_model* m, const l2a_rollout_params* p, double discount, const double* low,
                                  const double* high, const l2a_plan_opts* opts, l2a_plan** out) {
  if (!c || !m || !p || !low || !high || !opts || !out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (p->n_candidates < 1 || p->n_envs < 1 || p->horizon < 1) return fail(L2A_ERR_INVALID, "n_candidates/n_envs/horizon must be >= 1");
  if (opts->sampler != L2A_SAMPLER_PHILOX && opts->sampler != L2A_SAMPLER_MT19937) return fail(L2A_ERR_INVALID, "sampler %d", opts->sampler);
  const int world = opts->shard_world < 1 ? 1 : opts->shard_world;
  if (opts->shard_rank < 0 || opts->shard_rank >= world || world > 64) return fail(L2A_ERR_INVALID, "shard rank %d / world %d", opts->shard_rank, world);
  if (opts->planner != L2A_PLANNER_RS && opts->planner != L2A_PLANNER_CEM) return fail(L2A_ERR_INVALID, "planner %d", opts->planner);
  if (opts->planner == L2A_PLANNER_CEM) {
    if (world > 1) return fail(L2A_ERR_UNSUPPORTED, "the CEM planner is not sharded across GPUs");
    if (opts->cem_iters < 1 || opts->cem_num_elites < 1 || opts->cem_num_elites > p->n_candidates)
      return fail(L2A_ERR_INVALID, "cem_iters %d / cem_num_elites %d", opts->cem_iters, opts->cem_num_elites);
    if (!opts->cem_compat && p->n_envs > 1)
      return fail(L2A_ERR_UNSUPPORTED, "corrected CEM (cem_compat = 0) is defined for one env per call (see l2a_cem_refit)");
  }
  const int n_total = opts->n_candidates_total > 0 ? opts->n_candidates_total : p->n_candidates;
  if (opts->shard_offset < 0 || opts->shard_offset + p->n_candidates > n_total)
    return fail(L2A_ERR_INVALID, "shard [%lld, %lld) outside the %d candidates", (long long)opts->shard_offset,
                (long long)opts->shard_offset + p->n_candidates, n_total);
  CUDA_TRY(cudaSetDevice(c->device));
  l2a_plan* pl = new (std::nothrow) l2a_plan();
  if (!pl) return fail(L2A_ERR_INVALID, "out of host memory");
  pl->model = m;
  pl->p = *p;
  pl->o = *opts;
  pl->o.shard_world = world;
  pl->o.n_candidates_total = n_total;
  pl->D = m->dims.obs_dim;
  pl->A = m->dims.act_dim;
  pl->rec = 2 + pl->A;
  const int mm = p->n_envs, A = pl->A, H = p->horizon, D = pl->D;
  const long long rows = (long long)p->n_candidates * mm;
  pl->p.act_stride_t = rows * A;                       // the plan owns the candidate tensor: [H, m*N, A] (mpc_controller.py:114)
  pl->p.act_stride_row = A;
  if (opts->planner == L2A_PLANNER_CEM) {              // CEM: [N, m, H*A] viewed as (N*m, H, A) (:85-89)
    pl->p.act_stride_t = A;
    pl->p.act_stride_row = (long long)H * A;
  }
  const bool mt = opts->sampler == L2A_SAMPLER_MT19937;
  // input block: obs f32 [m, D] | call index u64 | MT key u32 [624] | MT pos i32
  size_t off = sizeof(float) * (size_t)mm * D;
  off = align_up(off, 8);  pl->in_call = off;  off += 8;
  pl->in_key = off;  pl->in_pos = off + offsetof(MtStateBlock, pos);  off += mt ? sizeof(MtStateBlock) : 0;
  pl->in_obs64 = off;  off += sizeof(double) * (size_t)mm * D;
  pl->in_bytes = align_up(off, 16);
  // output block: best_ret f32 [m] | best_idx i32 [m] | best_act f32 [m, A] | local record f64 [m, rec] | final record f64 [m, rec] |
  //               MT key u32 [624] | MT pos i32
  off = 0;
  pl->out_ret = off;  off += sizeof(float) * mm;
  pl->out_idx = off;  off += sizeof(int32_t) * mm;
  pl->out_act = off;  off += sizeof(float) * (size_t)mm * A;
  off = align_up(off, 8);
  pl->out_rec = off;  off += sizeof(double) * (size_t)mm * pl->rec;
  pl->out_final = off;  off += sizeof(double) * (size_t)mm * pl->rec;
  pl->out_key = off;  pl->out_pos = off + offsetof(MtStateBlock, pos);  off += mt ? sizeof(MtStateBlock) : 0;
  const bool cem = opts->planner == L2A_PLANNER_CEM;
  const int ha = H * A;
  pl->out_mean = off;  off += cem ? sizeof(double) * 2 * (size_t)mm * ha : 0;
  pl->out_meta = off;  off += cem ? sizeof(int32_t) * 2 * (size_t)std::max(1, opts->cem_iters) : 0;
  pl->out_bytes = align_up(off, 16);
  pl->use_graph = getenv("L2A_NO_GRAPH") == nullptr;
  // a plan that forces the single-CTA tcgen05 kernel while AUTO prefers the CTA pair re-tiles stale sets on demand inside
  // l2a_rollout (launch_prep): that decision is taken per call on the host, so such a plan issues its launches directly
  if (p->kernel == L2A_KERNEL_TCGEN05 && m->tc2_ok && prefer_pair()) pl->use_graph = false;
  std::vector<float> consts((size_t)2 * A + H + (cem ? 2 * (size_t)ha : 0));
  std::vector<double> consts64((size_t)2 * A);
  if (cem)
    for (int t = 0; t < H; ++t)
      for (int j = 0; j < A; ++j) {                                      // np.concatenate([low] * h) (:81-82)
        consts[(size_t)2 * A + H + t * A + j] = (float)low[j];
        consts[(size_t)2 * A + H + ha + t * A + j] = (float)high[j];
      }
  for (int j = 0; j < A; ++j) {
    consts[j] = (float)low[j]; consts[A + j] = (float)high[j];
    consts64[j] = low[j]; consts64[A + j] = high[j] - low[j];      // numpy: range = high - low in float64 (RandomState.uniform)
  }
  double pw = 1.0;
  for (int t = 0; t < H; ++t) { consts[2 * A + t] = (float)pw; pw *= discount; }   // discount**t (mpc_controller.py:126)
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_in, cudaEventDisableTiming);
  if (e == cudaSuccess && opts->planner == L2A_PLANNER_CEM && opts->sampler == L2A_SAMPLER_MT19937 && getenv("L2A_CEM_SERIAL") == nullptr) {
    e = cudaStreamCreateWithFlags(&pl->stream_mt, cudaStreamNonBlocking);
    for (int i = 0; i < 2 * opts->cem_iters + 2 && e == cudaSuccess; ++i) {
      cudaEvent_t ev = nullptr;
      e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e == cudaSuccess) pl->ev_mt.push_back(ev);
    }
  }
  if (e == cudaSuccess) e = cudaMallocHost(&pl->in_host, pl->in_bytes);
  if (e == cudaSuccess) e = cudaMallocHost(&pl->out_host, pl->out_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&pl->in_dev, pl->in_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&pl->out_dev, pl->out_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&pl->actions, sizeof(float) * (size_t)H * rows * A);
  if (e == cudaSuccess) e = cudaMalloc(&pl->consts, sizeof(float) * consts.size());
  if (e == cudaSuccess) e = cudaMemcpy(pl->consts, consts.data(), sizeof(float) * consts.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&pl->consts64, sizeof(double) * consts64.size());
  if (e == cudaSuccess) e = cudaMemcpy(pl->consts64, consts64.data(), sizeof(double) * consts64.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) { memset(pl->in_host, 0, pl->in_bytes); memset(pl->out_host, 0, pl->out_bytes); }
  if (cem && e == cudaSuccess) {
    const size_t tot = (size_t)rows * ha;
    e = cudaMalloc(&pl->z64, sizeof(double) * tot);
    if (e == cudaSuccess) e = cudaMalloc(&pl->clipped, sizeof(double) * tot);
    if (e == cudaSuccess) e = cudaMalloc(&pl->returns, sizeof(float) * (size_t)rows);
    if (e == cudaSuccess) e = cudaMalloc(&pl->rank, sizeof(int32_t) * (size_t)rows);
    if (e == cudaSuccess) e = cudaMalloc(&pl->mean, sizeof(double) * 2 * (size_t)mm * ha);
    if (e == cudaSuccess) e = cudaMalloc(&pl->first64, sizeof(double) * (size_t)rows * A);
    if (mt && e == cudaSuccess) {
      const long long pairs = ((long long)tot + 1) / 2;
      pl->attempts = pairs + pairs * 3 / 10 + 1024;                       // acceptance pi/4: mean 1.273 x pairs, >= 30 sigma of margin
      pl->mt_words = 4ll * pl->attempts;
      const long long blocks = (pl->mt_words + kMtN) / kMtN + 2;
      e = cudaMalloc(&pl->mt_raw, sizeof(uint32_t) * (size_t)blocks * kMtN);
      if (e == cudaSuccess) e = cudaMalloc(&pl->gflags, (size_t)pl->attempts);
      if (e == cudaSuccess) e = cudaMalloc(&pl->gvals, sizeof(double) * 2 * (size_t)pl->attempts);
      if (e == cudaSuccess) e = cudaMalloc(&pl->gcounts, sizeof(int) * (size_t)((pl->attempts + kGaussChunk - 1) / kGaussChunk + 1));
      if (e == cudaSuccess) e = cudaMalloc(&pl->mt_scratch, 2 * sizeof(MtStateBlock) + 64);
    }
  }
  if (mt && !cem && e == cudaSuccess) {
    pl->mt_words = 2ll * H * mm * (long long)n_total * A;                 // two 32-bit words per double, the reference's FULL tensor
    const long long blocks = (pl->mt_words + kMtN) / kMtN + 2;
    e = cudaMalloc(&pl->mt_raw, sizeof(uint32_t) * (size_t)blocks * kMtN);
    if (e == cudaSuccess) e = cudaMalloc(&pl->act64_t0, sizeof(double) * (size_t)rows * A);
  }
  if (world > 1 && e == cudaSuccess) {
    // two independent exchange areas in one allocation: [0] for l2a_plan_run_ex, [1] for l2a_plan_exchange_resident
    pl->xarea_bytes = align_up(align_up((size_t)world * 8, 256) + sizeof(double) * 2 * (size_t)world * mm * pl->rec, 256);
    pl->xbuf_bytes = 2 * pl->xarea_bytes;
    e = cudaMalloc(&pl->xbuf, pl->xbuf_bytes);
    if (e == cudaSuccess) e = cudaMemset(pl->xbuf, 0, pl->xbuf_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&pl->peers_dev, sizeof(void*) * 2 * world);
    if (e == cudaSuccess) e = cudaMalloc(&pl->res_scratch, sizeof(double) * (size_t)mm * pl->rec + 64);
    if (e == cudaSuccess) e = cudaMemset(pl->res_scratch, 0, sizeof(double) * (size_t)mm * pl->rec + 64);
  }
  if (e != cudaSuccess) {
    l2a_plan_destroy(c, pl);
    return fail(L2A_ERR_CUDA, "l2a_plan_create: %s", cudaGetErrorString(e));
  }
  *out = pl;
  return L2A_OK;
}

extern "C" int l2a_plan_create(l2a_ctx* c, l2a_model* m, const l2a_rollout_params* p, float discount, const float* low,
                               const float* high, uint64_t seed, l2a_plan** out) {
  if (!m || !low || !high) return fail(L2A_ERR_INVALID, "NULL argument");
  std::vector<double> lo(m->dims.act_dim), hi(m->dims.act_dim);
  for (int j = 0; j < m->dims.act_dim; ++j) { lo[j] = low[j]; hi[j] = high[j]; }
  l2a_plan_opts o;
  memset(&o, 0, sizeof(o));
  o.sampler = L2A_SAMPLER_PHILOX;
  o.seed = seed;
  o.shard_world = 1;
  return l2a_plan_create_ex(c, m, p, (double)discount, lo.data(), hi.data(), &o, out);
}

// ---- peer memory plumbing for the candidate shard: the host side exchanges the 64-byte IPC handles of the ranks' exchange
// buffers (e.g. with torch.distributed.all_gather_object), opens the peers' and hands the pointers to the plan.
extern "C" int l2a_plan_exchange_buffer(l2a_plan* pl, void** ptr_out, uint64_t* bytes_out) {
  if (!pl || !ptr_out || !bytes_out) return fail(L2A_ERR_INVALID, "NULL argument");
  *ptr_out = pl->xbuf;
  *bytes_out = pl->xbuf_bytes;
  return L2A_OK;
}
extern "C" int l2a_ipc_get_handle(l2a_ctx* c, void* dev_ptr, void* handle64_out) {
  if (!c || !dev_ptr || !handle64_out) return fail(L2A_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, dev_ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64_out, &h, 64);
  return L2A_OK;
}
extern "C" int l2a_ipc_open_handle(l2a_ctx* c, const void* handle64, void** dev_ptr_out) {
  if (!c || !handle64 || !dev_ptr_out) return fail(L2A_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return L2A_OK;
}
extern "C" int l2a_ipc_close_handle(l2a_ctx* c, void* dev_ptr) {
  if (!c || !dev_ptr) return fail(L2A_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return L2A_OK;
}
extern "C" int l2a_plan_attach_peers(l2a_ctx* c, l2a_plan* pl, void* const* peer_bufs) {
  if (!c || !pl || !peer_bufs) return fail(L2A_ERR_INVALID, "NULL argument");
  if (pl->o.shard_world < 2) return fail(L2A_ERR_INVALID, "the plan is not sharded");
  CUDA_TRY(cudaSetDevice(c->device));
  const int world = pl->o.shard_world;
  std::vector<void*> ptrs(2 * (size_t)world);
  for (int g = 0; g < world; ++g) {
    ptrs[g] = (g == pl->o.shard_rank) ? (void*)pl->xbuf : peer_bufs[g];
    if (!ptrs[g]) return fail(L2A_ERR_INVALID, "peer buffer %d is NULL", g);
    ptrs[world + g] = (uint8_t*)ptrs[g] + pl->xarea_bytes;
  }
  CUDA_TRY(cudaMemcpy(pl->peers_dev, ptrs.data(), sizeof(void*) * ptrs.size(), cudaMemcpyHostToDevice));
  pl->peers_attached = true;
  for (int f = 0; f < 4; ++f)
    if (pl->exec_flags[f]) { cudaGraphExecDestroy(pl->exec_flags[f]); pl->exec_flags[f] = nullptr; }
  return L2A_OK;
}

extern "C" int l2a_sample_uniform(l2a_ctx* c, const float* low, const float* high, float* out, int64_t rows, int A, uint64_t seed,
                                  uint64_t call_index, void* stream) {
  if (!c || !low || !high || !out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (rows < 1 || A < 1) return fail(L2A_ERR_INVALID, "rows and A must be >= 1");
  CUDA_TRY(cudaSetDevice(c->device));
  const long long total = (long long)rows * A;
  const long long blocks = (total + 4 * 256 - 1) / (4 * 256);
  sample_uniform_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(low, high, out, total, A, seed, nullptr, call_index);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

// CEM winner record: (return, candidate index, float64 first action of the winning row of the (N*m, H, A) view)
__global__ void plan_record_cem_kernel(const float* __restrict__ best_ret, const int* __restrict__ best_idx, const double* __restrict__ first64,
                                       int n, int m, int A, double* __restrict__ rec_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  double* r = rec_out + (size_t)e * (2 + A);
  const int bi = best_idx[e];
  r[0] = (double)best_ret[e];
  r[1] = (double)bi;
  for (int j = 0; j < A; ++j) r[2 + j] = first64[((size_t)e * n + bi) * A + j];
}

static int plan_enqueue_cem(l2a_ctx* c, l2a_plan* pl) {
  const int mm = pl->p.n_envs, A = pl->A, H = pl->p.horizon, n = pl->p.n_candidates, ha = H * A;
  const long long tot = (long long)n * mm * ha;
  cudaStream_t st = pl->stream;
  const bool mt = pl->o.sampler == L2A_SAMPLER_MT19937;
  const uint32_t* call_dev = reinterpret_cast<const uint32_t*>(pl->in_dev + pl->in_call);
  double* mean = pl->mean;
  double* std_ = pl->mean + (size_t)mm * ha;
  cem_init_kernel<<<(mm * ha + 255) / 256, 256, 0, st>>>(mean, std_, mm * ha);                       // :79-80
  c->launches++;
  float* best_ret = reinterpret_cast<float*>(pl->out_dev + pl->out_ret);
  int32_t* best_idx = reinterpret_cast<int32_t*>(pl->out_dev + pl->out_idx);
  float* best_act = reinterpret_cast<float*>(pl->out_dev + pl->out_act);
  const float* clip_low = pl->consts + 2 * A + H;
  const float* clip_high = clip_low + ha;
  const MtStateBlock* state_in = reinterpret_cast<const MtStateBlock*>(pl->in_dev + pl->in_key);
  MtStateBlock* scratch = reinterpret_cast<MtStateBlock*>(pl->mt_scratch);
  int32_t* meta_out = reinterpret_cast<int32_t*>(pl->out_dev + pl->out_meta);
  // numpy's stream: generator chain on its own stream (see l2a_plan), forked behind the H2D copy of the uploaded state
  const bool fork = mt && pl->stream_mt != nullptr;
  cudaStream_t smt = fork ? pl->stream_mt : st;
  if (fork) {
    CUDA_TRY(cudaEventRecord(pl->ev_mt[0], st));
    CUDA_TRY(cudaStreamWaitEvent(smt, pl->ev_mt[0], 0));
  }
  auto enqueue_normals = [&](int it) -> int {
    const bool last_it = (it + 1 == pl->o.cem_iters);
    // np.random.normal(size=(n, m, h*A)) (:85) continued from where the previous iteration left the generator
    MtStateBlock* state_out = last_it ? reinterpret_cast<MtStateBlock*>(pl->out_dev + pl->out_key) : &scratch[it & 1];
    mt19937_raw_kernel<<<1, kMtRawThreads, 0, smt>>>(state_in->key, &state_in->pos, pl->mt_words, pl->mt_raw);
    mt19937_gauss_kernel<<<(unsigned)((pl->attempts + 255) / 256), 256, 0, smt>>>(pl->mt_raw, &state_in->pos, pl->attempts, pl->gflags, pl->gvals);
    const int n_chunks = (int)((pl->attempts + kGaussChunk - 1) / kGaussChunk);
    CUDA_TRY(cudaMemsetAsync(meta_out + 2 * it, 0xFF, 2 * sizeof(int32_t), smt));                 // attempts consumed = -1 until found
    mt19937_gauss_count_kernel<<<n_chunks, 256, 0, smt>>>(pl->gflags, pl->attempts, pl->gcounts);
    mt19937_gauss_scan_kernel<<<1, 1024, 0, smt>>>(pl->gcounts, n_chunks);
    if (fork && it > 0) CUDA_TRY(cudaStreamWaitEvent(smt, pl->ev_mt[2 * (it - 1) + 2], 0));       // z64 of the previous iteration has been consumed
    mt19937_gauss_scatter_kernel<<<n_chunks, 256, 0, smt>>>(pl->gflags, pl->gvals, pl->gcounts, pl->attempts, tot, &state_in->has_gauss,
                                                           &state_in->cached, pl->z64, meta_out + 2 * it, &state_out->cached);
    if (fork) CUDA_TRY(cudaEventRecord(pl->ev_mt[2 * it + 1], smt));                               // normals of iteration `it` ready
    mt19937_state_out_kernel<<<1, 256, 0, smt>>>(pl->mt_raw, &state_in->pos, 0, meta_out + 2 * it, state_out->key, &state_out->pos);
    CUDA_TRY(cudaMemcpyAsync(&state_out->has_gauss, meta_out + 2 * it + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, smt));
    c->launches += 6;
    state_in = state_out;
    return L2A_OK;
  };
  if (fork) {                                        // iteration 0's draw; iteration i+1's is enqueued as soon as z64 of iteration i is consumed
    int rc = enqueue_normals(0);
    if (rc) return rc;
    if (pl->o.cem_iters == 1) CUDA_TRY(cudaEventRecord(pl->ev_mt[2 * pl->o.cem_iters + 1], smt));
  }
  for (int it = 0; it < pl->o.cem_iters; ++it) {
    const bool last = (it + 1 == pl->o.cem_iters);
    if (mt) {
      if (fork) CUDA_TRY(cudaStreamWaitEvent(st, pl->ev_mt[2 * it + 1], 0));
      else { int rc = enqueue_normals(it); if (rc) return rc; }
    } else {
      sample_normal_kernel<<<(unsigned)((tot + 4 * 256 - 1) / (4 * 256)), 256, 0, st>>>(pl->z64, tot, pl->o.seed, call_dev, (uint32_t)it);
      c->launches++;
    }
    const int blocks = (int)std::min<long long>((tot + 255) / 256, (long long)c->num_sms * 8);
    cem_sample64_kernel<<<blocks, 256, 0, st>>>(pl->z64, mean, std_, clip_low, clip_high, n, mm, ha, A, pl->actions, pl->clipped, pl->first64);   // :86-87
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    if (fork) {
      CUDA_TRY(cudaEventRecord(pl->ev_mt[2 * it + 2], st));                                          // z64 consumed
      if (!last) {                                                                                   // next iteration's draw, beside this iteration's K1
        int rc2 = enqueue_normals(it + 1);
        if (rc2) return rc2;
        if (it + 2 == pl->o.cem_iters) CUDA_TRY(cudaEventRecord(pl->ev_mt[2 * pl->o.cem_iters + 1], smt));
      }
    }
    int rc = l2a_rollout(c, pl->model, &pl->p, reinterpret_cast<const float*>(pl->in_dev), pl->actions, pl->consts + 2 * A, pl->returns,
                         best_ret, best_idx, best_act, st);                                          // :88-100
    if (rc) return rc;
    if (last) {
      plan_record_cem_kernel<<<(mm + 127) / 128, 128, 0, st>>>(best_ret, best_idx, pl->first64, n, mm, A,
                                                               reinterpret_cast<double*>(pl->out_dev + pl->out_final));   // :106
      c->launches++;
    }
    dim3 g1((n + 255) / 256, mm, kRankSplit);
    CUDA_TRY(cudaMemsetAsync(pl->rank, 0, sizeof(int32_t) * (size_t)n * mm, st));
    cem_rank_kernel<<<g1, 256, 0, st>>>(pl->returns, n, pl->rank);                                   // :101
    cem_refit_kernel<double><<<ha, 256, 0, st>>>(pl->rank, pl->clipped, n, mm, ha, pl->o.cem_num_elites, pl->o.cem_alpha, pl->o.cem_compat, mean, std_);   // :102-104
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
  }
  if (fork) CUDA_TRY(cudaStreamWaitEvent(st, pl->ev_mt[2 * pl->o.cem_iters + 1], 0));                 // join: the generator chain is complete
  CUDA_TRY(cudaMemcpyAsync(pl->out_dev + pl->out_mean, mean, sizeof(double) * 2 * (size_t)mm * ha, cudaMemcpyDeviceToDevice, st));
  return L2A_OK;
}

static int plan_enqueue_rs(l2a_ctx* c, l2a_plan* pl);

// the stream-ordered body of one planning call (captured into the graph, or issued directly):
//   H2D -> [GrBAL: window gather -> K2 adapt -> re-tile]  -> planner (sampling, K1, records)  -> [window push of (obs, action)] -> D2H
static int plan_enqueue(l2a_ctx* c, l2a_plan* pl) {
  cudaStream_t st = pl->stream;
  CUDA_TRY(cudaMemcpyAsync(pl->in_dev, pl->in_host, pl->in_bytes, cudaMemcpyHostToDevice, st));
  if (pl->cur_flags & L2A_PLAN_ADAPT) {
    // dynamics_model.switch_to_pre_adapt(); dynamics_model.adapt(obs[-M-1:-1], act[-M-1:-1], obs[-M:])   (samplers/sampler.py:82-90)
    const float *x = nullptr, *target = nullptr;
    int K = 0, M = 0;
    int rc = window_gather_own(c, pl->window, &x, &target, &K, &M, st);
    if (rc) return rc;
    rc = adapt_impl(c, pl->model, x, target, K, M, pl->win_lr, pl->win_src_set, pl->win_dst_first_set, st);
    if (rc) return rc;
  }
  int rc = (pl->o.planner == L2A_PLANNER_CEM) ? plan_enqueue_cem(c, pl) : plan_enqueue_rs(c, pl);
  if (rc) return rc;
  if (pl->cur_flags & L2A_PLAN_PUSH) {
    // running_paths[idx]["observations"].append(obs); ["actions"].append(action)   (sampler.py:109-110), float64 like the host lists
    const double* fin = reinterpret_cast<const double*>(pl->out_dev + pl->out_final);
    rc = window_push_strided(c, pl->window, reinterpret_cast<const double*>(pl->in_dev + pl->in_obs64), fin + 2, pl->rec,
                             /*count_on_host=*/false, st);
    if (rc) return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(pl->out_host, pl->out_dev, pl->out_bytes, cudaMemcpyDeviceToHost, st));
  return L2A_OK;
}

static int plan_enqueue_rs(l2a_ctx* c, l2a_plan* pl) {
  const int mm = pl->p.n_envs, A = pl->A, H = pl->p.horizon;
  const long long total = (long long)H * pl->p.n_candidates * mm * A;
  cudaStream_t st = pl->stream;
  const uint32_t* call_dev = reinterpret_cast<const uint32_t*>(pl->in_dev + pl->in_call);
  const bool mt = pl->o.sampler == L2A_SAMPLER_MT19937;
  if (mt) {
    const uint32_t* key = reinterpret_cast<const uint32_t*>(pl->in_dev + pl->in_key);
    const int* pos = reinterpret_cast<const int*>(pl->in_dev + pl->in_pos);
    mt19937_raw_kernel<<<1, kMtRawThreads, 0, st>>>(key, pos, pl->mt_words, pl->mt_raw);
    c->launches++;
    const long long blocks = (total + 255) / 256;
    mt19937_uniform_slice_kernel<<<(unsigned)blocks, 256, 0, st>>>(pl->mt_raw, pos, pl->consts64, pl->consts64 + A, total,
                                                                  pl->o.n_candidates_total, (int)pl->o.shard_offset, pl->p.n_candidates,
                                                                  mm, A, pl->actions, pl->act64_t0);
    c->launches++;
  } else {
    const long long blocks = (total + 4 * 256 - 1) / (4 * 256);
    sample_uniform_kernel<<<(unsigned)blocks, 256, 0, st>>>(pl->consts, pl->consts + A, pl->actions, total, A,
                                                            pl->o.seed + 0x9E3779B97F4A7C15ull * (uint64_t)pl->o.shard_rank, call_dev, 0ull);
    c->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  float* best_ret = reinterpret_cast<float*>(pl->out_dev + pl->out_ret);
  int32_t* best_idx = reinterpret_cast<int32_t*>(pl->out_dev + pl->out_idx);
  float* best_act = reinterpret_cast<float*>(pl->out_dev + pl->out_act);
  int rc = l2a_rollout(c, pl->model, &pl->p, reinterpret_cast<const float*>(pl->in_dev), pl->actions, pl->consts + 2 * A, nullptr,
                       best_ret, best_idx, best_act, st);
  if (rc) return rc;
  double* rec = reinterpret_cast<double*>(pl->out_dev + pl->out_rec);
  double* fin = reinterpret_cast<double*>(pl->out_dev + pl->out_final);
  plan_record_kernel<<<(mm + 127) / 128, 128, 0, st>>>(best_ret, best_idx, best_act, mt ? pl->act64_t0 : nullptr, pl->o.shard_offset,
                                                       pl->p.n_candidates, mm, A, pl->o.shard_world > 1 ? rec : fin);
  c->launches++;
  if (pl->o.shard_world > 1) {
    ShardXArgs xa;
    xa.peers = pl->peers_dev;
    xa.rank = pl->o.shard_rank;
    xa.world = pl->o.shard_world;
    xa.m = mm;
    xa.rec = pl->rec;
    xa.call_index = call_dev;
    xa.own_seq = nullptr;
    xa.mine = rec;
    xa.final_rec = fin;
    shard_exchange_kernel<<<1, 128, 0, st>>>(xa);
    c->launches++;
  }
  if (mt) {
    mt19937_state_out_kernel<<<1, 256, 0, st>>>(pl->mt_raw, reinterpret_cast<const int*>(pl->in_dev + pl->in_pos), pl->mt_words, nullptr,
                                                reinterpret_cast<uint32_t*>(pl->out_dev + pl->out_key),
                                                reinterpret_cast<int*>(pl->out_dev + pl->out_pos));
    c->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

extern "C" int l2a_plan_run_ex(l2a_ctx* c, l2a_plan* pl, const double* obs, l2a_plan_io* io, void* stream) {
  if (!c || !pl || !obs || !io || !io->act_out) return fail(L2A_ERR_INVALID, "NULL argument");
  const bool mt = pl->o.sampler == L2A_SAMPLER_MT19937;
  if (mt && (!io->mt_key || !io->mt_pos)) return fail(L2A_ERR_INVALID, "the MT19937 sampler needs the generator state (mt_key, mt_pos)");
  if (mt && (*io->mt_pos < 0 || *io->mt_pos > kMtN)) return fail(L2A_ERR_INVALID, "mt_pos %d not in [0, 624]", *io->mt_pos);
  if (pl->o.shard_world > 1 && !pl->peers_attached) return fail(L2A_ERR_INVALID, "sharded plan without peers (l2a_plan_attach_peers)");
  CUDA_TRY(cudaSetDevice(c->device));
  const int mm = pl->p.n_envs, A = pl->A, D = pl->D;
  float* obs32 = reinterpret_cast<float*>(pl->in_host);
  for (int i = 0; i < mm * D; ++i) obs32[i] = (float)obs[i];                          // the float32 feed of mlp_dynamics.py:212-214
  memcpy(pl->in_host + pl->in_call, &pl->calls, 8);
  const bool cem = pl->o.planner == L2A_PLANNER_CEM;
  if (mt && cem && (!io->mt_has_gauss || !io->mt_cached)) return fail(L2A_ERR_INVALID, "MT19937 + CEM needs mt_has_gauss / mt_cached");
  if (mt) {
    MtStateBlock* sb = reinterpret_cast<MtStateBlock*>(pl->in_host + pl->in_key);
    memcpy(sb->key, io->mt_key, sizeof(uint32_t) * kMtN);
    sb->pos = *io->mt_pos;
    sb->has_gauss = cem ? *io->mt_has_gauss : 0;
    sb->cached = cem ? *io->mt_cached : 0.0;
  }
  const int flags = io->flags & 3;
  if (flags && !pl->window) return fail(L2A_ERR_INVALID, "flags %d need an attached adaptation window (l2a_plan_attach_window)", flags);
  if (flags) {
    double* o64 = reinterpret_cast<double*>(pl->in_host + pl->in_obs64);
    for (int i = 0; i < mm * D; ++i) o64[i] = obs[i];
  }
  pl->cur_flags = flags;
  // everything the caller queued on its stream (weight uploads, adapt) happens before this call
  CUDA_TRY(cudaEventRecord(pl->ev_in, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamWaitEvent(pl->stream, pl->ev_in, 0));
  bool done = false;
  cudaGraphExec_t& exec = pl->exec_flags[flags];
  if (pl->graph_epoch != c->ws_epoch) {                           // a workspace the graphs point into was reallocated
    for (int f = 0; f < 4; ++f)
      if (pl->exec_flags[f]) { cudaGraphExecDestroy(pl->exec_flags[f]); pl->exec_flags[f] = nullptr; }
    pl->graph_epoch = c->ws_epoch;
  }
  if (pl->use_graph && pl->calls_flags[flags] > 0) {
    if (!exec) {
      // (the first call with these flags ran uncaptured and sized every workspace, so no allocation happens inside the capture)
      cudaGraph_t g = nullptr;
      const long long launches0 = c->launches, epoch0 = c->ws_epoch;
      cudaError_t e = cudaStreamBeginCapture(pl->stream, cudaStreamCaptureModeThreadLocal);
      int rc = L2A_OK;
      if (e == cudaSuccess) {
        rc = plan_enqueue(c, pl);
        e = cudaStreamEndCapture(pl->stream, &g);
      }
      pl->launches_flags[flags] = (int)(c->launches - launches0);
      c->launches = launches0;
      if (e == cudaSuccess && rc == L2A_OK && g && c->ws_epoch == epoch0) e = cudaGraphInstantiate(&exec, g, 0);
      if (g) cudaGraphDestroy(g);
      if (e != cudaSuccess || rc != L2A_OK || !exec || c->ws_epoch != epoch0) {
        if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; }
        cudaGetLastError();
        pl->use_graph = false;                                    // direct launches from now on (same kernels)
      }
    }
    if (exec) {
      CUDA_TRY(cudaGraphLaunch(exec, pl->stream));
      c->launches += pl->launches_flags[flags];
      done = true;
    }
  }
  if (!done) {
    int rc = plan_enqueue(c, pl);
    if (rc) return rc;
    pl->graph_epoch = c->ws_epoch;
  }
  pl->calls_flags[flags]++;
  if (flags & L2A_PLAN_PUSH) window_count_push(pl->window);
  CUDA_TRY(cudaStreamSynchronize(pl->stream));
  pl->calls++;
  const double* fin = reinterpret_cast<const double*>(pl->out_host + pl->out_final);
  for (int e = 0; e < mm; ++e) {
    if (io->ret_out) io->ret_out[e] = (float)fin[(size_t)e * pl->rec];
    if (io->idx_out) io->idx_out[e] = (int64_t)fin[(size_t)e * pl->rec + 1];
    for (int j = 0; j < A; ++j) io->act_out[e * A + j] = fin[(size_t)e * pl->rec + 2 + j];
  }
  if (cem && mt) {
    const int32_t* meta = reinterpret_cast<const int32_t*>(pl->out_host + pl->out_meta);
    for (int it = 0; it < pl->o.cem_iters; ++it)
      if (meta[2 * it] < 0)
        return fail(L2A_ERR_UNSUPPORTED, "CEM iteration %d: the normal draw needed more than %lld polar-method attempts", it, pl->attempts);
  }
  if (mt) {
    const MtStateBlock* sb = reinterpret_cast<const MtStateBlock*>(pl->out_host + pl->out_key);
    memcpy(io->mt_key, sb->key, sizeof(uint32_t) * kMtN);
    *io->mt_pos = sb->pos;
    if (cem) { *io->mt_has_gauss = sb->has_gauss; *io->mt_cached = sb->cached; }
  }
  if (cem) {
    const size_t cnt = (size_t)mm * pl->p.horizon * A;
    const double* ms = reinterpret_cast<const double*>(pl->out_host + pl->out_mean);
    if (io->cem_mean_out) memcpy(io->cem_mean_out, ms, sizeof(double) * cnt);
    if (io->cem_std_out) memcpy(io->cem_std_out, ms + cnt, sizeof(double) * cnt);
  }
  return L2A_OK;
}

extern "C" int l2a_plan_run(l2a_ctx* c, l2a_plan* pl, const double* obs, double* act_out, float* ret_out, int32_t* idx_out,
                            void* stream) {
  if (!pl || !act_out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (pl->o.sampler != L2A_SAMPLER_PHILOX) return fail(L2A_ERR_INVALID, "l2a_plan_run serves Philox plans; use l2a_plan_run_ex");
  std::vector<int64_t> idx64(pl->p.n_envs);
  l2a_plan_io io;
  memset(&io, 0, sizeof(io));
  io.act_out = act_out;
  io.ret_out = ret_out;
  io.idx_out = idx64.data();
  const int rc = l2a_plan_run_ex(c, pl, obs, &io, stream);
  if (rc == L2A_OK && idx_out)
    for (int e = 0; e < pl->p.n_envs; ++e) idx_out[e] = (int32_t)idx64[e];
  return rc;
}

extern "C" int l2a_plan_uses_graph(const l2a_plan* pl) {
  if (!pl) return 0;
  for (int f = 0; f < 4; ++f)
    if (pl->exec_flags[f]) return 1;
  return 0;
}

// The shard exchange alone, on DEVICE-resident per-rank results (what l2a_rollout wrote): record -> peer exchange -> winner
// records final_rec_out [m, 2 + A] float64 (return, global index, action) on every rank.  Uses its own exchange buffers and
// sequence counter (a second set next to the ones l2a_plan_run_ex uses), so it may be interleaved with plan runs as long as
// every rank makes the same sequence of calls.
extern "C" int l2a_plan_exchange_resident(l2a_ctx* c, l2a_plan* pl, const float* best_ret, const int32_t* best_idx, const float* best_act,
                                          double* final_rec_out, void* stream) {
  if (!c || !pl || !best_ret || !best_idx || !best_act || !final_rec_out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (pl->o.shard_world < 2 || !pl->peers_attached) return fail(L2A_ERR_INVALID, "needs a sharded plan with attached peers");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int mm = pl->p.n_envs, A = pl->A;
  double* rec = reinterpret_cast<double*>(pl->res_scratch);
  plan_record_kernel<<<(mm + 127) / 128, 128, 0, st>>>(best_ret, best_idx, best_act, nullptr, pl->o.shard_offset, pl->p.n_candidates, mm, A, rec);
  ShardXArgs xa;
  xa.peers = pl->peers_dev + pl->o.shard_world;          // the second set of exchange buffers
  xa.rank = pl->o.shard_rank;
  xa.world = pl->o.shard_world;
  xa.m = mm;
  xa.rec = pl->rec;
  xa.call_index = nullptr;
  xa.own_seq = reinterpret_cast<unsigned long long*>(pl->res_scratch + sizeof(double) * (size_t)mm * pl->rec);
  xa.mine = rec;
  xa.final_rec = final_rec_out;
  shard_exchange_kernel<<<1, 128, 0, st>>>(xa);
  c->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

extern "C" int l2a_plan_io_bytes(const l2a_plan* pl, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  if (!pl || !h2d_bytes || !d2h_bytes) return fail(L2A_ERR_INVALID, "NULL argument");
  *h2d_bytes = pl->in_bytes;
  *d2h_bytes = pl->out_bytes;
  return L2A_OK;
}

extern "C" int l2a_plan_copy_returns(l2a_ctx* c, l2a_plan* pl, float* host_out) {
  if (!c || !pl || !host_out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (!pl->returns) return fail(L2A_ERR_INVALID, "this plan keeps no per-candidate returns (random-shooting plans reduce them on the fly)");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(pl->stream));
  CUDA_TRY(cudaMemcpy(host_out, pl->returns, sizeof(float) * (size_t)pl->p.n_candidates * pl->p.n_envs, cudaMemcpyDeviceToHost));
  return L2A_OK;
}

extern "C" int l2a_plan_copy_candidates(l2a_ctx* c, l2a_plan* pl, float* host_out) {
  if (!c || !pl || !host_out) return fail(L2A_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(pl->stream));
  const size_t n = (size_t)pl->p.horizon * pl->p.n_candidates * pl->p.n_envs * pl->A;
  CUDA_TRY(cudaMemcpy(host_out, pl->actions, sizeof(float) * n, cudaMemcpyDeviceToHost));
  return L2A_OK;
}

// --------------------------------------------------------------------------------------------- predict
extern "C" int l2a_predict(l2a_ctx* c, l2a_model* m, int set_mode, int first_set, int n_sets, const float* obs,
                           const float* act, int n, float* delta_out, float* next_out, int kernel, void* stream) {
  if (!c || !m || !obs || !act) return fail(L2A_ERR_INVALID, "NULL argument");
  if (!delta_out && !next_out) return fail(L2A_ERR_INVALID, "both outputs are NULL");
  if (!m->norm_set) return fail(L2A_ERR_INVALID, "normalization not set");
  if (n < 1) return fail(L2A_ERR_INVALID, "n must be >= 1");
  if (kernel == L2A_KERNEL_TCGEN05) return fail(L2A_ERR_UNSUPPORTED, "one-step predict runs on the SIMT kernel");
  int groups = 1, rows = n, last_set = first_set;
  if (set_mode == L2A_SETS_PER_ENV) {
    if (n_sets < 1 || n % n_sets != 0) return fail(L2A_ERR_INVALID, "n=%d is not divisible by n_sets=%d", n, n_sets);
    groups = n_sets;
    rows = n / n_sets;
    last_set = first_set + n_sets - 1;
  } else if (set_mode == L2A_SETS_ENSEMBLE_MEAN) {
    if (n_sets < 1) return fail(L2A_ERR_INVALID, "n_sets must be >= 1");
    last_set = first_set + n_sets - 1;
  } else if (set_mode != L2A_SETS_SHARED) return fail(L2A_ERR_INVALID, "set_mode %d", set_mode);
  if (first_set < 0 || last_set >= m->desc.n_sets) return fail(L2A_ERR_INVALID, "weight sets [%d,%d] out of range", first_set, last_set);
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  SimtArgs sa;
  memset(&sa, 0, sizeof(sa));
  sa.dims = m->dims;
  sa.norm = m->norm_dev();
  sa.params = m->params;
  sa.obs = obs;
  sa.actions = act;
  sa.act_stride_t = 0;
  sa.act_stride_row = m->dims.act_dim;
  sa.rows_per_group = rows;
  sa.n_groups = groups;
  sa.horizon = 1;
  sa.set_mode = set_mode;
  sa.first_set = first_set;
  sa.n_sets = n_sets;
  sa.delta_out = delta_out;
  sa.next_out = next_out;
  const size_t smem = simt_smem_bytes(m->dims);
  if ((int)smem > c->max_smem_optin) return fail(L2A_ERR_UNSUPPORTED, "predict needs %zu B shared memory (> %d)", smem, c->max_smem_optin);
  CUDA_TRY(cudaFuncSetAttribute(rollout_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (rows + kSimtRT - 1) / kSimtRT;
  rollout_simt_kernel<true><<<(unsigned)(tiles * groups), kSimtThreads, smem, st>>>(sa);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

// --------------------------------------------------------------------------------------------- adapt
static int adapt_impl(l2a_ctx* c, l2a_model* m, const float* x, const float* target, int K, int M, float inner_lr,
                      int src_set, int dst_first_set, void* stream) {
  if (!c || !m || !x || !target) return fail(L2A_ERR_INVALID, "NULL argument");
  if (K < 1 || M < 1 || M > kAdaptMaxM) return fail(L2A_ERR_INVALID, "K=%d, M=%d: need K >= 1 and 1 <= M <= %d", K, M, kAdaptMaxM);
  if (src_set < 0 || src_set >= m->desc.n_sets || dst_first_set < 0 || dst_first_set + K > m->desc.n_sets)
    return fail(L2A_ERR_INVALID, "src_set %d / dst sets [%d,%d) out of range [0,%d)", src_set, dst_first_set, dst_first_set + K, m->desc.n_sets);
  if (src_set >= dst_first_set && src_set < dst_first_set + K) return fail(L2A_ERR_INVALID, "src_set inside the destination range");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const MlpDims& md = m->dims;
  AdaptArgs aa;
  memset(&aa, 0, sizeof(aa));
  aa.dims = md;
  int ao = 0, go = 0;
  // every h_l / g_l block (and every task's block) starts on a 16-byte boundary: the kernels read them with float4 loads
  // (an odd obs_dim such as Ant's 41 with M not a multiple of 4 would otherwise misalign task k >= 1)
  for (int l = 0; l < md.n_layers; ++l) {
    aa.act_off[l] = ao;
    ao = (ao + md.dims[l] * M + 3) & ~3;
    aa.grad_off[l] = go;
    go = (go + md.dims[l + 1] * M + 3) & ~3;
  }
  aa.act_off[md.n_layers] = ao;
  aa.grad_off[md.n_layers] = go;
  const size_t need_a = (size_t)K * ao, need_g = (size_t)K * go;
  if (need_a > c->adapt_acts_cap) {
    cudaFree(c->adapt_acts);
    c->adapt_acts = nullptr;
    c->adapt_acts_cap = 0;
    CUDA_TRY(cudaMalloc(&c->adapt_acts, need_a * 2 * sizeof(float)));
    c->adapt_acts_cap = need_a * 2;
  }
  if (need_g > c->adapt_grads_cap) {
    cudaFree(c->adapt_grads);
    c->adapt_grads = nullptr;
    c->adapt_grads_cap = 0;
    CUDA_TRY(cudaMalloc(&c->adapt_grads, need_g * 2 * sizeof(float)));
    c->adapt_grads_cap = need_g * 2;
  }
  aa.params = m->params;
  aa.params_out = m->params;
  aa.src_set = src_set;
  aa.dst_first_set = dst_first_set;
  aa.x = x;
  aa.target = target;
  aa.K = K;
  aa.M = M;
  aa.lr = inner_lr;
  aa.acts = c->adapt_acts;
  aa.grads = c->adapt_grads;
  {
    const int mr = (M <= 16) ? 16 : 32;
    const size_t smem = adapt_smem_bytes(md, mr);
    if ((int)smem > c->max_smem_optin) return fail(L2A_ERR_UNSUPPORTED, "adapt needs %zu B shared memory (layer width %d)", smem, md.max_width);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    // 16 CTAs per task (non-portable cluster size: one cluster per GPC) while every task's cluster is resident at once; measured
    // at K = 5, M = 16, 512^3: 134 -> 111 us for the three adapt kernels; K = 10 needs two waves of 16 and is faster with 8
    aa.csize = kAdaptCluster;
    if (K <= 8) {
      // ... if the device can hold K clusters of 16 at once (asked once per context and context-row variant)
      int& cached = c->adapt_max_clusters16[mr == 16 ? 0 : 1];
      if (cached < 0) {
        cudaLaunchConfig_t q;
        memset(&q, 0, sizeof(q));
        q.gridDim = dim3(8 * kAdaptClusterMax);
        q.blockDim = dim3(kAdaptThreads);
        q.dynamicSmemBytes = smem;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = kAdaptClusterMax;
        qa[0].val.clusterDim.y = 1;
        qa[0].val.clusterDim.z = 1;
        q.attrs = qa;
        q.numAttrs = 1;
        int n = 0;
        cudaError_t e1 = (mr == 16) ? cudaFuncSetAttribute(adapt_fwd_bwd_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)
                                    : cudaFuncSetAttribute(adapt_fwd_bwd_kernel<32>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaError_t e2 = (mr == 16) ? cudaFuncSetAttribute(adapt_fwd_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                    : cudaFuncSetAttribute(adapt_fwd_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaError_t e3 = (mr == 16) ? cudaOccupancyMaxActiveClusters(&n, adapt_fwd_bwd_kernel<16>, &q)
                                    : cudaOccupancyMaxActiveClusters(&n, adapt_fwd_bwd_kernel<32>, &q);
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { cudaGetLastError(); n = 0; }
        cached = n;
      }
      if (cached >= K) aa.csize = kAdaptClusterMax;
    }
    cfg.gridDim = dim3((unsigned)(K * aa.csize));
    cfg.blockDim = dim3(kAdaptThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)aa.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (aa.csize > 8) {
      CUDA_TRY(cudaFuncSetAttribute(adapt_fwd_bwd_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      CUDA_TRY(cudaFuncSetAttribute(adapt_fwd_bwd_kernel<32>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    if (mr == 16) {
      CUDA_TRY(cudaFuncSetAttribute(adapt_fwd_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_TRY(cudaLaunchKernelEx(&cfg, adapt_fwd_bwd_kernel<16>, aa));
    } else {
      CUDA_TRY(cudaFuncSetAttribute(adapt_fwd_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_TRY(cudaLaunchKernelEx(&cfg, adapt_fwd_bwd_kernel<32>, aa));
    }
    c->launches++;
  }
  int max_tiles = 0;
  for (int l = 0; l < md.n_layers; ++l)
    max_tiles = std::max(max_tiles, ((md.dims[l + 1] + kUpdTJ - 1) / kUpdTJ) * ((md.dims[l] + kUpdTI - 1) / kUpdTI + 1));
  dim3 grid((unsigned)max_tiles, md.n_layers, K);
  adapt_update_kernel<<<grid, 256, 0, st>>>(aa);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return launch_prep(c, m, dst_first_set, K, st);
}

// --------------------------------------------------------------------------------------------- CEM
extern "C" int l2a_adapt(l2a_ctx* c, l2a_model* m, const float* x, const float* target, int K, int M, float inner_lr,
                         int src_set, int dst_first_set, void* stream) {
  return adapt_impl(c, m, x, target, K, M, inner_lr, src_set, dst_first_set, stream);
}

// ------------------------------------------------------------------------------------------------ f3: adaptation window
struct l2a_window {
  WindowDev dev;
  double* norm = nullptr;      // device copy of the float64 statistics
  float* x = nullptr;          // [n_envs, M, D+A] normalised inputs
  float* target = nullptr;     // [n_envs, M, D] normalised targets
  int M = 0;
  bool norm_set = false;
  std::vector<int> length;     // host mirror of dev.count (push / reset are deterministic)
};

extern "C" int l2a_window_create(l2a_ctx* c, int n_envs, int M, int obs_dim, int act_dim, l2a_window** out) {
  if (!c || !out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (n_envs < 1 || M < 1 || M > kAdaptMaxM || obs_dim < 1 || act_dim < 1)
    return fail(L2A_ERR_INVALID, "n_envs=%d M=%d obs_dim=%d act_dim=%d: need n_envs >= 1, 1 <= M <= %d, dims >= 1", n_envs, M, obs_dim, act_dim, kAdaptMaxM);
  CUDA_TRY(cudaSetDevice(c->device));
  l2a_window* w = new l2a_window();
  memset(&w->dev, 0, sizeof(w->dev));
  w->M = M;
  w->dev.n_envs = n_envs;
  w->dev.cap = M + 1;
  w->dev.D = obs_dim;
  w->dev.A = act_dim;
  w->length.assign(n_envs, 0);
  const size_t pairs = (size_t)n_envs * (M + 1);
  cudaError_t e = cudaMalloc(&w->dev.obs, pairs * obs_dim * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&w->dev.act, pairs * act_dim * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&w->dev.count, n_envs * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&w->norm, (size_t)(4 * obs_dim + 2 * act_dim) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&w->x, (size_t)n_envs * M * (obs_dim + act_dim) * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&w->target, (size_t)n_envs * M * obs_dim * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(w->dev.count, 0, n_envs * sizeof(int));
  if (e != cudaSuccess) {
    l2a_window_destroy(c, w);
    return fail(L2A_ERR_CUDA, "window allocation failed: %s", cudaGetErrorString(e));
  }
  w->dev.norm = w->norm;
  *out = w;
  return L2A_OK;
}

extern "C" int l2a_window_destroy(l2a_ctx* c, l2a_window* w) {
  if (!c || !w) return fail(L2A_ERR_INVALID, "NULL argument");
  cudaSetDevice(c->device);
  cudaFree(w->dev.obs);
  cudaFree(w->dev.act);
  cudaFree(w->dev.count);
  cudaFree(w->norm);
  cudaFree(w->x);
  cudaFree(w->target);
  delete w;
  return L2A_OK;
}

extern "C" int l2a_window_set_normalization(l2a_ctx* c, l2a_window* w, const double* obs_mean, const double* obs_std,
                                            const double* act_mean, const double* act_std, const double* delta_mean,
                                            const double* delta_std, void* stream) {
  if (!c || !w || !obs_mean || !obs_std || !act_mean || !act_std || !delta_mean || !delta_std) return fail(L2A_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int D = w->dev.D, A = w->dev.A;
  std::vector<double> h((size_t)(4 * D + 2 * A));
  memcpy(h.data(), obs_mean, D * sizeof(double));
  memcpy(h.data() + D, obs_std, D * sizeof(double));
  memcpy(h.data() + 2 * D, act_mean, A * sizeof(double));
  memcpy(h.data() + 2 * D + A, act_std, A * sizeof(double));
  memcpy(h.data() + 2 * D + 2 * A, delta_mean, D * sizeof(double));
  memcpy(h.data() + 3 * D + 2 * A, delta_std, D * sizeof(double));
  CUDA_TRY(cudaMemcpyAsync(w->norm, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));     // h goes out of scope
  w->norm_set = true;
  return L2A_OK;
}

extern "C" int l2a_window_push(l2a_ctx* c, l2a_window* w, const double* obs, const double* act, void* stream) {
  if (!c || !w || !obs || !act) return fail(L2A_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  window_push_kernel<<<w->dev.n_envs, 64, 0, (cudaStream_t)stream>>>(w->dev, obs, act, w->dev.A);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  for (int& l : w->length) l += 1;
  return L2A_OK;
}

extern "C" int l2a_window_reset(l2a_ctx* c, l2a_window* w, int env, void* stream) {
  if (!c || !w) return fail(L2A_ERR_INVALID, "NULL argument");
  if (env >= w->dev.n_envs) return fail(L2A_ERR_INVALID, "env %d out of range [0,%d)", env, w->dev.n_envs);
  CUDA_TRY(cudaSetDevice(c->device));
  window_reset_kernel<<<(w->dev.n_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w->dev, env);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  for (int i = 0; i < w->dev.n_envs; ++i)
    if (env < 0 || env == i) w->length[i] = 0;
  return L2A_OK;
}

extern "C" int l2a_window_length(l2a_ctx* c, const l2a_window* w, int env) {
  if (!c || !w) return fail(L2A_ERR_INVALID, "NULL argument");
  if (env < 0 || env >= w->dev.n_envs) return fail(L2A_ERR_INVALID, "env %d out of range [0,%d)", env, w->dev.n_envs);
  return w->length[env];
}

extern "C" int l2a_window_gather(l2a_ctx* c, l2a_window* w, float* x, float* target, void* stream) {
  if (!c || !w || !x || !target) return fail(L2A_ERR_INVALID, "NULL argument");
  if (!w->norm_set) return fail(L2A_ERR_INVALID, "window normalisation statistics are not set");
  for (int i = 0; i < w->dev.n_envs; ++i)
    if (w->length[i] < w->M + 1)
      return fail(L2A_ERR_INVALID, "env %d has %d transitions on its running path, the adapt window needs %d", i, w->length[i], w->M + 1);
  CUDA_TRY(cudaSetDevice(c->device));
  window_gather_kernel<<<w->dev.n_envs, 128, 0, (cudaStream_t)stream>>>(w->dev, w->M, x, target);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  return L2A_OK;
}

extern "C" int l2a_adapt_from_window(l2a_ctx* c, l2a_model* m, l2a_window* w, float inner_lr, int src_set, int dst_first_set,
                                     void* stream) {
  if (!c || !m || !w) return fail(L2A_ERR_INVALID, "NULL argument");
  if (w->dev.D != m->dims.obs_dim || w->dev.A != m->dims.act_dim)
    return fail(L2A_ERR_INVALID, "window dims (%d, %d) do not match the model's (%d, %d)", w->dev.D, w->dev.A, m->dims.obs_dim, m->dims.act_dim);
  const int rc = l2a_window_gather(c, w, w->x, w->target, stream);
  if (rc != L2A_OK) return rc;
  return adapt_impl(c, m, w->x, w->target, w->dev.n_envs, w->M, inner_lr, src_set, dst_first_set, stream);
}

// ---- helpers of the host-buffer planning call (GrBAL step inside the plan's graph)
static bool window_matches(const l2a_window* w, int n_envs, int D, int A) {
  return w && w->dev.n_envs == n_envs && w->dev.D == D && w->dev.A == A;
}
static void window_count_push(l2a_window* w) {
  for (int& l : w->length) l += 1;
}
static int window_gather_own(l2a_ctx* c, l2a_window* w, const float** x_out, const float** target_out, int* K_out, int* M_out, cudaStream_t st) {
  const int rc = l2a_window_gather(c, w, w->x, w->target, st);
  if (rc != L2A_OK) return rc;
  *x_out = w->x;
  *target_out = w->target;
  *K_out = w->dev.n_envs;
  *M_out = w->M;
  return L2A_OK;
}
static int window_push_strided(l2a_ctx* c, l2a_window* w, const double* obs, const double* act, int act_stride, bool count_on_host, cudaStream_t st) {
  window_push_kernel<<<w->dev.n_envs, 64, 0, st>>>(w->dev, obs, act, act_stride);
  CUDA_TRY(cudaGetLastError());
  c->launches++;
  if (count_on_host)
    for (int& l : w->length) l += 1;
  return L2A_OK;
}

extern "C" int l2a_plan_attach_window(l2a_ctx* c, l2a_plan* pl, l2a_window* w, float inner_lr, int src_set, int dst_first_set) {
  if (!c || !pl || !w) return fail(L2A_ERR_INVALID, "NULL argument");
  if (!window_matches(w, pl->p.n_envs, pl->D, pl->A))
    return fail(L2A_ERR_INVALID, "window (envs %d, dims %d/%d) does not match the plan (envs %d, dims %d/%d)", w->dev.n_envs, w->dev.D,
                w->dev.A, pl->p.n_envs, pl->D, pl->A);
  pl->window = w;
  pl->win_lr = inner_lr;
  pl->win_src_set = src_set;
  pl->win_dst_first_set = dst_first_set;
  for (int f = 1; f < 4; ++f)
    if (pl->exec_flags[f]) { cudaGraphExecDestroy(pl->exec_flags[f]); pl->exec_flags[f] = nullptr; }
  return L2A_OK;
}

extern "C" int l2a_cem_sample(l2a_ctx* c, const float* z, const double* mean, const double* std_, const float* clip_low,
                              const float* clip_high, int n, int m, int ha, float* samples, float* clipped, void* stream) {
  if (!c || !z || !mean || !std_ || !clip_low || !clip_high || !samples || !clipped) return fail(L2A_ERR_INVALID, "NULL argument");
  if (n < 1 || m < 1 || ha < 1) return fail(L2A_ERR_INVALID, "n/m/ha must be >= 1");
  CUDA_TRY(cudaSetDevice(c->device));
  const long long total = (long long)n * m * ha;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)c->num_sms * 8);
  cem_sample_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(z, mean, std_, clip_low, clip_high, n, m, ha, samples, clipped);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

extern "C" int l2a_cem_refit(l2a_ctx* c, const float* returns, const float* clipped, int n, int m, int ha, int num_elites,
                             double alpha, int compat, int32_t* rank_scratch, double* mean, double* std_, void* stream) {
  if (!c || !returns || !clipped || !rank_scratch || !mean || !std_) return fail(L2A_ERR_INVALID, "NULL argument");
  if (n < 1 || m < 1 || ha < 1 || num_elites < 1 || num_elites > n) return fail(L2A_ERR_INVALID, "bad n/m/ha/num_elites");
  if (!compat && m > 1)
    return fail(L2A_ERR_UNSUPPORTED, "corrected CEM (compat = 0) is defined for one env per call: with m > 1 the reference's sample -> env "
                                     "layout is itself inconsistent (mpc_controller.py:85-102), only the bug-compatible mode reproduces it");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g1((n + 255) / 256, m, kRankSplit);
  CUDA_TRY(cudaMemsetAsync(rank_scratch, 0, sizeof(int32_t) * (size_t)n * m, st));
  cem_rank_kernel<<<g1, 256, 0, st>>>(returns, n, rank_scratch);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  cem_refit_kernel<float><<<ha, 256, 0, st>>>(rank_scratch, clipped, n, m, ha, num_elites, alpha, compat, mean, std_);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

#ifdef L2A_DEBUG_KERNELS
// --------------------------------------------------------------------------------------------- diagnostics
extern "C" int l2a_debug_umma_tile(l2a_ctx* c, const float* A, const float* B, float* C, int n, int k, int variant, void* stream) {
  if (!c || !A || !B || !C) return fail(L2A_ERR_INVALID, "NULL argument");
  if (n < 16 || n > 128 || n % 16 != 0 || k < 64 || k > 256 || k % 64 != 0) return fail(L2A_ERR_INVALID, "need n in {16..128} %% 16, k in {64..256} %% 64");
  CUDA_TRY(cudaSetDevice(c->device));
  const int nkc = k / 64;
  const size_t smem = (size_t)2 * nkc * 16384 + (size_t)2 * nkc * n * 128 + 64 + 1024;
  CUDA_TRY(cudaFuncSetAttribute(debug_umma_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  debug_umma_tile_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, C, n, k, variant);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

extern "C" int l2a_debug_stream(l2a_ctx* c, const void* blob, int n_tiles_per_pass, int passes, int stages, int tile_bytes,
                                int hold_cycles, int producers, int consumers, int grid, long long* cycles_out, void* stream) {
  if (!c || !blob || !cycles_out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (stages < 1 || stages > 13 || tile_bytes % 1024 != 0 || grid < 1) return fail(L2A_ERR_INVALID, "bad stages/tile_bytes/grid");
  if (producers < 1 || producers > 2 || consumers < 1 || consumers > 2) return fail(L2A_ERR_INVALID, "producers/consumers must be 1 or 2");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t smem = (size_t)stages * tile_bytes + 2 * stages * sizeof(uint64_t) + 1024 + 64;
  if ((int)smem > c->max_smem_optin) return fail(L2A_ERR_UNSUPPORTED, "needs %zu B shared memory", smem);
  CUDA_TRY(cudaFuncSetAttribute(debug_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  debug_stream_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>((const uint8_t*)blob, n_tiles_per_pass, passes, stages, tile_bytes,
                                                                hold_cycles, producers, consumers, cycles_out);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

extern "C" int l2a_debug_set_timeline(l2a_ctx* c, long long* buf128) {
  if (!c) return fail(L2A_ERR_INVALID, "NULL ctx");
  c->timeline = buf128;
  c->ws_epoch++;
  return L2A_OK;
}

#endif  // L2A_DEBUG_KERNELS

// --------------------------------------------------------------------------------------------- K3 shard glue
extern "C" int l2a_shard_pack(l2a_ctx* c, const float* best_ret, const int32_t* best_idx, const float* best_act,
                              int64_t idx_offset, int m, int A, float* packed, void* stream) {
  if (!c || !best_ret || !best_idx || !best_act || !packed) return fail(L2A_ERR_INVALID, "NULL argument");
  if (m < 1 || A < 1) return fail(L2A_ERR_INVALID, "m and A must be >= 1");
  CUDA_TRY(cudaSetDevice(c->device));
  shard_pack_kernel<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(best_ret, best_idx, best_act, idx_offset, m, A, packed);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

extern "C" int l2a_shard_select(l2a_ctx* c, const float* gathered, int G, int m, int A, float* best_ret, int64_t* best_idx,
                                float* best_act, void* stream) {
  if (!c || !gathered || !best_ret || !best_idx || !best_act) return fail(L2A_ERR_INVALID, "NULL argument");
  if (G < 1 || m < 1 || A < 1) return fail(L2A_ERR_INVALID, "G, m and A must be >= 1");
  CUDA_TRY(cudaSetDevice(c->device));
  shard_select_kernel<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gathered, G, m, A, best_ret, (long long*)best_idx, best_act);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

#ifdef L2A_DEBUG_KERNELS
extern "C" int l2a_debug_pair(l2a_ctx* c, int mode, int iters, int copy_bytes, long long* cycles_out, void* stream) {
  if (!c || !cycles_out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (mode < 0 || mode > 3 || iters < 1 || copy_bytes < 16 || copy_bytes > 16384 || copy_bytes % 16 != 0)
    return fail(L2A_ERR_INVALID, "bad mode/iters/copy_bytes");
  if ((long long)iters * copy_bytes >= (1 << 20)) return fail(L2A_ERR_INVALID, "iters * copy_bytes must stay below the mbarrier tx-count range (1 MiB)");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t smem = 4 * 16384 + 2 * 80 * 128 + 32768 + 64;
  CUDA_TRY(cudaFuncSetAttribute(debug_pair_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, debug_pair_kernel<80>, mode, iters, copy_bytes, cycles_out));
  c->launches++;
  return L2A_OK;
}

extern "C" int l2a_debug_mma_rate(l2a_ctx* c, int nc, int mode, int iters, long long* cycles_out, void* stream) {
  if (!c || !cycles_out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (mode < 0 || mode > 15 || iters < 1) return fail(L2A_ERR_INVALID, "bad mode/iters");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t smem = 4 * 16384 + 2 * (size_t)nc * 128 + 64 + ((mode == 9 || mode == 10) ? 2 * 32768 + 2048 : 0);
  if (nc == 80) {
    CUDA_TRY(cudaFuncSetAttribute(debug_mma_rate_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    debug_mma_rate_kernel<80><<<1, 128, smem, (cudaStream_t)stream>>>(mode, iters, cycles_out);
  } else if (nc == 64) {
    CUDA_TRY(cudaFuncSetAttribute(debug_mma_rate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    debug_mma_rate_kernel<64><<<1, 128, smem, (cudaStream_t)stream>>>(mode, iters, cycles_out);
  } else if (nc == 128) {
    CUDA_TRY(cudaFuncSetAttribute(debug_mma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    debug_mma_rate_kernel<128><<<1, 128, smem, (cudaStream_t)stream>>>(mode, iters, cycles_out);
  } else {
    return fail(L2A_ERR_INVALID, "nc must be 64, 80 or 128");
  }
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

#endif  // L2A_DEBUG_KERNELS

// --------------------------------------------------------------------------------------------- ReBAL (LSTM) model
struct l2a_rnn_model {
  RnnDims dims;
  RnnTcPlan plan;
  bool tc_ok = false;
  uint8_t* blob = nullptr;     // tensor-core tile pairs (rnn_tc_prep_kernel)
  float* params = nullptr;
  float* norm = nullptr;
  bool norm_set = false;
  NormDev norm_dev() const {
    const int D = dims.obs_dim, A = dims.act_dim;
    NormDev n;
    n.obs_mean = norm; n.obs_den = norm + D; n.act_mean = norm + 2 * D; n.act_den = norm + 2 * D + A;
    n.delta_mean = norm + 2 * D + 2 * A; n.delta_scale = norm + 3 * D + 2 * A;
    return n;
  }
};

extern "C" int l2a_rnn_model_create(l2a_ctx* c, int obs_dim, int act_dim, int hidden, l2a_rnn_model** out) {
  if (!c || !out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (obs_dim < 3 || act_dim < 1 || hidden < 1) return fail(L2A_ERR_INVALID, "bad obs_dim/act_dim/hidden");
  CUDA_TRY(cudaSetDevice(c->device));
  l2a_rnn_model* m = new (std::nothrow) l2a_rnn_model();
  if (!m) return fail(L2A_ERR_INVALID, "out of host memory");
  RnnDims& rd = m->dims;
  rd.obs_dim = obs_dim; rd.act_dim = act_dim; rd.hidden = hidden;
  int off = 0;
  rd.wk_off = off; off += (obs_dim + act_dim + hidden) * 4 * hidden;
  rd.bk_off = off; off += 4 * hidden;
  rd.wo_off = off; off += hidden * obs_dim;
  rd.bo_off = off; off += obs_dim;
  rd.total = off;
  if (rnn_smem_bytes(rd) > (size_t)c->max_smem_optin) { delete m; return fail(L2A_ERR_UNSUPPORTED, "LSTM width %d needs %zu B shared memory", hidden, rnn_smem_bytes(rd)); }
  if (cudaMalloc(&m->params, sizeof(float) * (size_t)off) != cudaSuccess) { delete m; return fail(L2A_ERR_CUDA, "cudaMalloc(rnn params)"); }
  if (cudaMalloc(&m->norm, sizeof(float) * (size_t)(4 * obs_dim + 2 * act_dim)) != cudaSuccess) { cudaFree(m->params); delete m; return fail(L2A_ERR_CUDA, "cudaMalloc(rnn norm)"); }
  m->tc_ok = rnn_tc_make_plan(obs_dim, act_dim, hidden, &m->plan);
  if (m->tc_ok && cudaMalloc(&m->blob, (size_t)m->plan.pairs_per_set * 2 * kTcTileBytes) != cudaSuccess) {
    cudaFree(m->params); cudaFree(m->norm); delete m; return fail(L2A_ERR_CUDA, "cudaMalloc(rnn blob)");
  }
  *out = m;
  return L2A_OK;
}

extern "C" int l2a_rnn_model_destroy(l2a_ctx* c, l2a_rnn_model* m) {
  if (!m) return L2A_OK;
  if (c) cudaSetDevice(c->device);
  cudaFree(m->params);
  cudaFree(m->norm);
  cudaFree(m->blob);
  delete m;
  return L2A_OK;
}

extern "C" int l2a_rnn_model_set_params(l2a_ctx* c, l2a_rnn_model* m, const float* cell_kernel, const float* cell_bias,
                                        const float* out_kernel, const float* out_bias, void* stream) {
  if (!c || !m || !cell_kernel || !cell_bias || !out_kernel || !out_bias) return fail(L2A_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const RnnDims& rd = m->dims;
  CUDA_TRY(cudaMemcpyAsync(m->params + rd.wk_off, cell_kernel, sizeof(float) * (size_t)(rd.obs_dim + rd.act_dim + rd.hidden) * 4 * rd.hidden, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(m->params + rd.bk_off, cell_bias, sizeof(float) * (size_t)4 * rd.hidden, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(m->params + rd.wo_off, out_kernel, sizeof(float) * (size_t)rd.hidden * rd.obs_dim, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(m->params + rd.bo_off, out_bias, sizeof(float) * (size_t)rd.obs_dim, cudaMemcpyDeviceToDevice, st));
  if (m->tc_ok) {
    RnnPrepArgs pa;
    pa.plan = m->plan;
    pa.D = rd.obs_dim;
    pa.A = rd.act_dim;
    pa.wk = m->params + rd.wk_off;
    pa.wo = m->params + rd.wo_off;
    pa.blob = m->blob;
    rnn_tc_prep_kernel<<<m->plan.pairs_per_set, 256, 0, st>>>(pa);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  return L2A_OK;
}

extern "C" int l2a_rnn_model_set_normalization(l2a_ctx* c, l2a_rnn_model* m, const float* obs_mean, const float* obs_den,
                                               const float* act_mean, const float* act_den, const float* delta_mean,
                                               const float* delta_scale, void* stream) {
  if (!c || !m || !obs_mean || !obs_den || !act_mean || !act_den || !delta_mean || !delta_scale) return fail(L2A_ERR_INVALID, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int D = m->dims.obs_dim, A = m->dims.act_dim;
  NormDev n = m->norm_dev();
  CUDA_TRY(cudaMemcpyAsync((void*)n.obs_mean, obs_mean, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync((void*)n.obs_den, obs_den, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync((void*)n.act_mean, act_mean, sizeof(float) * A, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync((void*)n.act_den, act_den, sizeof(float) * A, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync((void*)n.delta_mean, delta_mean, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync((void*)n.delta_scale, delta_scale, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
  m->norm_set = true;
  return L2A_OK;
}

extern "C" int l2a_rnn_rollout(l2a_ctx* c, l2a_rnn_model* m, const l2a_rollout_params* p, const float* obs0, const float* hidden_c,
                               const float* hidden_h, const float* actions, const float* discount_pow, float* returns,
                               float* best_ret, int32_t* best_idx, float* best_act, void* stream) {
  if (!c || !m || !p || !obs0 || !hidden_c || !hidden_h || !actions || !discount_pow || !best_ret || !best_idx || !best_act)
    return fail(L2A_ERR_INVALID, "NULL argument");
  if (!m->norm_set) return fail(L2A_ERR_INVALID, "normalization not set");
  if (p->n_candidates < 1 || p->n_envs < 1 || p->horizon < 1) return fail(L2A_ERR_INVALID, "n_candidates/n_envs/horizon must be >= 1");
  if (p->reward_kind < 0 || p->reward_kind > 2 || !(p->dt > 0.f)) return fail(L2A_ERR_INVALID, "bad reward_kind/dt");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  int kernel = p->kernel;
  if (kernel == L2A_KERNEL_AUTO) kernel = m->tc_ok ? L2A_KERNEL_TCGEN05 : L2A_KERNEL_SIMT;
  if (kernel == L2A_KERNEL_TCGEN05) {
    if (!m->tc_ok) return fail(L2A_ERR_UNSUPPORTED, "tcgen05 LSTM rollout needs hidden in {128, 256}, obs_dim <= 48, pad8(obs)+act <= 64");
    constexpr int NC = 64;
    const int groups = (p->n_candidates + NC - 1) / NC;
    int rc2 = ensure_reduce_ws(c, (size_t)groups * p->n_envs, p->n_envs, st);
    if (rc2) return rc2;
    RnnTcArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.plan = m->plan;
    ta.D = m->dims.obs_dim; ta.A = m->dims.act_dim;
    ta.norm = m->norm_dev();
    ta.bk = m->params + m->dims.bk_off; ta.bo = m->params + m->dims.bo_off;
    ta.blob = m->blob;
    ta.obs0 = obs0; ta.c0 = hidden_c; ta.h0 = hidden_h;
    ta.actions = actions; ta.act_stride_t = p->act_stride_t; ta.act_stride_row = p->act_stride_row;
    ta.discount_pow = discount_pow;
    ta.n_candidates = p->n_candidates; ta.n_envs = p->n_envs; ta.horizon = p->horizon; ta.reward_kind = p->reward_kind; ta.dt = p->dt;
    ta.groups_per_env = groups;
    ta.returns = returns;
    ta.red.part_ret = c->part_ret; ta.red.part_idx = c->part_idx; ta.red.counters = c->counters;
    ta.red.best_ret = best_ret; ta.red.best_idx = best_idx; ta.red.best_act = best_act;
    ta.red.actions = actions; ta.red.act_stride_row = p->act_stride_row; ta.red.act_dim = m->dims.act_dim;
    ta.red.n_candidates = p->n_candidates; ta.red.tiles_per_env = groups;
    const size_t smem_tc = RnnTcSmem<NC>::total(m->dims.obs_dim, m->dims.act_dim, m->dims.hidden);
    if ((int)smem_tc > c->max_smem_optin) return fail(L2A_ERR_UNSUPPORTED, "tcgen05 LSTM rollout needs %zu B shared memory", smem_tc);
    CUDA_TRY(cudaFuncSetAttribute(rollout_rnn_tc_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
    rollout_rnn_tc_kernel<NC><<<(unsigned)(groups * p->n_envs), kRnnTcThreads, smem_tc, st>>>(ta);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return L2A_OK;
  }
  const int tiles = (p->n_candidates + kRnnRT - 1) / kRnnRT;
  int rc = ensure_reduce_ws(c, (size_t)tiles * p->n_envs, p->n_envs, st);
  if (rc) return rc;
  RnnArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.dims = m->dims;
  ra.norm = m->norm_dev();
  ra.params = m->params;
  ra.obs = obs0; ra.c0 = hidden_c; ra.h0 = hidden_h;
  ra.actions = actions;
  ra.act_stride_t = p->act_stride_t; ra.act_stride_row = p->act_stride_row;
  ra.discount_pow = discount_pow;
  ra.rows_per_group = p->n_candidates; ra.n_groups = p->n_envs; ra.horizon = p->horizon;
  ra.reward_kind = p->reward_kind; ra.dt = p->dt;
  ra.returns = returns;
  ra.red.part_ret = c->part_ret; ra.red.part_idx = c->part_idx; ra.red.counters = c->counters;
  ra.red.best_ret = best_ret; ra.red.best_idx = best_idx; ra.red.best_act = best_act;
  ra.red.actions = actions; ra.red.act_stride_row = p->act_stride_row; ra.red.act_dim = m->dims.act_dim;
  ra.red.n_candidates = p->n_candidates; ra.red.tiles_per_env = tiles;
  const size_t smem = rnn_smem_bytes(m->dims);
  CUDA_TRY(cudaFuncSetAttribute(rollout_rnn_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rollout_rnn_simt_kernel<false><<<(unsigned)(tiles * p->n_envs), kRnnThreads, smem, st>>>(ra);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}

extern "C" int l2a_rnn_predict(l2a_ctx* c, l2a_rnn_model* m, const float* obs, const float* act, const float* hidden_c,
                               const float* hidden_h, int n, float* delta_out, float* c_out, float* h_out, void* stream) {
  if (!c || !m || !obs || !act || !hidden_c || !hidden_h || !delta_out || !c_out || !h_out) return fail(L2A_ERR_INVALID, "NULL argument");
  if (!m->norm_set) return fail(L2A_ERR_INVALID, "normalization not set");
  if (n < 1) return fail(L2A_ERR_INVALID, "n must be >= 1");
  CUDA_TRY(cudaSetDevice(c->device));
  RnnArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.dims = m->dims;
  ra.norm = m->norm_dev();
  ra.params = m->params;
  ra.obs = obs; ra.c0 = hidden_c; ra.h0 = hidden_h;
  ra.actions = act; ra.act_stride_t = 0; ra.act_stride_row = m->dims.act_dim;
  ra.rows_per_group = n; ra.n_groups = 1; ra.horizon = 1;
  ra.delta_out = delta_out; ra.c_out = c_out; ra.h_out = h_out;
  const size_t smem = rnn_smem_bytes(m->dims);
  CUDA_TRY(cudaFuncSetAttribute(rollout_rnn_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rollout_rnn_simt_kernel<true><<<(unsigned)((n + kRnnRT - 1) / kRnnRT), kRnnThreads, smem, (cudaStream_t)stream>>>(ra);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return L2A_OK;
}
