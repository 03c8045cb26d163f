// Error reporting, device queries, gather (pure copy) and small utility kernels.
#include <stdarg.h>

#include "common.cuh"

namespace rs {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}
}  // namespace rs

RS_API int rs_version(void) { return 100; }
RS_API const char *rs_last_error(void) { return rs::g_err; }

// ------------------------------------------------------------------ gather
struct GatherParams {
  const float *base[RS_MAX_FIELDS];
  int64_t rows[RS_MAX_FIELDS];
};

// One 16-byte element per thread-iteration; consecutive threads walk one row, so a warp reads
// (32 / (W/4)) rows of one sample at a time with fully coalesced 128-bit loads.
__global__ void __launch_bounds__(256) gather_rows_kernel(const __grid_constant__ GatherParams P, const int64_t *__restrict__ ids,
                                                         int64_t n_lookups, int F, int wv /* float4 per row */,
                                                         float *__restrict__ out, int32_t *status) {
  __shared__ const float *s_base[RS_MAX_FIELDS];
  __shared__ int64_t s_rows[RS_MAX_FIELDS];
  for (int i = threadIdx.x; i < F; i += blockDim.x) {
    s_base[i] = P.base[i];
    s_rows[i] = P.rows[i];
  }
  __syncthreads();
  int64_t total = n_lookups * wv;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = e / wv;
    int v = (int)(e - p * wv);
    int f = (int)(p % F);
    int64_t id = rs::clamp_id(ids[p], s_rows[f], status);
    float4 val = rs::ldg_nc_f4(s_base[f] + (id * wv + v) * 4);
    rs::stg_f4(out + e * 4, val);
  }
}

__global__ void gather_rows_scalar_kernel(const __grid_constant__ GatherParams P, const int64_t *__restrict__ ids, int64_t n_lookups,
                                          int F, int W, float *__restrict__ out, int32_t *status) {
  int64_t total = n_lookups * W;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = e / W;
    int v = (int)(e - p * W);
    int f = (int)(p % F);
    int64_t id = rs::clamp_id(ids[p], P.rows[f], status);
    out[e] = P.base[f][id * W + v];
  }
}

RS_API int rs_gather_rows(const rs_tables *T, const int64_t *ids, int64_t B, float *out, int32_t *status, void *stream) {
  RS_CHECK_ARG(T && ids && out, RS_E_ARG, "rs_gather_rows: null argument");
  RS_CHECK_ARG(T->num_fields >= 1 && T->num_fields <= RS_MAX_FIELDS && T->width >= 1, RS_E_SHAPE, "rs_gather_rows: bad F/width");
  if (B == 0) return RS_OK;
  GatherParams P;
  for (int f = 0; f < T->num_fields; ++f) {
    P.base[f] = T->base[f];
    P.rows[f] = T->rows[f];
  }
  int64_t n = B * T->num_fields;
  cudaStream_t st = (cudaStream_t)stream;
  if (T->width % 4 == 0) {
    int wv = T->width / 4;
    int64_t total = n * wv;
    int blocks = (int)((total + 255) / 256);
    int cap = rs::num_sms() * 16;
    if (blocks > cap) blocks = cap;
    gather_rows_kernel<<<blocks, 256, 0, st>>>(P, ids, n, T->num_fields, wv, out, status);
  } else {
    int64_t total = n * T->width;
    int blocks = (int)((total + 255) / 256);
    int cap = rs::num_sms() * 16;
    if (blocks > cap) blocks = cap;
    gather_rows_scalar_kernel<<<blocks, 256, 0, st>>>(P, ids, n, T->num_fields, T->width, out, status);
  }
  RS_CHECK_LAUNCH();
  return RS_OK;
}

// ------------------------------------------------------------------ gather fused with the row all-to-all
namespace rs {
int fill_routes(Routes &R, const rs_routes *r, const char *who) {
  RS_CHECK_ARG(r && r->n >= 1 && r->n <= RS_MAX_RANKS, RS_E_ARG, "%s: bad routes", who);
  R.n = r->n;
  for (int k = 0; k < r->n; ++k) {
    R.start[k] = r->start[k];
    R.base[k] = r->base[k];
    R.row0[k] = r->row0[k];
  }
  R.start[r->n] = r->start[r->n];
  R.dyn_start = r->dyn_start;
  R.dyn_row0 = r->dyn_row0;
  R.cap_rows = r->cap_rows;
  R.self = r->self;
  RS_CHECK_ARG(!r->dyn_start || r->dyn_row0, RS_E_ARG, "%s: dyn_start without dyn_row0", who);
  return RS_OK;
}
}  // namespace rs

// Thread e copies one 16-byte piece of one requested row into the requester's block through its peer-mapped
// pointer: coalesced 128-bit loads from the local shard, coalesced 128-bit stores over NVLink.
__global__ void __launch_bounds__(256) gather_rows_peer_kernel(const float *__restrict__ table, int64_t rows, int wv,
                                                              const int64_t *__restrict__ idx, int64_t m,
                                                              const __grid_constant__ rs::Routes R, int32_t *status) {
  // Remote stores are latency bound, so every thread keeps four independent 16-byte pieces in flight: the four
  // loads are issued before the four stores.
  __shared__ int64_t rtab[rs::ROUTE_TAB];
  rs::route_tab_load(rtab, R);
  __syncthreads();
  const int64_t total = m * wv;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < total; e0 += 4 * stride) {
    float4 val[4];
    float *dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t e = e0 + u * stride;
      dst[u] = nullptr;
      if (e < total) {
        const int64_t i = e / wv;
        const int v = (int)(e - i * wv);
        const int64_t id = rs::clamp_id(idx[i], rows, status);
        val[u] = rs::ldg_nc_f4(table + (id * wv + v) * 4);
        dst[u] = rs::route_row(R, rtab, i, wv * 4);
        if (dst[u]) dst[u] += v * 4;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (dst[u]) rs::stg_f4(dst[u], val[u]);
  }
}

RS_API int rs_gather_rows_peer(const float *table, int64_t rows, int32_t width, const int64_t *idx, int64_t m, const rs_routes *routes,
                               int32_t *status, void *stream) {
  RS_CHECK_ARG(table && rows > 0 && width >= 4 && width % 4 == 0, RS_E_ARG, "rs_gather_rows_peer: bad table/width");
  rs::Routes R;
  int rc = rs::fill_routes(R, routes, "rs_gather_rows_peer");
  if (rc) return rc;
  if (m == 0) return RS_OK;
  RS_CHECK_ARG(idx, RS_E_ARG, "rs_gather_rows_peer: null idx");
  const int wv = width / 4;
  int64_t blocks64 = (m * wv + 255) / 256;
  int cap = rs::num_sms() * 8;
  gather_rows_peer_kernel<<<(int)(blocks64 < cap ? blocks64 : cap), 256, 0, (cudaStream_t)stream>>>(table, rows, wv, idx, m, R, status);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

// ------------------------------------------------------------------ x[:, col].long()
__global__ void xcol_to_ids_kernel(const float *__restrict__ x, int64_t B, int xcols, int col, int64_t *__restrict__ ids) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) ids[b] = (int64_t)x[b * xcols + col];  // truncation toward zero == Tensor.long()
}
RS_API int rs_xcol_to_ids(const float *x, int64_t B, int32_t xcols, int32_t col, int64_t *ids, void *stream) {
  RS_CHECK_ARG(x && ids && col >= 0 && col < xcols, RS_E_ARG, "rs_xcol_to_ids: bad argument");
  if (B == 0) return RS_OK;
  xcol_to_ids_kernel<<<(int)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, B, xcols, col, ids);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

// ------------------------------------------------------------------ sigmoid + BCE (mean) fwd/bwd
// pass 1: per-block partial sums in a fixed tree order; pass 2: one warp adds the partials (lane-strided, then a fixed
// shuffle tree): deterministic.  Optional: the logit is cross[b] + *bias (the models' `cross + self.bias`), and the sum of
// d loss / d logit (= the bias gradient) is reduced alongside the loss.
__global__ void __launch_bounds__(256) sigmoid_bce_kernel(const float *__restrict__ logit, const float *__restrict__ bias,
                                                         const float *__restrict__ y, int64_t B, float *__restrict__ pred,
                                                         float *__restrict__ g_logit, float *__restrict__ partial, int want_gsum) {
  __shared__ float red[8], redg[8];
  float acc = 0.f, accg = 0.f;
  const float invB = 1.0f / (float)B;
  const float bz = bias ? __ldg(bias) : 0.f;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    float z = bias ? logit[b] + bz : logit[b];
    float p = 1.0f / (1.0f + expf(-z));
    float t = y[b];
    float lp = fmaxf(logf(p), -100.f);
    float l1p = fmaxf(logf(1.0f - p), -100.f);
    acc -= t * lp + (1.0f - t) * l1p;
    if (pred) pred[b] = p;
    // autograd's exact sequence (binary_cross_entropy_backward then sigmoid_backward), so that saturated
    // probabilities (p == 0 or 1 in fp32) give the same gradient as the reference, not the analytic (p-t)/B
    const float g = (invB * (p - t) / fmaxf((1.0f - p) * p, 1e-12f)) * ((1.0f - p) * p);
    if (g_logit) g_logit[b] = g;
    accg += g;
  }
  acc = rs::warp_sum(acc);
  accg = rs::warp_sum(accg);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc, redg[threadIdx.x >> 5] = accg;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, sg = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i], sg += redg[i];
    partial[blockIdx.x] = s;
    if (want_gsum) partial[1024 + blockIdx.x] = sg;
  }
}
__global__ void __launch_bounds__(32) bce_finish_kernel(const float *__restrict__ partial, int nparts, int64_t B, float *__restrict__ loss,
                                                       float *__restrict__ g_sum) {
  float s = 0.f, sg = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 32) {
    s += partial[i];
    if (g_sum) sg += partial[1024 + i];
  }
  s = rs::warp_sum(s);
  sg = rs::warp_sum(sg);
  if (threadIdx.x == 0) {
    *loss = s / (float)B;
    if (g_sum) *g_sum = sg;
  }
}
RS_API int rs_sigmoid_bce_bias(const float *cross, const float *bias, const float *y, int64_t B, float *pred, float *g_logit,
                               float *loss_mean, float *g_sum, float *ws, void *stream) {
  RS_CHECK_ARG(cross && y && loss_mean && ws && B > 0, RS_E_ARG, "rs_sigmoid_bce: bad argument");
  int blocks = (int)((B + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  cudaStream_t st = (cudaStream_t)stream;
  sigmoid_bce_kernel<<<blocks, 256, 0, st>>>(cross, bias, y, B, pred, g_logit, ws, g_sum ? 1 : 0);
  RS_CHECK_LAUNCH();
  bce_finish_kernel<<<1, 32, 0, st>>>(ws, blocks, B, loss_mean, g_sum);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
RS_API int rs_sigmoid_bce(const float *logit, const float *y, int64_t B, float *pred, float *g_logit, float *loss_mean, float *ws,
                          void *stream) {
  return rs_sigmoid_bce_bias(logit, nullptr, y, B, pred, g_logit, loss_mean, nullptr, ws, stream);
}
