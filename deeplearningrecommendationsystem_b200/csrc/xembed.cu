// MovieLens feature-vector front end: x (B, xcols) f32 -> per-slot embeddings E (B, slots, W).
// Replaces (reference) the mixed nn.Embedding / one-hot-matmul "lookups" at model/deepfm.py:45-51,
// model/ffm.py:48-59, model/afm.py:43-48, model/nfm.py:45-51, model/pnn.py:113-118.
#include "common.cuh"

namespace {

struct XParams {
  int32_t num_slots, W, xcols;
  int32_t col[RS_MAX_FIELDS], ncols[RS_MAX_FIELDS], kind[RS_MAX_FIELDS];
  const float *table[RS_MAX_FIELDS];
  int64_t rows[RS_MAX_FIELDS];
  int32_t bag_row0[RS_MAX_FIELDS];  // first row of this bag slot inside the concatenated partial buffer
  int32_t bag_rows_total;
};

constexpr int XW = 8;  // warps per block

// warp per sample; lanes stride the W columns.  Bag sums visit k ascending and skip x == 0 terms.
__global__ void __launch_bounds__(XW * 32) xembed_fwd_kernel(const __grid_constant__ XParams P, const float *__restrict__ x, int64_t B,
                                                            float *__restrict__ E, int32_t *status) {
  __shared__ float s_x[XW][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nw = (int64_t)gridDim.x * XW;
  for (int64_t b = (int64_t)blockIdx.x * XW + warp; b < B; b += nw) {
    for (int c = lane; c < P.xcols; c += 32) s_x[warp][c] = x[b * P.xcols + c];
    __syncwarp();
    for (int t = 0; t < P.num_slots; ++t) {
      float *out = E + (b * P.num_slots + t) * (int64_t)P.W;
      const int col = P.col[t];
      if (P.kind[t] == 0) {
        const int64_t id = rs::clamp_id((int64_t)s_x[warp][col], P.rows[t], status);
        const float *src = P.table[t] + id * P.W;
        for (int d = lane; d < P.W; d += 32) out[d] = src[d];
      } else if (P.kind[t] == 1) {
        for (int d = lane; d < P.W; d += 32) {
          float acc = 0.f;
          for (int k = 0; k < P.ncols[t]; ++k) {
            const float xv = s_x[warp][col + k];
            if (xv != 0.f) acc = fmaf(xv, P.table[t][k * P.W + d], acc);
          }
          out[d] = acc;
        }
      } else {
        const float xv = s_x[warp][col];
        for (int d = lane; d < P.W; d += 32) out[d] = xv;
      }
    }
    __syncwarp();
  }
}

// pass 1: block `blk` owns samples [blk*SPB, (blk+1)*SPB); thread d owns column d of every bag row and walks
// the block's samples in order -> partial[blk][row][d].   pass 2 adds the block partials in block order.
constexpr int SPB = 128;
__global__ void bag_bwd_partial_kernel(const __grid_constant__ XParams P, const float *__restrict__ x, const float *__restrict__ dE,
                                       int64_t B, float *__restrict__ partial) {
  extern __shared__ float s_acc[];  // [bag_rows_total][W]
  const int d = threadIdx.x;
  if (d >= P.W) return;
  for (int r = 0; r < P.bag_rows_total; ++r) s_acc[r * P.W + d] = 0.f;
  const int64_t b0 = (int64_t)blockIdx.x * SPB;
  const int64_t b1 = b0 + SPB < B ? b0 + SPB : B;
  for (int64_t b = b0; b < b1; ++b) {
    const float *xr = x + b * P.xcols;
    for (int t = 0; t < P.num_slots; ++t) {
      if (P.kind[t] != 1) continue;
      const float g = dE[(b * P.num_slots + t) * (int64_t)P.W + d];
      for (int k = 0; k < P.ncols[t]; ++k) {
        const float xv = __ldg(xr + P.col[t] + k);
        if (xv != 0.f) s_acc[(P.bag_row0[t] + k) * P.W + d] = fmaf(xv, g, s_acc[(P.bag_row0[t] + k) * P.W + d]);
      }
    }
  }
  float *dst = partial + (int64_t)blockIdx.x * P.bag_rows_total * P.W;
  for (int r = 0; r < P.bag_rows_total; ++r) dst[r * P.W + d] = s_acc[r * P.W + d];
}

struct DwPtrs {
  float *p[RS_MAX_FIELDS];
};
__global__ void bag_bwd_finish_kernel(const __grid_constant__ XParams P, const float *__restrict__ partial, int nblocks,
                                      const __grid_constant__ DwPtrs DW) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // (row, d) flattened
  if (e >= P.bag_rows_total * P.W) return;
  float acc = 0.f;
  for (int blk = 0; blk < nblocks; ++blk) acc += partial[(int64_t)blk * P.bag_rows_total * P.W + e];
  const int r = e / P.W, d = e - r * P.W;
  for (int t = 0; t < P.num_slots; ++t)
    if (P.kind[t] == 1 && r >= P.bag_row0[t] && r < P.bag_row0[t] + P.ncols[t]) DW.p[t][(r - P.bag_row0[t]) * P.W + d] = acc;
}

int fill(XParams &P, const rs_xslots *S, const char *who) {
  RS_CHECK_ARG(S, RS_E_ARG, "%s: null slots", who);
  RS_CHECK_ARG(S->num_slots >= 1 && S->num_slots <= RS_MAX_FIELDS && S->width >= 1 && S->xcols >= 1 && S->xcols <= 128, RS_E_SHAPE,
               "%s: bad slots/width/xcols", who);
  P.num_slots = S->num_slots;
  P.W = S->width;
  P.xcols = S->xcols;
  int row0 = 0;
  for (int t = 0; t < S->num_slots; ++t) {
    P.col[t] = S->col[t];
    P.ncols[t] = S->ncols[t];
    P.kind[t] = S->kind[t];
    P.table[t] = S->table[t];
    P.rows[t] = S->rows[t];
    RS_CHECK_ARG(S->kind[t] >= 0 && S->kind[t] <= 2, RS_E_ARG, "%s: slot %d bad kind", who, t);
    RS_CHECK_ARG(S->col[t] >= 0 && S->col[t] + (S->kind[t] == 1 ? S->ncols[t] : 1) <= S->xcols, RS_E_SHAPE, "%s: slot %d columns out of x",
                 who, t);
    if (S->kind[t] != 2) RS_CHECK_ARG(S->table[t] != nullptr, RS_E_ARG, "%s: slot %d has no table", who, t);
    if (S->kind[t] == 1) RS_CHECK_ARG(S->rows[t] == S->ncols[t], RS_E_SHAPE, "%s: bag slot %d: table rows != ncols", who, t);
    P.bag_row0[t] = row0;
    if (S->kind[t] == 1) row0 += S->ncols[t];
  }
  P.bag_rows_total = row0;
  return RS_OK;
}

}  // namespace

RS_API int rs_xembed_fwd(const rs_xslots *S, const float *x, int64_t B, float *E, int32_t *status, void *stream) {
  XParams P = {};
  int rc = fill(P, S, "rs_xembed_fwd");
  if (rc) return rc;
  RS_CHECK_ARG(x && E, RS_E_ARG, "rs_xembed_fwd: null argument");
  if (B == 0) return RS_OK;
  int64_t blocks64 = (B + XW - 1) / XW;
  int cap = rs::num_sms() * 8;
  int blocks = (int)(blocks64 < cap ? blocks64 : cap);
  xembed_fwd_kernel<<<blocks, XW * 32, 0, (cudaStream_t)stream>>>(P, x, B, E, status);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

RS_API int rs_xembed_bag_ws_bytes(const rs_xslots *S, int64_t B, size_t *bytes) {
  XParams P = {};
  int rc = fill(P, S, "rs_xembed_bag_ws_bytes");
  if (rc) return rc;
  RS_CHECK_ARG(bytes, RS_E_ARG, "rs_xembed_bag_ws_bytes: null bytes");
  int64_t nblocks = (B + SPB - 1) / SPB;
  *bytes = (size_t)(nblocks > 0 ? nblocks : 1) * P.bag_rows_total * P.W * 4 + 16;
  return RS_OK;
}

RS_API int rs_xembed_bag_bwd(const rs_xslots *S, const float *x, const float *dE, int64_t B, float *const *dW, float *ws, size_t ws_bytes,
                             void *stream) {
  XParams P = {};
  int rc = fill(P, S, "rs_xembed_bag_bwd");
  if (rc) return rc;
  RS_CHECK_ARG(x && dE && dW && ws, RS_E_ARG, "rs_xembed_bag_bwd: null argument");
  RS_CHECK_ARG(P.W <= 1024, RS_E_UNSUPPORTED, "rs_xembed_bag_bwd: width > 1024");
  if (P.bag_rows_total == 0 || B == 0) return RS_OK;
  int nblocks = (int)((B + SPB - 1) / SPB);
  size_t need = (size_t)nblocks * P.bag_rows_total * P.W * 4;
  RS_CHECK_ARG(ws_bytes >= need, RS_E_WORKSPACE, "rs_xembed_bag_bwd: workspace too small");
  DwPtrs DW = {};
  for (int t = 0; t < P.num_slots; ++t) {
    DW.p[t] = dW[t];
    if (P.kind[t] == 1) RS_CHECK_ARG(dW[t] != nullptr, RS_E_ARG, "rs_xembed_bag_bwd: dW[%d] is NULL", t);
  }
  size_t smem = (size_t)P.bag_rows_total * P.W * 4;
  RS_CHECK_ARG(smem <= 200 * 1024, RS_E_UNSUPPORTED, "rs_xembed_bag_bwd: bag rows * width too large");
  RS_CUDA(cudaFuncSetAttribute(bag_bwd_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int threads = (P.W + 31) / 32 * 32;
  cudaStream_t st = (cudaStream_t)stream;
  bag_bwd_partial_kernel<<<nblocks, threads, smem, st>>>(P, x, dE, B, ws);
  RS_CHECK_LAUNCH();
  int total = P.bag_rows_total * P.W;
  bag_bwd_finish_kernel<<<(total + 255) / 256, 256, 0, st>>>(P, ws, nblocks, DW);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
