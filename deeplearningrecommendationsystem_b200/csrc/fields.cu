// Fused multi-field embedding lookup + feature interaction, forward and backward.
//
// One warp owns one sample.  The sample's F rows (F*D floats) are pulled global->shared with 16-byte cp.async
// (LDGSTS) into a per-warp double buffer, so the rows of sample n+1 are in flight while sample n is reduced:
// every lane has ceil(F*D/128) independent 128-bit requests outstanding per buffer and no registers are tied up.
// Lane l owns the float4 slots q = l, l+32, ... of the flattened (F, D) tile; because D/4 is a power of two all
// of a lane's slots share the same column group, so the field sum S needs only log2(32/(D/4)) xor-shuffles.
//
// Replaces (reference): FM second order model/deepfm.py:71-77, NFM bi-interaction model/nfm.py:58-62, PNN inner
// products model/pnn.py:61-66, the field concat model/deepfm.py:54 / model/pnn.py:55 / model/neuralcf.py:46,
// MF dot model/mf.py:26 and GMF Hadamard model/neuralcf.py:39.
#include "common.cuh"

namespace {

constexpr int WARPS = 8;

struct FieldsParams {
  const float *base[RS_MAX_FIELDS];
  int64_t rows[RS_MAX_FIELDS];
  const int64_t *ids;
  const float *dense_in;
  float *cross, *bi, *pairs, *concat, *stash, *dot2, *had2;
  // backward
  const float *g_cross, *g_bi, *g_pairs, *g_concat, *g_dot2, *g_had2;
  float *dE;
  int64_t B;
  int F, D, dv, dvs;  // dv = D/4 (power of two), dvs = log2(dv)
  int FQ, NP;         // F*dv float4 per sample, number of pairs
  int32_t *status;
};

__device__ __forceinline__ void issue_sample(const FieldsParams &P, int64_t b, float4 *buf, int64_t *s_ids, const float *const *s_base,
                                             const int64_t *s_rows, int lane) {
  if (P.ids) {
    for (int f = lane; f < P.F; f += 32) s_ids[f] = rs::clamp_id(P.ids[b * P.F + f], s_rows[f], P.status);
    __syncwarp();
    for (int q = lane; q < P.FQ; q += 32) {
      const int f = q >> P.dvs, d4 = q & (P.dv - 1);
      rs::cp_async16(buf + q, s_base[f] + (s_ids[f] * P.dv + d4) * 4);
    }
  } else {
    const float *src = P.dense_in + b * (int64_t)P.FQ * 4;
    for (int q = lane; q < P.FQ; q += 32) rs::cp_async16(buf + q, src + q * 4);
  }
}

// field sum S and sum of squares Q for this lane's column group(s); valid on every lane after the shuffles
__device__ __forceinline__ void field_sums(const FieldsParams &P, const float4 *buf, int lane, float4 &S0, float4 &S1, float4 &Q0,
                                           float4 &Q1) {
  S0 = S1 = Q0 = Q1 = rs::f4_zero();
  int k = 0;
  for (int q = lane; q < P.FQ; q += 32, ++k) {
    const float4 v = buf[q];
    if (P.dv == 64 && (k & 1)) {  // warp-uniform: slot q covers columns 32..63 of its row
      S1 = rs::f4_add(S1, v);
      Q1 = rs::f4_fma(v, v, Q1);
    } else {
      S0 = rs::f4_add(S0, v);
      Q0 = rs::f4_fma(v, v, Q0);
    }
  }
  for (int o = 16; o >= P.dv; o >>= 1) {
    S0 = rs::f4_add(S0, rs::f4_shfl_xor(S0, o));
    Q0 = rs::f4_add(Q0, rs::f4_shfl_xor(Q0, o));
  }
}

__global__ void __launch_bounds__(WARPS * 32) fields_fwd_kernel(const __grid_constant__ FieldsParams P) {
  extern __shared__ float4 smem[];
  __shared__ const float *s_base[RS_MAX_FIELDS];
  __shared__ int64_t s_rows[RS_MAX_FIELDS];
  __shared__ int64_t s_ids_all[WARPS][RS_MAX_FIELDS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < P.F; i += blockDim.x) {
    s_base[i] = P.base[i];
    s_rows[i] = P.rows[i];
  }
  // pair table (i << 8 | j) in nested-loop order, after the per-warp tile buffers
  uint16_t *s_pair = reinterpret_cast<uint16_t *>(smem + (size_t)WARPS * 2 * P.FQ);
  if (P.pairs) {
    for (int pr = threadIdx.x; pr < P.NP; pr += blockDim.x) {
      int i = 0, rem = pr;
      while (rem >= P.F - 1 - i) {
        rem -= P.F - 1 - i;
        ++i;
      }
      s_pair[pr] = (uint16_t)((i << 8) | (i + 1 + rem));
    }
  }
  __syncthreads();
  float4 *buf0 = smem + (size_t)warp * 2 * P.FQ;
  int64_t *s_ids = s_ids_all[warp];
  const int64_t stride = (int64_t)gridDim.x * WARPS;
  int64_t b = (int64_t)blockIdx.x * WARPS + warp;
  if (b < P.B) issue_sample(P, b, buf0, s_ids, s_base, s_rows, lane);
  rs::cp_async_commit();
  const int dvm = P.dv < 32 ? P.dv : 32;
  for (int it = 0; b < P.B; b += stride, it ^= 1) {
    const int64_t bn = b + stride;
    if (bn < P.B) issue_sample(P, bn, buf0 + (it ^ 1) * P.FQ, s_ids, s_base, s_rows, lane);
    rs::cp_async_commit();
    rs::cp_async_wait<1>();
    __syncwarp();
    const float4 *buf = buf0 + it * P.FQ;
    const int64_t tile = b * (int64_t)P.FQ;

    if (P.cross || P.bi || P.stash) {
      float4 S0, S1, Q0, Q1;
      field_sums(P, buf, lane, S0, S1, Q0, Q1);
      float4 bi0 = rs::f4_scale(rs::f4_sub(rs::f4_mul(S0, S0), Q0), 0.5f);
      float4 bi1 = rs::f4_scale(rs::f4_sub(rs::f4_mul(S1, S1), Q1), 0.5f);
      if (P.cross) {
        float t = lane < dvm ? rs::f4_hsum(bi0) + (P.dv == 64 ? rs::f4_hsum(bi1) : 0.f) : 0.f;
        t = rs::warp_sum(t);
        if (lane == 0) P.cross[b] = t;
      }
      if (P.bi && lane < dvm) {
        rs::stg_f4(P.bi + b * P.D + lane * 4, bi0);
        if (P.dv == 64) rs::stg_f4(P.bi + b * P.D + (lane + 32) * 4, bi1);
      }
      if (P.stash) {
        int k = 0;
        for (int q = lane; q < P.FQ; q += 32, ++k) {
          const float4 s = (P.dv == 64 && (k & 1)) ? S1 : S0;
          rs::stg_cs_f4(P.stash + (tile + q) * 4, rs::f4_sub(s, buf[q]));
        }
      }
    }
    if (P.concat)
      for (int q = lane; q < P.FQ; q += 32) rs::stg_f4(P.concat + (tile + q) * 4, buf[q]);
    if (P.dot2 || P.had2) {
      float acc = 0.f;
      for (int q = lane; q < P.dv; q += 32) {
        const float4 h = rs::f4_mul(buf[q], buf[P.dv + q]);
        if (P.had2) rs::stg_f4(P.had2 + b * P.D + q * 4, h);
        acc += rs::f4_hsum(h);
      }
      if (P.dot2) {
        acc = rs::warp_sum(acc);
        if (lane == 0) P.dot2[b] = acc;
      }
    }
    if (P.pairs) {
      const float *e = reinterpret_cast<const float *>(buf);
      for (int pr = lane; pr < P.NP; pr += 32) {
        const int ij = s_pair[pr];
        const float *ei = e + (ij >> 8) * P.D, *ej = e + (ij & 255) * P.D;
        float acc = 0.f;
        for (int t = 0; t < P.D; ++t) {
          const int d = (t + lane) & (P.D - 1);  // rotate the start column per lane: conflict-free banks
          acc = fmaf(ei[d], ej[d], acc);
        }
        P.pairs[b * P.NP + pr] = acc;
      }
    }
    __syncwarp();
  }
  rs::cp_async_wait<0>();
}

__device__ __forceinline__ int pair_index(int F, int i, int j) { return i * F - (i * (i + 1)) / 2 + (j - i - 1); }

__global__ void __launch_bounds__(WARPS * 32) fields_bwd_kernel(const __grid_constant__ FieldsParams P) {
  extern __shared__ float4 smem[];
  __shared__ const float *s_base[RS_MAX_FIELDS];
  __shared__ int64_t s_rows[RS_MAX_FIELDS];
  __shared__ int64_t s_ids_all[WARPS][RS_MAX_FIELDS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < P.F; i += blockDim.x) {
    s_base[i] = P.base[i];
    s_rows[i] = P.rows[i];
  }
  __syncthreads();
  float4 *buf = smem + (size_t)warp * P.FQ;
  float *s_gp = reinterpret_cast<float *>(smem + (size_t)WARPS * P.FQ) + (size_t)warp * P.NP;
  int64_t *s_ids = s_ids_all[warp];
  const int64_t stride = (int64_t)gridDim.x * WARPS;
  for (int64_t b = (int64_t)blockIdx.x * WARPS + warp; b < P.B; b += stride) {
    issue_sample(P, b, buf, s_ids, s_base, s_rows, lane);
    rs::cp_async_commit();
    if (P.g_pairs)
      for (int pr = lane; pr < P.NP; pr += 32) s_gp[pr] = P.g_pairs[b * P.NP + pr];
    rs::cp_async_wait<0>();
    __syncwarp();
    const int64_t tile = b * (int64_t)P.FQ;
    float4 S0 = rs::f4_zero(), S1 = rs::f4_zero(), Q0, Q1;
    const bool fm = P.g_cross || P.g_bi;
    if (fm) field_sums(P, buf, lane, S0, S1, Q0, Q1);
    const float gc = P.g_cross ? P.g_cross[b] : 0.f;
    int k = 0;
    for (int q = lane; q < P.FQ; q += 32, ++k) {
      const int f = q >> P.dvs, d4 = q & (P.dv - 1);
      const float4 e = buf[q];
      float4 g = rs::f4_zero();
      if (fm) {
        const float4 s = (P.dv == 64 && (k & 1)) ? S1 : S0;
        float4 u = make_float4(gc, gc, gc, gc);
        if (P.g_bi) u = rs::f4_add(u, rs::ldg_f4(P.g_bi + b * P.D + d4 * 4));
        g = rs::f4_mul(u, rs::f4_sub(s, e));
      }
      if (P.g_pairs) {
        for (int j = 0; j < P.F; ++j) {
          if (j == f) continue;
          const float w = s_gp[j > f ? pair_index(P.F, f, j) : pair_index(P.F, j, f)];
          g = rs::f4_fmas(buf[j * P.dv + d4], w, g);
        }
      }
      if (P.g_concat) g = rs::f4_add(g, rs::ldg_f4(P.g_concat + (tile + q) * 4));
      if (P.g_dot2 || P.g_had2) {  // F == 2
        const float4 other = buf[(1 - f) * P.dv + d4];
        if (P.g_dot2) g = rs::f4_fmas(other, P.g_dot2[b], g);
        if (P.g_had2) g = rs::f4_fma(other, rs::ldg_f4(P.g_had2 + b * P.D + d4 * 4), g);
      }
      rs::stg_f4(P.dE + (tile + q) * 4, g);
    }
    __syncwarp();
  }
}

int fill_common(FieldsParams &P, const rs_tables *T, const int64_t *ids, const float *dense_in, int64_t B, const char *who) {
  RS_CHECK_ARG(T, RS_E_ARG, "%s: null tables", who);
  const int F = T->num_fields, D = T->width;
  RS_CHECK_ARG(F >= 1 && F <= RS_MAX_FIELDS, RS_E_SHAPE, "%s: F=%d out of range", who, F);
  RS_CHECK_ARG(D >= 4 && D <= 256 && (D & (D - 1)) == 0, RS_E_UNSUPPORTED, "%s: D=%d must be a power of two in [4,256]", who, D);
  RS_CHECK_ARG((ids != nullptr) != (dense_in != nullptr), RS_E_ARG, "%s: exactly one of ids / dense_in", who);
  for (int f = 0; f < F; ++f) {
    P.base[f] = T->base[f];
    P.rows[f] = T->rows[f];
    if (ids) RS_CHECK_ARG(T->base[f] && T->rows[f] > 0, RS_E_ARG, "%s: table %d missing", who, f);
  }
  P.ids = ids;
  P.dense_in = dense_in;
  P.B = B;
  P.F = F;
  P.D = D;
  P.dv = D / 4;
  P.dvs = 0;
  while ((1 << P.dvs) < P.dv) ++P.dvs;
  P.FQ = F * P.dv;
  P.NP = F * (F - 1) / 2;
  return RS_OK;
}

int grid_for(int64_t B, size_t smem, const void *kernel) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS * 32, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t want = (B + WARPS - 1) / WARPS;
  int64_t cap = (int64_t)rs::num_sms() * per_sm;
  return (int)(want < cap ? want : cap);
}

}  // namespace

RS_API int rs_fields_fwd(const rs_tables *T, const rs_fields_io *io, int64_t B, int32_t *status, void *stream) {
  RS_CHECK_ARG(io, RS_E_ARG, "rs_fields_fwd: null io");
  FieldsParams P = {};
  int rc = fill_common(P, T, io->ids, io->dense_in, B, "rs_fields_fwd");
  if (rc) return rc;
  P.cross = io->cross;
  P.bi = io->bi;
  P.pairs = io->pairs;
  P.concat = io->concat;
  P.stash = io->stash;
  P.dot2 = io->dot2;
  P.had2 = io->had2;
  P.status = status;
  RS_CHECK_ARG(!(P.dot2 || P.had2) || P.F == 2, RS_E_SHAPE, "rs_fields_fwd: dot2/had2 need F == 2");
  RS_CHECK_ARG(!P.pairs || P.F <= 255, RS_E_SHAPE, "rs_fields_fwd: pairs needs F <= 255");
  if (B == 0) return RS_OK;
  size_t smem = (size_t)WARPS * 2 * P.FQ * 16 + (P.pairs ? ((size_t)P.NP * 2 + 15) / 16 * 16 : 0);
  RS_CHECK_ARG(smem <= 200 * 1024, RS_E_UNSUPPORTED, "rs_fields_fwd: F*D=%d too large for the shared-memory tile", P.F * P.D);
  RS_CUDA(cudaFuncSetAttribute(fields_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = grid_for(B, smem, (const void *)fields_fwd_kernel);
  fields_fwd_kernel<<<grid, WARPS * 32, smem, (cudaStream_t)stream>>>(P);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

RS_API int rs_fields_bwd(const rs_tables *T, const rs_fields_grad *g, int64_t B, void *stream) {
  RS_CHECK_ARG(g && g->dE, RS_E_ARG, "rs_fields_bwd: null argument");
  FieldsParams P = {};
  int rc = fill_common(P, T, g->ids, g->dense_in, B, "rs_fields_bwd");
  if (rc) return rc;
  P.g_cross = g->g_cross;
  P.g_bi = g->g_bi;
  P.g_pairs = g->g_pairs;
  P.g_concat = g->g_concat;
  P.g_dot2 = g->g_dot2;
  P.g_had2 = g->g_had2;
  P.dE = g->dE;
  P.status = nullptr;
  RS_CHECK_ARG(!(P.g_dot2 || P.g_had2) || P.F == 2, RS_E_SHAPE, "rs_fields_bwd: dot2/had2 need F == 2");
  if (B == 0) return RS_OK;
  size_t smem = (size_t)WARPS * P.FQ * 16 + (P.g_pairs ? (size_t)WARPS * P.NP * 4 : 0);
  smem = (smem + 15) / 16 * 16;
  RS_CHECK_ARG(smem <= 200 * 1024, RS_E_UNSUPPORTED, "rs_fields_bwd: F*D=%d too large for the shared-memory tile", P.F * P.D);
  RS_CUDA(cudaFuncSetAttribute(fields_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = grid_for(B, smem, (const void *)fields_bwd_kernel);
  fields_bwd_kernel<<<grid, WARPS * 32, smem, (cudaStream_t)stream>>>(P);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
