// Field-aware factorisation machine: lookup + pairwise field-aware interaction (reference model/ffm.py:46-82,
// generalised from the six MovieLens features to F id-fields).
//
// rs_ffm_fwd -- HBM-bound streaming kernel.  A sample needs F table rows of F*D floats (43 KB at F=26, D=16).
// A persistent CTA per SM runs a producer warp that issues one cp.async.bulk (TMA bulk copy, SASS UBLKCP) per
// row into a ring of shared-memory stages guarded by full/empty mbarriers; four consumer warps reduce
// sum_{i<j} <T[i][j], T[j][i]> from shared memory and stream out the transposed tile
// stash[b][i][j] = T[j][i] (the Jacobian d cross / d v_{i,j}) that the segment-reduce/update kernel later
// scales by dL/dcross[b].  Rows are padded in shared memory so the transposed reads are bank-conflict free.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int NCW = 8;                 // consumer warps
constexpr int NPW = 2;                 // producer warps: one warp issues a bulk copy every ~70 cycles (profiles/tma_rate_micro.cu:
                                       // 44 ns per 1664-byte row from one warp, 31-35 ns from two), so the rows of a sample are
                                       // split over two issuing warps
constexpr int NTHREADS = (NCW + NPW) * 32;
constexpr int MAX_STAGES = 8;
constexpr int64_t HOT_TABLE_BYTES = 32ll << 20;

struct FfmParams {
  const float *base[RS_MAX_FIELDS];
  int64_t rows[RS_MAX_FIELDS];
  const int64_t *ids;
  float *cross, *stash;
  int64_t B;
  int F, D, dv, dvs;
  int rowv;    // float4 per table row = F*dv
  int pitchv;  // float4 per padded row in shared memory
  int nst;     // ring stages
  int32_t *status;
  // row-sharded "direct" fields (rs_ffm_fwd_peer): read from the owner's shard over NVLink, ids are global rows
  int world;
  uint64_t direct_mask;
  const float *shard[RS_MAX_RANKS];
  int64_t total_rows;
  // "cold-slice" stash (rs_ffm_fwd_train): mini[b][i][c][:] = v_{cold_c, i}(b) for the nC fields flagged cold
  float *mini;
  // split stash (rs_peer_tables.stash_split): rows of the fields in split_mask go to stash2, the others to stash
  float *stash2;
  int F1, F2;                            // number of fields whose rows go to stash / stash2
  unsigned char st_slot[RS_MAX_FIELDS];  // position of field i inside its stash; bit 7 set = stash2
  int hint;   // use the L2 eviction hints (RS_FFM_L2HINT=0 turns them off)
  int nC;
  signed char cidx[RS_MAX_FIELDS];   // index of field j among the cold fields, -1 = not cold
};

__global__ void __launch_bounds__(NTHREADS, 1) ffm_fwd_kernel(const __grid_constant__ FfmParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES];
  __shared__ float s_part[MAX_STAGES][NCW];
  __shared__ const float *s_base[RS_MAX_FIELDS];
  __shared__ int64_t s_rows[RS_MAX_FIELDS];
  __shared__ int s_cidx[RS_MAX_FIELDS];
  __shared__ int s_slot[RS_MAX_FIELDS];
  for (int i = threadIdx.x; i < P.F; i += blockDim.x) {
    s_base[i] = P.base[i];
    s_rows[i] = P.rows[i];
    s_cidx[i] = P.cidx[i];
    s_slot[i] = P.st_slot[i];
  }
  float4 *tiles = reinterpret_cast<float4 *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stage_v = P.F * P.pitchv;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.nst; ++s) {
      rs::mbar_init(&full_bar[s], NPW);
      rs::mbar_init(&empty_bar[s], NCW);
    }
    rs::mbar_fence_init();
  }
  __syncthreads();
  const uint32_t row_bytes = (uint32_t)P.rowv * 16u;

  if (warp < NPW) {
    // ===== producers: one bulk copy per table row; lane l of producer warp w owns field f = l * NPW + w =====
    // The id reads sit on the producers' critical path (one dependent global load per sample), so the id of sample
    // k+1 is requested before the copies of sample k are issued.
    const int f = lane * NPW + warp;
    const bool has = f < P.F;
    const uint32_t nmine = (uint32_t)((P.F - warp + NPW - 1) / NPW);
    // L2 residency hint per field: a table of at most HOT_TABLE_BYTES is re-read many times per batch (keep it), a large
    // one is touched about once per row (let it go first); the stash stores are streaming (st.global.cs) already
    const bool hot = has && !((P.direct_mask >> f) & 1ull) && P.hint && s_rows[f] * (int64_t)P.rowv * 16 <= HOT_TABLE_BYTES;
    const uint64_t policy = hot ? rs::l2_policy_evict_last() : rs::l2_policy_evict_first();
    int k = 0;
    int64_t b = blockIdx.x;
    int64_t id_cur = 0, id_nxt = 0;
    if (b < P.B && has) id_cur = P.ids[b * P.F + f];
    for (; b < P.B; b += gridDim.x, ++k) {
      const int64_t bn = b + gridDim.x;
      if (bn < P.B && has) id_nxt = P.ids[bn * P.F + f];
      const int s = k % P.nst;
      const uint32_t ph = (uint32_t)(k / P.nst) & 1u;
      rs::mbar_wait(&empty_bar[s], ph ^ 1u);
      if (lane == 0) rs::mbar_arrive_expect_tx(&full_bar[s], row_bytes * nmine);
      __syncwarp();
      float4 *dst = tiles + (size_t)s * stage_v;
      if (has) {
        const float *src;
        if ((P.direct_mask >> f) & 1ull) {   // global row g of a direct field: rank g % world holds it at local row g / world
          const int64_t g = rs::clamp_id(id_cur, P.total_rows, P.status);
          const int64_t l = g / P.world;
          src = P.shard[(int)(g - l * P.world)] + l * (int64_t)P.rowv * 4;
        } else {
          src = s_base[f] + rs::clamp_id(id_cur, s_rows[f], P.status) * (int64_t)P.rowv * 4;
        }
        if (P.hint)
          rs::bulk_g2s_hint(dst + (size_t)f * P.pitchv, src, row_bytes, &full_bar[s], policy);
        else
          rs::bulk_g2s(dst + (size_t)f * P.pitchv, src, row_bytes, &full_bar[s]);
      }
      id_cur = id_nxt;
    }
  } else {
    // ===== consumers =====
    const int cw = warp - NPW;
    const int ct = threadIdx.x - 32 * NPW;
    int k = 0;
    for (int64_t b = blockIdx.x; b < P.B; b += gridDim.x, ++k) {
      const int s = k % P.nst;
      const uint32_t ph = (uint32_t)(k / P.nst) & 1u;
      rs::mbar_wait(&full_bar[s], ph);
      const float4 *T = tiles + (size_t)s * stage_v;
      float4 *st_out = P.stash ? reinterpret_cast<float4 *>(P.stash) + b * (int64_t)P.F * P.rowv : nullptr;
      float4 *mini_b = P.mini ? reinterpret_cast<float4 *>(P.mini) + b * (int64_t)P.F * P.nC * P.dv : nullptr;
      float acc = 0.f;
      for (int i = cw; i < P.F; i += NCW) {
        const float4 *Ti = T + (size_t)i * P.pitchv;
        const float4 *Tcol = T + i * P.dv;  // + j*pitchv + d4 -> v_{j,i}
        float4 *out_i = st_out ? st_out + (size_t)i * P.rowv : nullptr;
        if (P.stash2) {   // split stash: (B, F1, W) and (B, F2, W)
          const int sl = s_slot[i];
          out_i = (sl & 128) ? reinterpret_cast<float4 *>(P.stash2) + (b * P.F2 + (sl & 127)) * (int64_t)P.rowv
                             : (P.stash ? reinterpret_cast<float4 *>(P.stash) + (b * P.F1 + sl) * (int64_t)P.rowv : nullptr);
        }
        float4 *mini_i = mini_b ? mini_b + (size_t)i * P.nC * P.dv : nullptr;
        for (int w0 = lane; w0 < P.rowv; w0 += 128) {
          float4 tr[4], own[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {  // four independent shared-memory reads per lane before any use
            const int w = w0 + 32 * u;
            const int j = w >> P.dvs, d4 = w & (P.dv - 1);
            const bool live = w < P.rowv;
            tr[u] = (live && j != i) ? Tcol[(size_t)j * P.pitchv + d4] : rs::f4_zero();
            own[u] = (live && j > i) ? Ti[w] : rs::f4_zero();
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int w = w0 + 32 * u;
            acc += rs::f4_dot(own[u], tr[u]);
            if (out_i && w < P.rowv) rs::stg_cs_f4(reinterpret_cast<float *>(out_i + w), tr[u]);
            if (mini_i && w < P.rowv) {
              const int j = w >> P.dvs;
              const int c = s_cidx[j];
              if (c >= 0 && j != i) rs::stg_cs_f4(reinterpret_cast<float *>(mini_i + c * P.dv + (w & (P.dv - 1))), tr[u]);
            }
          }
        }
      }
      acc = rs::warp_sum(acc);
      if (lane == 0) s_part[s][cw] = acc;
      // all four consumer warps: partials visible, tile reads done
      asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
      if (ct == 0) {
        float t = 0.f;
#pragma unroll
        for (int c = 0; c < NCW; ++c) t += s_part[s][c];
        P.cross[b] = t;
      }
      if (lane == 0) rs::mbar_arrive(&empty_bar[s]);
    }
  }
}

// ------------------------------------------------------------------ dense small-F variant (MovieLens FFM)
struct FieldOf {
  int32_t fo[RS_MAX_FIELDS];
};

// warp per sample; Tin (B, F, NF, D).  mode 0: cross;  mode 1: dT = g * d cross / dT
__global__ void __launch_bounds__(256) ffm_dense_kernel(const float *__restrict__ Tin, const float *__restrict__ g_cross, int64_t B, int F,
                                                       int NF, int D, const __grid_constant__ FieldOf FO, float *__restrict__ cross,
                                                       float *__restrict__ dT) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t tile = (int64_t)F * NF * D;
  for (int64_t b = warp_global; b < B; b += nwarps) {
    const float *T = Tin + b * tile;
    if (cross) {
      float acc = 0.f;
      for (int i = 0; i < F; ++i)
        for (int j = i + 1; j < F; ++j) {
          const float *a = T + ((int64_t)i * NF + FO.fo[j]) * D;
          const float *c = T + ((int64_t)j * NF + FO.fo[i]) * D;
          for (int d = lane; d < D; d += 32) acc = fmaf(a[d], c[d], acc);
        }
      acc = rs::warp_sum(acc);
      if (lane == 0) cross[b] = acc;
    }
    if (dT) {
      const float g = g_cross[b];
      for (int i = 0; i < F; ++i)
        for (int c = 0; c < NF; ++c)
          for (int d = lane; d < D; d += 32) {
            float acc = 0.f;
            for (int j = 0; j < F; ++j)
              if (j != i && FO.fo[j] == c) acc += T[((int64_t)j * NF + FO.fo[i]) * D + d];
            dT[b * tile + ((int64_t)i * NF + c) * D + d] = g * acc;
          }
    }
  }
}

int dense_launch(const float *Tin, const float *g, int64_t B, int F, int NF, int D, const int32_t *field_of, float *cross, float *dT,
                 void *stream) {
  RS_CHECK_ARG(Tin && field_of, RS_E_ARG, "rs_ffm_dense: null argument");
  RS_CHECK_ARG(F >= 2 && F <= RS_MAX_FIELDS && NF >= 1 && NF <= RS_MAX_FIELDS && D >= 1, RS_E_SHAPE, "rs_ffm_dense: bad shape");
  FieldOf FO;
  for (int f = 0; f < RS_MAX_FIELDS; ++f) FO.fo[f] = 0;
  for (int f = 0; f < F; ++f) {
    RS_CHECK_ARG(field_of[f] >= 0 && field_of[f] < NF, RS_E_ARG, "rs_ffm_dense: field_of[%d] out of range", f);
    FO.fo[f] = field_of[f];
  }
  if (B == 0) return RS_OK;
  int64_t blocks64 = (B + 7) / 8;
  int cap = rs::num_sms() * 8;
  int blocks = (int)(blocks64 < cap ? blocks64 : cap);
  ffm_dense_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(Tin, g, B, F, NF, D, FO, cross, dT);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

}  // namespace

static int ffm_fwd_launch(const rs_tables *T, const int64_t *ids, int64_t B, int32_t D, const rs_peer_tables *PT, float *cross,
                          float *stash, uint64_t cold_mask, float *mini, int32_t *status, void *stream);

RS_API int rs_ffm_fwd(const rs_tables *T, const int64_t *ids, int64_t B, int32_t D, float *cross, float *stash, int32_t *status,
                      void *stream) {
  return ffm_fwd_launch(T, ids, B, D, nullptr, cross, stash, 0, nullptr, status, stream);
}

RS_API int rs_ffm_fwd_peer(const rs_tables *T, const int64_t *ids, int64_t B, int32_t D, const rs_peer_tables *PT, float *cross,
                           float *stash, int32_t *status, void *stream) {
  return ffm_fwd_launch(T, ids, B, D, PT, cross, stash, 0, nullptr, status, stream);
}

RS_API int rs_ffm_fwd_train(const rs_tables *T, const int64_t *ids, int64_t B, int32_t D, uint64_t cold_mask, float *cross,
                            float *cold_stash, int32_t *status, void *stream) {
  RS_CHECK_ARG(!cold_mask || cold_stash, RS_E_ARG, "rs_ffm_fwd_train: cold_stash is NULL");
  return ffm_fwd_launch(T, ids, B, D, nullptr, cross, nullptr, cold_mask, cold_stash, status, stream);
}

static int ffm_fwd_launch(const rs_tables *T, const int64_t *ids, int64_t B, int32_t D, const rs_peer_tables *PT, float *cross,
                          float *stash, uint64_t cold_mask, float *mini, int32_t *status, void *stream) {
  RS_CHECK_ARG(T && ids && cross, RS_E_ARG, "rs_ffm_fwd: null argument");
  const int F = T->num_fields;
  RS_CHECK_ARG(F >= 2 && F <= RS_MAX_FIELDS, RS_E_SHAPE, "rs_ffm_fwd: F=%d out of range", F);
  RS_CHECK_ARG(D >= 4 && D <= 256 && (D & (D - 1)) == 0, RS_E_UNSUPPORTED, "rs_ffm_fwd: D=%d must be a power of two in [4,256]", D);
  RS_CHECK_ARG(T->width == F * D, RS_E_SHAPE, "rs_ffm_fwd: table width %d != F*D = %d", T->width, F * D);
  if (B == 0) return RS_OK;
  FfmParams P = {};
  P.world = 1;
  if (PT && (PT->direct_mask || PT->stash_split)) {
    RS_CHECK_ARG(PT->world >= 1 && PT->world <= RS_MAX_RANKS && PT->total_rows > 0, RS_E_ARG, "rs_ffm_fwd_peer: bad world / total_rows");
    P.world = PT->world;
    P.direct_mask = PT->direct_mask;
    P.total_rows = PT->total_rows;
    for (int k = 0; k < PT->world; ++k) {
      RS_CHECK_ARG(PT->shard[k], RS_E_ARG, "rs_ffm_fwd_peer: null shard pointer for rank %d", k);
      P.shard[k] = PT->shard[k];
    }
  }
  for (int f = 0; f < F; ++f) {
    const bool direct = (P.direct_mask >> f) & 1ull;
    RS_CHECK_ARG(direct || (T->base[f] && T->rows[f] > 0), RS_E_ARG, "rs_ffm_fwd: table %d missing", f);
    P.base[f] = T->base[f];
    P.rows[f] = T->rows[f];
  }
  P.ids = ids;
  P.cross = cross;
  P.stash = stash;
  P.stash2 = nullptr;
  P.F1 = F, P.F2 = 0;
  if (PT && PT->stash_split) {
    P.stash2 = PT->stash_split;
    P.F1 = P.F2 = 0;
    for (int f = 0; f < F; ++f) {
      if ((PT->split_mask >> f) & 1ull) P.st_slot[f] = (unsigned char)(128 | P.F2++); else P.st_slot[f] = (unsigned char)P.F1++;
    }
    RS_CHECK_ARG(P.F2 > 0 && P.F2 < 128 && (P.F1 == 0 || stash), RS_E_ARG, "rs_ffm_fwd_peer: bad stash split");
  }
  if (F < 64) cold_mask &= (1ull << F) - 1ull;
  P.mini = cold_mask ? mini : nullptr;
  {
    const char *e = getenv("RS_FFM_L2HINT");
    P.hint = !(e && e[0] == '0');
  }
  P.nC = 0;
  for (int f = 0; f < RS_MAX_FIELDS; ++f) P.cidx[f] = (f < F && ((cold_mask >> f) & 1ull)) ? (signed char)P.nC++ : (signed char)-1;
  P.B = B;
  P.F = F;
  P.D = D;
  P.dv = D / 4;
  P.dvs = 0;
  while ((1 << P.dvs) < P.dv) ++P.dvs;
  P.rowv = F * P.dv;
  const int row_bytes = P.rowv * 16;
  int pad = 0;
  if (P.dv < 8) pad = (((P.dv * 16 - row_bytes) % 128) + 128) % 128;  // rows of consecutive j land 16*dv bytes apart mod 128
  P.pitchv = (row_bytes + pad) / 16;
  const size_t stage_bytes = (size_t)F * P.pitchv * 16;
  // Shared memory: 227 KB per SM.  TWO resident CTAs per SM with a 2-stage ring each (four producer and sixteen consumer
  // warps per SM) beat one CTA with a 5-stage ring: 0.73 vs 0.88 ms at the C2 shape -- one producer warp issues a bulk copy
  // every ~70 cycles and each stage hand-shake costs ~430 ns (profiles/tma_rate_micro.cu), so issue parallelism matters
  // more than ring depth.  RS_FFM_CTAS=1 restores the single deep ring.
  int want_ctas = 2;
  if (const char *e = getenv("RS_FFM_CTAS")) want_ctas = atoi(e) == 1 ? 1 : 2;
  const size_t budget = 232448 / want_ctas - 2048 - 1024;      // static shared memory + the per-CTA reservation
  int nst = (int)(budget / stage_bytes);
  if (nst < 2 && want_ctas == 2) nst = (int)((232448 - 2048 - 1024) / stage_bytes);   // tile too large for two CTAs: one deep ring
  RS_CHECK_ARG(nst >= 1, RS_E_UNSUPPORTED, "rs_ffm_fwd: F*F*D tile (%zu B) does not fit in shared memory", stage_bytes);
  if (nst > MAX_STAGES) nst = MAX_STAGES;
  P.nst = nst;
  P.status = status;
  const size_t smem = stage_bytes * nst;
  RS_CUDA(cudaFuncSetAttribute(ffm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ffm_fwd_kernel, NTHREADS, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t cap = (int64_t)rs::num_sms() * per_sm;
  int grid = (int)(B < cap ? B : cap);
  ffm_fwd_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(P);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

RS_API int rs_ffm_dense_fwd(const float *Tin, int64_t B, int32_t F, int32_t NF, int32_t D, const int32_t *field_of, float *cross,
                            void *stream) {
  RS_CHECK_ARG(cross, RS_E_ARG, "rs_ffm_dense_fwd: null output");
  return dense_launch(Tin, nullptr, B, F, NF, D, field_of, cross, nullptr, stream);
}

RS_API int rs_ffm_dense_bwd(const float *Tin, const float *g_cross, int64_t B, int32_t F, int32_t NF, int32_t D,
                            const int32_t *field_of, float *dT, void *stream) {
  RS_CHECK_ARG(g_cross && dT, RS_E_ARG, "rs_ffm_dense_bwd: null argument");
  return dense_launch(Tin, g_cross, B, F, NF, D, field_of, nullptr, dT, stream);
}
