// FFM backward + row update WITHOUT a Jacobian stash (reference: autograd of model/ffm.py:61-82 generalised to F fields,
// embedding_dense_backward and optimizer.step() at trainer/trainer.py:38-39).
//
// The gradient of table row (b, i) is g_b * [v_{j,i}(b)]_j: slot i of the 25 other rows of sample b.  rs_ffm_fwd can write
// that transposed tile to HBM (the "stash", 1664 B per lookup written and read back: 5.6 of the 8.8 GB an FFM step moves
// at F = 26, D = 16, B = 65536).  Here it is recomputed from the table instead, which is possible because the lookups
// split into two classes with different hazards:
//
//   rows of the COLD fields (cold_mask: the large tables) looked up ONCE in the batch -- nearly all of their rows: only
//     their own sample reads them, so the CTA that holds the sample's tile in shared memory (same TMA ring as the
//     forward) updates them in place -- ffm_single_kernel; traffic = tile read + row write, nothing else;
//   every other row (looked up several times, or in a small field): its lookups are reduced in sorted (= ascending
//     position) order, RS_CHUNK at a time exactly as rs_segment_update does, but each lookup's gradient row is GATHERED
//     slice by slice (ffm_multi_kernel): the slices that live in rows of the small fields straight from the table
//     (the lookups of one field need slot i of every small table: a few MB, L2 resident), the slices that live in rows
//     of the cold fields from the "cold-slice stash" rs_ffm_fwd_train wrote (nC of F slices per lookup; a random 64-byte
//     read into a 56 GB table costs a TLB miss each -- measured 1.2 ms for the 10 M of them at the C2 shape).  Results
//     go to a small gradient buffer, because a table row may still be read by other chunks;
//   ffm_apply_kernel + the combine pass of segment.cu then apply the buffered gradients.
//
// Order: multi (reads only) -> single (writes cold-field rows nobody else reads) -> apply (all reads are over).  The
// arithmetic per element is the stash path's (explicit mul, add in sorted order, same chunking, same combine), so the
// two paths produce identical bits.
#include "segment.cuh"

namespace {

using rs::UpdParams;

constexpr int NCW = 8;                 // consumer warps of the single-row pass
constexpr int NPW = 2;                 // producer warps (see ffm.cu)
constexpr int NTHREADS = (NCW + NPW) * 32;
constexpr int MAX_STAGES = 8;

struct TrainParams {
  const float *base[RS_MAX_FIELDS];
  int64_t rows[RS_MAX_FIELDS];
  const int64_t *ids;          // (B, F)
  const float *g;              // (B) dL/dcross
  const int32_t *sorted_pos, *seg_first_chunk, *chunk_start, *chunk_seg, *n_chunks;
  const int64_t *uniq;
  float *partial;              // chunk partials of the multi-chunk segments (rs_segments.partial)
  float *gbuf;                 // (list capacity, W): reduced gradient of list entry k when its segment is a single chunk
  int32_t *list, *n_list;      // chunks that belong to rows looked up more than once
  unsigned long long *single_mask;   // (B): bit i set <=> lookup (b, i) is the only lookup of its row
  float *table;                // concatenated (total_rows, W)
  int64_t B;
  int F, D, dv, dvs, rowv, W, pitchv, nst;
  float lr, wd;
  int32_t *status;
  unsigned long long cold_mask;      // fields whose once-looked-up rows are updated in place / whose slices come from `mini`
  const float *mini;                 // (B, F, nC, D) cold-slice stash written by rs_ffm_fwd_train
  int nC;
  signed char cidx[RS_MAX_FIELDS];
};

__device__ __forceinline__ float upd_sgd1(float w, float g, float lr, float wd) { return w - lr * (g + wd * w); }
__device__ __forceinline__ float4 upd_sgd4(float4 w, float4 g, float lr, float wd) {
  return make_float4(upd_sgd1(w.x, g.x, lr, wd), upd_sgd1(w.y, g.y, lr, wd), upd_sgd1(w.z, g.z, lr, wd), upd_sgd1(w.w, g.w, lr, wd));
}
// acc + v * s with separate roundings (no fma contraction): the stash path's per-lookup arithmetic
__device__ __forceinline__ float4 f4_mul_add_rn(float4 acc, float4 v, float s) {
  return make_float4(__fadd_rn(acc.x, __fmul_rn(v.x, s)), __fadd_rn(acc.y, __fmul_rn(v.y, s)), __fadd_rn(acc.z, __fmul_rn(v.z, s)),
                     __fadd_rn(acc.w, __fmul_rn(v.w, s)));
}

// ------------------------------------------------------------------ classify the chunks
// One thread per chunk: a chunk that is a whole one-lookup segment marks its lookup in single_mask; every other chunk
// joins the work list of the gather pass (warp-aggregated append; the order of the list does not affect any result).
__global__ void __launch_bounds__(256) ffm_classify_kernel(const __grid_constant__ TrainParams P) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = c < *P.n_chunks;
  bool multi = false;
  if (valid) {
    const int s0 = P.chunk_start[c], s1 = P.chunk_start[c + 1];
    const int g = P.chunk_seg[c];
    const bool one_chunk = (P.seg_first_chunk[g + 1] - P.seg_first_chunk[g]) == 1;
    multi = true;
    if (one_chunk && s1 - s0 == 1) {
      const int p = P.sorted_pos[s0];
      const int b = p / P.F;
      const int i = p - b * P.F;
      if ((P.cold_mask >> i) & 1ull) {
        atomicOr(P.single_mask + b, 1ull << i);
        multi = false;
      }
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, multi);
  if (m) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(P.n_list, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (multi) P.list[base + __popc(m & ((1u << lane) - 1u))] = c;
  }
}

// ------------------------------------------------------------------ rows looked up more than once: gather + ordered reduce
// One warp per listed chunk.  Lane l owns float4 columns l, l+32, ... of the gradient row (column w = slot j = w / dv, part
// q = w % dv).  For lookup (b, i) of the chunk, column (j, q) is part q of slot i of the row sample b uses in field j:
// lanes < F fetch the sample's ids and form the row pointers, the others get them by shuffle, then every lane issues its
// NA slice loads; UNR lookups are in flight before their (ordered) adds.
template <int NA, int UNR, bool BIGF>
__global__ void __launch_bounds__(256, (NA * UNR <= 8) ? 3 : 2) ffm_multi_kernel(const __grid_constant__ TrainParams P) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nlist = *P.n_list;
  int jj[NA], qq[NA], cj[NA];
  bool inrow[NA];
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    const int w = lane + 32 * a;
    jj[a] = w >> P.dvs;
    qq[a] = w & (P.dv - 1);
    inrow[a] = w < P.rowv;
    cj[a] = inrow[a] ? P.cidx[jj[a]] : -1;     // >= 0: this column's slices come from the cold-slice stash
  }
  const float *base0 = lane < P.F ? P.base[lane] : nullptr;
  const int64_t rows0 = lane < P.F ? P.rows[lane] : 1;
  const float *base1 = (BIGF && lane + 32 < P.F) ? P.base[lane + 32] : nullptr;
  const int64_t rows1 = (BIGF && lane + 32 < P.F) ? P.rows[lane + 32] : 1;

  for (int k = warp_global; k < nlist; k += nwarps) {
    const int c = P.list[k];
    const int s0 = P.chunk_start[c], s1 = P.chunk_start[c + 1];
    const int gseg = P.chunk_seg[c];
    const bool one_chunk = (P.seg_first_chunk[gseg + 1] - P.seg_first_chunk[gseg]) == 1;
    float4 acc[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) acc[a] = rs::f4_zero();
    for (int sb = s0; sb < s1; sb += 32) {
      const int s = sb + lane;
      const int my_p = s < s1 ? P.sorted_pos[s] : 0;
      const int my_b = my_p / P.F;
      const int my_i = my_p - my_b * P.F;
      const float my_g = s < s1 ? __ldg(P.g + my_b) : 0.f;
      const int cnt = (s1 - sb) < 32 ? (s1 - sb) : 32;
      for (int u0 = 0; u0 < cnt; u0 += UNR) {
        const float *ptr0[UNR], *ptr1[UNR];
        const float *mu[UNR];
        int iu[UNR];
        float gu[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int b = __shfl_sync(0xffffffffu, my_b, u0 + u);
          iu[u] = __shfl_sync(0xffffffffu, my_i, u0 + u);
          gu[u] = __shfl_sync(0xffffffffu, my_g, u0 + u);
          mu[u] = P.mini + ((int64_t)b * P.F + iu[u]) * P.nC * P.D;
          ptr0[u] = ptr1[u] = nullptr;
          if (u0 + u < cnt) {   // (lanes of cold fields need no row pointer)
            if (lane < P.F && !((P.cold_mask >> lane) & 1ull))
              ptr0[u] = base0 + rs::clamp_id(__ldg(P.ids + (int64_t)b * P.F + lane), rows0, P.status) * P.W;
            if (BIGF && lane + 32 < P.F && !((P.cold_mask >> (lane + 32)) & 1ull))
              ptr1[u] = base1 + rs::clamp_id(__ldg(P.ids + (int64_t)b * P.F + lane + 32), rows1, P.status) * P.W;
          }
        }
        float4 val[UNR][NA];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            unsigned long long src = __shfl_sync(0xffffffffu, (unsigned long long)ptr0[u], jj[a] & 31);
            if (BIGF) {
              const unsigned long long hi = __shfl_sync(0xffffffffu, (unsigned long long)ptr1[u], jj[a] & 31);
              if (jj[a] >= 32) src = hi;
            }
            const bool ok = (u0 + u < cnt) && inrow[a] && jj[a] != iu[u];
            const float *from = cj[a] >= 0 ? mu[u] + (cj[a] * P.dv + qq[a]) * 4
                                           : reinterpret_cast<const float *>(src) + (iu[u] * P.dv + qq[a]) * 4;
            val[u][a] = ok ? rs::ldg_nc_f4(from) : rs::f4_zero();
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          if (u0 + u < cnt) {
#pragma unroll
            for (int a = 0; a < NA; ++a) acc[a] = f4_mul_add_rn(acc[a], val[u][a], gu[u]);
          }
        }
      }
    }
    float *dst = one_chunk ? P.gbuf + (int64_t)k * P.W : P.partial + (int64_t)rs::partial_slot(s0, s1 - s0) * P.W;
#pragma unroll
    for (int a = 0; a < NA; ++a)
      if (inrow[a]) rs::stg_f4(dst + (lane + 32 * a) * 4, acc[a]);
  }
}

// ------------------------------------------------------------------ rows looked up once: in-place update from the sample's tile
// Same producer / ring as ffm_fwd_kernel; samples without such a row are skipped.  A stage carries the sample's mask,
// dL/dcross and the global address of each of its rows.
__global__ void __launch_bounds__(NTHREADS, 1) ffm_single_kernel(const __grid_constant__ TrainParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES];
  __shared__ unsigned long long s_mask[MAX_STAGES];
  __shared__ float s_g[MAX_STAGES];
  __shared__ float *s_row[MAX_STAGES][RS_MAX_FIELDS];
  __shared__ const float *s_base[RS_MAX_FIELDS];
  __shared__ int64_t s_rows[RS_MAX_FIELDS];
  for (int i = threadIdx.x; i < P.F; i += blockDim.x) {
    s_base[i] = P.base[i];
    s_rows[i] = P.rows[i];
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
    // ===== producers: lane l of producer warp w owns field f = l * NPW + w =====
    const int f = lane * NPW + warp;
    const bool has = f < P.F;
    const uint32_t nmine = (uint32_t)((P.F - warp + NPW - 1) / NPW);
    int k = 0;      // stages handed out so far
    int64_t b = blockIdx.x;
    int64_t id_cur = 0, id_nxt = 0;
    unsigned long long m_cur = 0, m_nxt = 0;
    float g_cur = 0.f, g_nxt = 0.f;
    if (b < P.B) {
      if (has) id_cur = P.ids[b * P.F + f];
      m_cur = P.single_mask[b];
      g_cur = P.g[b];
    }
    for (; b < P.B; b += gridDim.x) {
      const int64_t bn = b + gridDim.x;
      if (bn < P.B) {
        if (has) id_nxt = P.ids[bn * P.F + f];
        m_nxt = P.single_mask[bn];
        g_nxt = P.g[bn];
      }
      if (m_cur) {
        const int s = k % P.nst;
        const uint32_t ph = (uint32_t)(k / P.nst) & 1u;
        ++k;
        rs::mbar_wait(&empty_bar[s], ph ^ 1u);
        const float *src = nullptr;
        if (has) {
          src = s_base[f] + rs::clamp_id(id_cur, s_rows[f], P.status) * (int64_t)P.rowv * 4;
          s_row[s][f] = const_cast<float *>(src);
        }
        if (warp == 0 && lane == 0) {
          s_mask[s] = m_cur;
          s_g[s] = g_cur;
        }
        __syncwarp();
        if (lane == 0) rs::mbar_arrive_expect_tx(&full_bar[s], row_bytes * nmine);
        __syncwarp();
        if (has) rs::bulk_g2s(tiles + (size_t)s * stage_v + (size_t)f * P.pitchv, src, row_bytes, &full_bar[s]);
      }
      id_cur = id_nxt;
      m_cur = m_nxt;
      g_cur = g_nxt;
    }
    // terminator
    const int s = k % P.nst;
    const uint32_t ph = (uint32_t)(k / P.nst) & 1u;
    rs::mbar_wait(&empty_bar[s], ph ^ 1u);
    if (lane == 0) {
      if (warp == 0) s_mask[s] = 0ull;
      rs::mbar_arrive(&full_bar[s]);
    }
  } else {
    // ===== consumers: warp cw takes the cw-th, (cw + NCW)-th, ... row of the mask =====
    const int cw = warp - NPW;
    for (int k = 0;; ++k) {
      const int s = k % P.nst;
      const uint32_t ph = (uint32_t)(k / P.nst) & 1u;
      rs::mbar_wait(&full_bar[s], ph);
      unsigned long long m = s_mask[s];
      if (!m) break;
      const float g = s_g[s];
      const float4 *T = tiles + (size_t)s * stage_v;
      for (int r = 0; r < cw && m; ++r) m &= m - 1;     // drop the rows of the warps before this one
      while (m) {
        const int i = __ffsll((long long)m) - 1;
        const float4 *Ti = T + (size_t)i * P.pitchv;
        const float4 *Tcol = T + i * P.dv;               // + j*pitchv + q -> v_{j,i}
        float *out = s_row[s][i];
        for (int w0 = lane; w0 < P.rowv; w0 += 128) {
          float4 tr[4], own[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int w = w0 + 32 * u;
            const int j = w >> P.dvs, q = w & (P.dv - 1);
            const bool live = w < P.rowv;
            tr[u] = (live && j != i) ? Tcol[(size_t)j * P.pitchv + q] : rs::f4_zero();
            own[u] = live ? Ti[w] : rs::f4_zero();
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int w = w0 + 32 * u;
            if (w < P.rowv) rs::stg_f4(out + w * 4, upd_sgd4(own[u], f4_mul_add_rn(rs::f4_zero(), tr[u], g), P.lr, P.wd));
          }
        }
        for (int r = 0; r < NCW && m; ++r) m &= m - 1;   // next row of this warp
      }
      __syncwarp();
      if (lane == 0) rs::mbar_arrive(&empty_bar[s]);
    }
  }
}

// ------------------------------------------------------------------ apply the buffered gradients of single-chunk segments
template <int NA>
__global__ void __launch_bounds__(256) ffm_apply_kernel(const __grid_constant__ TrainParams P) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nlist = *P.n_list;
  for (int k = warp_global; k < nlist; k += nwarps) {
    const int c = P.list[k];
    const int gseg = P.chunk_seg[c];
    if ((P.seg_first_chunk[gseg + 1] - P.seg_first_chunk[gseg]) != 1) continue;   // multi-chunk: the combine pass applies it
    float *row = P.table + P.uniq[gseg] * P.W;
    const float *gr = P.gbuf + (int64_t)k * P.W;
    float4 w[NA], g[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const int e = lane + 32 * a;
      if (e < P.rowv) {
        w[a] = rs::ldg_f4(row + e * 4);
        g[a] = rs::ldg_f4(gr + e * 4);
      }
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const int e = lane + 32 * a;
      if (e < P.rowv) rs::stg_f4(row + e * 4, upd_sgd4(w[a], g[a], P.lr, P.wd));
    }
  }
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
struct TrainWs {
  size_t mask, n_list, list, gbuf, total;
  int64_t list_cap;
};
TrainWs train_layout(int64_t n, int64_t B, int W) {
  TrainWs L;
  size_t o = 0;
  L.mask = o, o += align_up((size_t)B * 8);
  L.n_list = o, o += 256;
  L.list = o, o += align_up((size_t)n * 4);
  // a listed chunk holds >= 2 lookups unless it is the tail of a multi-chunk segment (<= n / RS_CHUNK of those)
  L.list_cap = n / 2 + n / RS_CHUNK + 2;
  L.gbuf = o, o += align_up((size_t)L.list_cap * W * 4);
  L.total = o;
  return L;
}

template <int NA>
int launch_train(const TrainParams &P, const UpdParams &U, int64_t n, int mode, cudaStream_t st) {
  const int sms = rs::num_sms();
  RS_CUDA(cudaMemsetAsync(P.single_mask, 0, (size_t)P.B * 8, st));
  RS_CUDA(cudaMemsetAsync(P.n_list, 0, 4, st));
  ffm_classify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P);
  RS_CHECK_LAUNCH();
  constexpr int UNR = NA <= 2 ? 4 : (NA <= 4 ? 4 : 2);
  if (P.F > 32)
    ffm_multi_kernel<NA, UNR, true><<<sms * 2, 256, 0, st>>>(P);
  else
    ffm_multi_kernel<NA, UNR, false><<<sms * ((NA * UNR <= 8) ? 3 : 2), 256, 0, st>>>(P);
  RS_CHECK_LAUNCH();
  const size_t stage_bytes = (size_t)P.F * P.pitchv * 16;
  const size_t smem = stage_bytes * P.nst;
  RS_CUDA(cudaFuncSetAttribute(ffm_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (P.cold_mask) {
    ffm_single_kernel<<<(int)(P.B < sms ? P.B : sms), NTHREADS, smem, st>>>(P);
    RS_CHECK_LAUNCH();
  }
  ffm_apply_kernel<NA><<<sms * 8, 256, 0, st>>>(P);
  RS_CHECK_LAUNCH();
  return rs::dispatch_update(U, mode, n, st);   // combine pass only (U.combine_only): multi-chunk segments, in place
}

}  // namespace

RS_API int rs_ffm_bwd_ws_bytes(int64_t n, int64_t B, int32_t width, size_t *bytes) {
  RS_CHECK_ARG(bytes && n > 0 && B > 0 && width >= 4, RS_E_ARG, "rs_ffm_bwd_ws_bytes: bad argument");
  *bytes = train_layout(n, B, width).total;
  return RS_OK;
}

RS_API int rs_ffm_bwd_update(const rs_tables *T, const int64_t *ids, int64_t B, int32_t D, uint64_t cold_mask, const float *cold_stash,
                             const rs_segments *seg, const rs_update *u, void *ws, size_t ws_bytes, int32_t *status, void *stream) {
  RS_CHECK_ARG(T && ids && seg && u && ws, RS_E_ARG, "rs_ffm_bwd_update: null argument");
  const int F = T->num_fields;
  RS_CHECK_ARG(F >= 2 && F <= RS_MAX_FIELDS, RS_E_SHAPE, "rs_ffm_bwd_update: F=%d out of range", F);
  RS_CHECK_ARG(D >= 4 && D <= 256 && (D & (D - 1)) == 0, RS_E_UNSUPPORTED, "rs_ffm_bwd_update: D=%d must be a power of two in [4,256]", D);
  RS_CHECK_ARG(T->width == F * D && u->width == F * D && u->F == F, RS_E_SHAPE, "rs_ffm_bwd_update: table width must be F*D = %d", F * D);
  RS_CHECK_ARG(!cold_mask || cold_stash, RS_E_ARG, "rs_ffm_bwd_update: cold fields need the cold-slice stash of rs_ffm_fwd_train");
  RS_CHECK_ARG(u->mode == RS_UPD_SGD, RS_E_UNSUPPORTED, "rs_ffm_bwd_update: only RS_UPD_SGD (use the stash + rs_segment_update for other modes)");
  RS_CHECK_ARG(u->table && u->scale && !u->stash && !u->dense, RS_E_ARG,
               "rs_ffm_bwd_update: needs table and scale (= dL/dcross per sample); stash / dense must be NULL");
  const int rowv = F * D / 4;
  RS_CHECK_ARG(rowv <= 8 * 32, RS_E_UNSUPPORTED, "rs_ffm_bwd_update: rows wider than 1024 floats are not built (F*D = %d)", F * D);
  if (B == 0) return RS_OK;
  const int64_t n = B * F;
  RS_CHECK_ARG(n < (1ll << 31), RS_E_SHAPE, "rs_ffm_bwd_update: too many lookups");
  const TrainWs L = train_layout(n, B, F * D);
  RS_CHECK_ARG(ws_bytes >= L.total, RS_E_WORKSPACE, "rs_ffm_bwd_update: workspace too small (%zu < %zu)", ws_bytes, L.total);
  UpdParams U;
  int rc = rs::make_upd_params(seg, n, u, U);
  if (rc) return rc;
  U.combine_only = 1;
  U.scale = nullptr;

  TrainParams P = {};
  for (int f = 0; f < F; ++f) {
    RS_CHECK_ARG(T->base[f] && T->rows[f] > 0, RS_E_ARG, "rs_ffm_bwd_update: table %d missing", f);
    P.base[f] = T->base[f];
    P.rows[f] = T->rows[f];
  }
  char *w = (char *)ws;
  P.ids = ids;
  P.g = u->scale;
  P.sorted_pos = seg->sorted_pos;
  P.seg_first_chunk = seg->seg_first_chunk;
  P.chunk_start = seg->chunk_start;
  P.chunk_seg = seg->chunk_seg;
  P.n_chunks = seg->n_chunks;
  P.uniq = seg->uniq;
  P.partial = seg->partial;
  P.single_mask = (unsigned long long *)(w + L.mask);
  P.n_list = (int32_t *)(w + L.n_list);
  P.list = (int32_t *)(w + L.list);
  P.gbuf = (float *)(w + L.gbuf);
  P.table = u->table;
  P.B = B;
  P.F = F;
  P.D = D;
  P.dv = D / 4;
  P.dvs = 0;
  while ((1 << P.dvs) < P.dv) ++P.dvs;
  P.rowv = rowv;
  P.W = F * D;
  const int row_bytes = rowv * 16;
  int pad = 0;
  if (P.dv < 8) pad = (((P.dv * 16 - row_bytes) % 128) + 128) % 128;   // as ffm_fwd_kernel: conflict-free transposed reads
  P.pitchv = (row_bytes + pad) / 16;
  const size_t stage_bytes = (size_t)F * P.pitchv * 16;
  int nst = (int)((232448 - 6144) / stage_bytes);   // 227 KB per CTA minus the static barriers / pointer tables (5.3 KB)
  RS_CHECK_ARG(nst >= 1, RS_E_UNSUPPORTED, "rs_ffm_bwd_update: F*F*D tile (%zu B) does not fit in shared memory", stage_bytes);
  P.nst = nst > MAX_STAGES ? MAX_STAGES : nst;
  P.lr = u->lr;
  P.wd = u->wd;
  P.status = status;
  P.cold_mask = F < 64 ? (cold_mask & ((1ull << F) - 1ull)) : cold_mask;
  P.mini = cold_stash;
  P.nC = 0;
  for (int f = 0; f < RS_MAX_FIELDS; ++f) P.cidx[f] = (f < F && ((P.cold_mask >> f) & 1ull)) ? (signed char)P.nC++ : (signed char)-1;
  cudaStream_t st = (cudaStream_t)stream;
  if (rowv <= 32) return launch_train<1>(P, U, n, u->mode, st);
  if (rowv <= 64) return launch_train<2>(P, U, n, u->mode, st);
  if (rowv <= 128) return launch_train<4>(P, U, n, u->mode, st);
  return launch_train<8>(P, U, n, u->mode, st);
}
