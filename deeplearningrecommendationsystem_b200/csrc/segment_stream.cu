// Streaming segment-reduce + row update for wide rows (FFM: 416 floats = 1664 B per row).
//
// The generic lane-group kernel in segment.cu is latency bound on wide rows: each warp walks
// metadata -> position -> gradient row -> table row with one or two rows in flight.  Here a persistent CTA per SM
// decouples memory from arithmetic: a producer warp reads one 16-byte record per sorted lookup (lookup_desc, built by
// rs_dedup_sort) and issues cp.async.bulk (TMA bulk copy, SASS UBLKCP) of the 1.6 KB gradient row -- and, for a
// segment that fits one chunk, of the table row it will update -- into a ring of shared-memory stages guarded by
// full/empty mbarriers; four consumer warps own one float4 column each, add the rows of a chunk in sorted
// (= ascending position) order, and at the chunk end write the updated table row (SGD), the dense gradient row,
// or the chunk partial.  Work is handed out in units of ~256 sorted lookups through an atomic counter; units start
// at chunk boundaries, so the summation order -- and therefore every bit of the result -- does not depend on the
// schedule.
#include <cstddef>
#include <stdlib.h>

#include "segment.cuh"

namespace rs {
namespace {

constexpr int SR = 8;          // gradient rows per stage (and table-row slots per stage)
constexpr int NCW = 4;         // consumer warps
constexpr int NCT = NCW * 32;  // consumer threads: thread t owns float4 columns t, t+128, ...
constexpr int NTHREADS = NCT + 32;
constexpr int UNIT = RS_UNIT;   // sorted lookups per work unit
constexpr int MAX_ST = 12;
constexpr int OUT_SLOTS = 64;    // most staging rows of the pusher warp (routed RS_UPD_GRAD); runs of out_slots / 4 rows

struct __align__(16) StageHdr {
  int flags[SR];      // 1 first | 2 last lookup of its chunk | 4 segment is a single chunk
  float scale[SR];
  uint32_t row[SR];   // global table row
  int pslot[SR];      // chunk-partial slot (multi-chunk segments)
  int tslot[SR];      // table-row slot inside the stage (single-chunk segments, SGD)
  float *dst[SR];     // where the finished row of a single-chunk segment goes (table row / gradient row / routed peer row);
                      // resolved once by the producer lane instead of by every consumer thread
  int count, done, pad0, pad1;
};

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ int4 lds_i4(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// out_slots > 0 (routed RS_UPD_GRAD, NA == 1): reduced rows leave through a PUSHER warp.  Remote (NVLink peer) stores issued
// by the consumer warps themselves stall those warps until the link accepts them, which backs the whole ring up -- the
// kernel then costs HBM time PLUS link time (measured at N = 2: 1.7-2.0 ms against 0.95 ms unsharded).  Instead the
// consumers drop each finished row into a shared-memory staging ring and the pusher warp sends RUNS of rows with one
// cp.async.bulk shared -> (peer) global each: consecutive unique rows of a work unit are consecutive rows of the owner's
// buffer, and a bulk store to peer memory costs ~1.4 us of the SM's copy engine whatever its size (measured: one store
// per 1664-byte row ran at 175 GB/s chip-wide), so the rows have to leave several at a time.
// BATCH = stages the producer fills per iteration (BATCH x 8 lanes each own one lookup); the ring needs BATCH + 1 stages
template <int NA, int MODE, bool PUSH, int BATCH>
__global__ void __launch_bounds__(NTHREADS + (PUSH ? 32 : 0), 1) seg_stream_kernel(const __grid_constant__ UpdParams P, int n, int nst, int out_slots) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[MAX_ST], empty_bar[MAX_ST];
  __shared__ uint64_t out_full[OUT_SLOTS], out_empty[OUT_SLOTS];
  __shared__ float *out_dst[OUT_SLOTS];
  __shared__ StageHdr hdr[MAX_ST];
  __shared__ int64_t rtab[ROUTE_TAB];
  if (MODE == RS_UPD_GRAD && P.routes.n > 0) route_tab_load(rtab, P.routes);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int WV = P.W >> 2;
  const uint32_t row_bytes = (uint32_t)P.W * 4u;
  const size_t stage_floats = (size_t)2 * SR * P.W;  // SR gradient rows then SR table rows
  float *ring = reinterpret_cast<float *>(smem_raw);
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NCW);
    }
    for (int s = 0; s < OUT_SLOTS; ++s) {
      mbar_init(&out_full[s], NCW);
      mbar_init(&out_empty[s], 1);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const float *src = P.stash ? P.stash : P.dense;
  float *oring = ring + (size_t)nst * stage_floats;        // OUT_SLOTS staging rows behind the stage ring
  float *const OUT_DONE = reinterpret_cast<float *>(uintptr_t(1));
  constexpr bool pushing = PUSH;   // compile-time: the handshake code must not weigh on the plain instantiations

  if (warp == NCW + 1) {
    // ===================== pusher =====================
    if (pushing) {
      if (lane == 0) {
        const int RUN = out_slots / 4;       // rows per bulk store at most; <= 2 RUN rows are unreleased at any time, so
        int oc = 0, rel = 0;                 // the consumers always find free slots and the pusher can wait for rows
        int run_len = 0, run_slot = 0, prev_rows = 0;
        float *run_dst = nullptr;
        auto flush = [&]() {
          if (run_len == 0) return;
          if (run_dst) bulk_s2g(run_dst, oring + (size_t)run_slot * P.W, row_bytes * (uint32_t)run_len);
          bulk_commit();
          bulk_wait_read<1>();               // every store but the newest has read its staging rows: recycle them
          for (int i = 0; i < prev_rows; ++i, rel = (rel + 1 == out_slots ? 0 : rel + 1)) mbar_arrive(&out_empty[rel]);
          prev_rows = run_len;
          run_len = 0;
        };
        int os = 0;
        uint32_t oph = 0;
        for (;; ++oc, os = (os + 1 == out_slots ? 0 : os + 1), oph ^= (os == 0 ? 1u : 0u)) {
          mbar_wait(&out_full[os], oph);
          float *dst = *reinterpret_cast<float *volatile *>(&out_dst[os]);
          if (dst == OUT_DONE) break;
          const bool extend = run_len > 0 && run_len < RUN && os != 0 && run_dst != nullptr && dst == run_dst + (size_t)run_len * P.W;
          if (extend) {
            ++run_len;
          } else {
            flush();
            run_slot = os, run_dst = dst, run_len = 1;
          }
        }
        flush();
      }
      __syncwarp();
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp == NCW) {
    // ===================== producer =====================
    // The producer is latency bound on its own metadata reads, so they are software pipelined: the record (and
    // the pre-permuted scale) of batch i+1 and the boundaries of the next work unit are requested before batch i
    // is issued.
    int pst = 0;        // next stage of the ring and its phase, advanced without divisions
    uint32_t pph = 0;
    int u = 0;
    if (lane == 0) u = atomicAdd(P.work_counter, 1);
    u = __shfl_sync(0xffffffffu, u, 0);
    const int nunits = n / UNIT + 1;
    // routed rows: every rank starts its walk over the (owner-major) units at its own share, so that at any moment the
    // ranks push to different owners instead of all to the same one (rs_routes.self)
    const int urot = (MODE == RS_UPD_GRAD && P.routes.n > 0) ? (int)route_rotation(P.routes, nunits) : 0;
    while (u < nunits) {
      const int uu = u + urot < nunits ? u + urot : u + urot - nunits;
      const int sA = P.unit_start[uu], sB = P.unit_start[uu + 1];
      int un = 0;                                    // claim the next unit now; its latency hides behind this one
      if (lane == 0) un = atomicAdd(P.work_counter, 1);
      int4 d_nxt = make_int4(0, 0, 0, 0);
      float sc_nxt = 1.0f;
      const bool mine = lane < SR * BATCH;            // lanes beyond the batch carry no lookup
      if (mine && sA + lane < sB) {
        d_nxt = P.lookup_desc[sA + lane];
        if (P.scale) sc_nxt = P.scale_sorted[sA + lane];
      }
      for (int s0 = sA; s0 < sB; s0 += SR * BATCH) {
        const int s = s0 + lane;
        const bool valid = mine && s < sB;
        const int4 d = d_nxt;
        const float sc = sc_nxt;
        const int sn = s + SR * BATCH;               // prefetch the next batch of this unit
        if (mine && sn < sB) {
          d_nxt = P.lookup_desc[sn];
          if (P.scale) sc_nxt = P.scale_sorted[sn];
        }
        const bool want_table = valid && MODE == RS_UPD_SGD && (d.y & 6) == 6;  // last lookup of a single-chunk segment
        float *dptr = nullptr;
        if (valid && (d.y & 6) == 6) {
          const int64_t row = (int64_t)(uint32_t)d.z;
          if (MODE == RS_UPD_GRAD)
            dptr = P.routes.n > 0 ? route_row(P.routes, rtab, row, P.W) : P.dense_grad + row * P.W;
          else
            dptr = P.table + row * P.W;
        }
        const int sub = lane >> 3, e = lane & 7;                                 // stage within the batch, entry within the stage
        const unsigned tmask = __ballot_sync(0xffffffffu, want_table);
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        const unsigned grp = 0xffu << (sub * 8);
        const int tslot = __popc(tmask & grp & ((1u << lane) - 1u));
        const int ntab = __popc(tmask & grp), nval = __popc(vmask & grp);
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          // every lane walks the four stages in order so the empty-waits are warp-uniform
          const int st = pst;
          const uint32_t ph = pph;
          pst = (pst + 1 == nst) ? 0 : pst + 1;
          pph ^= (pst == 0) ? 1u : 0u;
          mbar_wait(&empty_bar[st], ph ^ 1u);
          if (sub == b) {
            if (valid) {
              hdr[st].flags[e] = d.y;
              hdr[st].scale[e] = sc;
              hdr[st].row[e] = (uint32_t)d.z;
              hdr[st].pslot[e] = d.w;
              hdr[st].tslot[e] = tslot;
              hdr[st].dst[e] = dptr;
            }
            if (e == 0) {
              hdr[st].count = nval;
              hdr[st].done = 0;
            }
          }
          __syncwarp();
          if (sub == b && e == 0) mbar_arrive_expect_tx(&full_bar[st], row_bytes * (uint32_t)(nval + ntab));
          __syncwarp();
          if (sub == b && valid) {
            float *dst = ring + (size_t)st * stage_floats;
            bulk_g2s(dst + (size_t)e * P.W, src + (int64_t)d.x * P.W, row_bytes, &full_bar[st]);
            if (want_table) bulk_g2s(dst + (size_t)(SR + tslot) * P.W, P.table + (int64_t)(uint32_t)d.z * P.W, row_bytes, &full_bar[st]);
          }
        }
      }
      u = __shfl_sync(0xffffffffu, un, 0);
    }
    // terminator stage
    const int st = pst;
    const uint32_t ph = pph;
    mbar_wait(&empty_bar[st], ph ^ 1u);
    if (lane == 0) {
      hdr[st].count = 0;
      hdr[st].done = 1;
      mbar_arrive(&full_bar[st]);
    }
  } else {
    // ===================== consumers =====================
    // Everything of a stage that does not depend on the running sums is read first (flags, scales, the eight
    // gradient values of this thread's column): 4 + 8 independent LDS.128, then a short FMUL/FADD chain per entry.
    const int t = threadIdx.x;
    const uint32_t ring_s = smem_u32(ring), hdr_s = smem_u32(&hdr[0]);
    const uint32_t stage_bytes = (uint32_t)stage_floats * 4u;
    float4 acc[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) acc[a] = f4_zero();
    int os = 0;                                  // next staging slot handed to the pusher and its phase (same in every thread)
    uint32_t oph = 0;
    const uint32_t oring_s = smem_u32(oring);
    int st = 0;
    uint32_t ph = 0;
    for (;; st = (st + 1 == nst ? 0 : st + 1), ph ^= (st == 0 ? 1u : 0u)) {
      mbar_wait(&full_bar[st], ph);
      const uint32_t h = hdr_s + (uint32_t)st * (uint32_t)sizeof(StageHdr);
      const int4 tail = lds_i4(h + offsetof(StageHdr, count));
      if (tail.y) break;
      const int count = tail.x;
      int fl[SR];
      float sc[SR];
      *reinterpret_cast<int4 *>(&fl[0]) = lds_i4(h + offsetof(StageHdr, flags));
      *reinterpret_cast<int4 *>(&fl[4]) = lds_i4(h + offsetof(StageHdr, flags) + 16);
      *reinterpret_cast<float4 *>(&sc[0]) = lds_f4(h + offsetof(StageHdr, scale));
      *reinterpret_cast<float4 *>(&sc[4]) = lds_f4(h + offsetof(StageHdr, scale) + 16);
      const uint32_t rows_s = ring_s + (uint32_t)st * stage_bytes;
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const int col = t + a * NCT;
        const bool active = col < WV;
        if (!active && !pushing) continue;       // (when pushing, every consumer thread takes part in the slot handshake)
        const uint32_t col_s = rows_s + (uint32_t)col * 16u;
        float4 v[SR];
#pragma unroll
        for (int e = 0; e < SR; ++e) v[e] = (active && e < count) ? lds_f4(col_s + (uint32_t)e * row_bytes) : f4_zero();
#pragma unroll
        for (int e = 0; e < SR; ++e) {
          if (e < count) {
            // explicit mul then add (no fma contraction): same bits as scale*stash summed sequentially
            acc[a].x = __fadd_rn(acc[a].x, __fmul_rn(v[e].x, sc[e]));
            acc[a].y = __fadd_rn(acc[a].y, __fmul_rn(v[e].y, sc[e]));
            acc[a].z = __fadd_rn(acc[a].z, __fmul_rn(v[e].z, sc[e]));
            acc[a].w = __fadd_rn(acc[a].w, __fmul_rn(v[e].w, sc[e]));
            if (fl[e] & 2) {  // chunk ends here
              if (fl[e] & 4) {
                float *dst;
                asm volatile("ld.shared.u64 %0, [%1];" : "=l"(dst) : "r"(h + (uint32_t)offsetof(StageHdr, dst) + 8u * e));
                if (pushing) {
                  mbar_wait(&out_empty[os], oph ^ 1u);
                  if (active)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(oring_s + (uint32_t)os * row_bytes + (uint32_t)col * 16u),
                                 "f"(acc[a].x), "f"(acc[a].y), "f"(acc[a].z), "f"(acc[a].w)
                                 : "memory");
                  if (t == 0) out_dst[os] = dst;
                  fence_proxy_async();
                  __syncwarp();
                  if (lane == 0) mbar_arrive(&out_full[os]);
                  os = (os + 1 == out_slots) ? 0 : os + 1;
                  oph ^= (os == 0) ? 1u : 0u;
                  acc[a] = f4_zero();
                  continue;
                }
                if (dst) dst += col * 4;   // nullptr: the routed destination is beyond the receiver's capacity (flagged there)
                if (MODE == RS_UPD_SGD) {
                  int ts;
                  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(ts) : "r"(h + (uint32_t)offsetof(StageHdr, tslot) + 4u * e));
                  const float4 w = lds_f4(col_s + (uint32_t)(SR + ts) * row_bytes);
                  stg_f4(dst, upd_sgd(w, acc[a], P));
                } else if (MODE == RS_UPD_GRAD) {
                  if (dst) stg_f4(dst, acc[a]);
                } else {
                  uint32_t row;
                  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(row) : "r"(h + (uint32_t)offsetof(StageHdr, row) + 4u * e));
                  float wv[4], mv[4], vv[4], gv[4];
                  *reinterpret_cast<float4 *>(wv) = ldg_f4(dst);
                  *reinterpret_cast<float4 *>(mv) = ldg_f4(P.m + (int64_t)row * P.W + col * 4);
                  *reinterpret_cast<float4 *>(vv) = ldg_f4(P.v + (int64_t)row * P.W + col * 4);
                  *reinterpret_cast<float4 *>(gv) = acc[a];
#pragma unroll
                  for (int i = 0; i < 4; ++i) adam1(wv[i], mv[i], vv[i], gv[i], P);
                  stg_f4(dst, *reinterpret_cast<float4 *>(wv));
                  stg_f4(P.m + (int64_t)row * P.W + col * 4, *reinterpret_cast<float4 *>(mv));
                  stg_f4(P.v + (int64_t)row * P.W + col * 4, *reinterpret_cast<float4 *>(vv));
                }
              } else {
                int ps;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(ps) : "r"(h + (uint32_t)offsetof(StageHdr, pslot) + 4u * e));
                if (active) stg_f4(P.partial + (int64_t)ps * P.W + col * 4, acc[a]);
              }
              acc[a] = f4_zero();
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
    }
    if (pushing) {   // tell the pusher warp that no more rows will come
      mbar_wait(&out_empty[os], oph ^ 1u);
      if (t == 0) out_dst[os] = OUT_DONE;
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_full[os]);
    }
  }
}

__global__ void __launch_bounds__(256) scale_sorted_kernel(const int4 *__restrict__ desc, const float *__restrict__ scale, int F, int n,
                                                          float *__restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) out[s] = __ldg(scale + desc[s].x / F);
}

template <int NA>
int launch_na(const UpdParams &P, int n, int mode, cudaStream_t st) {
  const size_t stage_bytes = (size_t)2 * SR * P.W * 4;
  // RS_PUSHER=1: routed gradient rows leave through the pusher warp (see the kernel).  Measured at N = 2 on the C2 shape the
  // consumers' own 128-bit stores are faster (1.07 ms vs 1.43 ms: the per-row staging handshake costs more than the link
  // stalls it removes), so direct stores are the default.
  int out_slots = 0;
  if (mode == RS_UPD_GRAD && NA == 1 && P.routes.n > 0 && getenv("RS_PUSHER")) {
    out_slots = (int)((26 * 1024) / ((size_t)P.W * 4)) / 4 * 4;          // ~26 KB of staging rows, a multiple of 4
    out_slots = out_slots > OUT_SLOTS ? OUT_SLOTS : (out_slots < 8 ? 8 : out_slots);
  }
  const size_t out_bytes = (size_t)out_slots * P.W * 4;
  // Two resident CTAs per SM (each: one producer warp, four consumer warps, a ring of >= 3 stages filled two at a time) when
  // two such rings fit in the 227 KB, else one CTA with a deep ring filled four stages at a time.  One producer warp
  // issues a bulk copy every ~70 cycles and pays ~430 ns per stage hand-shake (profiles/tma_rate_micro.cu): two issuing
  // warps per SM stream the 1664-byte rows faster than one, whatever the ring depth.  RS_SEG_CTAS=1 forces one CTA.
  const size_t static_bytes = 6 * 1024;
  int per_sm = (out_slots == 0) ? 2 : 1;
  if (const char *e = getenv("RS_SEG_CTAS")) per_sm = (atoi(e) == 1) ? 1 : per_sm;
  int nst = 0;
  bool small_ring = per_sm == 2;           // half-size ring filled two stages at a time
  if (per_sm == 2) {
    nst = (int)((232448 / 2 - static_bytes - 1024) / stage_bytes);
    if (nst < 3) per_sm = 1, small_ring = false;
    else if (P.half_sm) per_sm = 1;        // same ring, one CTA per SM: the other half of the SM is left to a concurrent kernel
  }
  if (!small_ring) nst = (int)((200 * 1024 - out_bytes) / stage_bytes);
  if (nst > MAX_ST) nst = MAX_ST;
  if (nst < (small_ring ? 3 : 5)) {
    set_error("rs_segment_update: row of %d floats too wide for the streaming kernel", P.W);
    return RS_E_UNSUPPORTED;
  }
  const size_t smem = stage_bytes * nst + out_bytes;
  RS_CUDA(cudaMemsetAsync(P.work_counter, 0, sizeof(int32_t), st));
  if (P.scale) {  // per-sample scale gathered into sorted-lookup order: the producer then reads it coalesced
    scale_sorted_kernel<<<(n + 255) / 256, 256, 0, st>>>(P.lookup_desc, P.scale, P.F, n, P.scale_sorted);
    RS_CHECK_LAUNCH();
  }
  const int grid = num_sms() * per_sm;
#define RS_LAUNCH_STREAM(M, PUSH, BT)                                                                                          \
  do {                                                                                                                         \
    RS_CUDA(cudaFuncSetAttribute(seg_stream_kernel<NA, M, PUSH, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    seg_stream_kernel<NA, M, PUSH, BT><<<grid, NTHREADS + (PUSH ? 32 : 0), smem, st>>>(P, n, nst, out_slots);                  \
  } while (0)
#define RS_LAUNCH_BT(M, PUSH)                                      \
  do {                                                             \
    if (small_ring) RS_LAUNCH_STREAM(M, PUSH, 2); else RS_LAUNCH_STREAM(M, PUSH, 4); \
  } while (0)
  switch (mode) {
    case RS_UPD_GRAD:
      if (NA == 1 && out_slots > 0) RS_LAUNCH_STREAM(RS_UPD_GRAD, (NA == 1), 4); else RS_LAUNCH_BT(RS_UPD_GRAD, false);
      break;
    case RS_UPD_SGD: RS_LAUNCH_BT(RS_UPD_SGD, false); break;
    default: RS_LAUNCH_BT(RS_UPD_ADAM, false); break;
  }
#undef RS_LAUNCH_BT
#undef RS_LAUNCH_STREAM
  return RS_OK;
}

}  // namespace

int launch_seg_stream(const UpdParams &P, int64_t n, int mode, cudaStream_t st) {
  const int wv = P.W / 4;
  if (wv <= NCT) return launch_na<1>(P, (int)n, mode, st);
  if (wv <= 2 * NCT) return launch_na<2>(P, (int)n, mode, st);
  return launch_na<4>(P, (int)n, mode, st);
}

}  // namespace rs
