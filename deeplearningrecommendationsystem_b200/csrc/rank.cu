// Catalogue ranking: segmented descending top-k / full argsort, and the fused MF score + rank.
//
// Replaces the per-user python loop of the reference's recommendation() methods
// (model/deepfm.py:85-95 `torch.topk(scores, k, dim=0)` once per user; model/mf.py:28-35 `U @ V^T` + topk).
// One CTA owns one user (segment): the scores become 64-bit keys (order-preserving score bits, position) in shared
// memory, a bitonic network sorts them, and the first k positions are written out.  Ties rank the lower position
// first and NaN ranks above everything (torch.topk's convention), so the result is a pure function of the scores.
// Segments longer than the slot buffer are streamed through it, carrying the running best k.
// When k is a small part of the segment the full sort is replaced by a radix select (8-bit digits of the same keys,
// warp-aggregated shared-memory histogram) that leaves the k best unordered, and only those k are sorted.
#include <stdlib.h>

#include "common.cuh"

namespace rs {
namespace {

constexpr int RK_MAX_SLOTS = 16384;  // 128 KB of keys
constexpr uint64_t RK_PAD = ~0ull;

__device__ __forceinline__ uint64_t rank_key(float s, uint32_t pos) {
  uint32_t o;
  if (s != s) {
    o = 0xffffffffu;
  } else {
    uint32_t u = __float_as_uint(s + 0.0f);  // -0 -> +0
    o = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
  return ((uint64_t)(~o) << 32) | pos;
}
__device__ __forceinline__ float key_score(uint64_t key) {
  uint32_t o = ~(uint32_t)(key >> 32);
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

// the cooperating threads of a sort / select: the whole CTA, or one warp (the tiled MF kernel runs one user per warp)
struct BlockGroup {
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};
struct WarpGroup {
  __device__ __forceinline__ int tid() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int size() const { return 32; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};

template <class G>
__device__ void bitonic_sort(const G g, uint64_t *a, int P) {
  for (int k2 = 2; k2 <= P; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = g.tid(); t < (P >> 1); t += g.size()) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int l = i | j;
        bool up = (i & k2) == 0;
        uint64_t x = a[i], y = a[l];
        if ((x > y) == up) {
          a[i] = y;
          a[l] = x;
        }
      }
      g.sync();
    }
  }
}

// ---- radix select: out[0..k) = the k smallest keys of a[0..n) in no particular order (keys are distinct; n >= k).
// sh: 256 histogram bins + 4 words of scratch (per group).
template <class G>
__device__ void select_smallest(const G g, const uint64_t *a, int n, int k, uint64_t *out, int *sh) {
  int *hist = sh, *misc = sh + 256;
  const int lane = threadIdx.x & 31;
  if (n == k) {
    for (int i = g.tid(); i < n; i += g.size()) out[i] = a[i];
    g.sync();
    return;
  }
  uint64_t prefix = 0;  // the digits chosen so far (the top 64 - shift bits of the threshold key)
  int need = k, shift = 64;
  const int n32 = (n + 31) & ~31;
  while (true) {
    shift -= 8;
    for (int i = g.tid(); i < 256; i += g.size()) hist[i] = 0;
    g.sync();
    for (int i = g.tid(); i < n32; i += g.size()) {
      int bin = 256 + lane;  // lanes with nothing to count form singleton groups
      if (i < n) {
        uint64_t key = a[i];
        if (shift == 56 || (key >> (shift + 8)) == prefix) bin = (int)((key >> shift) & 255);
      }
      unsigned peers = __match_any_sync(0xffffffffu, bin);
      if (bin < 256 && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
    }
    g.sync();
    if (g.tid() < 32) {  // one warp locates the bin holding the need-th smallest candidate
      int loc[8], s = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        loc[j] = hist[8 * lane + j];
        s += loc[j];
      }
      int inc = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      int before = inc - s;
      if (before < need && need <= inc) {  // exactly one lane
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (need <= before + loc[j]) {
            misc[0] = 8 * lane + j;
            misc[1] = before;
            misc[2] = loc[j];
            break;
          }
          before += loc[j];
        }
      }
    }
    g.sync();
    const int bin = misc[0], cnt = misc[2];
    need -= misc[1];
    prefix = (prefix << 8) | (uint64_t)bin;
    g.sync();
    if (cnt == need || shift == 0) break;  // everything in that bin is wanted
  }
  if (g.tid() == 0) misc[3] = 0;
  g.sync();
  for (int i = g.tid(); i < n; i += g.size()) {
    uint64_t key = a[i];
    if ((key >> shift) <= prefix) out[atomicAdd(&misc[3], 1)] = key;
  }
  g.sync();
}

struct ArrayScorer {
  const float *s;  // this segment's scores
  __device__ __forceinline__ void fill(uint64_t *dst, int64_t first, int take) const {
    for (int i = threadIdx.x; i < take; i += blockDim.x) dst[i] = rank_key(s[first + i], (uint32_t)(first + i));
  }
};

// dot(user row in smem, item row): 8 lanes per item row, float4 loads, fixed shuffle tree
struct DotScorer {
  const float *u;     // smem, W floats
  const float *rows;  // item table
  int W;
  __device__ __forceinline__ void fill(uint64_t *dst, int64_t first, int take) const {
    const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3, ngrp = blockDim.x >> 3;
    const int nv = (W & 3) ? 0 : (W >> 2);  // rows are 16-byte aligned only when W % 4 == 0
    const int rounds = (take + ngrp - 1) / ngrp;
    for (int r = 0; r < rounds; ++r) {
      int i = r * ngrp + grp;
      float acc = 0.f;
      if (i < take) {
        const float *row = rows + (first + i) * (int64_t)W;
        for (int v = sub; v < nv; v += 8) acc += f4_dot(ldg_f4(row + 4 * v), *reinterpret_cast<const float4 *>(u + 4 * v));
        for (int d = 4 * nv + sub; d < W; d += 8) acc = fmaf(row[d], u[d], acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (i < take && sub == 0) dst[i] = rank_key(acc, (uint32_t)(first + i));
    }
  }
};

// P slots; P2 == 0: sort everything (full ranking).  P2 > 0: select the best k of every round into `best` (P2 >= k
// keys), sort only those at the end.
template <class Scorer>
__device__ void rank_one(const Scorer &sc, int64_t len, int k, int P, int P2, uint64_t *slots, uint64_t *best, int *sh,
                         int64_t *out_idx, float *out_val, int32_t *status) {
  if (len < k) {  // torch.topk raises "selected index k out of range"; the host wrapper does the same from status
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      out_idx[i] = -1;
      if (out_val) out_val[i] = 0.f;
    }
    if (threadIdx.x == 0) atomicOr(status, 2);
    return;
  }
  int64_t done = 0;
  int keep = 0;
  while (done < len) {
    const int room = P - keep;
    if (room <= 0) {  // segment longer than the declared bound with k == P
      if (threadIdx.x == 0) atomicOr(status, 4);
      break;
    }
    const int take = (int)((len - done) < (int64_t)room ? (len - done) : (int64_t)room);
    sc.fill(slots + keep, done, take);
    done += take;
    if (P2 == 0) {
      for (int i = keep + take + threadIdx.x; i < P; i += blockDim.x) slots[i] = RK_PAD;
      __syncthreads();
      bitonic_sort(BlockGroup(), slots, P);
      keep = keep + take < k ? keep + take : k;
    } else {
      __syncthreads();
      select_smallest(BlockGroup(), slots, keep + take, k, best, sh);
      keep = k;
      if (done < len) {  // carry the running best into the next round
        for (int i = threadIdx.x; i < k; i += blockDim.x) slots[i] = best[i];
        __syncthreads();
      }
    }
  }
  const uint64_t *res = slots;
  if (P2 > 0) {
    for (int i = k + threadIdx.x; i < P2; i += blockDim.x) best[i] = RK_PAD;
    __syncthreads();
    bitonic_sort(BlockGroup(), best, P2);
    res = best;
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    uint64_t key = res[i];
    out_idx[i] = (int64_t)(uint32_t)key;
    if (out_val) out_val[i] = key_score(key);
  }
}

__global__ void __launch_bounds__(1024) rank_segments_kernel(const float *__restrict__ scores, const int64_t *__restrict__ seg_start,
                                                             int64_t uniform_len, int k, int P, int P2,
                                                             int64_t *__restrict__ out_idx, float *__restrict__ out_val,
                                                             int32_t *status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *slots = reinterpret_cast<uint64_t *>(smem_raw);
  uint64_t *best = slots + P;
  int *sh = reinterpret_cast<int *>(best + P2);
  const int64_t s = blockIdx.x;
  const int64_t b = seg_start ? seg_start[s] : s * uniform_len;
  const int64_t e = seg_start ? seg_start[s + 1] : b + uniform_len;
  ArrayScorer sc{scores + b};
  rank_one(sc, e - b, k, P, P2, slots, best, sh, out_idx + s * k, out_val ? out_val + s * k : nullptr, status);
}

__global__ void __launch_bounds__(1024) mf_rank_kernel(const float *__restrict__ users, const float *__restrict__ items, int64_t num_items,
                                                       int W, int k, int P, int P2, int64_t *__restrict__ out_idx,
                                                       float *__restrict__ out_val, int32_t *status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *slots = reinterpret_cast<uint64_t *>(smem_raw);
  uint64_t *best = slots + P;
  int *sh = reinterpret_cast<int *>(best + P2);
  float *u = reinterpret_cast<float *>(sh + 264);  // 16-byte aligned: P, P2 are multiples of 64
  const int64_t s = blockIdx.x;
  for (int d = threadIdx.x; d < W; d += blockDim.x) u[d] = users[s * W + d];
  __syncthreads();
  DotScorer sc{u, items, W};
  rank_one(sc, num_items, k, P, P2, slots, best, sh, out_idx + s * k, out_val ? out_val + s * k : nullptr, status);
}

// ---- tiled MF top-k: 8 users per CTA share every item load, so the item table crosses L2 -> SM once per 8 users.
// Items come from a transposed copy (W, ldt) so that a warp's loads of one coordinate are one coalesced line; each
// thread scores 2 items x 8 users per pass from registers, and a score is kept only if it beats the user's current
// k-th best (tau), so after the first few hundred items appends are rare.  Warp w owns user w's candidate list: when
// the list cannot take another pass it is cut back to the best k by the warp-level radix select.
constexpr int MT_USERS = 8, MT_THREADS = 256, MT_CHUNK = 2 * MT_THREADS;

__global__ void transpose_pad_kernel(const float *__restrict__ in, int64_t rows, int W, float *__restrict__ out, int64_t ldt) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = r0 + j;
    int c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < W) ? in[r * W + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j;
    int64_t r = r0 + threadIdx.x;
    if (c < W && r < ldt) out[(int64_t)c * ldt + r] = tile[threadIdx.x][j];
  }
}

__device__ __forceinline__ void mt_shrink(uint64_t *cand, uint64_t *best, int *sh, int *cnt, uint64_t *tau, int k) {
  const int lane = threadIdx.x & 31;
  const int n = *cnt;
  __syncwarp();
  if (n <= k) return;
  select_smallest(WarpGroup(), cand, n, k, best, sh);
  uint64_t worst = 0;
  for (int i = lane; i < k; i += 32) {
    uint64_t key = best[i];
    cand[i] = key;
    worst = key > worst ? key : worst;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t t = __shfl_xor_sync(0xffffffffu, worst, o);
    worst = t > worst ? t : worst;
  }
  if (lane == 0) {
    *cnt = k;
    *tau = worst;  // a later item must beat the current k-th best
  }
  __syncwarp();
}

__global__ void __launch_bounds__(MT_THREADS) mf_topk_tiled_kernel(const float *__restrict__ users, const float *__restrict__ itemsT,
                                                                    int64_t ldt, int64_t num_users, int64_t num_items, int W, int k,
                                                                    int P2, int Pu, int64_t *__restrict__ out_idx,
                                                                    float *__restrict__ out_val) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *cand = reinterpret_cast<uint64_t *>(smem_raw);            // [MT_USERS][Pu]
  uint64_t *best = cand + (size_t)MT_USERS * Pu;                      // [MT_USERS][P2]
  uint64_t *tau = best + (size_t)MT_USERS * P2;                       // [MT_USERS]
  int *sh = reinterpret_cast<int *>(tau + MT_USERS);                  // [MT_USERS][264]
  int *cnt = sh + MT_USERS * 264;                                     // [MT_USERS]
  float *uvec = reinterpret_cast<float *>(cnt + MT_USERS);            // [MT_USERS][W]
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t u0 = (int64_t)blockIdx.x * MT_USERS;
  for (int i = tid; i < MT_USERS * W; i += MT_THREADS) {
    int64_t u = u0 + i / W;
    uvec[i] = u < num_users ? users[u * W + i % W] : 0.f;
  }
  if (tid < MT_USERS) {
    cnt[tid] = 0;
    tau[tid] = (u0 + tid < num_users) ? RK_PAD : 0ull;  // absent users accept nothing
  }
  __syncthreads();
  for (int64_t base = 0; base < num_items; base += MT_CHUNK) {
    if (cnt[warp] + MT_CHUNK > Pu) mt_shrink(cand + (size_t)warp * Pu, best + (size_t)warp * P2, sh + warp * 264, cnt + warp, tau + warp, k);
    __syncthreads();
    const int64_t i0 = base + tid, i1 = i0 + MT_THREADS;
    float a0[MT_USERS], a1[MT_USERS];
#pragma unroll
    for (int u = 0; u < MT_USERS; ++u) a0[u] = a1[u] = 0.f;
    const float *col = itemsT + i0;
#pragma unroll 2
    for (int d = 0; d < W; d += 4) {
      float x0[4], x1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        x0[j] = __ldg(col + (int64_t)(d + j) * ldt);
        x1[j] = __ldg(col + (int64_t)(d + j) * ldt + MT_THREADS);
      }
#pragma unroll
      for (int u = 0; u < MT_USERS; ++u) {
        const float4 uv = *reinterpret_cast<const float4 *>(uvec + u * W + d);
        a0[u] = fmaf(x0[3], uv.w, fmaf(x0[2], uv.z, fmaf(x0[1], uv.y, fmaf(x0[0], uv.x, a0[u]))));
        a1[u] = fmaf(x1[3], uv.w, fmaf(x1[2], uv.z, fmaf(x1[1], uv.y, fmaf(x1[0], uv.x, a1[u]))));
      }
    }
#pragma unroll
    for (int u = 0; u < MT_USERS; ++u) {
      const uint64_t t = tau[u];
      if (i0 < num_items) {
        uint64_t key = rank_key(a0[u], (uint32_t)i0);
        if (key < t) cand[(size_t)u * Pu + atomicAdd(&cnt[u], 1)] = key;
      }
      if (i1 < num_items) {
        uint64_t key = rank_key(a1[u], (uint32_t)i1);
        if (key < t) cand[(size_t)u * Pu + atomicAdd(&cnt[u], 1)] = key;
      }
    }
    __syncthreads();
  }
  // warp w finishes user w: cut to k, sort, write
  uint64_t *mine = cand + (size_t)warp * Pu, *mybest = best + (size_t)warp * P2;
  mt_shrink(mine, mybest, sh + warp * 264, cnt + warp, tau + warp, k);
  const int lane = tid & 31;
  for (int i = lane; i < P2; i += 32) mybest[i] = i < k ? mine[i] : RK_PAD;
  __syncwarp();
  bitonic_sort(WarpGroup(), mybest, P2);
  if (u0 + warp < num_users) {
    for (int i = lane; i < k; i += 32) {
      uint64_t key = mybest[i];
      out_idx[(u0 + warp) * k + i] = (int64_t)(uint32_t)key;
      if (out_val) out_val[(u0 + warp) * k + i] = key_score(key);
    }
  }
}

struct TiledPlan {
  bool use;
  int P2, Pu;
  int64_t ldt;
  size_t smem, ws_bytes;
};
// the tiled kernel serves "few of many": k <= 512, k a small part of the catalogue, enough users to fill the machine
TiledPlan plan_tiled(int64_t num_users, int64_t num_items, int width, int k) {
  TiledPlan t = {};
  if (k < 1 || k > 512 || num_items < 4096 || num_users < 64 || (width & 3) || width > 512 || getenv("RS_RANK_NO_TILED") ||
      getenv("RS_RANK_NO_SELECT"))
    return t;
  int P2 = 64;
  while (P2 < k) P2 <<= 1;
  if ((int64_t)P2 * 16 > num_items) return t;
  t.use = true;
  t.P2 = P2;
  t.Pu = 4 * P2 > 1024 ? 4 * P2 : 1024;
  t.ldt = (num_items + MT_CHUNK - 1) / MT_CHUNK * MT_CHUNK;
  t.smem = (size_t)MT_USERS * (t.Pu + P2 + 1) * sizeof(uint64_t) + (size_t)MT_USERS * 265 * sizeof(int) +
           (size_t)MT_USERS * width * sizeof(float);
  t.ws_bytes = (size_t)width * t.ldt * sizeof(float);
  return t;
}

// slot buffer P, "best" buffer P2 (0 = full sort), threads, dynamic shared memory without the scorer's own part
int pick_slots(int64_t max_len, int k, int *P_out, int *P2_out, int *threads_out, size_t *smem_out, const char *who) {
  RS_CHECK_ARG(k >= 1, RS_E_ARG, "%s: k=%d must be >= 1", who, k);
  RS_CHECK_ARG(max_len >= 1, RS_E_ARG, "%s: max_len=%lld must be >= 1", who, (long long)max_len);
  int64_t want = max_len > k ? max_len : k;
  int P = 64;
  while (P < want && P < RK_MAX_SLOTS) P <<= 1;
  // a segment longer than the slot buffer is streamed through it, which needs room beside the running best k
  RS_CHECK_ARG(want <= P || k <= P / 2, RS_E_UNSUPPORTED,
               "%s: k=%d with segments of %lld needs more than %d slots (k <= %d when segments exceed %d)", who, k,
               (long long)max_len, RK_MAX_SLOTS, RK_MAX_SLOTS / 2, RK_MAX_SLOTS);
  int P2 = 64;
  while (P2 < k) P2 <<= 1;
  if (P2 * 4 > P || getenv("RS_RANK_NO_SELECT")) P2 = 0;  // selecting pays when k is a small part of the segment
  *P_out = P;
  *P2_out = P2;
  int t = P / 2;
  t = t > 1024 ? 1024 : t;
  *threads_out = t < 64 ? 64 : t;
  *smem_out = (size_t)(P + P2) * sizeof(uint64_t) + 264 * sizeof(int);
  return 0;
}

}  // namespace
}  // namespace rs

RS_API int rs_rank_segments(const float *scores, const int64_t *seg_start, int64_t num_segments, int64_t max_len, int32_t k,
                            int64_t *out_idx, float *out_val, int32_t *status, void *stream) {
  using namespace rs;
  RS_CHECK_ARG(scores && out_idx && status, RS_E_ARG, "rs_rank_segments: null scores/out_idx/status");
  RS_CHECK_ARG(num_segments >= 0 && num_segments < (1ll << 31), RS_E_ARG, "rs_rank_segments: num_segments=%lld",
               (long long)num_segments);
  if (num_segments == 0) return 0;
  int P, P2, threads;
  size_t smem;
  if (int rc = pick_slots(max_len, k, &P, &P2, &threads, &smem, "rs_rank_segments")) return rc;
  RS_CUDA(cudaFuncSetAttribute(rank_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rank_segments_kernel<<<(unsigned)num_segments, threads, smem, (cudaStream_t)stream>>>(scores, seg_start, max_len, k, P, P2,
                                                                                         out_idx, out_val, status);
  RS_CHECK_LAUNCH();
  return 0;
}

RS_API int rs_mf_rank_ws_bytes(int64_t num_users, int64_t num_items, int32_t width, int32_t k, size_t *bytes) {
  using namespace rs;
  RS_CHECK_ARG(bytes, RS_E_ARG, "rs_mf_rank_ws_bytes: null bytes");
  *bytes = plan_tiled(num_users, num_items, width, k).ws_bytes;
  return 0;
}

RS_API int rs_mf_rank(const float *user_rows, const float *item_rows, int64_t num_users, int64_t num_items, int32_t width, int32_t k,
                      int64_t *out_idx, float *out_val, int32_t *status, void *ws, size_t ws_bytes, void *stream) {
  using namespace rs;
  RS_CHECK_ARG(user_rows && item_rows && out_idx && status, RS_E_ARG, "rs_mf_rank: null rows/out_idx/status");
  RS_CHECK_ARG(width >= 1 && width <= 4096, RS_E_UNSUPPORTED, "rs_mf_rank: width=%d outside [1, 4096]", width);
  RS_CHECK_ARG(num_users >= 0 && num_users < (1ll << 31) && num_items < (1ll << 32) - 1, RS_E_ARG,
               "rs_mf_rank: num_users=%lld num_items=%lld", (long long)num_users, (long long)num_items);
  if (num_users == 0) return 0;
  TiledPlan t = plan_tiled(num_users, num_items, width, k);
  if (t.use && num_items >= k) {
    RS_CHECK_ARG(ws && ws_bytes >= t.ws_bytes, RS_E_WORKSPACE, "rs_mf_rank: workspace of %zu bytes needed, got %zu", t.ws_bytes,
                 ws_bytes);
    float *itemsT = static_cast<float *>(ws);
    dim3 tg((unsigned)(t.ldt / 32), (unsigned)((width + 31) / 32));
    transpose_pad_kernel<<<tg, dim3(32, 8), 0, (cudaStream_t)stream>>>(item_rows, num_items, width, itemsT, t.ldt);
    RS_CHECK_LAUNCH();
    RS_CUDA(cudaFuncSetAttribute(mf_topk_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem));
    unsigned grid = (unsigned)((num_users + MT_USERS - 1) / MT_USERS);
    mf_topk_tiled_kernel<<<grid, MT_THREADS, t.smem, (cudaStream_t)stream>>>(user_rows, itemsT, t.ldt, num_users, num_items, width, k,
                                                                             t.P2, t.Pu, out_idx, out_val);
    RS_CHECK_LAUNCH();
    return 0;
  }
  int P, P2, threads;
  size_t smem;
  if (int rc = pick_slots(num_items, k, &P, &P2, &threads, &smem, "rs_mf_rank")) return rc;
  smem += (size_t)width * sizeof(float);
  RS_CUDA(cudaFuncSetAttribute(mf_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mf_rank_kernel<<<(unsigned)num_users, threads, smem, (cudaStream_t)stream>>>(user_rows, item_rows, num_items, width, k, P, P2,
                                                                                out_idx, out_val, status);
  RS_CHECK_LAUNCH();
  return 0;
}
