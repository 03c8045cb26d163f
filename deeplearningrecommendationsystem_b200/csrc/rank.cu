// Catalogue ranking: segmented descending top-k / full argsort, and the fused MF score + rank.
//
// Replaces the per-user python loop of the reference's recommendation() methods
// (model/deepfm.py:85-95 `torch.topk(scores, k, dim=0)` once per user; model/mf.py:28-35 `U @ V^T` + topk).
// One CTA owns one user (segment): the scores become 64-bit keys (order-preserving score bits, position) in shared
// memory, a bitonic network sorts them, and the first k positions are written out.  Ties rank the lower position
// first and NaN ranks above everything (torch.topk's convention), so the result is a pure function of the scores.
// Segments longer than the slot buffer are streamed through it, carrying the running best k.
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

__device__ void bitonic_sort(uint64_t *a, int P) {
  for (int k2 = 2; k2 <= P; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int l = i | j;
        bool up = (i & k2) == 0;
        uint64_t x = a[i], y = a[l];
        if ((x > y) == up) {
          a[i] = y;
          a[l] = x;
        }
      }
      __syncthreads();
    }
  }
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

template <class Scorer>
__device__ void rank_one(const Scorer &sc, int64_t len, int k, int P, uint64_t *slots, int64_t *out_idx, float *out_val,
                         int32_t *status) {
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
    for (int i = keep + take + threadIdx.x; i < P; i += blockDim.x) slots[i] = RK_PAD;
    __syncthreads();
    bitonic_sort(slots, P);
    done += take;
    keep = keep + take < k ? keep + take : k;
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    uint64_t key = slots[i];
    out_idx[i] = (int64_t)(uint32_t)key;
    if (out_val) out_val[i] = key_score(key);
  }
}

__global__ void __launch_bounds__(1024) rank_segments_kernel(const float *__restrict__ scores, const int64_t *__restrict__ seg_start,
                                                             int64_t uniform_len, int k, int P, int64_t *__restrict__ out_idx,
                                                             float *__restrict__ out_val, int32_t *status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *slots = reinterpret_cast<uint64_t *>(smem_raw);
  const int64_t s = blockIdx.x;
  const int64_t b = seg_start ? seg_start[s] : s * uniform_len;
  const int64_t e = seg_start ? seg_start[s + 1] : b + uniform_len;
  ArrayScorer sc{scores + b};
  rank_one(sc, e - b, k, P, slots, out_idx + s * k, out_val ? out_val + s * k : nullptr, status);
}

__global__ void __launch_bounds__(1024) mf_rank_kernel(const float *__restrict__ users, const float *__restrict__ items, int64_t num_items,
                                                       int W, int k, int P, int64_t *__restrict__ out_idx, float *__restrict__ out_val,
                                                       int32_t *status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *slots = reinterpret_cast<uint64_t *>(smem_raw);
  float *u = reinterpret_cast<float *>(slots + P);
  const int64_t s = blockIdx.x;
  for (int d = threadIdx.x; d < W; d += blockDim.x) u[d] = users[s * W + d];
  __syncthreads();
  DotScorer sc{u, items, W};
  rank_one(sc, num_items, k, P, slots, out_idx + s * k, out_val ? out_val + s * k : nullptr, status);
}

int pick_slots(int64_t max_len, int k, int *P_out, int *threads_out, const char *who) {
  RS_CHECK_ARG(k >= 1, RS_E_ARG, "%s: k=%d must be >= 1", who, k);
  RS_CHECK_ARG(max_len >= 1, RS_E_ARG, "%s: max_len=%lld must be >= 1", who, (long long)max_len);
  int64_t want = max_len > k ? max_len : k;
  int P = 64;
  while (P < want && P < RK_MAX_SLOTS) P <<= 1;
  // a segment longer than the slot buffer is streamed through it, which needs room beside the running best k
  RS_CHECK_ARG(want <= P || k <= P / 2, RS_E_UNSUPPORTED,
               "%s: k=%d with segments of %lld needs more than %d slots (k <= %d when segments exceed %d)", who, k,
               (long long)max_len, RK_MAX_SLOTS, RK_MAX_SLOTS / 2, RK_MAX_SLOTS);
  *P_out = P;
  int t = P / 2;
  *threads_out = t > 1024 ? 1024 : t;
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
  int P, threads;
  if (int rc = pick_slots(max_len, k, &P, &threads, "rs_rank_segments")) return rc;
  size_t smem = (size_t)P * sizeof(uint64_t);
  RS_CUDA(cudaFuncSetAttribute(rank_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rank_segments_kernel<<<(unsigned)num_segments, threads, smem, (cudaStream_t)stream>>>(scores, seg_start, max_len, k, P, out_idx,
                                                                                         out_val, status);
  RS_CHECK_LAUNCH();
  return 0;
}

RS_API int rs_mf_rank(const float *user_rows, const float *item_rows, int64_t num_users, int64_t num_items, int32_t width, int32_t k,
                      int64_t *out_idx, float *out_val, int32_t *status, void *stream) {
  using namespace rs;
  RS_CHECK_ARG(user_rows && item_rows && out_idx && status, RS_E_ARG, "rs_mf_rank: null rows/out_idx/status");
  RS_CHECK_ARG(width >= 1 && width <= 4096, RS_E_UNSUPPORTED, "rs_mf_rank: width=%d outside [1, 4096]", width);
  RS_CHECK_ARG(num_users >= 0 && num_users < (1ll << 31) && num_items < (1ll << 32) - 1, RS_E_ARG,
               "rs_mf_rank: num_users=%lld num_items=%lld", (long long)num_users, (long long)num_items);
  if (num_users == 0) return 0;
  int P, threads;
  if (int rc = pick_slots(num_items, k, &P, &threads, "rs_mf_rank")) return rc;
  if (threads < 64) threads = 64;
  size_t smem = (size_t)P * sizeof(uint64_t) + (size_t)width * sizeof(float);
  RS_CUDA(cudaFuncSetAttribute(mf_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mf_rank_kernel<<<(unsigned)num_users, threads, smem, (cudaStream_t)stream>>>(user_rows, item_rows, num_items, width, k, P, out_idx,
                                                                                out_val, status);
  RS_CHECK_LAUNCH();
  return 0;
}
