// AFM attention pooling forward on the tensor cores (reference model/afm.py:55-65):
//   P_p = e_i * e_j (p = (i<j) pair),  s_p = h . relu(P_p W + b),  w = softmax_p(s),  pooled = sum_p w_p P_p.
// The projection P W over the F(F-1)/2 pairs of a sample is a (pairs x D) x (D x A) GEMM; here it runs as tcgen05
// 3xTF32 MMAs, 128 pairs per tile.  A persistent CTA holds two warpgroups that take alternate SAMPLES; each keeps the
// sample's F x D embeddings in shared memory (next sample prefetched with cp.async), forms the pair products of a
// tile directly in the A-operand layout (thread r = pair r of the tile, tf32 hi/lo split), issues that tile's MMAs
// into one of two TMEM accumulators, and reads the previous tile's accumulator back (bias, ReLU, dot with h) while
// they run.  Softmax over the sample's pairs and the weighted sum stay on the CUDA cores (they are O(P D), the
// projection is O(P D A)).  W is resident as an (N = A, K = D) hi/lo core-matrix image.
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace {

using namespace rs::tc;
constexpr int NTH = 256;

struct AfmTcParams {
  const float *E, *W, *bvec, *h;
  float *pooled, *attw;
  int64_t B;
  int F, D, A, NP, tmem_cols;
};

__device__ __forceinline__ void pair_of(int F, int p, int &i, int &j) {
  i = 0;
  while (p >= F - 1 - i) {
    p -= F - 1 - i;
    ++i;
  }
  j = i + 1 + p;
}

__global__ void __launch_bounds__(NTH, 1) afm_fwd_tc_kernel(const __grid_constant__ AfmTcParams P) {
  extern __shared__ __align__(128) uint32_t sm[];
  __shared__ uint64_t bar[2][2];   // [warpgroup][accumulator / operand buffer]
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = threadIdx.x >> 7, tid = threadIdx.x & (MT - 1), w4 = warp & 3;
  const int F = P.F, D = P.D, A = P.A, NP = P.NP;
  const int DP = D + 4;                        // padded embedding row: conflict-free 16-byte reads of e_i, e_j
  const int nt = (NP + MT - 1) / MT;           // tiles per sample
  const int abuf_words = 2 * D * MT;           // hi | lo of one (128 x D) operand tile
  // ---- shared memory carve-up
  uint32_t *wh = sm, *wl = wh + A * D;
  float *bs = reinterpret_cast<float *>(wl + A * D), *hs = bs + A;
  uint16_t *pair = reinterpret_cast<uint16_t *>(hs + A);
  uint32_t *gbase = reinterpret_cast<uint32_t *>(pair) + ((NP + 1) / 2 + 3) / 4 * 4;
  const int per_group = 2 * abuf_words + (2 * F * DP + 3) / 4 * 4 + nt * MT + MT;
  uint32_t *mine = gbase + (size_t)grp * per_group;
  uint32_t *abuf = mine;                                                   // [2][hi | lo][D * MT]
  float *Es = reinterpret_cast<float *>(abuf + 2 * abuf_words);            // [2][F][DP]
  float *score = Es + (2 * F * DP + 3) / 4 * 4;                            // [nt * MT]
  float *part = score + nt * MT;                                           // [MT] pooling partials
  if (warp == 0) tmem_alloc(&tmem_base_s, P.tmem_cols);
  if (threadIdx.x == 0) {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) rs::mbar_init(&bar[a][b], 1);
    rs::mbar_fence_init();
  }
  for (int e = threadIdx.x; e < A * D; e += NTH) {
    const int k = e / A, n = e - k * A;        // W is (D, A): element (k = d, n = a), coalesced read
    const float x = P.W[e];
    const uint32_t hh = to_tf32(x);
    wh[tile_off(A, n, k)] = hh;
    wl[tile_off(A, n, k)] = to_tf32(x - __uint_as_float(hh));
  }
  for (int e = threadIdx.x; e < A; e += NTH) {
    bs[e] = P.bvec[e];
    hs[e] = P.h[e];
  }
  for (int p = threadIdx.x; p < NP; p += NTH) {
    int i, j;
    pair_of(F, p, i, j);
    pair[p] = (uint16_t)((i << 8) | j);
  }
  rs::fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s + (uint32_t)grp * (uint32_t)(P.tmem_cols / 2);
  const uint32_t idesc = idesc_tf32(A);
  const uint32_t lbo_a = MT * 16, lbo_w = (uint32_t)A * 16, sbo = 128;
  const int64_t first = (int64_t)blockIdx.x * 2 + grp, step = (int64_t)gridDim.x * 2;
  const int pieces = F * D / 4;                // 16-byte pieces of one sample's embeddings
  auto fetch = [&](int64_t b, int which) {     // cp.async the (F, D) block of sample b into Es[which]
    const float *src = P.E + b * (int64_t)F * D;
    float *dst = Es + (size_t)which * F * DP;
    for (int e = tid; e < pieces; e += MT) {
      const int f = e / (D / 4), q = e - f * (D / 4);
      rs::cp_async16(dst + f * DP + 4 * q, src + 4 * e);
    }
    rs::cp_async_commit();
  };
  uint32_t cnt[2] = {0, 0};                    // commits so far on bar[grp][0 / 1]
  uint32_t it = 0;                             // running tile counter of this group: buffer = it & 1
  int which = 0;
  if (first < P.B) fetch(first, 0);
  for (int64_t b = first; b < P.B; b += step, which ^= 1) {
    if (b + step < P.B) {
      fetch(b + step, which ^ 1);
      rs::cp_async_wait<1>();
    } else {
      rs::cp_async_wait<0>();
    }
    group_sync(grp);                           // this sample's embeddings are in Es[which]; score[] is free again
    const float *Eb = Es + (size_t)which * F * DP;
    int pend_tile = -1, pend_ub = 0;
    auto epilogue = [&](int t, int ub) {       // accumulator ub holds tile t: scores of its 128 pairs
      rs::mbar_wait(&bar[grp][ub], (cnt[ub] - 1) & 1u);
      fence_after_sync();
      float s = 0.f;
      for (int c0 = 0; c0 < A; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + (uint32_t)(ub * A), warp, c0, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {   // b and h come as warp-wide 16-byte broadcasts
          const float4 b4 = *reinterpret_cast<const float4 *>(bs + c0 + 4 * q), h4 = *reinterpret_cast<const float4 *>(hs + c0 + 4 * q);
          s = fmaf(h4.x, fmaxf(__uint_as_float(v[4 * q + 0]) + b4.x, 0.f), s);
          s = fmaf(h4.y, fmaxf(__uint_as_float(v[4 * q + 1]) + b4.y, 0.f), s);
          s = fmaf(h4.z, fmaxf(__uint_as_float(v[4 * q + 2]) + b4.z, 0.f), s);
          s = fmaf(h4.w, fmaxf(__uint_as_float(v[4 * q + 3]) + b4.w, 0.f), s);
        }
      }
      const int p = t * MT + tid;
      score[p] = p < NP ? s : -INFINITY;
    };
    for (int t = 0; t < nt; ++t, ++it) {
      const int ub = it & 1;
      uint32_t *ah = abuf + (size_t)ub * abuf_words, *al = ah + D * MT;
      const int p = t * MT + tid;
      const int ij = p < NP ? pair[p] : 0;
      const float *ei = Eb + (ij >> 8) * DP, *ej = Eb + (ij & 255) * DP;
      for (int q = 0; q < D / 4; ++q) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < NP) x = rs::f4_mul(*reinterpret_cast<const float4 *>(ei + 4 * q), *reinterpret_cast<const float4 *>(ej + 4 * q));
        uint4 hh, ll;
        split4(x, hh, ll);
        *reinterpret_cast<uint4 *>(ah + (q * MT + tid) * 4) = hh;
        *reinterpret_cast<uint4 *>(al + (q * MT + tid) * 4) = ll;
      }
      rs::fence_proxy_async();
      fence_before_sync();
      group_sync(grp);                         // operand tile complete; every warp has drained accumulator ub (tile it-2)
      if (tid == 0) {
        fence_after_sync();
        for (int s = 0; s < D / 8; ++s) {
          const uint64_t dah = smem_desc(rs::smem_u32(ah) + s * 2 * lbo_a, lbo_a, sbo);
          const uint64_t dal = smem_desc(rs::smem_u32(al) + s * 2 * lbo_a, lbo_a, sbo);
          const uint64_t dbh = smem_desc(rs::smem_u32(wh) + (uint32_t)(s * 2) * lbo_w, lbo_w, sbo);
          const uint64_t dbl = smem_desc(rs::smem_u32(wl) + (uint32_t)(s * 2) * lbo_w, lbo_w, sbo);
          mma_tf32(tmem + (uint32_t)(ub * A), dal, dbh, idesc, s == 0 ? 0u : 1u);
          mma_tf32(tmem + (uint32_t)(ub * A), dah, dbl, idesc, 1u);
          mma_tf32(tmem + (uint32_t)(ub * A), dah, dbh, idesc, 1u);
        }
        commit(&bar[grp][ub]);
      }
      __syncwarp();
      cnt[ub]++;
      if (pend_tile >= 0) epilogue(pend_tile, pend_ub);   // overlaps the MMAs just issued
      pend_tile = t, pend_ub = ub;
    }
    epilogue(pend_tile, pend_ub);
    fence_before_sync();
    group_sync(grp);                           // all scores of the sample are in score[]
    // ---- softmax over the NP pairs (group-wide)
    float m = -INFINITY;
    for (int p = tid; p < NP; p += MT) m = fmaxf(m, score[p]);
    m = rs::warp_max(m);
    if (lane == 0) red[grp][w4] = m;
    group_sync(grp);
    m = fmaxf(fmaxf(red[grp][0], red[grp][1]), fmaxf(red[grp][2], red[grp][3]));
    float sum = 0.f;
    for (int p = tid; p < NP; p += MT) {
      const float e = expf(score[p] - m);
      score[p] = e;
      sum += e;
    }
    sum = rs::warp_sum(sum);
    if (lane == 0) red[grp][4 + w4] = sum;
    group_sync(grp);
    const float inv = 1.f / ((red[grp][4] + red[grp][5]) + (red[grp][6] + red[grp][7]));
    for (int p = tid; p < NP; p += MT) {
      const float w = score[p] * inv;
      score[p] = w;
      if (P.attw) P.attw[b * NP + p] = w;
    }
    group_sync(grp);
    // ---- pooled[d] = sum_p w_p e_i[d] e_j[d].  Thread r adds its own pairs (r, r+128, ...) for all d in registers,
    // then the 128 partial rows are summed through shared memory (the free operand buffer) in a fixed order.
    {
      float acc[32];   // D <= 32 in this kernel
#pragma unroll
      for (int d = 0; d < 32; ++d) acc[d] = 0.f;
      for (int p = tid; p < NP; p += MT) {
        const int ij = pair[p];
        const float w = score[p];
        const float *ei = Eb + (ij >> 8) * DP, *ej = Eb + (ij & 255) * DP;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (4 * q < D) {
            const float4 a = *reinterpret_cast<const float4 *>(ei + 4 * q), c = *reinterpret_cast<const float4 *>(ej + 4 * q);
            acc[4 * q + 0] = fmaf(w, a.x * c.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(w, a.y * c.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(w, a.z * c.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(w, a.w * c.w, acc[4 * q + 3]);
          }
        }
      }
      float *scr = reinterpret_cast<float *>(abuf);   // [MT][D + 1]; no MMA is in flight (all epilogues are done)
#pragma unroll
      for (int d = 0; d < 32; ++d)
        if (d < D) scr[tid * (D + 1) + d] = acc[d];
      group_sync(grp);
      const int nsl = MT / D, d = tid % D, sl = tid / D, rows_per = MT / nsl;
      float tot = 0.f;
      for (int r2 = sl * rows_per; r2 < (sl + 1) * rows_per; ++r2) tot += scr[r2 * (D + 1) + d];
      part[tid] = tot;
      group_sync(grp);
      if (tid < D) {
        float o = 0.f;
        for (int s2 = 0; s2 < nsl; ++s2) o += part[s2 * D + tid];
        P.pooled[b * D + tid] = o;
      }
      group_sync(grp);   // scr (= the operand buffer) is free again before the next sample's first tile
    }
  }
  __syncthreads();
  if (warp == 0) tmem_free(tmem_base_s, P.tmem_cols);
}

size_t afm_tc_smem(int F, int D, int A, int NP) {
  const int DP = D + 4, nt = (NP + MT - 1) / MT;
  size_t words = (size_t)2 * A * D + 2 * A + ((NP + 1) / 2 + 3) / 4 * 4;
  words += (size_t)2 * (2 * 2 * D * MT + (2 * F * DP + 3) / 4 * 4 + nt * MT + MT);
  return words * 4;
}

}  // namespace

namespace rs {
// RS_OK when the tensor-core forward handled the call, 1 when the shape is outside what it is built for
int afm_fwd_tc_try(const float *E, int64_t B, int F, int D, int A, const float *W, const float *bvec, const float *h, float *pooled,
                   float *attw, cudaStream_t st) {
  if (!(D == 16 || D == 32) || !(A == 32 || A == 64 || A == 128) || F < 2 || F > 255 || getenv("RS_AFM_NO_TC")) return 1;
  const int NP = F * (F - 1) / 2;
  if (B < 2 * num_sms() || NP < MT) return 1;   // too little work to fill persistent CTAs / a tile
  const size_t smem = afm_tc_smem(F, D, A, NP);
  if (smem > 220 * 1024) return 1;
  AfmTcParams P = {};
  P.E = E, P.W = W, P.bvec = bvec, P.h = h, P.pooled = pooled, P.attw = attw;
  P.B = B, P.F = F, P.D = D, P.A = A, P.NP = NP;
  P.tmem_cols = 32;
  while (P.tmem_cols < 2 * A) P.tmem_cols <<= 1;
  P.tmem_cols *= 2;
  RS_CUDA(cudaFuncSetAttribute(afm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t pairs = (B + 1) / 2;
  afm_fwd_tc_kernel<<<(unsigned)(pairs < num_sms() ? pairs : num_sms()), NTH, smem, st>>>(P);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
}  // namespace rs
