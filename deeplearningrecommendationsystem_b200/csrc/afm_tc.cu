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
constexpr int NG3 = 3;   // warpgroups per CTA in the backward kernels (384 threads): more latency hiding per SM

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
#pragma unroll
      for (int q = 0; q < 8; ++q) {   // D <= 32: unrolled with a predicate so the eight products overlap
        if (4 * q < D) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p < NP) x = rs::f4_mul(*reinterpret_cast<const float4 *>(ei + 4 * q), *reinterpret_cast<const float4 *>(ej + 4 * q));
          uint4 hh, ll;
          split4(x, hh, ll);
          *reinterpret_cast<uint4 *>(ah + (q * MT + tid) * 4) = hh;
          *reinterpret_cast<uint4 *>(al + (q * MT + tid) * 4) = ll;
        }
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

// ---------------------------------------------------------------------------------------------------- backward
// Three kernels, each with one simple data flow (workspace: ds (B, P), ReLU masks (B, P, A/32) bits, dP (B, P, D)):
//   afm_bwd_chain_tc_kernel   ds = softmax backward;  z = P W + b (MMA) -> mask m = [z > 0];  dP = ds (m (h*W)^T) (MMA) + w g
//   afm_dw_tc_kernel          U[d][a] = sum over all pairs of P[.][d] * (ds [z > 0])[.][a]  -- a GEMM whose K runs over the
//                             pairs, so both operands are built K(=pair)-major by "column owner" threads straight from
//                             E, ds and the mask bits; the accumulator lives in TMEM for the whole kernel.  Also
//                             m1[a] = sum ds [z > 0].  Then  dW = U * h,  db = h * m1,  dh[a] = sum_d W[d][a] U[d][a] + b[a] m1[a].
//   afm_de_kernel             dE_i += dP_p * e_j, dE_j += dP_p * e_i  (CUDA cores, lane = coordinate d, pairs in order)
struct AfmTcBwdParams {
  const float *E, *W, *bvec, *h, *attw, *g;
  float *ds;          // (B, NP)
  uint32_t *mask;     // (B, NP, A / 32)
  float *dP;          // (B, NP, D)
  float *U_part;      // (parts, D, A)
  float *m1_part;     // (parts, A)
  float *dE;          // (B, F, D)
  int64_t B;
  int F, D, A, NP, tmem_cols, tmem_slot;   // tmem_slot: accumulator columns reserved per warpgroup
};

__global__ void __launch_bounds__(NG3 * MT, 1) afm_bwd_chain_tc_kernel(const __grid_constant__ AfmTcBwdParams P) {
  extern __shared__ __align__(128) uint32_t sm[];
  __shared__ uint64_t bar[NG3];
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[NG3][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = threadIdx.x >> 7, tid = threadIdx.x & (MT - 1), w4 = warp & 3;
  const int F = P.F, D = P.D, A = P.A, NP = P.NP, AW = A / 32;
  const int DP = D + 4, nt = (NP + MT - 1) / MT;
  uint32_t *wh = sm, *wl = wh + A * D;                     // z GEMM:  N = A rows, K = D
  uint32_t *vh = wl + A * D, *vl = vh + A * D;             // dP GEMM: N = D rows, K = A
  float *bs = reinterpret_cast<float *>(vl + A * D), *hs = bs + A;
  uint16_t *pair = reinterpret_cast<uint16_t *>(hs + A);
  uint32_t *gbase = reinterpret_cast<uint32_t *>(pair) + ((NP + 1) / 2 + 3) / 4 * 4;
  const int ewords = (2 * F * DP + 3) / 4 * 4;
  const int op_words = 2 * (D > KC ? D : KC) * MT;           // the P tile and the dz chunks take turns in one buffer
  const int per_group = op_words + ewords + 2 * nt * MT + 64;
  uint32_t *mine = gbase + (size_t)grp * per_group;
  uint32_t *opP = mine, *opZ = mine;                       // [hi | lo]; dz is built only after the z MMAs have read P
  float *Es = reinterpret_cast<float *>(mine + op_words);
  float *ds_s = Es + ewords, *w_s = ds_s + nt * MT, *g_s = w_s + nt * MT;
  if (warp == 0) tmem_alloc(&tmem_base_s, P.tmem_cols);
  if (threadIdx.x == 0) {
    for (int x = 0; x < NG3; ++x) rs::mbar_init(&bar[x], 1);
    rs::mbar_fence_init();
  }
  for (int e = threadIdx.x; e < A * D; e += NG3 * MT) {
    const int d = e / A, a = e - d * A;                    // W is (D, A)
    const float x = P.W[e];
    const uint32_t hh = to_tf32(x), ll = to_tf32(x - __uint_as_float(hh));
    wh[tile_off(A, a, d)] = hh, wl[tile_off(A, a, d)] = ll;
    const float y = x * P.h[a];                              // dP = ds * ([z > 0] . (h * W)^T): h is folded into the operand
    const uint32_t yh = to_tf32(y);
    vh[tile_off(D, d, a)] = yh, vl[tile_off(D, d, a)] = to_tf32(y - __uint_as_float(yh));
  }
  for (int e = threadIdx.x; e < A; e += NG3 * MT) bs[e] = P.bvec[e], hs[e] = P.h[e];
  for (int p = threadIdx.x; p < NP; p += NG3 * MT) {
    int i, j;
    pair_of(F, p, i, j);
    pair[p] = (uint16_t)((i << 8) | j);
  }
  rs::fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s + (uint32_t)grp * (uint32_t)P.tmem_slot;
  const uint32_t idescZ = idesc_tf32(A), idescP = idesc_tf32(D);
  const uint32_t lbo_a = MT * 16, lbo_w = (uint32_t)A * 16, lbo_v = (uint32_t)D * 16, sbo = 128;
  const int64_t first = (int64_t)blockIdx.x * NG3 + grp, step = (int64_t)gridDim.x * NG3;
  const int pieces = F * D / 4;
  auto fetch = [&](int64_t b, int which) {
    const float *src = P.E + b * (int64_t)F * D;
    float *dst = Es + (size_t)which * F * DP;
    for (int e = tid; e < pieces; e += MT) {
      const int f = e / (D / 4), q = e - f * (D / 4);
      rs::cp_async16(dst + f * DP + 4 * q, src + 4 * e);
    }
    rs::cp_async_commit();
  };
  uint32_t cnt = 0;   // commits on bar[grp]; every commit is waited before the next one is issued
  int which = 0;
  if (first < P.B) fetch(first, 0);
  for (int64_t b = first; b < P.B; b += step, which ^= 1) {
    if (b + step < P.B) {
      fetch(b + step, which ^ 1);
      rs::cp_async_wait<1>();
    } else {
      rs::cp_async_wait<0>();
    }
    for (int p = tid; p < nt * MT; p += MT) w_s[p] = p < NP ? P.attw[b * NP + p] : 0.f;
    if (tid < D) g_s[tid] = P.g[b * D + tid];
    group_sync(grp);
    const float *Eb = Es + (size_t)which * F * DP;
    // ---- softmax backward: dw_p = <g, P_p>,  ds_p = w_p (dw_p - sum_q w_q dw_q)
    float tsum = 0.f;
    for (int p = tid; p < NP; p += MT) {
      const int ij = pair[p];
      const float *ei = Eb + (ij >> 8) * DP, *ej = Eb + (ij & 255) * DP;
      float dw = 0.f;
#pragma unroll 4
      for (int q = 0; q < D / 4; ++q)
        dw += rs::f4_dot(rs::f4_mul(*reinterpret_cast<const float4 *>(ei + 4 * q), *reinterpret_cast<const float4 *>(ej + 4 * q)),
                         *reinterpret_cast<const float4 *>(g_s + 4 * q));
      ds_s[p] = dw;
      tsum = fmaf(w_s[p], dw, tsum);
    }
    tsum = rs::warp_sum(tsum);
    if (lane == 0) red[grp][w4] = tsum;
    group_sync(grp);
    tsum = (red[grp][0] + red[grp][1]) + (red[grp][2] + red[grp][3]);
    for (int p = tid; p < nt * MT; p += MT) {
      const float v = p < NP ? w_s[p] * (ds_s[p] - tsum) : 0.f;
      ds_s[p] = v;
      if (p < NP) P.ds[b * NP + p] = v;
    }
    group_sync(grp);
    for (int t = 0; t < nt; ++t) {
      const int p = t * MT + tid;
      const bool ok = p < NP;
      const int ij = ok ? pair[p] : 0;
      const float *ei = Eb + (ij >> 8) * DP, *ej = Eb + (ij & 255) * DP;
      // ---- z = P W: pair products straight into the A operand
#pragma unroll
      for (int q = 0; q < 8; ++q) {   // D <= 32
        if (4 * q < D) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) x = rs::f4_mul(*reinterpret_cast<const float4 *>(ei + 4 * q), *reinterpret_cast<const float4 *>(ej + 4 * q));
          uint4 hh, ll;
          split4(x, hh, ll);
          *reinterpret_cast<uint4 *>(opP + (q * MT + tid) * 4) = hh;
          *reinterpret_cast<uint4 *>(opP + D * MT + (q * MT + tid) * 4) = ll;
        }
      }
      rs::fence_proxy_async();
      fence_before_sync();
      group_sync(grp);
      if (tid == 0) {
        fence_after_sync();
        for (int s = 0; s < D / 8; ++s) {
          const uint64_t dah = smem_desc(rs::smem_u32(opP) + s * 2 * lbo_a, lbo_a, sbo);
          const uint64_t dal = smem_desc(rs::smem_u32(opP + D * MT) + s * 2 * lbo_a, lbo_a, sbo);
          const uint64_t dbh = smem_desc(rs::smem_u32(wh) + (uint32_t)(s * 2) * lbo_w, lbo_w, sbo);
          const uint64_t dbl = smem_desc(rs::smem_u32(wl) + (uint32_t)(s * 2) * lbo_w, lbo_w, sbo);
          mma_tf32(tmem, dal, dbh, idescZ, s == 0 ? 0u : 1u);
          mma_tf32(tmem, dah, dbl, idescZ, 1u);
          mma_tf32(tmem, dah, dbh, idescZ, 1u);
        }
        commit(&bar[grp]);
      }
      __syncwarp();
      ++cnt;
      rs::mbar_wait(&bar[grp], (cnt - 1) & 1u);
      fence_after_sync();
      // ---- [z > 0], 32 columns at a time = one K chunk of  dP = ds * ([z > 0] (h * W)^T).  The 0/1 mask is exact in
      // tf32: no lo part, no split, two MMAs per K step -- and two chunks fit in the operand buffer at once, so they
      // share one barrier round trip.
      const float dsp = ds_s[p];
      for (int c0 = 0; c0 < AW; c0 += 2) {
        const int nc = AW - c0 < 2 ? AW - c0 : 2;
        for (int cl = 0; cl < nc; ++cl) {
          const int c = c0 + cl;
          uint32_t *om = opZ + (size_t)cl * KC * MT;
          uint32_t v[32];
          tmem_ld32(tmem, warp, c * 32, v);
          uint32_t bits = 0;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b4 = *reinterpret_cast<const float4 *>(bs + c * 32 + 4 * q);
            const bool m0 = __uint_as_float(v[4 * q + 0]) + b4.x > 0.f, m1 = __uint_as_float(v[4 * q + 1]) + b4.y > 0.f;
            const bool m2 = __uint_as_float(v[4 * q + 2]) + b4.z > 0.f, m3 = __uint_as_float(v[4 * q + 3]) + b4.w > 0.f;
            bits |= (m0 ? 1u : 0u) << (4 * q) | (m1 ? 1u : 0u) << (4 * q + 1) | (m2 ? 1u : 0u) << (4 * q + 2) | (m3 ? 1u : 0u) << (4 * q + 3);
            *reinterpret_cast<uint4 *>(om + (q * MT + tid) * 4) =
                make_uint4(m0 ? 0x3f800000u : 0u, m1 ? 0x3f800000u : 0u, m2 ? 0x3f800000u : 0u, m3 ? 0x3f800000u : 0u);
          }
          if (ok) P.mask[(b * NP + p) * AW + c] = bits;
        }
        rs::fence_proxy_async();
        fence_before_sync();
        group_sync(grp);
        if (tid == 0) {
          fence_after_sync();
          for (int cl = 0; cl < nc; ++cl) {
            const uint32_t om = rs::smem_u32(opZ + (size_t)cl * KC * MT);
#pragma unroll
            for (int s = 0; s < KC / 8; ++s) {
              const uint32_t kb = (uint32_t)((c0 + cl) * (KC / 4) + s * 2);
              const uint64_t dam = smem_desc(om + s * 2 * lbo_a, lbo_a, sbo);
              const uint64_t dbh = smem_desc(rs::smem_u32(vh) + kb * lbo_v, lbo_v, sbo);
              const uint64_t dbl = smem_desc(rs::smem_u32(vl) + kb * lbo_v, lbo_v, sbo);
              mma_tf32(tmem + (uint32_t)A, dam, dbl, idescP, (c0 + cl == 0 && s == 0) ? 0u : 1u);
              mma_tf32(tmem + (uint32_t)A, dam, dbh, idescP, 1u);
            }
          }
          commit(&bar[grp]);
        }
        __syncwarp();
        ++cnt;
        rs::mbar_wait(&bar[grp], (cnt - 1) & 1u);   // the operand buffer is reused: these MMAs must have read it
        fence_after_sync();
      }
      // ---- dP_p = ds_p * acc + w_p g  -> workspace
      {
        uint32_t v[32];
        tmem_ld32(tmem, warp, A, v);
        if (ok) {
          const float wp = w_s[p];
          float *dst = P.dP + (b * NP + p) * (int64_t)D;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (4 * q < D) {
              const float4 g4 = *reinterpret_cast<const float4 *>(g_s + 4 * q);
              rs::stg_cs_f4(dst + 4 * q, make_float4(fmaf(wp, g4.x, dsp * __uint_as_float(v[4 * q + 0])), fmaf(wp, g4.y, dsp * __uint_as_float(v[4 * q + 1])),
                                                     fmaf(wp, g4.z, dsp * __uint_as_float(v[4 * q + 2])), fmaf(wp, g4.w, dsp * __uint_as_float(v[4 * q + 3]))));
            }
          }
        }
      }
      fence_before_sync();
      group_sync(grp);   // accumulators drained before the next tile's MMAs
      fence_after_sync();
    }
  }
  __syncthreads();
  if (warp == 0) tmem_free(tmem_base_s, P.tmem_cols);
}

// dE from the dP workspace: warp per sample, lane = coordinate d; pairs are walked in their (i < j) order, so the row
// of feature i accumulates in a register while every partner j is updated in shared memory (deterministic, no atomics)
__global__ void __launch_bounds__(256) afm_de_kernel(const __grid_constant__ AfmTcBwdParams P) {
  extern __shared__ __align__(16) float smf[];
  const int F = P.F, D = P.D, NP = P.NP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float *dEw = smf + (size_t)warp * F * D;   // only the accumulators live in shared memory: more warps per SM
  for (int64_t b = (int64_t)blockIdx.x * nw + warp; b < P.B; b += (int64_t)gridDim.x * nw) {
    const float *Ew = P.E + b * (int64_t)F * D;   // 5 KB per sample, re-read through L1
    for (int e = lane; e < F * D; e += 32) dEw[e] = 0.f;
    __syncwarp();
    if (lane < D) {
      const int d = lane;
      const float *dp = P.dP + b * (int64_t)NP * D + d;
      for (int i = 0; i + 1 < F; ++i) {
        const float ei = __ldg(Ew + i * D + d);
        float acc = 0.f;
        int j = i + 1;
        for (; j + 8 <= F; j += 8) {   // eight independent loads in flight per lane
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = dp[u * D];
          dp += 8 * D;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            acc = fmaf(v[u], __ldg(Ew + (j + u) * D + d), acc);
            dEw[(j + u) * D + d] = fmaf(v[u], ei, dEw[(j + u) * D + d]);
          }
        }
        for (; j + 4 <= F; j += 4) {
          const float v0 = dp[0], v1 = dp[D], v2 = dp[2 * D], v3 = dp[3 * D];   // four loads in flight
          dp += 4 * D;
          acc = fmaf(v0, __ldg(Ew + (j + 0) * D + d), acc), dEw[(j + 0) * D + d] = fmaf(v0, ei, dEw[(j + 0) * D + d]);
          acc = fmaf(v1, __ldg(Ew + (j + 1) * D + d), acc), dEw[(j + 1) * D + d] = fmaf(v1, ei, dEw[(j + 1) * D + d]);
          acc = fmaf(v2, __ldg(Ew + (j + 2) * D + d), acc), dEw[(j + 2) * D + d] = fmaf(v2, ei, dEw[(j + 2) * D + d]);
          acc = fmaf(v3, __ldg(Ew + (j + 3) * D + d), acc), dEw[(j + 3) * D + d] = fmaf(v3, ei, dEw[(j + 3) * D + d]);
        }
        for (; j < F; ++j) {
          const float v = dp[0];
          dp += D;
          acc = fmaf(v, __ldg(Ew + j * D + d), acc), dEw[j * D + d] = fmaf(v, ei, dEw[j * D + d]);
        }
        dEw[i * D + d] += acc;
      }
    }
    __syncwarp();
    for (int e = lane; e < F * D; e += 32) P.dE[b * (int64_t)F * D + e] = dEw[e];
    __syncwarp();
  }
}

// U^T[a][d] = sum over pairs of mask[pair][a] * (ds[pair] P[pair][d]),  and with one extra operand column holding ds
// itself, U^T[a][D] = m1[a].  The A operand is the 0/1 ReLU mask: exact in tf32, so it needs no lo part and each K step
// costs two MMAs (mask x Q_lo, mask x Q_hi) instead of three.
__global__ void __launch_bounds__(NG3 * MT, 1) afm_dw_tc_kernel(const __grid_constant__ AfmTcBwdParams P) {
  extern __shared__ __align__(128) uint32_t sm[];
  __shared__ uint64_t bar[NG3][2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const int grp = threadIdx.x >> 7, tid = threadIdx.x & (MT - 1);
  const int F = P.F, D = P.D, A = P.A, NP = P.NP, AW = A / 32;
  const int ND = D + 16;                                                   // Q columns: D products, ds, zero padding
  const int DP = D + 4, nt = (NP + MT - 1) / MT, rows_pad = nt * MT;
  uint16_t *pair = reinterpret_cast<uint16_t *>(sm);
  uint32_t *gbase = sm + ((NP + 1) / 2 + 3) / 4 * 4;
  const int ewords = (F * DP + 3) / 4 * 4;                                 // one sample's embeddings (single buffer)
  const int opA_words = KC * MT, opB_words = 2 * KC * ND;                  // mask (hi only); Q hi | lo
  const int per_group = 2 * opA_words + 2 * opB_words + ewords + rows_pad + rows_pad * AW;
  uint32_t *mine = gbase + (size_t)grp * per_group;
  uint32_t *opA = mine, *opB = opA + 2 * opA_words;
  float *Es = reinterpret_cast<float *>(opB + 2 * opB_words);
  float *ds_s = Es + ewords;
  uint32_t *mask_s = reinterpret_cast<uint32_t *>(ds_s + rows_pad);
  if (warp == 0) tmem_alloc(&tmem_base_s, P.tmem_cols);
  if (threadIdx.x == 0) {
    for (int x = 0; x < NG3; ++x)
      for (int y = 0; y < 2; ++y) rs::mbar_init(&bar[x][y], 1);
    rs::mbar_fence_init();
  }
  for (int p = threadIdx.x; p < NP; p += NG3 * MT) {
    int i, j;
    pair_of(F, p, i, j);
    pair[p] = (uint16_t)((i << 8) | j);
  }
  // operand rows that never change: mask rows a >= A, Q columns d > D
  for (int e = tid; e < 2 * opA_words + 2 * opB_words; e += MT) opA[e] = 0u;
  rs::fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s + (uint32_t)grp * (uint32_t)P.tmem_slot;
  const uint32_t idesc = idesc_tf32(ND);
  const uint32_t lbo_a = MT * 16, lbo_b = (uint32_t)ND * 16, sbo = 128;
  const int64_t first = (int64_t)blockIdx.x * NG3 + grp, step = (int64_t)gridDim.x * NG3;
  const int pieces = F * D / 4;
  auto fetch = [&](int64_t b) {
    const float *src = P.E + b * (int64_t)F * D;
    float *dst = Es;
    for (int e = tid; e < pieces; e += MT) {
      const int f = e / (D / 4), q = e - f * (D / 4);
      rs::cp_async16(dst + f * DP + 4 * q, src + 4 * e);
    }
    rs::cp_async_commit();
  };
  // column-owner roles: A operand (mask^T): column a, row groups [ja0, ja0 + jan);  B operand (Q^T): coordinate d
  const int a_col = tid % A, a_sub = tid / A, a_nsub = MT / A;     // A in {32, 64, 128}
  const int jan = 8 / a_nsub, ja0 = a_sub * jan;
  const int b_col = tid % D, b_sub = tid / D, b_nsub = MT / D;     // D in {16, 32}
  const int jbn = 8 / b_nsub, jb0 = b_sub * jbn;
  uint32_t cnt[2] = {0, 0}, it = 0;
  for (int64_t b = first; b < P.B; b += step) {
    fetch(b);
    for (int p = tid; p < rows_pad; p += MT) ds_s[p] = p < NP ? P.ds[b * NP + p] : 0.f;
    for (int e = tid; e < rows_pad * AW; e += MT) mask_s[e] = e < NP * AW ? P.mask[b * (int64_t)NP * AW + e] : 0u;
    rs::cp_async_wait<0>();
    group_sync(grp);
    const float *Eb = Es;
    for (int ck = 0; ck < nt * 4; ++ck, ++it) {      // 32 pairs per chunk
      const int ub = it & 1, r0 = ck * KC;
      uint32_t *am = opA + (size_t)ub * opA_words;
      uint32_t *bh = opB + (size_t)ub * opB_words, *bl = bh + KC * ND;
      if (cnt[ub] > 0) rs::mbar_wait(&bar[grp][ub], (cnt[ub] - 1) & 1u);   // the MMAs that read this buffer pair are done
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {   // jan <= 8: unrolled with a predicate so the units' loads overlap
        if (jj >= jan) break;
        const int j = ja0 + jj, r = r0 + 4 * j;
        uint32_t mv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) mv[i] = ((mask_s[(r + i) * AW + (a_col >> 5)] >> (a_col & 31)) & 1u) ? 0x3f800000u : 0u;
        *reinterpret_cast<uint4 *>(am + (j * MT + a_col) * 4) = make_uint4(mv[0], mv[1], mv[2], mv[3]);
      }
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {   // jbn <= 2
        if (jj >= jbn) break;
        const int j = jb0 + jj, r = r0 + 4 * j;
        float x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int p = r + i;
          const int ij = p < NP ? pair[p] : 0;
          x[i] = ds_s[p] * (Eb[(ij >> 8) * DP + b_col] * Eb[(ij & 255) * DP + b_col]);   // ds is 0 on padding rows
        }
        uint4 hh, ll;
        split4(make_float4(x[0], x[1], x[2], x[3]), hh, ll);
        *reinterpret_cast<uint4 *>(bh + (j * ND + b_col) * 4) = hh;
        *reinterpret_cast<uint4 *>(bl + (j * ND + b_col) * 4) = ll;
      }
      if (tid < 8) {   // column D of Q: ds itself -> m1 = mask^T ds comes out of the same accumulator
        const int r = r0 + 4 * tid;
        uint4 hh, ll;
        split4(make_float4(ds_s[r], ds_s[r + 1], ds_s[r + 2], ds_s[r + 3]), hh, ll);
        *reinterpret_cast<uint4 *>(bh + (tid * ND + D) * 4) = hh;
        *reinterpret_cast<uint4 *>(bl + (tid * ND + D) * 4) = ll;
      }
      rs::fence_proxy_async();
      group_sync(grp);
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int s = 0; s < KC / 8; ++s) {
          const uint64_t dam = smem_desc(rs::smem_u32(am) + s * 2 * lbo_a, lbo_a, sbo);
          const uint64_t dbh = smem_desc(rs::smem_u32(bh) + s * 2 * lbo_b, lbo_b, sbo);
          const uint64_t dbl = smem_desc(rs::smem_u32(bl) + s * 2 * lbo_b, lbo_b, sbo);
          mma_tf32(tmem, dam, dbl, idesc, (it == 0 && s == 0) ? 0u : 1u);
          mma_tf32(tmem, dam, dbh, idesc, 1u);
        }
        commit(&bar[grp][ub]);
      }
      __syncwarp();
      cnt[ub]++;
    }
    group_sync(grp);   // ds / mask / E of this sample are no longer read by this group's threads
  }
  // ---- drain: accumulator lane = a, columns 0..D-1 = U[.][a], column D = m1[a] -> this group's partial
  for (int ub = 0; ub < 2; ++ub)
    if (cnt[ub] > 0) rs::mbar_wait(&bar[grp][ub], (cnt[ub] - 1) & 1u);
  fence_after_sync();
  const int part = blockIdx.x * NG3 + grp;
  for (int c0 = 0; c0 < ND; c0 += 32) {
    uint32_t v[32];
    if (it > 0) {
      tmem_ld32(tmem, warp, c0, v);
    } else {
#pragma unroll
      for (int d = 0; d < 32; ++d) v[d] = 0u;
    }
    if (tid < A) {
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        if (c0 + d < D) P.U_part[((int64_t)part * D + c0 + d) * A + tid] = __uint_as_float(v[d]);
        if (c0 + d == D) P.m1_part[(int64_t)part * A + tid] = __uint_as_float(v[d]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free(tmem_base_s, P.tmem_cols);
}

size_t afm_chain_smem(int F, int D, int A, int NP) {
  const int DP = D + 4, nt = (NP + MT - 1) / MT;
  size_t words = (size_t)4 * A * D + 2 * A + ((NP + 1) / 2 + 3) / 4 * 4;
  words += (size_t)NG3 * (2 * (D > KC ? D : KC) * MT + (2 * F * DP + 3) / 4 * 4 + 2 * nt * MT + 64);
  return words * 4;
}
size_t afm_dw_smem(int F, int D, int A, int NP) {
  const int DP = D + 4, nt = (NP + MT - 1) / MT, rows_pad = nt * MT, ND = D + 16;
  size_t words = ((NP + 1) / 2 + 3) / 4 * 4;
  words += (size_t)NG3 * (2 * KC * MT + 2 * 2 * KC * ND + (F * DP + 3) / 4 * 4 + rows_pad + rows_pad * (A / 32));
  return words * 4;
}
size_t ws_region(size_t bytes) { return (bytes + 255) / 256 * 256; }
bool afm_tc_shape_ok(int64_t B, int F, int D, int A) {
  const int NP = F * (F - 1) / 2;
  return (D == 16 || D == 32) && (A == 32 || A == 64 || A == 128) && F >= 2 && F <= 255 && NP >= MT && B >= 2 * rs::num_sms();
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
  if (!afm_tc_shape_ok(B, F, D, A) || getenv("RS_AFM_NO_TC")) return 1;   // else: too little work for persistent CTAs / a tile
  const int NP = F * (F - 1) / 2;
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

// ---- C ABI of the tensor-core backward
RS_API int rs_afm_bwd_tc_plan(int64_t B, int32_t F, int32_t D, int32_t A, int32_t *num_parts, size_t *ws_bytes) {
  RS_CHECK_ARG(num_parts && ws_bytes, RS_E_ARG, "rs_afm_bwd_tc_plan: null output");
  *num_parts = 0, *ws_bytes = 0;
  if (!afm_tc_shape_ok(B, F, D, A) || getenv("RS_AFM_NO_TC")) return RS_OK;   // 0 parts: use rs_afm_bwd
  const int NP = F * (F - 1) / 2;
  if (afm_chain_smem(F, D, A, NP) > 220 * 1024 || afm_dw_smem(F, D, A, NP) > 220 * 1024) return RS_OK;
  const int64_t ctas = (B + NG3 - 1) / NG3;
  *num_parts = NG3 * (int)(ctas < rs::num_sms() ? ctas : rs::num_sms());
  *ws_bytes = ws_region((size_t)B * NP * 4) + ws_region((size_t)B * NP * 4 * (A / 32)) + ws_region((size_t)B * NP * 4 * D);
  return RS_OK;
}

RS_API int rs_afm_bwd_tc(const float *E, int64_t B, int32_t F, int32_t D, int32_t A, const float *W, const float *bvec, const float *h,
                         const float *attw, const float *g_pooled, float *dE, float *U_part, float *m1_part, int32_t num_parts,
                         void *ws, size_t ws_bytes, void *stream) {
  RS_CHECK_ARG(E && W && bvec && h && attw && g_pooled && dE && U_part && m1_part && ws, RS_E_ARG, "rs_afm_bwd_tc: null argument");
  int32_t parts;
  size_t need;
  if (int rc = rs_afm_bwd_tc_plan(B, F, D, A, &parts, &need)) return rc;
  RS_CHECK_ARG(parts > 0, RS_E_UNSUPPORTED, "rs_afm_bwd_tc: shape (B=%lld, F=%d, D=%d, A=%d) is served by rs_afm_bwd", (long long)B, F, D, A);
  RS_CHECK_ARG(num_parts == parts, RS_E_ARG, "rs_afm_bwd_tc: num_parts %d != plan %d", num_parts, parts);
  RS_CHECK_ARG(ws_bytes >= need, RS_E_WORKSPACE, "rs_afm_bwd_tc: workspace of %zu bytes needed, got %zu", need, ws_bytes);
  const int NP = F * (F - 1) / 2;
  cudaStream_t st = (cudaStream_t)stream;
  AfmTcBwdParams P = {};
  P.E = E, P.W = W, P.bvec = bvec, P.h = h, P.attw = attw, P.g = g_pooled, P.dE = dE, P.U_part = U_part, P.m1_part = m1_part;
  char *wsb = static_cast<char *>(ws);
  P.ds = reinterpret_cast<float *>(wsb);
  P.mask = reinterpret_cast<uint32_t *>(wsb + ws_region((size_t)B * NP * 4));
  P.dP = reinterpret_cast<float *>(wsb + ws_region((size_t)B * NP * 4) + ws_region((size_t)B * NP * 4 * (A / 32)));
  P.B = B, P.F = F, P.D = D, P.A = A, P.NP = NP;
  const int grid = parts / NG3;
  {
    P.tmem_slot = A + 32;   // z (A columns) + dP (up to 32)
    P.tmem_cols = 32;
    while (P.tmem_cols < NG3 * P.tmem_slot) P.tmem_cols <<= 1;
    const size_t smem = afm_chain_smem(F, D, A, NP);
    RS_CUDA(cudaFuncSetAttribute(afm_bwd_chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    afm_bwd_chain_tc_kernel<<<grid, NG3 * MT, smem, st>>>(P);
    RS_CHECK_LAUNCH();
  }
  {
    const int nw = 8;
    const size_t smem = (size_t)nw * F * D * 4;
    RS_CUDA(cudaFuncSetAttribute(afm_de_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (B + nw - 1) / nw;
    const int64_t cap = (int64_t)rs::num_sms() * 5;
    afm_de_kernel<<<(unsigned)(blocks < cap ? blocks : cap), nw * 32, smem, st>>>(P);
    RS_CHECK_LAUNCH();
  }
  {
    P.tmem_slot = 64;       // D + 16 <= 48 accumulator columns
    P.tmem_cols = 256;
    const size_t smem = afm_dw_smem(F, D, A, NP);
    RS_CUDA(cudaFuncSetAttribute(afm_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    afm_dw_tc_kernel<<<grid, NG3 * MT, smem, st>>>(P);
    RS_CHECK_LAUNCH();
  }
  return RS_OK;
}
