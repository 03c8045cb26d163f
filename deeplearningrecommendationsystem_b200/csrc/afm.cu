// AFM attention pooling over the pairwise Hadamard products (reference model/afm.py:55-65), forward and backward.
//
//   P_p = e_i * e_j (i<j),  s_p = h . relu(P_p W + b),  w = softmax_p(s),  pooled = sum_p w_p P_p
//
// The (B, P, D) pair tensor the reference materialises (3.1 GB at F = 39, D = 32, B = 32768) never exists here:
// a warp owns a sample, keeps its F x D tile in shared memory and forms products four pairs at a time.  W (and its
// transpose, for the backward) live in shared memory for the whole CTA.  Parameter gradients are accumulated per
// warp in a fixed order (shared memory for dW, registers for db / dh) and written as per-warp partials that the
// host adds in warp order -> deterministic.
#include "common.cuh"

namespace rs {
int afm_fwd_tc_try(const float *E, int64_t B, int F, int D, int A, const float *W, const float *bvec, const float *h, float *pooled,
                   float *attw, cudaStream_t st);
}

namespace {

constexpr int AW = 4;    // max warps per CTA (fewer when the per-warp shared-memory tiles are large)
constexpr int PT = 4;    // pairs processed together
constexpr int MAXC = 4;  // attention columns per lane (A <= 128)
constexpr int MAXQ = 8;  // embedding columns per lane (D <= 256)

struct AfmParams {
  const float *E, *W, *bvec, *h, *attw_in, *g_pooled;
  float *pooled, *attw_out, *dE, *dW_part, *db_part, *dh_part;
  int64_t B;
  int F, D, A, NP, aw;
};

__host__ __device__ __forceinline__ int r4(int x) { return (x + 3) & ~3; }

__device__ __forceinline__ void pair_of(int F, int p, int &i, int &j) {
  i = 0;
  while (p >= F - 1 - i) {
    p -= F - 1 - i;
    ++i;
  }
  j = i + 1 + p;
}

struct Smem {
  float *W, *Wt, *b, *h;  // CTA-wide
  float *E, *score, *prod, *dz, *dEt, *dW;  // this warp's
  uint16_t *pair;
};

__device__ __forceinline__ Smem carve(float *base, const AfmParams &P, int warp, bool bwd) {
  Smem s;
  float *p = base;
  s.W = p;
  p += r4(P.D * P.A);
  s.Wt = p;
  p += bwd ? r4(P.D * P.A) : 0;
  s.b = p;
  p += r4(P.A);
  s.h = p;
  p += r4(P.A);
  s.pair = reinterpret_cast<uint16_t *>(p);
  p += r4((P.NP + 1) / 2);
  const int per_warp = r4(P.F * P.D) + r4(P.NP) + PT * P.D + (bwd ? r4(PT * P.A) + r4(P.F * P.D) + r4(P.D * P.A) : 0);
  float *w = p + (size_t)warp * per_warp;
  s.E = w;
  w += r4(P.F * P.D);
  s.score = w;
  w += r4(P.NP);
  s.prod = w;
  w += PT * P.D;
  s.dz = w;
  w += bwd ? r4(PT * P.A) : 0;
  s.dEt = w;
  w += bwd ? r4(P.F * P.D) : 0;
  s.dW = w;
  return s;
}

size_t smem_bytes(const AfmParams &P, bool bwd, int aw) {
  size_t f = (size_t)r4(P.D * P.A) * (bwd ? 2 : 1) + 2 * r4(P.A) + r4((P.NP + 1) / 2);
  f += (size_t)aw * (r4(P.F * P.D) + r4(P.NP) + PT * P.D + (bwd ? r4(PT * P.A) + r4(P.F * P.D) + r4(P.D * P.A) : 0));
  return f * 4;
}
// the backward's per-warp dW tile decides how many warps fit; forward and backward must agree on the partition
int pick_warps(const AfmParams &P) {
  for (int aw = AW; aw >= 1; aw >>= 1)
    if (smem_bytes(P, true, aw) <= 216 * 1024) return aw;
  return 0;
}

// z[pt][c] = b[a] + sum_d prod[pt][d] * W[d][a] for this lane's columns a = lane + 32c
__device__ __forceinline__ void project(const AfmParams &P, const Smem &s, int lane, int npt, float z[PT][MAXC]) {
  const int NC = (P.A + 31) >> 5;
#pragma unroll
  for (int pt = 0; pt < PT; ++pt)
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int a = lane + 32 * c;
      z[pt][c] = (c < NC && a < P.A) ? s.b[a] : 0.f;
    }
  for (int d = 0; d < P.D; d += 4) {
    float4 pv[PT];
#pragma unroll
    for (int pt = 0; pt < PT; ++pt) pv[pt] = *reinterpret_cast<const float4 *>(s.prod + pt * P.D + d);  // broadcast
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int a = lane + 32 * c;
      if (c < NC && a < P.A) {
        const float w0 = s.W[(d + 0) * P.A + a], w1 = s.W[(d + 1) * P.A + a], w2 = s.W[(d + 2) * P.A + a], w3 = s.W[(d + 3) * P.A + a];
#pragma unroll
        for (int pt = 0; pt < PT; ++pt) {
          z[pt][c] = fmaf(pv[pt].x, w0, z[pt][c]);
          z[pt][c] = fmaf(pv[pt].y, w1, z[pt][c]);
          z[pt][c] = fmaf(pv[pt].z, w2, z[pt][c]);
          z[pt][c] = fmaf(pv[pt].w, w3, z[pt][c]);
        }
      }
    }
  }
  (void)npt;
}

__device__ __forceinline__ void load_common(const AfmParams &P, const Smem &s, bool bwd) {
  for (int e = threadIdx.x; e < P.D * P.A; e += blockDim.x) {
    const float w = P.W[e];
    s.W[e] = w;
    if (bwd) {
      const int d = e / P.A, a = e - d * P.A;
      s.Wt[a * P.D + d] = w;
    }
  }
  for (int e = threadIdx.x; e < P.A; e += blockDim.x) {
    s.b[e] = P.bvec[e];
    s.h[e] = P.h[e];
  }
  for (int p = threadIdx.x; p < P.NP; p += blockDim.x) {
    int i, j;
    pair_of(P.F, p, i, j);
    s.pair[p] = (uint16_t)((i << 8) | j);
  }
}

__device__ __forceinline__ void make_products(const AfmParams &P, const Smem &s, int lane, int p0, int npt) {
  for (int pt = 0; pt < PT; ++pt) {
    const int ij = pt < npt ? s.pair[p0 + pt] : 0;
    const float *ei = s.E + (ij >> 8) * P.D, *ej = s.E + (ij & 255) * P.D;
    for (int d = lane; d < P.D; d += 32) s.prod[pt * P.D + d] = pt < npt ? ei[d] * ej[d] : 0.f;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(AW * 32) afm_fwd_kernel(const __grid_constant__ AfmParams P) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const Smem s = carve(smem, P, warp, false);
  load_common(P, s, false);
  __syncthreads();
  const int NC = (P.A + 31) >> 5;
  const int64_t stride = (int64_t)gridDim.x * P.aw;
  for (int64_t b = (int64_t)blockIdx.x * P.aw + warp; b < P.B; b += stride) {
    const float *Eg = P.E + b * (int64_t)P.F * P.D;
    for (int e = lane; e < P.F * P.D; e += 32) s.E[e] = Eg[e];
    __syncwarp();
    for (int p0 = 0; p0 < P.NP; p0 += PT) {
      const int npt = P.NP - p0 < PT ? P.NP - p0 : PT;
      make_products(P, s, lane, p0, npt);
      float z[PT][MAXC];
      project(P, s, lane, npt, z);
#pragma unroll
      for (int pt = 0; pt < PT; ++pt) {
        float part = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          const int a = lane + 32 * c;
          if (c < NC && a < P.A) part = fmaf(s.h[a], fmaxf(z[pt][c], 0.f), part);
        }
        part = rs::warp_sum(part);
        if (lane == 0 && pt < npt) s.score[p0 + pt] = part;
      }
      __syncwarp();
    }
    // softmax over the pairs
    float mx = -INFINITY;
    for (int p = lane; p < P.NP; p += 32) mx = fmaxf(mx, s.score[p]);
    mx = rs::warp_max(mx);
    float sum = 0.f;
    for (int p = lane; p < P.NP; p += 32) {
      const float e = expf(s.score[p] - mx);
      s.score[p] = e;
      sum += e;
    }
    sum = rs::warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int p = lane; p < P.NP; p += 32) {
      const float w = s.score[p] * inv;
      s.score[p] = w;
      if (P.attw_out) P.attw_out[b * P.NP + p] = w;
    }
    __syncwarp();
    for (int d = lane; d < P.D; d += 32) {
      float acc = 0.f;
      for (int p = 0; p < P.NP; ++p) {
        const int ij = s.pair[p];
        acc = fmaf(s.score[p], s.E[(ij >> 8) * P.D + d] * s.E[(ij & 255) * P.D + d], acc);
      }
      P.pooled[b * P.D + d] = acc;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(AW * 32) afm_bwd_kernel(const __grid_constant__ AfmParams P) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const Smem s = carve(smem, P, warp, true);
  load_common(P, s, true);
  for (int e = lane; e < P.D * P.A; e += 32) s.dW[e] = 0.f;
  __syncthreads();
  const int NC = (P.A + 31) >> 5;
  float acc_db[MAXC], acc_dh[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) acc_db[c] = acc_dh[c] = 0.f;
  const int64_t stride = (int64_t)gridDim.x * P.aw;
  for (int64_t b = (int64_t)blockIdx.x * P.aw + warp; b < P.B; b += stride) {
    const float *Eg = P.E + b * (int64_t)P.F * P.D;
    for (int e = lane; e < P.F * P.D; e += 32) {
      s.E[e] = Eg[e];
      s.dEt[e] = 0.f;
    }
    __syncwarp();
    // ds_p = w_p * (<g, P_p> - sum_q w_q <g, P_q>)
    const float *g = P.g_pooled + b * P.D;
    float tsum = 0.f;
    for (int p = 0; p < P.NP; ++p) {
      const int ij = s.pair[p];
      float part = 0.f;
      for (int d = lane; d < P.D; d += 32) part = fmaf(g[d], s.E[(ij >> 8) * P.D + d] * s.E[(ij & 255) * P.D + d], part);
      part = rs::warp_sum(part);
      const float w = P.attw_in[b * P.NP + p];
      if (lane == 0) s.score[p] = part;
      tsum = fmaf(w, part, tsum);
    }
    __syncwarp();
    for (int p = lane; p < P.NP; p += 32) {
      const float w = P.attw_in[b * P.NP + p];
      s.score[p] = w * (s.score[p] - tsum);
    }
    __syncwarp();
    for (int p0 = 0; p0 < P.NP; p0 += PT) {
      const int npt = P.NP - p0 < PT ? P.NP - p0 : PT;
      make_products(P, s, lane, p0, npt);
      float z[PT][MAXC];
      project(P, s, lane, npt, z);
      // dz, dh, db for this lane's attention columns
#pragma unroll
      for (int pt = 0; pt < PT; ++pt) {
        const float ds = pt < npt ? s.score[p0 + pt] : 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          const int a = lane + 32 * c;
          if (c < NC && a < P.A) {
            const float zr = fmaxf(z[pt][c], 0.f);
            acc_dh[c] = fmaf(ds, zr, acc_dh[c]);
            const float dz = z[pt][c] > 0.f ? ds * s.h[a] : 0.f;
            acc_db[c] += dz;
            s.dz[pt * P.A + a] = dz;
          }
        }
      }
      __syncwarp();
      // dW[d][a] += P[pt][d] * dz[pt][a]  (lane owns columns a; sequential over d, pt -> fixed order)
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int a = lane + 32 * c;
        if (c < NC && a < P.A) {
          for (int d = 0; d < P.D; ++d) {
            float acc = s.dW[d * P.A + a];
#pragma unroll
            for (int pt = 0; pt < PT; ++pt) acc = fmaf(s.prod[pt * P.D + d], s.dz[pt * P.A + a], acc);
            s.dW[d * P.A + a] = acc;
          }
        }
      }
      // dP[pt][d] = w_p g[d] + sum_a W[d][a] dz[pt][a]; then dE_i += dP * e_j, dE_j += dP * e_i  (lane owns columns d)
      for (int d = lane; d < P.D; d += 32) {
        float dp[PT];
#pragma unroll
        for (int pt = 0; pt < PT; ++pt) dp[pt] = pt < npt ? P.attw_in[b * P.NP + p0 + pt] * g[d] : 0.f;
        for (int a = 0; a < P.A; ++a) {
          const float wt = s.Wt[a * P.D + d];
#pragma unroll
          for (int pt = 0; pt < PT; ++pt) dp[pt] = fmaf(wt, s.dz[pt * P.A + a], dp[pt]);
        }
#pragma unroll
        for (int pt = 0; pt < PT; ++pt) {
          if (pt < npt) {
            const int ij = s.pair[p0 + pt];
            const int i = ij >> 8, j = ij & 255;
            s.dEt[i * P.D + d] = fmaf(dp[pt], s.E[j * P.D + d], s.dEt[i * P.D + d]);
            s.dEt[j * P.D + d] = fmaf(dp[pt], s.E[i * P.D + d], s.dEt[j * P.D + d]);
          }
        }
      }
      __syncwarp();
    }
    float *dEg = P.dE + b * (int64_t)P.F * P.D;
    for (int e = lane; e < P.F * P.D; e += 32) dEg[e] = s.dEt[e];
    __syncwarp();
  }
  // per-warp partials, added by the host in warp order
  const int64_t part = (int64_t)blockIdx.x * P.aw + warp;
  for (int e = lane; e < P.D * P.A; e += 32) P.dW_part[part * P.D * P.A + e] = s.dW[e];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int a = lane + 32 * c;
    if (c < NC && a < P.A) {
      P.db_part[part * P.A + a] = acc_db[c];
      P.dh_part[part * P.A + a] = acc_dh[c];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Fast path for the shapes the configs use (D in {16,32,64}, A = 32*NC, D*NC <= 64): every lane keeps its NC columns
// of W (and, in the backward, of dW) in registers, so the projection is 8 broadcast LDS.128 per 64 FMAs and the
// weight-gradient accumulation never touches shared memory.  Eight pairs are processed together.
constexpr int PT8 = 8;

template <int D, int NC>
__device__ __forceinline__ void project_reg(const float (&Wr)[D][NC], const float (&br)[NC], const float *prod, float (&z)[PT8][NC]) {
#pragma unroll
  for (int pt = 0; pt < PT8; ++pt)
#pragma unroll
    for (int c = 0; c < NC; ++c) z[pt][c] = br[c];
#pragma unroll
  for (int d = 0; d < D; d += 4) {
    float4 pv[PT8];
#pragma unroll
    for (int pt = 0; pt < PT8; ++pt) pv[pt] = *reinterpret_cast<const float4 *>(prod + pt * D + d);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int pt = 0; pt < PT8; ++pt) {
        z[pt][c] = fmaf(pv[pt].x, Wr[d][c], z[pt][c]);
        z[pt][c] = fmaf(pv[pt].y, Wr[d + 1][c], z[pt][c]);
        z[pt][c] = fmaf(pv[pt].z, Wr[d + 2][c], z[pt][c]);
        z[pt][c] = fmaf(pv[pt].w, Wr[d + 3][c], z[pt][c]);
      }
  }
}

// same projection with W read from shared memory (W[d][a], lanes over a: conflict free) -- used by the backward,
// whose registers are taken by the dW accumulators
template <int D, int NC>
__device__ __forceinline__ void project_smem(const float *Ws, const float (&br)[NC], const float *prod, int lane, float (&z)[PT8][NC]) {
  constexpr int A = 32 * NC;
#pragma unroll
  for (int pt = 0; pt < PT8; ++pt)
#pragma unroll
    for (int c = 0; c < NC; ++c) z[pt][c] = br[c];
#pragma unroll 2
  for (int d = 0; d < D; d += 4) {
    float4 pv[PT8];
#pragma unroll
    for (int pt = 0; pt < PT8; ++pt) pv[pt] = *reinterpret_cast<const float4 *>(prod + pt * D + d);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float w0 = Ws[(d + 0) * A + lane + 32 * c], w1 = Ws[(d + 1) * A + lane + 32 * c];
      const float w2 = Ws[(d + 2) * A + lane + 32 * c], w3 = Ws[(d + 3) * A + lane + 32 * c];
#pragma unroll
      for (int pt = 0; pt < PT8; ++pt) {
        z[pt][c] = fmaf(pv[pt].x, w0, z[pt][c]);
        z[pt][c] = fmaf(pv[pt].y, w1, z[pt][c]);
        z[pt][c] = fmaf(pv[pt].z, w2, z[pt][c]);
        z[pt][c] = fmaf(pv[pt].w, w3, z[pt][c]);
      }
    }
  }
}

template <int D, int NC, bool BWD>
__global__ void __launch_bounds__(AW * 32) afm_reg_kernel(const __grid_constant__ AfmParams P) {
  constexpr int A = 32 * NC;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // CTA-wide: Wt [A][D] and W [D][A] (backward), pair table; per warp: E, score, prod[8][D], dz[8][A], dEt
  float *Wt = smem;
  float *Ws = Wt + (BWD ? A * D : 0);
  uint16_t *pair = reinterpret_cast<uint16_t *>(Ws + (BWD ? A * D : 0));
  float *wbase = reinterpret_cast<float *>(pair) + r4((P.NP + 1) / 2);
  const int per_warp = r4(P.F * D) + r4(P.NP) + PT8 * D + (BWD ? PT8 * A + r4(P.F * D) : 0);
  float *sE = wbase + (size_t)warp * per_warp, *score = sE + r4(P.F * D), *prod = score + r4(P.NP);
  float *dz = prod + PT8 * D, *dEt = dz + (BWD ? PT8 * A : 0);
  if (BWD)
    for (int e = threadIdx.x; e < D * A; e += blockDim.x) {
      Wt[(e % A) * D + e / A] = P.W[e];
      Ws[e] = P.W[e];
    }
  for (int p = threadIdx.x; p < P.NP; p += blockDim.x) {
    int i, j;
    pair_of(P.F, p, i, j);
    pair[p] = (uint16_t)((i << 8) | j);
  }
  float Wr[BWD ? 1 : D][NC], br[NC], hr[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    br[c] = P.bvec[lane + 32 * c];
    hr[c] = P.h[lane + 32 * c];
#pragma unroll
    for (int d = 0; d < (BWD ? 1 : D); ++d) Wr[d][c] = P.W[d * A + lane + 32 * c];
  }
  float dWr[BWD ? D : 1][NC], acc_db[NC], acc_dh[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    acc_db[c] = acc_dh[c] = 0.f;
#pragma unroll
    for (int d = 0; d < (BWD ? D : 1); ++d) dWr[d][c] = 0.f;
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * P.aw;
  for (int64_t b = (int64_t)blockIdx.x * P.aw + warp; b < P.B; b += stride) {
    const float *Eg = P.E + b * (int64_t)P.F * D;
    for (int e = lane; e < P.F * D; e += 32) {
      sE[e] = Eg[e];
      if (BWD) dEt[e] = 0.f;
    }
    __syncwarp();
    const float *g = BWD ? P.g_pooled + b * D : nullptr;
    if (BWD) {  // ds_p = w_p (<g, P_p> - sum_q w_q <g, P_q>), kept in score[]
      float tsum = 0.f;
      for (int p = 0; p < P.NP; ++p) {
        const int ij = pair[p];
        float part = 0.f;
        for (int d = lane; d < D; d += 32) part = fmaf(g[d], sE[(ij >> 8) * D + d] * sE[(ij & 255) * D + d], part);
        part = rs::warp_sum(part);
        if (lane == 0) score[p] = part;
        tsum = fmaf(P.attw_in[b * P.NP + p], part, tsum);
      }
      __syncwarp();
      for (int p = lane; p < P.NP; p += 32) score[p] = P.attw_in[b * P.NP + p] * (score[p] - tsum);
      __syncwarp();
    }
    for (int p0 = 0; p0 < P.NP; p0 += PT8) {
      const int npt = P.NP - p0 < PT8 ? P.NP - p0 : PT8;
      for (int e = lane; e < PT8 * D; e += 32) {
        const int pt = e / D, d = e - pt * D;
        const int ij = pt < npt ? pair[p0 + pt] : 0;
        prod[e] = pt < npt ? sE[(ij >> 8) * D + d] * sE[(ij & 255) * D + d] : 0.f;
      }
      __syncwarp();
      float z[PT8][NC];
      if constexpr (BWD)
        project_smem<D, NC>(Ws, br, prod, lane, z);
      else
        project_reg<D, NC>(Wr, br, prod, z);
      if (!BWD) {
#pragma unroll
        for (int pt = 0; pt < PT8; ++pt) {
          float part = 0.f;
#pragma unroll
          for (int c = 0; c < NC; ++c) part = fmaf(hr[c], fmaxf(z[pt][c], 0.f), part);
          part = rs::warp_sum(part);
          if (lane == 0 && pt < npt) score[p0 + pt] = part;
        }
      } else {
#pragma unroll
        for (int pt = 0; pt < PT8; ++pt) {
          const float ds = pt < npt ? score[p0 + pt] : 0.f;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            acc_dh[c] = fmaf(ds, fmaxf(z[pt][c], 0.f), acc_dh[c]);
            const float dzv = z[pt][c] > 0.f ? ds * hr[c] : 0.f;
            acc_db[c] += dzv;
            z[pt][c] = dzv;  // reuse the registers for dz
            dz[pt * A + lane + 32 * c] = dzv;
          }
        }
        // dW[d][a] += sum_pt P[pt][d] dz[pt][a]: registers only
#pragma unroll
        for (int d = 0; d < D; d += 4) {
          float4 pv[PT8];
#pragma unroll
          for (int pt = 0; pt < PT8; ++pt) pv[pt] = *reinterpret_cast<const float4 *>(prod + pt * D + d);
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int pt = 0; pt < PT8; ++pt) {
              dWr[BWD ? d : 0][c] = fmaf(pv[pt].x, z[pt][c], dWr[BWD ? d : 0][c]);
              dWr[BWD ? d + 1 : 0][c] = fmaf(pv[pt].y, z[pt][c], dWr[BWD ? d + 1 : 0][c]);
              dWr[BWD ? d + 2 : 0][c] = fmaf(pv[pt].z, z[pt][c], dWr[BWD ? d + 2 : 0][c]);
              dWr[BWD ? d + 3 : 0][c] = fmaf(pv[pt].w, z[pt][c], dWr[BWD ? d + 3 : 0][c]);
            }
        }
        __syncwarp();
        // dP[pt][d] = w_p g[d] + sum_a W[d][a] dz[pt][a]  (lane owns columns d); dE_i += dP e_j, dE_j += dP e_i
        for (int d = lane; d < D; d += 32) {
          float dp[PT8];
#pragma unroll
          for (int pt = 0; pt < PT8; ++pt) dp[pt] = pt < npt ? P.attw_in[b * P.NP + p0 + pt] * g[d] : 0.f;
          for (int a = 0; a < A; ++a) {
            const float wt = Wt[a * D + d];
#pragma unroll
            for (int pt = 0; pt < PT8; ++pt) dp[pt] = fmaf(wt, dz[pt * A + a], dp[pt]);
          }
#pragma unroll
          for (int pt = 0; pt < PT8; ++pt)
            if (pt < npt) {
              const int ij = pair[p0 + pt];
              const int i = ij >> 8, j = ij & 255;
              dEt[i * D + d] = fmaf(dp[pt], sE[j * D + d], dEt[i * D + d]);
              dEt[j * D + d] = fmaf(dp[pt], sE[i * D + d], dEt[j * D + d]);
            }
        }
      }
      __syncwarp();
    }
    if (!BWD) {
      float mx = -INFINITY;
      for (int p = lane; p < P.NP; p += 32) mx = fmaxf(mx, score[p]);
      mx = rs::warp_max(mx);
      float sum = 0.f;
      for (int p = lane; p < P.NP; p += 32) {
        const float e = expf(score[p] - mx);
        score[p] = e;
        sum += e;
      }
      sum = rs::warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int p = lane; p < P.NP; p += 32) {
        const float w = score[p] * inv;
        score[p] = w;
        if (P.attw_out) P.attw_out[b * P.NP + p] = w;
      }
      __syncwarp();
      for (int d = lane; d < D; d += 32) {
        float acc = 0.f;
        for (int p = 0; p < P.NP; ++p) {
          const int ij = pair[p];
          acc = fmaf(score[p], sE[(ij >> 8) * D + d] * sE[(ij & 255) * D + d], acc);
        }
        P.pooled[b * D + d] = acc;
      }
    } else {
      float *dEg = P.dE + b * (int64_t)P.F * D;
      for (int e = lane; e < P.F * D; e += 32) dEg[e] = dEt[e];
    }
    __syncwarp();
  }
  if (BWD) {
    const int64_t part = (int64_t)blockIdx.x * P.aw + warp;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int a = lane + 32 * c;
#pragma unroll
      for (int d = 0; d < D; ++d) P.dW_part[(part * D + d) * A + a] = dWr[BWD ? d : 0][c];
      P.db_part[part * A + a] = acc_db[c];
      P.dh_part[part * A + a] = acc_dh[c];
    }
  }
}

size_t reg_smem_bytes(const AfmParams &P, bool bwd, int aw) {
  size_t f = (bwd ? (size_t)2 * P.A * P.D : 0) + r4((P.NP + 1) / 2);
  f += (size_t)aw * (r4(P.F * P.D) + r4(P.NP) + PT8 * P.D + (bwd ? PT8 * P.A + r4(P.F * P.D) : 0));
  return f * 4;
}

template <int D, int NC>
int launch_reg(const AfmParams &P, bool bwd, int grid, cudaStream_t st) {
  const size_t smem = reg_smem_bytes(P, bwd, P.aw);
  if (bwd) {
    RS_CUDA(cudaFuncSetAttribute(afm_reg_kernel<D, NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    afm_reg_kernel<D, NC, true><<<grid, P.aw * 32, smem, st>>>(P);
  } else {
    RS_CUDA(cudaFuncSetAttribute(afm_reg_kernel<D, NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    afm_reg_kernel<D, NC, false><<<grid, P.aw * 32, smem, st>>>(P);
  }
  RS_CHECK_LAUNCH();
  return RS_OK;
}

// returns RS_OK if the register-resident fast path handled the launch, 1 if the shape is not covered
int try_reg(const AfmParams &P, bool bwd, int grid, cudaStream_t st) {
  if (reg_smem_bytes(P, true, P.aw) > 200 * 1024) return 1;
#define RS_AFM_CASE(d, nc) \
  if (P.D == d && P.A == 32 * nc) return launch_reg<d, nc>(P, bwd, grid, st);
  RS_AFM_CASE(32, 2)
  RS_AFM_CASE(32, 1)
  RS_AFM_CASE(16, 1)
  RS_AFM_CASE(16, 2)
  RS_AFM_CASE(64, 1)
#undef RS_AFM_CASE
  return 1;
}

int check(const AfmParams &P, const char *who) {
  RS_CHECK_ARG(P.F >= 2 && P.F <= 255, RS_E_SHAPE, "%s: F=%d out of range", who, P.F);
  RS_CHECK_ARG(P.D >= 4 && P.D % 4 == 0 && P.D <= 32 * MAXQ, RS_E_UNSUPPORTED, "%s: D=%d must be a multiple of 4, <= 256", who, P.D);
  RS_CHECK_ARG(P.A >= 1 && P.A <= 32 * MAXC, RS_E_UNSUPPORTED, "%s: attention dim %d must be <= 128", who, P.A);
  return RS_OK;
}

}  // namespace

static int grid_for(const AfmParams &P, int aw) {
  int64_t blocks = (P.B + aw - 1) / aw;
  int64_t cap = (int64_t)rs::num_sms() * 2;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

RS_API int rs_afm_num_parts(int64_t B, int32_t F, int32_t D, int32_t A, int32_t *parts) {
  RS_CHECK_ARG(parts, RS_E_ARG, "rs_afm_num_parts: null");
  AfmParams P = {};
  P.B = B;
  P.F = F;
  P.D = D;
  P.A = A;
  P.NP = F * (F - 1) / 2;
  int rc = check(P, "rs_afm_num_parts");
  if (rc) return rc;
  const int aw = pick_warps(P);
  RS_CHECK_ARG(aw >= 1, RS_E_UNSUPPORTED, "rs_afm: F=%d D=%d A=%d do not fit in shared memory", F, D, A);
  *parts = grid_for(P, aw) * aw;
  return RS_OK;
}

RS_API int rs_afm_fwd(const float *E, int64_t B, int32_t F, int32_t D, int32_t A, const float *W, const float *bvec, const float *h,
                      float *pooled, float *attw, void *stream) {
  RS_CHECK_ARG(E && W && bvec && h && pooled, RS_E_ARG, "rs_afm_fwd: null argument");
  AfmParams P = {};
  P.E = E;
  P.W = W;
  P.bvec = bvec;
  P.h = h;
  P.pooled = pooled;
  P.attw_out = attw;
  P.B = B;
  P.F = F;
  P.D = D;
  P.A = A;
  P.NP = F * (F - 1) / 2;
  int rc = check(P, "rs_afm_fwd");
  if (rc) return rc;
  if (B == 0) return RS_OK;
  P.aw = pick_warps(P);
  RS_CHECK_ARG(P.aw >= 1, RS_E_UNSUPPORTED, "rs_afm_fwd: F=%d D=%d A=%d do not fit in shared memory", F, D, A);
  const size_t smem = smem_bytes(P, false, P.aw);
  RS_CUDA(cudaFuncSetAttribute(afm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {  // large batches: the projection on the tensor cores (afm_tc.cu); same outputs
    int rc3 = rs::afm_fwd_tc_try(E, B, F, D, A, W, bvec, h, pooled, attw, (cudaStream_t)stream);
    if (rc3 != 1) return rc3;
  }
  {
    int rc2 = try_reg(P, false, grid_for(P, P.aw), (cudaStream_t)stream);
    if (rc2 != 1) return rc2;
  }
  afm_fwd_kernel<<<grid_for(P, P.aw), P.aw * 32, smem, (cudaStream_t)stream>>>(P);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

RS_API int rs_afm_bwd(const float *E, int64_t B, int32_t F, int32_t D, int32_t A, const float *W, const float *bvec, const float *h,
                      const float *attw, const float *g_pooled, float *dE, float *dW_part, float *db_part, float *dh_part,
                      int32_t num_parts, void *stream) {
  RS_CHECK_ARG(E && W && bvec && h && attw && g_pooled && dE && dW_part && db_part && dh_part, RS_E_ARG, "rs_afm_bwd: null argument");
  AfmParams P = {};
  P.E = E;
  P.W = W;
  P.bvec = bvec;
  P.h = h;
  P.attw_in = attw;
  P.g_pooled = g_pooled;
  P.dE = dE;
  P.dW_part = dW_part;
  P.db_part = db_part;
  P.dh_part = dh_part;
  P.B = B;
  P.F = F;
  P.D = D;
  P.A = A;
  P.NP = F * (F - 1) / 2;
  int rc = check(P, "rs_afm_bwd");
  if (rc) return rc;
  P.aw = pick_warps(P);
  RS_CHECK_ARG(P.aw >= 1, RS_E_UNSUPPORTED, "rs_afm_bwd: F=%d D=%d A=%d do not fit in shared memory", F, D, A);
  const int grid = grid_for(P, P.aw);
  RS_CHECK_ARG(num_parts == grid * P.aw, RS_E_ARG, "rs_afm_bwd: num_parts %d != rs_afm_num_parts() = %d", num_parts, grid * P.aw);
  const size_t smem = smem_bytes(P, true, P.aw);
  RS_CUDA(cudaFuncSetAttribute(afm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    int rc2 = try_reg(P, true, grid, (cudaStream_t)stream);
    if (rc2 != 1) return rc2;
  }
  afm_bwd_kernel<<<grid, P.aw * 32, smem, (cudaStream_t)stream>>>(P);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
