// C (M x N) = A^T B for A (K x M), B (K x N), both row-major fp32 -- the batch-collapsed outer product of PNN's
// "out" mode, p = S^T S with S = sum_f e_f (reference model/pnn.py:69-72), on the 5th-generation tensor cores.
//
// tcgen05.mma kind::tf32 alone gives ~1e-3 relative error; the path's bar is 1e-5, so every operand is split into
// hi = tf32(x) and lo = tf32(x - hi) and three MMAs (lo*hi, hi*lo, hi*hi) are accumulated in fp32 in tensor memory
// (3xTF32, ~1e-6).  A CTA owns a K-slab: 128 threads repack 32-deep operand chunks into the canonical K-major,
// non-swizzled core-matrix layout in shared memory (8 rows x 16 B core matrices; SBO = 128 B between 8-row groups,
// LBO = rows*16 B between the two 16-byte K halves of one K=8 MMA), one elected thread issues the MMAs and commits
// them to an mbarrier, and the four warps read the accumulator back with tcgen05.ld (warp w owns TMEM lanes
// 32w..32w+31).  Slab partials are added in slab order by a second kernel (deterministic).
#include "common.cuh"
#include "tc.cuh"

namespace {

using namespace rs::tc;
constexpr int NTHREADS = 128;

__global__ void __launch_bounds__(NTHREADS, 4) gemm_tn_3xtf32_kernel(const float *__restrict__ A, const float *__restrict__ B, int64_t K,
                                                                    int M, int N, int NP /* N padded to 16 */, int tmem_cols,
                                                                    int64_t slab, float *__restrict__ partial) {
  extern __shared__ __align__(128) uint32_t sm[];
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;
  uint32_t *a_hi = sm, *a_lo = a_hi + KC * MT, *b_hi = a_lo + KC * MT, *b_lo = b_hi + KC * NP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rs::smem_u32(&tmem_base_s)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    rs::mbar_init(&mma_bar, 1);
    rs::mbar_fence_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
  const int64_t k_lo = (int64_t)blockIdx.x * slab;
  const int64_t k_hi = (k_lo + slab < K) ? k_lo + slab : K;
  uint32_t phase = 0, first = 1;
  for (int64_t k0 = k_lo; k0 < k_hi; k0 += KC) {
    // repack chunk [k0, k0+KC) of A (-> MT rows) and B (-> NP rows), zero padded, split hi/lo.  A thread owns one
    // operand row and turns 4 consecutive k (four coalesced loads across the threads) into one 16-byte store, which
    // is exactly one row of a core matrix: conflict-free.  Fully unrolled: all loads of a row are in flight at once;
    // only the last chunk of a slab pays for bounds checks.
    const bool full = k0 + KC <= k_hi;
    {
      const float *ap = A + k0 * M + tid;
      float v[KC];
      if (tid < M && full) {
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) v[kk] = ap[(int64_t)kk * M];
      } else {
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) v[kk] = (tid < M && k0 + kk < k_hi) ? ap[(int64_t)kk * M] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < KC / 4; ++q) {
        uint4 h, l;
        split4(make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]), h, l);
        *reinterpret_cast<uint4 *>(a_hi + (q * MT + tid) * 4) = h;
        *reinterpret_cast<uint4 *>(a_lo + (q * MT + tid) * 4) = l;
      }
    }
    // B has NP <= 256 rows: with fewer rows than threads, NTHREADS / NP threads share a row and split its k range
    {
      const int nsub = NP < NTHREADS ? NTHREADS / NP : 1;   // NP is a multiple of 16: 1, 2, 4 or 8 when NP divides 128
      const int qper = (KC / 4) / nsub;
      for (int n = tid % (NP < NTHREADS ? NP : NTHREADS), sub = (NP < NTHREADS ? tid / NP : 0); n < NP && sub < nsub; n += NTHREADS) {
        const float *bp = B + k0 * N + n;
        float v[KC];
#pragma unroll
        for (int j = 0; j < KC; ++j) {   // j counts inside this thread's share of the row
          const int kk = sub * qper * 4 + j;
          v[j] = (j < qper * 4 && n < N && k0 + kk < k_hi) ? bp[(int64_t)kk * N] : 0.f;
        }
#pragma unroll
        for (int qq = 0; qq < KC / 4; ++qq) {
          if (qq < qper) {
            const int q = sub * qper + qq;
            uint4 h, l;
            split4(make_float4(v[4 * qq], v[4 * qq + 1], v[4 * qq + 2], v[4 * qq + 3]), h, l);
            *reinterpret_cast<uint4 *>(b_hi + (q * NP + n) * 4) = h;
            *reinterpret_cast<uint4 *>(b_lo + (q * NP + n) * 4) = l;
          }
        }
      }
    }
    rs::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's async proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lbo_a = MT * 16, lbo_b = NP * 16, sbo = 128;
#pragma unroll
      for (int s = 0; s < KC / 8; ++s) {  // one MMA covers K = 8 = two 16-byte K halves
        const uint64_t ah = smem_desc(rs::smem_u32(a_hi) + s * 2 * lbo_a, lbo_a, sbo);
        const uint64_t al = smem_desc(rs::smem_u32(a_lo) + s * 2 * lbo_a, lbo_a, sbo);
        const uint64_t bh = smem_desc(rs::smem_u32(b_hi) + s * 2 * lbo_b, lbo_b, sbo);
        const uint64_t bl = smem_desc(rs::smem_u32(b_lo) + s * 2 * lbo_b, lbo_b, sbo);
        mma_tf32(tmem, al, bh, idesc, first ? 0u : 1u);  // small terms first
        first = 0;
        mma_tf32(tmem, ah, bl, idesc, 1u);
        mma_tf32(tmem, ah, bh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rs::smem_u32(&mma_bar)) : "memory");
    }
    first = 0;
    rs::mbar_wait(&mma_bar, phase);  // MMAs of this chunk have read shared memory: safe to overwrite
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows of C), 8 columns at a time
  const int row = warp * 32 + lane;
  float *dst = partial + (int64_t)blockIdx.x * M * N;
  for (int c0 = 0; c0 < NP; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (k_lo < k_hi && row < M) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < N) dst[(int64_t)row * N + c0 + j] = __uint_as_float(r[j]);
    } else if (row < M) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < N) dst[(int64_t)row * N + c0 + j] = 0.f;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

__global__ void slab_reduce_kernel(const float *__restrict__ partial, int nslabs, int64_t mn, float *__restrict__ C) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= mn) return;
  float acc = 0.f;
  for (int s = 0; s < nslabs; ++s) acc += partial[(int64_t)s * mn + e];
  C[e] = acc;
}

// ------------------------------------------------------------------------------------------------------------------
// C (M x N) = epilogue(A B^T): A (M x K) rows K-contiguous, B (N x K) in torch Linear layout (out, in) -- the DIN /
// DIEN attention-unit layers (reference model/din.py:14-20,43) over all B*L (sample, position) rows at once.
// Persistent CTA per SM.  B is split into hi/lo and kept resident in shared memory in the canonical K-major core
// matrix layout; A streams through a double-buffered 128 x 32 chunk (thread r copies 128 contiguous bytes of row r
// -> conflict-free 16-byte stores); one thread issues the 3xTF32 MMAs of a chunk and commits them to that buffer's
// mbarrier, so the next chunk is repacked while the tensor core works.  Epilogue (warp w <-> TMEM lanes 32w..):
// + bias[n] + rowbias[r / rb_group][n], ReLU, multiply by (mask[r][n] > 0), store.
struct NtParams {
  const float *A, *B, *bias, *rowbias, *mask;
  float *C;
  int64_t M, lda, a_group, a_group_stride, ldb, rb_group, ldm, ldc;
  int K, N, NP, KP, relu, tmem_cols;
};

__global__ void __launch_bounds__(NTHREADS, 1) gemm_nt_3xtf32_kernel(const __grid_constant__ NtParams P) {
  extern __shared__ __align__(128) uint32_t sm[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunk = P.KP / KC;
  uint32_t *b_hi = sm, *b_lo = b_hi + (size_t)P.KP * P.NP;
  uint32_t *a_buf = b_lo + (size_t)P.KP * P.NP;  // [2][hi|lo][KC*MT]
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rs::smem_u32(&tmem_base_s)), "r"(P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    rs::mbar_init(&bar[0], 1);
    rs::mbar_init(&bar[1], 1);
    rs::mbar_fence_init();
  }
  // resident B: element (n, k) -> chunk k/4, row n
  for (int e = tid; e < P.KP * P.NP; e += NTHREADS) {
    const int k = e / P.NP, n = e - k * P.NP;
    const float x = (n < P.N && k < P.K) ? P.B[(int64_t)n * P.ldb + k] : 0.f;
    const uint32_t h = to_tf32(x);
    b_hi[tile_off(P.NP, n, k)] = h;
    b_lo[tile_off(P.NP, n, k)] = to_tf32(x - __uint_as_float(h));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.NP >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
  const uint32_t lbo_a = MT * 16, lbo_b = (uint32_t)P.NP * 16, sbo = 128;
  uint32_t uses[2] = {0, 0};  // how many times each A buffer has been committed (-> mbarrier parity)
  const int64_t ntiles = (P.M + MT - 1) / MT;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r = tile * MT + tid;  // this thread's row (A repack and epilogue use different row mappings)
    const float *arow = nullptr;
    if (r < P.M) arow = P.A + (P.a_group ? (r / P.a_group) * P.a_group_stride + (r % P.a_group) * P.lda : r * P.lda);
    for (int c = 0; c < nchunk; ++c) {
      const int bi = c & 1;
      uint32_t *ah = a_buf + (size_t)bi * 2 * KC * MT, *al = ah + KC * MT;
      if (uses[bi] > 0) rs::mbar_wait(&bar[bi], (uses[bi] - 1) & 1u);  // MMAs that read this buffer are done
#pragma unroll
      for (int q = 0; q < KC / 4; ++q) {
        const int k = c * KC + q * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (arow && k < P.K) v = rs::ldg_nc_f4(arow + k);
        uint4 h, l;
        h.x = to_tf32(v.x), h.y = to_tf32(v.y), h.z = to_tf32(v.z), h.w = to_tf32(v.w);
        l.x = to_tf32(v.x - __uint_as_float(h.x)), l.y = to_tf32(v.y - __uint_as_float(h.y));
        l.z = to_tf32(v.z - __uint_as_float(h.z)), l.w = to_tf32(v.w - __uint_as_float(h.w));
        *reinterpret_cast<uint4 *>(ah + (q * MT + tid) * 4) = h;
        *reinterpret_cast<uint4 *>(al + (q * MT + tid) * 4) = l;
      }
      rs::fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int s = 0; s < KC / 8; ++s) {
          const uint32_t kb = (uint32_t)(c * (KC / 4) + s * 2);  // 16-byte K chunk index inside resident B
          const uint64_t dah = smem_desc(rs::smem_u32(ah) + s * 2 * lbo_a, lbo_a, sbo);
          const uint64_t dal = smem_desc(rs::smem_u32(al) + s * 2 * lbo_a, lbo_a, sbo);
          const uint64_t dbh = smem_desc(rs::smem_u32(b_hi) + kb * lbo_b, lbo_b, sbo);
          const uint64_t dbl = smem_desc(rs::smem_u32(b_lo) + kb * lbo_b, lbo_b, sbo);
          mma_tf32(tmem, dal, dbh, idesc, (c == 0 && s == 0) ? 0u : 1u);
          mma_tf32(tmem, dah, dbl, idesc, 1u);
          mma_tf32(tmem, dah, dbh, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rs::smem_u32(&bar[bi])) : "memory");
      }
      uses[bi]++;
    }
    // all MMAs of this tile done?  the last commit covers every earlier MMA
    const int lb = (nchunk - 1) & 1;
    rs::mbar_wait(&bar[lb], (uses[lb] - 1) & 1u);
    if (nchunk > 1) rs::mbar_wait(&bar[lb ^ 1], (uses[lb ^ 1] - 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // Epilogue.  tcgen05.ld hands lane l the 32 columns [c0, c0+32) of row 32w+l; the warp transposes that 32x32
    // block through shared memory so that each store instruction writes 128 contiguous bytes of ONE row, and the
    // bias / per-group row bias / ReLU mask are applied with lane <-> column (coalesced reads of rowbias and mask).
    float *stg = reinterpret_cast<float *>(a_buf + (size_t)4 * KC * MT) + warp * 32 * 33;
    const int64_t row0 = tile * MT + warp * 32;
    for (int c0 = 0; c0 < P.NP; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      if (c0 + 32 <= P.NP) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
            "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
      } else {  // NP is a multiple of 16: a 16-column tail
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
#pragma unroll
        for (int j = 16; j < 32; ++j) v[j] = 0u;
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int n = c0 + lane;
      const float bn = (P.bias && n < P.N) ? P.bias[n] : 0.f;
      for (int rr = 0; rr < 32; ++rr) {
        const int64_t row = row0 + rr;
        if (row >= P.M) break;
        if (n < P.N) {
          float x = stg[rr * 33 + lane] + bn;
          if (P.rowbias) x += P.rowbias[(row / P.rb_group) * P.N + n];
          if (P.relu) x = fmaxf(x, 0.f);
          if (P.mask && !(P.mask[row * P.ldm + n] > 0.f)) x = 0.f;
          P.C[row * P.ldc + n] = x;
        }
      }
      __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // every warp has drained the accumulator before the next tile overwrites it
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(P.tmem_cols) : "memory");
}

// K is cut into slabs, one CTA each.  Several CTAs per SM (48 KB of shared memory, <= 128 TMEM columns each) keep loads
// in flight while another CTA's MMAs run; the slab partials are added in slab order afterwards.
int plan_slabs(int64_t K, int N, int64_t *slab) {
  int64_t chunks = (K + KC - 1) / KC;
  int cols = 32;
  while (cols < (N + 15) / 16 * 16) cols <<= 1;
  int per_sm = 512 / cols < 4 ? 512 / cols : 4;
  int64_t n = (int64_t)rs::num_sms() * per_sm;
  if (n > chunks) n = chunks;
  if (n < 1) n = 1;
  int64_t per = (chunks + n - 1) / n;  // chunks per slab
  *slab = per * KC;
  return (int)((chunks + per - 1) / per);
}

}  // namespace

RS_API int rs_gemm_tn_ws_bytes(int64_t K, int32_t M, int32_t N, size_t *bytes) {
  RS_CHECK_ARG(bytes && K >= 1 && M >= 1 && N >= 1, RS_E_ARG, "rs_gemm_tn_ws_bytes: bad argument");
  int64_t slab;
  *bytes = (size_t)plan_slabs(K, N, &slab) * M * N * 4;
  return RS_OK;
}

RS_API int rs_gemm_tn_3xtf32(const float *A, const float *B, int64_t K, int32_t M, int32_t N, float *C, float *ws, size_t ws_bytes,
                             void *stream) {
  RS_CHECK_ARG(A && B && C && ws && K >= 1, RS_E_ARG, "rs_gemm_tn_3xtf32: bad argument");
  RS_CHECK_ARG(M >= 1 && M <= MT && N >= 1 && N <= 256, RS_E_UNSUPPORTED, "rs_gemm_tn_3xtf32: need M <= 128 and N <= 256 (got %d, %d)", M, N);
  int64_t slab;
  const int nslabs = plan_slabs(K, N, &slab);
  RS_CHECK_ARG(ws_bytes >= (size_t)nslabs * M * N * 4, RS_E_WORKSPACE, "rs_gemm_tn_3xtf32: workspace too small");
  const int NP = (N + 15) / 16 * 16;
  int tmem_cols = 32;
  while (tmem_cols < NP) tmem_cols <<= 1;
  const size_t smem = (size_t)(2 * KC * MT + 2 * KC * NP) * 4;
  RS_CUDA(cudaFuncSetAttribute(gemm_tn_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = (cudaStream_t)stream;
  gemm_tn_3xtf32_kernel<<<nslabs, NTHREADS, smem, st>>>(A, B, K, M, N, NP, tmem_cols, slab, ws);
  RS_CHECK_LAUNCH();
  const int64_t mn = (int64_t)M * N;
  slab_reduce_kernel<<<(int)((mn + 255) / 256), 256, 0, st>>>(ws, nslabs, mn, C);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

RS_API int rs_gemm_nt_3xtf32(const rs_gemm_nt *g, void *stream) {
  RS_CHECK_ARG(g && g->A && g->B && g->C && g->M >= 0 && g->K >= 1 && g->N >= 1, RS_E_ARG, "rs_gemm_nt_3xtf32: bad argument");
  RS_CHECK_ARG(g->K % 4 == 0 && g->lda % 4 == 0 && (g->a_group == 0 || g->a_group_stride % 4 == 0), RS_E_UNSUPPORTED,
               "rs_gemm_nt_3xtf32: K, lda and a_group_stride must be multiples of 4 (16-byte row chunks)");
  RS_CHECK_ARG(g->N <= 256, RS_E_UNSUPPORTED, "rs_gemm_nt_3xtf32: N=%d > 256", g->N);
  RS_CHECK_ARG(!g->rowbias || g->rb_group >= 1, RS_E_ARG, "rs_gemm_nt_3xtf32: rowbias needs rb_group >= 1");
  if (g->M == 0) return RS_OK;
  NtParams P = {};
  P.A = g->A, P.B = g->B, P.bias = g->bias, P.rowbias = g->rowbias, P.mask = g->mask, P.C = g->C;
  P.M = g->M, P.lda = g->lda, P.a_group = g->a_group, P.a_group_stride = g->a_group_stride, P.ldb = g->ldb;
  P.rb_group = g->rb_group, P.ldm = g->ldm, P.ldc = g->ldc;
  P.K = g->K, P.N = g->N, P.relu = g->relu;
  P.NP = (g->N + 15) / 16 * 16;
  P.KP = (g->K + KC - 1) / KC * KC;
  P.tmem_cols = 32;
  while (P.tmem_cols < P.NP) P.tmem_cols <<= 1;
  const size_t smem = ((size_t)2 * P.KP * P.NP + (size_t)4 * KC * MT + 4 * 32 * 33) * 4;
  RS_CHECK_ARG(smem <= 220 * 1024, RS_E_UNSUPPORTED, "rs_gemm_nt_3xtf32: N=%d, K=%d need %zu B of shared memory for the resident B", g->N,
               g->K, smem);
  RS_CUDA(cudaFuncSetAttribute(gemm_nt_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t ntiles = (g->M + MT - 1) / MT;
  int grid = (int)(ntiles < rs::num_sms() ? ntiles : rs::num_sms());
  gemm_nt_3xtf32_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(P);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
