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

namespace {

constexpr int KC = 32;         // k depth per chunk (4 MMAs of K = 8)
constexpr int MT = 128;        // UMMA M
constexpr int NTHREADS = 128;

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// K-major, no swizzle (LayoutType::INTERLEAVE): ((8,n),2):((16 B,SBO),LBO)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// element (row, k) of a K-major core-matrix tile with `rows` rows: chunk k/4, row, k%4
__device__ __forceinline__ int tile_off(int rows, int row, int k) { return ((k >> 2) * rows + row) * 4 + (k & 3); }

__global__ void __launch_bounds__(NTHREADS, 1) gemm_tn_3xtf32_kernel(const float *__restrict__ A, const float *__restrict__ B, int64_t K,
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
    // repack chunk [k0, k0+KC) of A (-> MT rows) and B (-> NP rows), zero padded, split hi/lo
    for (int e = tid; e < KC * MT; e += NTHREADS) {
      const int kk = e / MT, m = e - kk * MT;
      const float x = (m < M && k0 + kk < k_hi) ? A[(k0 + kk) * M + m] : 0.f;
      const uint32_t h = to_tf32(x);
      a_hi[tile_off(MT, m, kk)] = h;
      a_lo[tile_off(MT, m, kk)] = to_tf32(x - __uint_as_float(h));
    }
    for (int e = tid; e < KC * NP; e += NTHREADS) {
      const int kk = e / NP, n = e - kk * NP;
      const float x = (n < N && k0 + kk < k_hi) ? B[(k0 + kk) * N + n] : 0.f;
      const uint32_t h = to_tf32(x);
      b_hi[tile_off(NP, n, kk)] = h;
      b_lo[tile_off(NP, n, kk)] = to_tf32(x - __uint_as_float(h));
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

int plan_slabs(int64_t K, int64_t *slab) {
  int64_t chunks = (K + KC - 1) / KC;
  int64_t n = rs::num_sms();
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
  *bytes = (size_t)plan_slabs(K, &slab) * M * N * 4;
  return RS_OK;
}

RS_API int rs_gemm_tn_3xtf32(const float *A, const float *B, int64_t K, int32_t M, int32_t N, float *C, float *ws, size_t ws_bytes,
                             void *stream) {
  RS_CHECK_ARG(A && B && C && ws && K >= 1, RS_E_ARG, "rs_gemm_tn_3xtf32: bad argument");
  RS_CHECK_ARG(M >= 1 && M <= MT && N >= 1 && N <= 256, RS_E_UNSUPPORTED, "rs_gemm_tn_3xtf32: need M <= 128 and N <= 256 (got %d, %d)", M, N);
  int64_t slab;
  const int nslabs = plan_slabs(K, &slab);
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
