// Micro-benchmark: how fast can ONE SM pull rows through cp.async.bulk (TMA bulk copy) vs cp.async (LDGSTS)?
// Persistent CTA per SM, a producer warp fills a ring of shared-memory stages with K copies of S bytes each from
// pseudo-random row addresses; one consumer warp releases the stages untouched.  Reports chip-wide GB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu && ./tma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *d, const void *s, uint32_t bytes, uint64_t *b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)), "l"(s), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

// NP producer warps; each stage = K copies of S bytes; copies issued by lanes (one copy per lane per round)
template <int NP>
__global__ void __launch_bounds__(32 * (NP + 1), 1) tma_kernel(const char *src, uint64_t nrows, int S, int K, int nst, int iters, int pitch) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[16], empty[16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) { mbar_init(&full[s], NP); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t stage_bytes = (size_t)K * S;
  if (warp < NP) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % nst; const uint32_t ph = (it / nst) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      // this warp's share of the K copies
      const int k0 = K * warp / NP, k1 = K * (warp + 1) / NP;
      if (lane == 0) mbar_expect(&full[s], (uint32_t)((k1 - k0) * S));
      __syncwarp();
      for (int k = k0 + lane; k < k1; k += 32) {
        const uint64_t r = mix(((uint64_t)blockIdx.x << 40) ^ ((uint64_t)it << 12) ^ k) % nrows;
        bulk_g2s(smem + s * stage_bytes + (size_t)k * S, src + r * pitch, S, &full[s]);
      }
    }
  } else {
    for (int it = 0; it < iters; ++it) {
      const int s = it % nst; const uint32_t ph = (it / nst) & 1u;
      mbar_wait(&full[s], ph);
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
}

// LDGSTS variant: NW warps, each keeps G groups of (S/16 per row) 16-byte cp.async in flight; row per warp-iteration
template <int G>
__global__ void __launch_bounds__(256, 1) ldgsts_kernel(const char *src, uint64_t nrows, int S, int rows_per_warp_iter, int iters, int pitch) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *mine = smem + (size_t)warp * G * rows_per_warp_iter * S;
  const int v = S / 16;
  for (int it = 0; it < iters + G; ++it) {
    if (it < iters) {
      unsigned char *dst = mine + (size_t)(it % G) * rows_per_warp_iter * S;
      for (int e = lane; e < rows_per_warp_iter * v; e += 32) {
        const int k = e / v, q = e - k * v;
        const uint64_t r = mix(((uint64_t)blockIdx.x << 40) ^ ((uint64_t)warp << 32) ^ ((uint64_t)it << 8) ^ k) % nrows;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + (size_t)k * S + q * 16)), "l"(src + r * pitch + q * 16) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(G - 1) : "memory");
  }
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t big = 8ull << 30, small = 48ull << 20;
  char *buf; cudaMalloc(&buf, big); cudaMemset(buf, 1, big);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int sizes[] = {256, 512, 1024, 1664, 2048, 4096, 8192, 16384};
  for (int resident = 0; resident < 2; ++resident) {
    const size_t span = resident ? small : big;
    for (int S : sizes) {
      const int pitch = S == 1664 ? 1664 : S;
      const uint64_t nrows = span / pitch;
      for (int np = 1; np <= 2; ++np) {
        for (int stage_kb : {16, 44}) {
          int K = stage_kb * 1024 / S; if (K < 1) K = 1;
          const size_t stage_bytes = (size_t)K * S;
          int nst = (int)((200 * 1024) / stage_bytes); if (nst > 16) nst = 16; if (nst < 2) continue;
          const size_t smem = stage_bytes * nst;
          const int iters = (int)((64ull << 20) / stage_bytes);   // 64 MB per SM
          float ms = 0;
          for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (np == 1) { cudaFuncSetAttribute(tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tma_kernel<1><<<sms, 64, smem>>>(buf, nrows, S, K, nst, iters, pitch); }
            else { cudaFuncSetAttribute(tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tma_kernel<2><<<sms, 96, smem>>>(buf, nrows, S, K, nst, iters, pitch); }
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
          }
          cudaError_t err = cudaGetLastError();
          printf("tma   %s S=%5d K=%3d nst=%2d producers=%d : %8.1f GB/s  (%.1f ns per copy per SM) %s\n", resident ? "L2 " : "HBM", S, K, nst, np,
                 (double)iters * stage_bytes * sms / ms / 1e6, ms * 1e6 / ((double)iters * K), err ? cudaGetErrorString(err) : "");
        }
      }
    }
    for (int S : {1664, 4096}) {
      const uint64_t nrows = span / S;
      const int rpw = 4;                       // rows per warp per group
      const size_t smem = (size_t)8 * 3 * rpw * S;
      if (smem > 200 * 1024) continue;
      const int iters = (int)((64ull << 20) / (8 * rpw * S));
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaFuncSetAttribute(ldgsts_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEventRecord(e0);
        ldgsts_kernel<3><<<sms, 256, smem>>>(buf, nrows, S, rpw, iters, S);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("ldgsts %s S=%5d 8 warps x 3 groups x %d rows: %8.1f GB/s %s\n", resident ? "L2 " : "HBM", S, rpw, (double)iters * 8 * rpw * S * sms / ms / 1e6,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
