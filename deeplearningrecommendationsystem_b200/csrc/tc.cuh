// tcgen05 building blocks shared by the tensor-core kernels (gemm_tc.cu, din_tc.cu): 3xTF32 operand split, the
// canonical K-major no-swizzle core-matrix layout, shared-memory / instruction descriptors, MMA issue, TMEM
// allocation and read-back.
#pragma once
#include "common.cuh"

namespace rs {
namespace tc {

constexpr int KC = 32;   // k depth of one staged operand chunk (4 MMAs of K = 8)
constexpr int MT = 128;  // UMMA M: one accumulator row per TMEM lane

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));   // pure: not volatile, so independent splits interleave
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

// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
__device__ __forceinline__ uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
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

__device__ __forceinline__ void commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one warp allocates / frees `cols` (power of two >= 32) TMEM columns for the CTA
__device__ __forceinline__ void tmem_alloc(uint32_t *base_smem, int cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(base_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t base, int cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

// lane l of warp w receives columns [col, col+32) of accumulator row 32(w % 4) + l (a warp can only reach the TMEM
// lane quarter given by its index inside its warpgroup)
__device__ __forceinline__ void tmem_ld32(uint32_t tmem, int warp, int col, uint32_t (&v)[32]) {
  const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// barrier over the 128 threads of one warpgroup (ids 1.. ; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }

// split four fp32 values into tf32 hi and lo parts (x = hi + lo up to ~2^-22 relative)
__device__ __forceinline__ void split4(float4 v, uint4 &h, uint4 &l) {
  h.x = to_tf32(v.x), h.y = to_tf32(v.y), h.z = to_tf32(v.z), h.w = to_tf32(v.w);
  l.x = to_tf32(v.x - __uint_as_float(h.x)), l.y = to_tf32(v.y - __uint_as_float(h.y));
  l.z = to_tf32(v.z - __uint_as_float(h.z)), l.w = to_tf32(v.w - __uint_as_float(h.w));
}

}  // namespace tc
}  // namespace rs
