// Shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/recsys_b200.h"

#define RS_API extern "C" __attribute__((visibility("default")))

namespace rs {

void set_error(const char *fmt, ...);
int num_sms();

#define RS_CHECK_ARG(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      rs::set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)

#define RS_CHECK_LAUNCH()                                                        \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      rs::set_error("%s:%d launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return (int)e__;                                                           \
    }                                                                            \
  } while (0)

#define RS_CUDA(call)                                                            \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      rs::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return (int)e__;                                                           \
    }                                                                            \
  } while (0)

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ float4 ldg_nc_f4(const float *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg_f4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void stg_f4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
// streaming store: written once, read by a later kernel -- do not keep in L1
__device__ __forceinline__ void stg_cs_f4(float *p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4_fmas(float4 a, float s, float4 c) {
  return make_float4(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w));
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
__device__ __forceinline__ float f4_hsum(float4 a) { return (a.x + a.y) + (a.z + a.w); }

__device__ __forceinline__ float4 f4_shfl_xor(float4 v, int m) {
  v.x = __shfl_xor_sync(0xffffffffu, v.x, m);
  v.y = __shfl_xor_sync(0xffffffffu, v.y, m);
  v.z = __shfl_xor_sync(0xffffffffu, v.z, m);
  v.w = __shfl_xor_sync(0xffffffffu, v.w, m);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- cp.async (LDGSTS) 16-byte global -> shared
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP)
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same copy with an L2 eviction-priority hint (createpolicy): rows of small, hot tables are worth keeping in L2
// (evict_last) while rows that are touched once per batch should leave first (evict_first)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
// shared -> global bulk store (bulk_group completion)
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// by-value copy of rs_routes for kernel parameters
struct Routes {
  int n;
  int64_t start[RS_MAX_RANKS + 1];
  float *base[RS_MAX_RANKS];
  int64_t row0[RS_MAX_RANKS];
  const int64_t *dyn_start, *dyn_row0;  // device-resident ranges (host-sync-free sharded step), else nullptr
  int64_t cap_rows;                     // > 0: destination rows >= cap_rows are dropped
  int self;                             // own index among the destinations (all-to-all schedule rotation), -1 unknown
};
// rotation of a walk over `count` items so that rank `self` of n starts self/n of the way through
__device__ __forceinline__ int64_t route_rotation(const Routes &R, int64_t count) {
  return (R.n > 1 && R.self > 0) ? count * R.self / R.n : 0;
}
// The ranges are staged ONCE per CTA into shared memory (route_tab_load + __syncthreads): the device-resident ones live in
// symmetric (peer-mapped) memory, where every load is a round trip to L2/HBM -- three dependent ones per routed row cost
// ~0.7 us per row when read in place.  tab[0..n] = start, tab[RS_MAX_RANKS+1 ..] = row0.
constexpr int ROUTE_TAB = 2 * RS_MAX_RANKS + 1;
__device__ __forceinline__ void route_tab_load(int64_t *tab, const Routes &R) {
  for (int k = threadIdx.x; k <= R.n; k += blockDim.x) tab[k] = R.dyn_start ? R.dyn_start[k] : R.start[k];
  for (int k = threadIdx.x; k < R.n; k += blockDim.x) tab[RS_MAX_RANKS + 1 + k] = R.dyn_row0 ? R.dyn_row0[k] : R.row0[k];
}
// destination of logical row r (floats): linear scan over at most n <= 64 ranges; nullptr = drop the row
__device__ __forceinline__ float *route_row(const Routes &R, const int64_t *tab, int64_t r, int W) {
  int k = 0;
  while (k + 1 < R.n && r >= tab[k + 1]) ++k;
  const int64_t dst = tab[RS_MAX_RANKS + 1 + k] + (r - tab[k]);
  if (R.cap_rows > 0 && dst >= R.cap_rows) return nullptr;
  return R.base[k] + dst * W;
}

__device__ __forceinline__ int64_t clamp_id(int64_t id, int64_t rows, int32_t *status) {
  if ((uint64_t)id >= (uint64_t)rows) {
    if (status) atomicOr(status, 1);
    id = id < 0 ? 0 : rows - 1;
  }
  return id;
}

}  // namespace rs
