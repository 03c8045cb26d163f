// Stable LSD radix sort of (uint32 key, int32 value) pairs and an int32 prefix sum: the two primitives under the
// deterministic dedup (rs_dedup_sort; the bookkeeping that embedding_dense_backward hides, trainer/trainer.py:38).
//
// 9-bit digits (26-bit row keys -> 3 passes).  Per pass: (1) per-tile digit histogram, (2) exclusive scan of the
// (digit-major, tile-minor) counts -> where each tile's run of each digit starts, (3) scatter: a warp walks its
// 512 contiguous elements 32 at a time, ranks equal digits with match.any (lower lanes first) on top of its running
// per-digit count, so equal keys keep their input order at every level (lane < round < warp < tile) -- the sort is
// stable, which is what makes the later segment sums run in ascending position.
#include "common.cuh"

namespace rs {
namespace {

constexpr int RBITS = 9, RBINS = 1 << RBITS;
constexpr int STILE = 4096;   // scan tile: 1024 threads x 4
constexpr int TILE = 4096, SWARPS = 8, PER_WARP = TILE / SWARPS, ROUNDS = PER_WARP / 32;   // 256 threads per tile

__global__ void __launch_bounds__(256) radix_hist_kernel(const uint32_t *__restrict__ keys, int64_t n, int shift, int nblocks,
                                                         int32_t *__restrict__ hist) {
  __shared__ int32_t h[RBINS];
  for (int i = threadIdx.x; i < RBINS; i += 256) h[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t base = (int64_t)blockIdx.x * TILE;
  for (int e = threadIdx.x; e < TILE; e += 256) {
    const int64_t idx = base + e;
    const int digit = idx < n ? (int)((keys[idx] >> shift) & (RBINS - 1)) : RBINS + lane;   // padding lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    if (digit < RBINS && lane == __ffs(peers) - 1) atomicAdd(&h[digit], __popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RBINS; i += 256) hist[(int64_t)i * nblocks + blockIdx.x] = h[i];
}

// exclusive scan of `m` ints in place by ONE block (the histogram table: RBINS x tiles entries)
__global__ void __launch_bounds__(1024) scan_exclusive_single_kernel(int32_t *__restrict__ a, int64_t m) {
  __shared__ int32_t wsum[32];
  __shared__ int32_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < m; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int32_t x = i < m ? a[i] : 0;
    int32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int32_t w = wsum[lane], winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      wsum[lane] = winc - w;   // exclusive offsets of the warps
    }
    __syncthreads();
    const int32_t carry = carry_s;
    if (i < m) a[i] = carry + wsum[warp] + inc - x;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wsum[31] + inc;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) radix_scatter_kernel(const uint32_t *__restrict__ kin, const int32_t *__restrict__ vin,
                                                            uint32_t *__restrict__ kout, int32_t *__restrict__ vout, int64_t n, int shift,
                                                            int nblocks, const int32_t *__restrict__ goff) {
  __shared__ int32_t cnt[SWARPS][RBINS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < SWARPS * RBINS; i += 256) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t sub = (int64_t)blockIdx.x * TILE + (int64_t)warp * PER_WARP;
  uint32_t key[ROUNDS];
  int32_t val[ROUNDS], loc[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int64_t idx = sub + r * 32 + lane;
    const bool ok = idx < n;
    key[r] = ok ? kin[idx] : 0u;
    val[r] = ok ? vin[idx] : 0;
  }
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const bool ok = sub + r * 32 + lane < n;
    const int digit = ok ? (int)((key[r] >> shift) & (RBINS - 1)) : RBINS + lane;
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const int before = __popc(peers & ((1u << lane) - 1u));
    loc[r] = ok ? cnt[warp][digit] + before : 0;          // read the running count ...
    __syncwarp();
    if (ok && before == 0) cnt[warp][digit] += __popc(peers);   // ... then the first lane of each digit group bumps it
    __syncwarp();
  }
  __syncthreads();
  // where this tile's run of digit d starts (global), then per-warp starts inside it, in warp order
  for (int d = threadIdx.x; d < RBINS; d += 256) {
    int32_t run = goff[(int64_t)d * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < SWARPS; ++w) {
      const int32_t t = cnt[w][d];
      cnt[w][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    if (sub + r * 32 + lane < n) {
      const int digit = (int)((key[r] >> shift) & (RBINS - 1));
      const int64_t pos = (int64_t)cnt[warp][digit] + loc[r];
      kout[pos] = key[r];
      vout[pos] = val[r];
    }
  }
}

// ---- inclusive prefix sum of n ints: tile scans + one-block scan of the tile sums + offset add
template <bool EXCLUSIVE>
__global__ void __launch_bounds__(1024) scan_tiles_kernel(const int32_t *in, int32_t *out, int64_t n, int32_t *__restrict__ tile_sum) {
  __shared__ int32_t wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i0 = (int64_t)blockIdx.x * STILE + threadIdx.x * 4;
  int32_t x[4], raw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) raw[j] = x[j] = i0 + j < n ? in[i0 + j] : 0;   // in may alias out: read everything first
  x[1] += x[0], x[2] += x[1], x[3] += x[2];
  int32_t inc = x[3];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int32_t w = wsum[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    wsum[lane] = winc - w;
    if (lane == 31) tile_sum[blockIdx.x] = winc;
  }
  __syncthreads();
  const int32_t before = wsum[warp] + inc - x[3];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i0 + j < n) out[i0 + j] = before + x[j] - (EXCLUSIVE ? raw[j] : 0);
}

__global__ void __launch_bounds__(1024) scan_add_kernel(int32_t *__restrict__ out, int64_t n, const int32_t *__restrict__ tile_off) {
  const int32_t off = tile_off[blockIdx.x];   // exclusive offsets
  const int64_t i0 = (int64_t)blockIdx.x * STILE + threadIdx.x * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i0 + j < n) out[i0 + j] += off;
}

}  // namespace

size_t radix_sort_ws_ints(int64_t n) {
  const size_t m = (size_t)RBINS * (size_t)((n + TILE - 1) / TILE);
  return m + (m + STILE - 1) / STILE + 1;   // count table + its tile sums
}
size_t scan_ws_ints(int64_t n) { return (size_t)((n + STILE - 1) / STILE) + 1; }

// keys/values of [0, end_bit) bits; the result lands in (kout, vout); (ktmp, vtmp) are scratch of n entries each
int radix_sort_pairs(const uint32_t *kin, const int32_t *vin, uint32_t *kout, int32_t *vout, uint32_t *ktmp, int32_t *vtmp,
                     int32_t *hist_ws, int64_t n, int end_bit, cudaStream_t st) {
  const int passes = (end_bit + RBITS - 1) / RBITS;
  const int nblocks = (int)((n + TILE - 1) / TILE);
  const uint32_t *ks = kin;
  const int32_t *vs = vin;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) & 1) == 0;   // alternate so that the last pass writes (kout, vout)
    uint32_t *kd = to_out ? kout : ktmp;
    int32_t *vd = to_out ? vout : vtmp;
    radix_hist_kernel<<<nblocks, 256, 0, st>>>(ks, n, p * RBITS, nblocks, hist_ws);
    RS_CHECK_LAUNCH();
    {  // exclusive scan of the (digit-major, tile-minor) counts, in place; tile sums behind the table
      const int64_t m = (int64_t)RBINS * nblocks;
      const int nt = (int)((m + STILE - 1) / STILE);
      int32_t *sums = hist_ws + m;
      scan_tiles_kernel<true><<<nt, 1024, 0, st>>>(hist_ws, hist_ws, m, sums);
      RS_CHECK_LAUNCH();
      if (nt > 1) {
        scan_exclusive_single_kernel<<<1, 1024, 0, st>>>(sums, nt);
        RS_CHECK_LAUNCH();
        scan_add_kernel<<<nt, 1024, 0, st>>>(hist_ws, m, sums);
        RS_CHECK_LAUNCH();
      }
    }
    radix_scatter_kernel<<<nblocks, 256, 0, st>>>(ks, vs, kd, vd, n, p * RBITS, nblocks, hist_ws);
    RS_CHECK_LAUNCH();
    ks = kd, vs = vd;
  }
  return RS_OK;
}

int inclusive_sum_i32(const int32_t *in, int32_t *out, int64_t n, int32_t *ws, cudaStream_t st) {
  const int nt = (int)((n + STILE - 1) / STILE);
  scan_tiles_kernel<false><<<nt, 1024, 0, st>>>(in, out, n, ws);
  RS_CHECK_LAUNCH();
  if (nt > 1) {
    scan_exclusive_single_kernel<<<1, 1024, 0, st>>>(ws, nt);
    RS_CHECK_LAUNCH();
    scan_add_kernel<<<nt, 1024, 0, st>>>(out, n, ws);
    RS_CHECK_LAUNCH();
  }
  return RS_OK;
}

}  // namespace rs
