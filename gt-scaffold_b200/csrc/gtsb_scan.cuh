// gtsb_scan.cuh -- device-wide exclusive prefix sum (reduce-then-scan), u32
// sums over u8 or u32 inputs.  Three launches: tile sums, scan of tile sums
// (one block), tile scan with carry-in.  out[n] receives the grand total.
#pragma once
#include "gtsb_common.cuh"
#include "gtsb_kernels.h"

namespace gtsb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;                          // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;    // 4096 elements per block

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if ((int) lane_id() >= d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across the block; returns the
// exclusive prefix and writes the block total to *total (valid in all threads)
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t block_total;
  const uint32_t incl = warp_incl_scan(v);
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) warp_sums[w] = incl;
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    uint32_t s = ((int) lane_id() < nw) ? warp_sums[lane_id()] : 0u;
    const uint32_t si = warp_incl_scan(s);
    if ((int) lane_id() < nw) warp_sums[lane_id()] = si - s;
    if ((int) lane_id() == nw - 1) block_total = si;
  }
  __syncthreads();
  const uint32_t r = incl - v + warp_sums[w];
  *total = block_total;
  __syncthreads();
  return r;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const T *__restrict__ in, uint64_t n,
                                                                 uint32_t *__restrict__ tile_sums) {
  const uint64_t base = (uint64_t) blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const uint64_t i = base + (uint64_t) k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += (uint32_t) in[i];
  }
  uint32_t total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of tile sums in place; grand total -> *total_out
static __global__ void __launch_bounds__(1024) k_scan_tile_offsets(uint32_t *__restrict__ tile_sums,
                                                            uint32_t ntiles,
                                                            uint32_t *__restrict__ total_out) {
  uint32_t carry = 0;
  for (uint32_t base = 0; base < ntiles; base += blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < ntiles ? tile_sums[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan(v, &total);
    if (i < ntiles) tile_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const T *__restrict__ in, uint64_t n,
                                                             const uint32_t *__restrict__ tile_offsets,
                                                             uint32_t *__restrict__ out) {
  // blocked arrangement: thread t owns SCAN_ITEMS consecutive elements
  const uint64_t base = (uint64_t) blockIdx.x * SCAN_TILE + (uint64_t) threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const uint64_t i = base + k;
    v[k] = i < n ? (uint32_t) in[i] : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = block_excl_scan(s, &total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const uint64_t i = base + k;
    if (i < n) out[i] = run;
    run += v[k];
  }
}

// out must hold n+1 elements; tile_scratch must hold ceil(n/SCAN_TILE)+1.
template <typename T>
inline void exclusive_scan(const T *in, uint64_t n, uint32_t *out, uint32_t *tile_scratch,
                           cudaStream_t stream) {
  if (n == 0) {
    cudaMemsetAsync(out, 0, sizeof(uint32_t), stream);
    return;
  }
  const uint32_t ntiles = (uint32_t) ((n + SCAN_TILE - 1) / SCAN_TILE);
  KernelTimer timer_(sizeof(T) == 1 ? "scan_u8(3 kernels)" : "scan_u32(3 kernels)", stream);
  k_scan_tile_sums<T><<<ntiles, SCAN_THREADS, 0, stream>>>(in, n, tile_scratch);
  k_scan_tile_offsets<<<1, 1024, 0, stream>>>(tile_scratch, ntiles, out + n);
  k_scan_tiles<T><<<ntiles, SCAN_THREADS, 0, stream>>>(in, n, tile_scratch, out);
}

inline uint64_t scan_scratch_elems(uint64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

}  // namespace gtsb
