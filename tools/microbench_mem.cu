// microbench_mem.cu -- B200 random-access primitives that bound the CSR build
// (dev tool; numbers recorded in profiles/).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// each thread reads CHUNK bytes (as uint32 x CHUNK/4... using 16B vectors) at a random aligned-to-ALIGN offset
template <int CHUNK>
__global__ void k_gather(const uint4 *__restrict__ src, uint64_t nchunks, uint64_t n, uint32_t *out, uint64_t seed) {
  uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
    uint64_t c = mix(i ^ seed) % nchunks;
    const uint4 *p = src + c * (CHUNK / 16);
#pragma unroll
    for (int k = 0; k < CHUNK / 16; k++) { uint4 v = p[k]; acc += v.x ^ v.y ^ v.z ^ v.w; }
  }
  if (acc == 0x12345678u) out[0] = acc;
}
// 4-byte random gather
__global__ void k_gather4(const uint32_t *__restrict__ src, uint64_t nelem, uint64_t n, uint32_t *out, uint64_t seed) {
  uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (; i < n; i += (uint64_t) gridDim.x * blockDim.x) acc += src[mix(i ^ seed) % nelem];
  if (acc == 0x12345678u) out[0] = acc;
}
// dependent: idx table (4B gather) -> 32B chunk
__global__ void k_gather_dep(const uint32_t *__restrict__ tab, uint64_t ntab, const uint4 *__restrict__ src, uint64_t nchunks,
                             uint64_t n, uint32_t *out, uint64_t seed) {
  uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
    uint32_t t = tab[mix(i ^ seed) % ntab];
    const uint4 *p = src + ((uint64_t) t % nchunks) * 2;
    uint4 a = p[0], b = p[1]; acc += a.x ^ b.y;
  }
  if (acc == 0x12345678u) out[0] = acc;
}
__global__ void k_atomic(uint32_t *dst, uint64_t nelem, uint64_t n, uint64_t seed) {
  uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += (uint64_t) gridDim.x * blockDim.x) atomicAdd(&dst[mix(i ^ seed) % nelem], 1u);
}
__global__ void k_atomic_ret(uint32_t *dst, uint64_t nelem, uint64_t n, uint64_t seed, uint32_t *out) {
  uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (; i < n; i += (uint64_t) gridDim.x * blockDim.x) acc += atomicAdd(&dst[mix(i ^ seed) % nelem], 1u);
  if (acc == 0x12345678u) out[0] = acc;
}
__global__ void k_scatter16(uint4 *dst, uint64_t nelem, uint64_t n, uint64_t seed) {
  uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += (uint64_t) gridDim.x * blockDim.x) dst[mix(i ^ seed) % nelem] = make_uint4((uint32_t) i, 1, 2, 3);
}
__global__ void k_scatter4(uint32_t *dst, uint64_t nelem, uint64_t n, uint64_t seed) {
  uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += (uint64_t) gridDim.x * blockDim.x) dst[mix(i ^ seed) % nelem] = (uint32_t) i;
}
__global__ void k_copy(const uint4 *__restrict__ a, uint4 *__restrict__ b, uint64_t n) {
  uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += (uint64_t) gridDim.x * blockDim.x) b[i] = a[i];
}

template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(0); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(a)); f(r + 1); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  const uint64_t BYTES = 2ull << 30;              // 2 GiB region (>> L2)
  uint4 *big, *big2; uint32_t *out, *small;
  CK(cudaMalloc(&big, BYTES)); CK(cudaMalloc(&big2, BYTES)); CK(cudaMalloc(&out, 64));
  CK(cudaMemset(big, 1, BYTES)); CK(cudaMemset(big2, 0, BYTES));
  const uint64_t SMALLB = 40ull << 20;            // 40 MiB (L2 resident)
  CK(cudaMalloc(&small, SMALLB)); CK(cudaMemset(small, 0, SMALLB));
  const int G = 148 * 16, T = 256;
  const uint64_t N = 1ull << 26;                  // 67M operations
  float ms;
  ms = timeit([&](int r) { k_copy<<<G, T>>>(big, big2, BYTES / 16); });
  printf("copy 2GiB                : %.3f ms  %.0f GB/s (r+w)\n", ms, 2.0 * BYTES / ms / 1e6);
  ms = timeit([&](int r) { k_gather4<<<G, T>>>((uint32_t *) big, BYTES / 4, N, out, r); });
  printf("gather 4B  from 2GiB     : %.3f ms  %.2f Gop/s  (sector traffic %.0f GB/s)\n", ms, N / ms / 1e6, 32.0 * N / ms / 1e6);
  ms = timeit([&](int r) { k_gather<32><<<G, T>>>(big, BYTES / 32, N, out, r); });
  printf("gather 32B from 2GiB     : %.3f ms  %.2f Gop/s  %.0f GB/s\n", ms, N / ms / 1e6, 32.0 * N / ms / 1e6);
  ms = timeit([&](int r) { k_gather<64><<<G, T>>>(big, BYTES / 64, N, out, r); });
  printf("gather 64B from 2GiB     : %.3f ms  %.2f Gop/s  %.0f GB/s\n", ms, N / ms / 1e6, 64.0 * N / ms / 1e6);
  ms = timeit([&](int r) { k_gather<128><<<G, T>>>(big, BYTES / 128, N, out, r); });
  printf("gather 128B from 2GiB    : %.3f ms  %.2f Gop/s  %.0f GB/s\n", ms, N / ms / 1e6, 128.0 * N / ms / 1e6);
  ms = timeit([&](int r) { k_gather4<<<G, T>>>(small, SMALLB / 4, N, out, r); });
  printf("gather 4B  from 40MiB(L2): %.3f ms  %.2f Gop/s\n", ms, N / ms / 1e6);
  ms = timeit([&](int r) { k_gather<32><<<G, T>>>((uint4 *) small, SMALLB / 32, N, out, r); });
  printf("gather 32B from 40MiB(L2): %.3f ms  %.2f Gop/s\n", ms, N / ms / 1e6);
  ms = timeit([&](int r) { k_gather4<<<G, T>>>((uint32_t *) big, (80ull << 20) / 4, N, out, r); });
  printf("gather 4B  from 80MiB    : %.3f ms  %.2f Gop/s\n", ms, N / ms / 1e6);
  ms = timeit([&](int r) { k_gather4<<<G, T>>>((uint32_t *) big, (320ull << 20) / 4, N, out, r); });
  printf("gather 4B  from 320MiB   : %.3f ms  %.2f Gop/s\n", ms, N / ms / 1e6);
  ms = timeit([&](int r) { k_gather_dep<<<G, T>>>(small, SMALLB / 4, big, BYTES / 32, N, out, r); });
  printf("dep: 4B(L2) -> 32B(2GiB) : %.3f ms  %.2f Gop/s\n", ms, N / ms / 1e6);
  ms = timeit([&](int r) { k_atomic<<<G, T>>>(small, SMALLB / 4, N, r); });
  printf("atomicAdd(red) 40MiB     : %.3f ms  %.2f Gop/s\n", ms, N / ms / 1e6);
  ms = timeit([&](int r) { k_atomic_ret<<<G, T>>>(small, SMALLB / 4, N, r, out); });
  printf("atomicAdd(ret) 40MiB     : %.3f ms  %.2f Gop/s\n", ms, N / ms / 1e6);
  ms = timeit([&](int r) { k_atomic<<<G, T>>>((uint32_t *) big2, (400ull << 20) / 4, N, r); });
  printf("atomicAdd(red) 400MiB    : %.3f ms  %.2f Gop/s\n", ms, N / ms / 1e6);
  ms = timeit([&](int r) { k_scatter16<<<G, T>>>(big2, BYTES / 16, N, r); });
  printf("scatter 16B into 2GiB    : %.3f ms  %.2f Gop/s\n", ms, N / ms / 1e6);
  ms = timeit([&](int r) { k_scatter4<<<G, T>>>((uint32_t *) big2, BYTES / 4, N, r); });
  printf("scatter 4B into 2GiB     : %.3f ms  %.2f Gop/s\n", ms, N / ms / 1e6);
  ms = timeit([&](int r) { k_scatter4<<<G, T>>>(small, SMALLB / 4, N, r); });
  printf("scatter 4B into 40MiB    : %.3f ms  %.2f Gop/s\n", ms, N / ms / 1e6);
  return 0;
}
