// microbench_scatter.cu -- what bounds k2_deliver: N 16-byte entries streamed in,
// each stored at a (pseudo)random slot of a WINDOW that slides over the output
// as the input is consumed (dev tool; nvcc -O3 -gencode arch=compute_100a,code=sm_100a).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// mode 0: store only; 1: atomic(ret) on a cursor of the slot's group of 8, store at slot; 2: atomic only
template <int MODE, int ILP>
__global__ void k_deliver(const uint4 *__restrict__ src, uint4 *__restrict__ dst, uint32_t *__restrict__ cursor,
                          uint64_t n, uint64_t win_entries, uint64_t seed, uint32_t *out) {
  const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (uint64_t e0 = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += stride * ILP) {
    uint4 v[ILP]; uint64_t slot[ILP]; uint32_t r[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      const uint64_t e = e0 + k * stride;
      if (e < n) {
        v[k] = src[e];
        slot[k] = (e / win_entries) * win_entries + mix(e ^ seed) % win_entries;
      }
    }
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      const uint64_t e = e0 + k * stride;
      r[k] = 0;
      if (e < n && MODE >= 1) r[k] = atomicAdd(&cursor[slot[k] >> 3], 1u);
    }
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      const uint64_t e = e0 + k * stride;
      if (e < n && MODE <= 1) { v[k].x += r[k]; dst[slot[k]] = v[k]; }
      acc += r[k];
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <typename F> float timeit(F f, int reps = 4) {
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
  const uint64_t N = 40000000;                       // entries (C3: 4e7 mail entries)
  uint4 *src, *dst; uint32_t *cursor, *out;
  CK(cudaMalloc(&src, N * 16)); CK(cudaMalloc(&dst, N * 16)); CK(cudaMalloc(&cursor, N / 2)); CK(cudaMalloc(&out, 64));
  CK(cudaMemset(src, 1, N * 16)); CK(cudaMemset(dst, 0, N * 16)); CK(cudaMemset(cursor, 0, N / 2));
  const int T = 256;
  const uint64_t wins[] = {4096, 65536, 625000, 5000000, N};     // window sizes in entries (x16 B)
  for (uint64_t w : wins) {
    for (int G : {148 * 8, 148 * 32}) {
      float a = timeit([&](int r) { k_deliver<0, 4><<<G, T>>>(src, dst, cursor, N, w, r, out); });
      float b = timeit([&](int r) { k_deliver<1, 4><<<G, T>>>(src, dst, cursor, N, w, r, out); });
      float c = timeit([&](int r) { k_deliver<2, 4><<<G, T>>>(src, dst, cursor, N, w, r, out); });
      float d = timeit([&](int r) { k_deliver<1, 1><<<G, T>>>(src, dst, cursor, N, w, r, out); });
      printf("window %9llu entries (%7.1f MB) grid %5d: store %.3f ms (%5.1f G/s) | atomic+store %.3f ms (%5.1f G/s) | atomic only %.3f ms | atomic+store ILP1 %.3f ms\n",
             (unsigned long long) w, w * 16 / 1e6, G, a, N / a / 1e6, b, N / b / 1e6, c, d);
    }
  }
  return 0;
}
