// gtsb_sort_core.h -- block-cooperative bitonic sort of a hub vertex's bucket (general build,
// gtsb_build.cu: k_resolve_large), blocked through shared memory.
//
// A bucket of n half-edge entries is padded to P = 2^m and sorted by the bitonic network.  The
// plain version runs all m(m+1)/2 steps over global memory -- 120 passes over 512 KB for a hub of
// 10^4 edges, and with ~10^3 blocks in flight the buckets do not stay in L2.  Here every step whose
// partner distance j is below CHUNK runs on a chunk held in shared memory: the first log2(CHUNK)
// stages sort each chunk on chip, every later stage k makes log2(k / CHUNK) passes over global
// memory (j >= CHUNK) and finishes (j < CHUNK) chunk by chunk on chip -- 16 array passes instead
// of 120 for P = 32768.
//
// The same source is the device code and a plain C++ loop (tests/emul/sort_emul.cpp): a "group"
// (the block, or one warp for the buckets k_resolve_mid sorts in shared memory) is a loop over t,
// a barrier is nothing, and the loop runs forwards or backwards -- valid because no thread reads
// what another one writes between two barriers (the pairs (i, i + j) of one step are disjoint;
// loads and stores of a chunk use one index per thread).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) || defined(CUSIM)   // CUSIM: the host model of the device language (tests/emul/cusim)
#define GTSB_SORT_FN __device__ __forceinline__
namespace gtsbs {
typedef uint4 Ent;
// who runs a loop together: the whole block, or one warp of it (k_resolve_mid: a bucket per warp)
struct BlockGroup {
  static __device__ __forceinline__ uint32_t rank() { return threadIdx.x; }
  static __device__ __forceinline__ uint32_t size() { return blockDim.x; }
  static __device__ __forceinline__ void sync() { __syncthreads(); }
};
struct WarpGroup {
  static __device__ __forceinline__ uint32_t rank() { return threadIdx.x & 31u; }
  static __device__ __forceinline__ uint32_t size() { return 32u; }
  static __device__ __forceinline__ void sync() { __syncwarp(); }
};
}
#define GTSB_FOR_GROUP(G, t, n) for (uint32_t t = G::rank(); t < (n); t += G::size())
#else
namespace gtsbs {
struct Ent { uint32_t x, y, z, w; };
extern int emul_reverse;                    // host loop order of a "group"
inline uint32_t emul_t(uint32_t t, uint32_t n) { return emul_reverse ? n - 1u - t : t; }
struct BlockGroup { static void sync() {} };
struct WarpGroup { static void sync() {} };
}
#define GTSB_SORT_FN inline
#define GTSB_FOR_GROUP(G, t, n) \
  for (uint32_t t##_i = 0, t = 0; t##_i < (n) && ((t = ::gtsbs::emul_t(t##_i, (n))), true); t##_i++)
#endif

namespace gtsbs {

constexpr uint32_t SORT_CHUNK = 2048;                 // entries of a chunk (32 KB + 8 KB of tags)
constexpr uint32_t SORT_OTHER_MASK = (1u << 27) - 1u; // = E_OTHER_MASK (gtsb_common.cuh)

// mode 0: key (other vertex, record index) -- groups one contig pair's records in file order;
// mode 1: key x (creating record) -- adjacency order = creation order (graph.c:166-167)
GTSB_SORT_FN bool ent_gt(const Ent &a, const Ent &b, int mode) {
  if (mode != 0) return a.x > b.x;
  const uint32_t oa = a.y & SORT_OTHER_MASK, ob = b.y & SORT_OTHER_MASK;
  return oa > ob || (oa == ob && a.x > b.x);
}

// step (k, j) of the network over n entries at a[] whose global indices start at c0
template <typename G>
GTSB_SORT_FN void bitonic_step(Ent *a, uint32_t *tag, uint32_t n, uint32_t c0, uint32_t k, uint32_t j, int mode) {
  GTSB_FOR_GROUP(G, t, n >> 1) {
    const uint32_t i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u)), l = i + j;
    const bool up = ((c0 + i) & k) == 0u;
    const Ent x = a[i], y = a[l];
    if (up ? ent_gt(x, y, mode) : ent_gt(y, x, mode)) {
      a[i] = y;
      a[l] = x;
      if (tag != nullptr) {
        const uint32_t tx = tag[i];
        tag[i] = tag[l];
        tag[l] = tx;
      }
    }
  }
  G::sync();
}

template <typename G>
GTSB_SORT_FN void chunk_load(const Ent *a, const uint32_t *tag, uint32_t c0, uint32_t ch, Ent *s_a, uint32_t *s_t) {
  GTSB_FOR_GROUP(G, t, ch) {
    s_a[t] = a[c0 + t];
    if (tag != nullptr) s_t[t] = tag[c0 + t];
  }
  G::sync();
}

template <typename G>
GTSB_SORT_FN void chunk_store(Ent *a, uint32_t *tag, uint32_t c0, uint32_t ch, const Ent *s_a, const uint32_t *s_t) {
  GTSB_FOR_GROUP(G, t, ch) {
    a[c0 + t] = s_a[t];
    if (tag != nullptr) tag[c0 + t] = s_t[t];
  }
  G::sync();
}

// ascending sort of a[0 .. P) where it lies (shared memory), P a power of two; tag (optional) moves along
template <typename G>
GTSB_SORT_FN void plain_bitonic(Ent *a, uint32_t *tag, uint32_t P, int mode) {
  for (uint32_t k = 2; k <= P; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) bitonic_step<G>(a, tag, P, 0u, k, j, mode);
}

// ascending sort of a[0 .. P) in global memory, P a power of two (<= 2^30); tag (optional) moves
// along.  s_a / s_t: SORT_CHUNK entries of shared memory.  Called by every thread of the block.
GTSB_SORT_FN void blocked_bitonic(Ent *a, uint32_t *tag, uint32_t P, int mode, Ent *s_a, uint32_t *s_t) {
  typedef BlockGroup G;
  const uint32_t ch = P < SORT_CHUNK ? P : SORT_CHUNK;
  uint32_t *st = tag != nullptr ? s_t : nullptr;
  for (uint32_t c0 = 0; c0 < P; c0 += ch) {           // stages k <= ch: every chunk on chip
    chunk_load<G>(a, tag, c0, ch, s_a, s_t);
    for (uint32_t k = 2; k <= ch; k <<= 1)
      for (uint32_t j = k >> 1; j > 0; j >>= 1) bitonic_step<G>(s_a, st, ch, c0, k, j, mode);
    chunk_store<G>(a, tag, c0, ch, s_a, s_t);
  }
  for (uint32_t k = ch << 1; k <= P && k != 0u; k <<= 1) {
    for (uint32_t j = k >> 1; j >= ch; j >>= 1) bitonic_step<G>(a, tag, P, 0u, k, j, mode);   // partners in other chunks
    for (uint32_t c0 = 0; c0 < P; c0 += ch) {         // the rest of stage k stays inside a chunk
      chunk_load<G>(a, tag, c0, ch, s_a, s_t);
      for (uint32_t j = ch >> 1; j > 0; j >>= 1) bitonic_step<G>(s_a, st, ch, c0, k, j, mode);
      chunk_store<G>(a, tag, c0, ch, s_a, s_t);
    }
  }
}

}  // namespace gtsbs
