// gtsb_filter2.cu -- bandwidth-oriented kernels of the filter (same closed form
// as gtsb_filter.cu, which keeps the hub paths and the fix-point helpers).
//
// Mapping: one block owns FSEG consecutive row positions; threads map to the
// SLOTS of those rows (not to rows), so that loads of the slot columns are
// coalesced, neighbouring threads run loops of equal length and the per-pair
// work of a row is spread over its own slots.  Per-vertex facts needed per
// neighbour are packed so that each pass makes ONE gather per slot:
//   vinfo[v] = {copy_num, seq_len | marked-on-entry << 31}          (pairs pass)
//   vres[v]  = polyTime (27 bits) | F[v,antisense] | F[v,sense] | repeat-pred
//                                                                   (final pass)
#include "gtsb_common.cuh"
#include "gtsb_scan.cuh"
#include "gtsb_kernels.h"

namespace gtsb {

constexpr int FSEG = 128;              // row positions per block
constexpr int FTHREADS = 256;
constexpr uint32_t FCAP = 1920;        // staged slots per chunk of the pairs pass (static smem < 48 KB)

constexpr uint32_t VI_MARKED = 1u << 31;
constexpr uint32_t VR_TIME_MASK = (1u << 27) - 1u;   // all ones = never
constexpr uint32_t VR_F0 = 1u << 27, VR_F1 = 1u << 28, VR_REP = 1u << 29;

__device__ __forceinline__ uint32_t vertex_at2(const GraphArgs &g, uint32_t p) {
  return g.vid != nullptr ? g.vid[p] : p;
}

// row of virtual slot i: largest j with s_off[j] <= i  (s_off has nrows+1 entries)
__device__ __forceinline__ uint32_t row_of(const uint32_t *s_off, uint32_t nrows, uint32_t i) {
  uint32_t lo = 0, hi = nrows;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (s_off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void warp_append2(bool pred, uint2 value, uint2 *list, uint32_t cap,
                                             uint32_t *count, uint32_t *overflow) {
  const unsigned mask = __ballot_sync(0xffffffffu, pred);
  if (mask == 0) return;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if ((int) lane_id() == leader) base = atomicAdd(count, (uint32_t) __popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) {
    const uint32_t at = base + __popc(mask & ((1u << lane_id()) - 1u));
    if (at < cap) list[at] = value; else atomicOr(overflow, 1u);
  }
}

// ------------------------------------------------------------------ packed vertex facts

__global__ void __launch_bounds__(256) k3_vinfo(uint32_t V, const VAttr *__restrict__ vattr,
                                                 const uint8_t *__restrict__ vstate,
                                                 uint2 *__restrict__ vinfo, uint32_t *__restrict__ err) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const VAttr a = vattr[v];
  if (a.seq_len & VI_MARKED) atomicOr(err, 4u);            // seq_len must fit 31 bits
  vinfo[v] = make_uint2(__float_as_uint(a.copy_num),
                        a.seq_len | (vertex_state_marked(vstate[v]) ? VI_MARKED : 0u));
}

__global__ void __launch_bounds__(256) k3_vres(uint32_t V, const uint32_t *__restrict__ poly,
                                                const uint8_t *__restrict__ fstat,
                                                const uint8_t *__restrict__ rep_pred,
                                                uint32_t *__restrict__ vres) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const uint32_t t = poly[v];
  const uint32_t f = fstat[v];
  vres[v] = (t == NO_TIME ? VR_TIME_MASK : t) | ((f & 1u) ? VR_F0 : 0u) | ((f & 2u) ? VR_F1 : 0u) |
            ((rep_pred != nullptr && rep_pred[v]) ? VR_REP : 0u);
}

// ------------------------------------------------------------------ phase 1 + static overlap

// Proposals of check_mark_polymorphic (algorithms.c:283-295) and, in the same
// sweep over the pairs, G0[v,s] = "some same-direction pair of edges that are
// unmarked on entry overlaps by more than ocutoff" (algorithms.c:301-324 before
// any mark of this filter run is taken into account; k_fire_init repairs the
// rows next to polymorphic vertices).
__global__ void __launch_bounds__(FTHREADS) k3_pairs(FilterArgs a) {
  const GraphArgs &g = a.g;
  __shared__ uint32_t s_r0[FSEG], s_off[FSEG + 1], s_v[FSEG];
  __shared__ uint8_t s_g0[FSEG], s_g1[FSEG];
  __shared__ uint32_t s_chunk_end;
  __shared__ int32_t s_dist[FCAP];
  __shared__ float s_std[FCAP], s_cn[FCAP];
  __shared__ uint32_t s_len[FCAP], s_dst[FCAP];
  __shared__ uint8_t s_fl[FCAP], s_prop[FCAP], s_row[FCAP];
  const uint32_t p0 = blockIdx.x * FSEG;
  const uint32_t nrows = min((uint32_t) FSEG, g.V - p0);
  {
    const uint32_t j = threadIdx.x;
    uint32_t d = 0;
    if (j < nrows) {
      const uint32_t p = p0 + j, v = vertex_at2(g, p);
      const uint32_t r0 = g.rs[p];
      d = g.re[p] - r0;
      s_r0[j] = r0;
      s_v[j] = v;
      s_g0[j] = s_g1[j] = 0;
      // rows that cannot propose or fire: marked on entry (algorithms.c:279), < 2 edges;
      // hubs go to the block-per-row kernels
      if (d < 2 || d > BIG_ROW || (a.vinfo[v].y & VI_MARKED)) d = 0;
    }
    uint32_t total;
    const uint32_t ex = block_excl_scan(d, &total);
    if (j < nrows) s_off[j] = ex;
    if (j == 0) s_off[nrows] = total;
  }
  __syncthreads();
  uint32_t jb = 0;
  while (jb < nrows) {
    // chunk [jb, je): as many rows as fit the staging buffers
    if (threadIdx.x == 0) {
      uint32_t je = jb + 1;
      while (je < nrows && s_off[je + 1] - s_off[jb] <= FCAP) je++;
      s_chunk_end = je;
    }
    __syncthreads();
    const uint32_t je = s_chunk_end;
    const uint32_t base = s_off[jb], n = s_off[je] - base;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t j = jb + row_of(s_off + jb, je - jb, base + i);
      const uint32_t slot = s_r0[j] + (base + i - s_off[j]);
      const uint32_t w = g.dst[slot];
      const uint2 vi = a.vinfo[w];
      const uint32_t f = g.flags[slot];
      const bool wm = (vi.y & VI_MARKED) != 0;
      const bool ok = a.fused_repeats ? !wm : !edge_state_marked(g.estate[slot]);
      s_row[i] = (uint8_t) (j - jb);
      s_dst[i] = w;
      s_dist[i] = g.dist[slot];
      s_std[i] = g.std_dev[slot];
      s_cn[i] = __uint_as_float(vi.x);
      s_len[i] = vi.y & ~VI_MARKED;
      s_fl[i] = (uint8_t) ((f & F_SENSE) | (ok ? 2u : 0u) | (wm ? 4u : 0u));
      s_prop[i] = 0;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t j = jb + s_row[i];
      const uint32_t end = s_off[j + 1] - base;
      const uint32_t fi = s_fl[i];
      const int32_t di = s_dist[i];
      const float si = s_std[i], ci = s_cn[i];
      const uint32_t li = s_len[i];
      bool fire = false;
      for (uint32_t k = i + 1; k < end; k++) {
        const uint32_t fk = s_fl[k];
        if ((fk ^ fi) & F_SENSE) continue;
        // check_mark_polymorphic, algorithms.c:232-238 (edge1 = earlier adjacency slot)
        if (ambiguous_order(di, si, s_dist[k], s_std[k], a.ambig) && __fadd_rn(ci, s_cn[k]) < a.cncutoff)
          s_prop[ci < s_cn[k] ? i : k] = 1;
        if ((fi & fk & 2u) && a.ocutoff >= 0)
          fire |= interval_overlap(di, li, s_dist[k], s_len[k]) > a.ocutoff;
      }
      if (fire) ((fi & F_SENSE) ? s_g1 : s_g0)[j] = 1;            // every writer stores the same value
    }
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
      const uint32_t i = i0 + threadIdx.x;
      const bool emit = i < n && s_prop[i] && !(s_fl[i] & 4u);               // algorithms.c:242
      const uint2 pr = emit ? make_uint2(s_v[jb + s_row[i]], s_dst[i]) : make_uint2(0u, 0u);
      warp_append2(emit, pr, a.proposals, a.proposals_cap, &g.counters[CNT_PROPOSALS],
                   &g.counters[CNT_OVERFLOW]);
    }
    __syncthreads();
    jb = je;
  }
  if (threadIdx.x < nrows) {
    const uint32_t p = p0 + threadIdx.x;
    if (g.re[p] - g.rs[p] <= BIG_ROW)
      a.gbits[s_v[threadIdx.x]] = (uint8_t) ((s_g0[threadIdx.x] ? 1u : 0u) | (s_g1[threadIdx.x] ? 2u : 0u));
  }
}

// rows whose static overlap answer may be stale: every neighbour of a vertex
// that became polymorphic, and that vertex itself
__global__ void __launch_bounds__(256) k3_dirty(FilterArgs a, uint32_t n) {
  const GraphArgs &g = a.g;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2 pr = a.proposals[i];
  if (a.poly_cur[pr.y] != pr.x) return;             // not the winning proposer
  const uint32_t p = g.pos != nullptr ? g.pos[pr.y] : pr.y;
  a.dirty[pr.y] = 1;
  for (uint32_t s = g.rs[p]; s < g.re[p]; s++) a.dirty[g.dst[s]] = 1;
}

// ------------------------------------------------------------------ fire, first (dense) round

__global__ void __launch_bounds__(FTHREADS) k3_fire_dense(FilterArgs a, uint32_t *__restrict__ work_out,
                                                           uint32_t *__restrict__ n_out) {
  const GraphArgs &g = a.g;
  __shared__ uint32_t s_r0[FSEG], s_off[FSEG + 1], s_v[FSEG], s_res[FSEG];
  __shared__ uint8_t s_und[FSEG];
  const uint32_t p0 = blockIdx.x * FSEG;
  const uint32_t nrows = min((uint32_t) FSEG, g.V - p0);
  {
    const uint32_t j = threadIdx.x;
    uint32_t d = 0;
    if (j < nrows) {
      const uint32_t p = p0 + j, v = vertex_at2(g, p);
      const uint32_t r0 = g.rs[p];
      const uint32_t und = (~(uint32_t) a.fstat[v] >> 2) & 3u;
      d = und ? g.re[p] - r0 : 0u;
      s_r0[j] = r0;
      s_v[j] = v;
      s_und[j] = (uint8_t) und;
      s_res[j] = 0;
    }
    uint32_t total;
    const uint32_t ex = block_excl_scan(d, &total);
    if (j < nrows) s_off[j] = ex;
    if (j == 0) s_off[nrows] = total;
  }
  __syncthreads();
  const uint32_t n = s_off[nrows];
  const volatile uint8_t *fstat = a.fstat;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t j = row_of(s_off, nrows, i);
    const uint32_t slot = s_r0[j] + (i - s_off[j]);
    const uint32_t u = g.dst[slot], v = s_v[j];
    if (u >= v) continue;
    const uint32_t f = g.flags[slot];
    const bool rs = (f & F_RSENSE) != 0, rm = (f & F_RSAME) != 0;
    const uint32_t s = twin_dir(rs, rm) ? 1u : 0u;
    if (!((s_und[j] >> s) & 1u)) continue;
    const uint32_t su = fstat[u], du = rs ? 1u : 0u;
    if ((su >> (2 + du)) & 1u) {
      if ((su >> du) & 1u) atomicOr(&s_res[j], 1u << s);          // a smaller neighbour fired into (v,s)
    } else {
      atomicOr(&s_res[j], 4u << s);                               // still pending
    }
  }
  __syncthreads();
  bool again = false;
  uint32_t p = 0;
  if (threadIdx.x < nrows && s_und[threadIdx.x]) {
    const uint32_t j = threadIdx.x, v = s_v[j], und = s_und[j], res = s_res[j];
    p = p0 + j;
    uint8_t st = a.fstat[v];
    for (uint32_t s = 0; s < 2; s++) {
      if (!((und >> s) & 1u)) continue;
      if ((res >> s) & 1u) st |= (uint8_t) (4u << s);
      else if (!((res >> (2 + s)) & 1u)) st |= (uint8_t) ((4u << s) | (1u << s));
    }
    a.fstat[v] = st;
    again = (st & 0x0C) != 0x0C;
  }
  warp_append(again, p, work_out, n_out);
}

// ------------------------------------------------------------------ final states

__global__ void __launch_bounds__(FTHREADS) k3_finalize(FilterArgs a) {
  const GraphArgs &g = a.g;
  __shared__ uint32_t s_r0[FSEG], s_off[FSEG + 1], s_v[FSEG], s_own[FSEG];
  __shared__ int s_inc[2 * FSEG];
  const uint32_t p0 = blockIdx.x * FSEG;
  const uint32_t nrows = min((uint32_t) FSEG, g.V - p0);
  {
    const uint32_t j = threadIdx.x;
    uint32_t d = 0;
    if (j < nrows) {
      const uint32_t p = p0 + j, v = vertex_at2(g, p);
      const uint32_t r0 = g.rs[p];
      d = g.re[p] - r0;
      const uint32_t own = a.vres[v];
      s_r0[j] = r0;
      s_v[j] = v;
      s_own[j] = own;
      s_inc[2 * j] = s_inc[2 * j + 1] = -1;
      if ((own & VR_TIME_MASK) != VR_TIME_MASK) g.vstate[v] = GIS_POLYMORPHIC;
    }
    uint32_t total;
    const uint32_t ex = block_excl_scan(d, &total);
    if (j < nrows) s_off[j] = ex;
    if (j == 0) s_off[nrows] = total;
  }
  __syncthreads();
  const uint32_t n = s_off[nrows];
  constexpr int KEEP = 8;                 // per-thread cache of the neighbour facts between passes
  uint32_t keep_ru[KEEP], keep_j[KEEP];
  // pass 1: latest neighbour that fired into (v, s)
  {
    int it = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x, it++) {
      const uint32_t j = row_of(s_off, nrows, i);
      const uint32_t slot = s_r0[j] + (i - s_off[j]);
      const uint32_t u = g.dst[slot], f = g.flags[slot];
      const bool rs = (f & F_RSENSE) != 0, rm = (f & F_RSAME) != 0;
      const uint32_t ru = a.vres[u];
      if (it < KEEP) {
        keep_ru[it] = ru;
        keep_j[it] = j;
      }
      if (ru & (rs ? VR_F1 : VR_F0)) atomicMax(&s_inc[2 * j + (twin_dir(rs, rm) ? 1 : 0)], (int) u);
    }
  }
  __syncthreads();
  // pass 2: last writer wins (INCONSISTENT on ties: phase 3 follows phase 1)
  int it = 0;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x, it++) {
    const uint32_t j = it < KEEP ? keep_j[it] : row_of(s_off, nrows, i);
    const uint32_t slot = s_r0[j] + (i - s_off[j]);
    const uint32_t u = g.dst[slot], f = g.flags[slot];
    const uint32_t ru = it < KEEP ? keep_ru[it] : a.vres[u], own = s_own[j];
    const uint32_t s = (f & F_SENSE) ? 1u : 0u;
    const uint32_t tu = ru & VR_TIME_MASK, tv = own & VR_TIME_MASK;
    const int pw = tu == VR_TIME_MASK ? -1 : (int) tu, pv = tv == VR_TIME_MASK ? -1 : (int) tv;
    const int tp = pv > pw ? pv : pw;
    int ti = s_inc[2 * j + s];
    if ((own & (s ? VR_F1 : VR_F0)) && (int) s_v[j] > ti) ti = (int) s_v[j];
    if (tp < 0 && ti < 0) {
      if (a.fused_repeats) g.estate[slot] = ((own | ru) & VR_REP) ? GIS_REPEAT : GIS_UNVISITED;
    } else {
      g.estate[slot] = ti >= tp ? GIS_INCONSISTENT : GIS_POLYMORPHIC;
    }
  }
}

// ------------------------------------------------------------------ launchers

void launch_vinfo(const FilterArgs &a, cudaStream_t s) {
  if (a.g.V == 0) return;
  KernelTimer t_("k3_vinfo", s);
  k3_vinfo<<<(a.g.V + 255) / 256, 256, 0, s>>>(a.g.V, a.g.vattr, a.g.vstate, a.vinfo,
                                                a.g.counters + CNT_ERROR);
}

void launch_pairs2(const FilterArgs &a, cudaStream_t s) {
  if (a.g.V == 0) return;
  KernelTimer t_("k3_pairs", s);
  k3_pairs<<<(a.g.V + FSEG - 1) / FSEG, FTHREADS, 0, s>>>(a);
}

void launch_dirty(const FilterArgs &a, uint32_t n_proposals, cudaStream_t s) {
  if (n_proposals == 0) return;
  KernelTimer t_("k3_dirty", s);
  k3_dirty<<<(n_proposals + 255) / 256, 256, 0, s>>>(a, n_proposals);
}

void launch_fire_dense(const FilterArgs &a, uint32_t *work_out, uint32_t *n_out, cudaStream_t s) {
  if (a.g.V == 0) return;
  KernelTimer t_("k3_fire_dense", s);
  k3_fire_dense<<<(a.g.V + FSEG - 1) / FSEG, FTHREADS, 0, s>>>(a, work_out, n_out);
}

void launch_finalize2(const FilterArgs &a, const uint8_t *rep_pred, cudaStream_t s) {
  if (a.g.V == 0) return;
  {
    KernelTimer t_("k3_vres", s);
    k3_vres<<<(a.g.V + 255) / 256, 256, 0, s>>>(a.g.V, a.poly_cur, a.fstat, rep_pred, a.vres);
  }
  KernelTimer t_("k3_finalize", s);
  k3_finalize<<<(a.g.V + FSEG - 1) / FSEG, FTHREADS, 0, s>>>(a);
}

}  // namespace gtsb
