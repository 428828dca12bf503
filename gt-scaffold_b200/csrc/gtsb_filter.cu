// gtsb_filter.cu -- mark_repeats and the polymorphic / inconsistent filter on the
// device-resident CSR graph.
//
// mark_repeats (reference gt_scaffolder_algorithms.c:160-166 with mark_vertex /
// mark_edge, :61-87) has the closed form
//     vstate[v]      = REPEAT  iff pred(v)
//     estate[v -> w] = REPEAT  iff pred(v) || pred(w)          (else untouched)
//
// gt_scaffolder_graph_filter (:261-343) is a sequential sweep over vertices in
// index order whose marks feed later iterations.  It is evaluated here through
// the equivalent closed form of SURVEY.md section 8(a) (validated against the
// compiled reference, tests/test_filter_closed_form.py), with "time" = index of
// the vertex being processed:
//
//   PC(t)       targets chosen by check_mark_polymorphic over the same-sense
//               pairs (i<j in adjacency order) of t                     (:283-295)
//   polyTime(p) = min{ t : A[t], p in PC(t) }  if p is unmarked on entry
//   A[t]        = t unmarked on entry  &&  !(polyTime(t) < t)           (:279)
//   U(v->w)     = edge unmarked on entry && !(polyTime(w) <= v)
//                 && !(some u<v fired into (v, sense(v->w)))            (:304-309)
//   F[v,s]      = A[v] && max(0, max overlap over U-pairs of direction s) > ocutoff
//   u "fires into" (x, td) for every edge u->x of a fired direction, with
//               td = sense ? !same : same  of THAT edge                (:326-338)
//   final edge state = last writer: INCONSISTENT at the largest firing time that
//               touches it, POLYMORPHIC at max(polyTime(v), polyTime(w)); a tie
//               goes to INCONSISTENT (phase 3 runs after phase 1).
//
// Requirement: the graph is "paired" -- every edge v->w has exactly one reverse
// edge w->v (true for every graph the reference's constructor can produce,
// parser.c:374-377); the build stores the reverse edge's sense/same per slot.
#include "gtsb_common.cuh"
#include "gtsb_scan.cuh"
#include "gtsb_kernels.h"

namespace gtsb {

__device__ __forceinline__ uint32_t vertex_at(const GraphArgs &g, uint32_t p) {
  return g.vid != nullptr ? g.vid[p] : p;
}

// ------------------------------------------------------------------ mark_repeats

__global__ void __launch_bounds__(256) k_repeat_vertices(uint32_t V, const float *__restrict__ astat,
                                                          const VAttr *__restrict__ vattr,
                                                          float copy_num_cutoff, float astat_cutoff,
                                                          int use_copy_num,
                                                          uint8_t *__restrict__ rep_pred,
                                                          uint8_t *__restrict__ vstate) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  // algorithms.c:163-164 (float compares)
  const bool pred = astat[v] <= astat_cutoff || (use_copy_num && vattr[v].copy_num < copy_num_cutoff);
  rep_pred[v] = pred ? 1 : 0;
  if (pred) vstate[v] = GIS_REPEAT;
}

__global__ void __launch_bounds__(256) k_repeat_edges(GraphArgs g, const uint8_t *__restrict__ rep_pred) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t *__restrict__ dst = g.dst;
  uint8_t *__restrict__ estate = g.estate;
  uint32_t r0 = 0, d = 0;
  bool pv = false;
  if (p < g.V) {
    r0 = g.rs[p];
    d = g.re[p] - r0;
    pv = rep_pred[vertex_at(g, p)] != 0;
  }
  const bool big = d > BIG_ROW;
  if (!big)
    for (uint32_t k = 0; k < d; k++)
      if (pv || rep_pred[dst[r0 + k]]) estate[r0 + k] = GIS_REPEAT;
  unsigned todo = __ballot_sync(0xffffffffu, big);
  while (todo) {
    const int l = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t rr = __shfl_sync(0xffffffffu, r0, l), dd = __shfl_sync(0xffffffffu, d, l);
    const bool pp = __shfl_sync(0xffffffffu, (int) pv, l) != 0;
    for (uint32_t k = lane_id(); k < dd; k += 32)
      if (pp || rep_pred[dst[rr + k]]) estate[rr + k] = GIS_REPEAT;
  }
}

void launch_mark_repeats(const GraphArgs &g, uint8_t *rep_pred, float copy_num_cutoff,
                         float astat_cutoff, int use_copy_num, cudaStream_t s) {
  if (g.V == 0) return;
  const uint32_t blocks = (g.V + 255) / 256;
  { KernelTimer t_("k_repeat_vertices", s);
  k_repeat_vertices<<<blocks, 256, 0, s>>>(g.V, g.astat, g.vattr, copy_num_cutoff, astat_cutoff,
                                           use_copy_num, rep_pred, g.vstate); }
  KernelTimer t2_("k_repeat_edges", s);
  k_repeat_edges<<<blocks, 256, 0, s>>>(g, rep_pred);
}

void launch_repeat_vertices(const GraphArgs &g, uint8_t *rep_pred, float copy_num_cutoff,
                            float astat_cutoff, int use_copy_num, cudaStream_t s) {
  if (g.V == 0) return;
  KernelTimer t_("k_repeat_vertices", s);
  k_repeat_vertices<<<(g.V + 255) / 256, 256, 0, s>>>(g.V, g.astat, g.vattr, copy_num_cutoff, astat_cutoff,
                                                      use_copy_num, rep_pred, g.vstate);
}

// ------------------------------------------------------------------ filter, phase 1

__device__ __forceinline__ void append_proposals(uint32_t t, uint32_t n, const uint32_t *targets,
                                                 uint2 *proposals, uint32_t cap, uint32_t *counters) {
  if (n == 0) return;
  const uint32_t base = atomicAdd(&counters[CNT_PROPOSALS], n);
  if (base + n > cap) {
    atomicOr(&counters[CNT_OVERFLOW], 1u);
    return;
  }
  for (uint32_t k = 0; k < n; k++) proposals[base + k] = make_uint2(t, targets[k]);
}

// thread per vertex, rows <= BIG_ROW: all same-sense pairs (i<j) of the row
__global__ void __launch_bounds__(128) k_pairs_small(FilterArgs a) {
  const GraphArgs &g = a.g;
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.V) return;
  const uint32_t r0 = g.rs[p], d = g.re[p] - r0;
  if (d < 2 || d > BIG_ROW) return;
  const uint32_t v = vertex_at(g, p);
  if (vertex_state_marked(g.vstate[v])) return;                    // algorithms.c:279
  int32_t dist[BIG_ROW];
  float sd[BIG_ROW], cn[BIG_ROW];
  uint32_t nb[BIG_ROW];
  uint32_t sense_mask = 0;
  for (uint32_t k = 0; k < d; k++) {
    dist[k] = g.dist[r0 + k];
    sd[k] = g.std_dev[r0 + k];
    nb[k] = g.dst[r0 + k];
    cn[k] = g.vattr[nb[k]].copy_num;
    if (g.flags[r0 + k] & F_SENSE) sense_mask |= 1u << k;
  }
  uint32_t prop = 0;
  for (uint32_t i = 0; i + 1 < d; i++) {
    const uint32_t si = (sense_mask >> i) & 1u;
    for (uint32_t j = i + 1; j < d; j++) {
      if (((sense_mask >> j) & 1u) != si) continue;
      // check_mark_polymorphic, algorithms.c:232-238
      if (ambiguous_order(dist[i], sd[i], dist[j], sd[j], a.ambig) &&
          __fadd_rn(cn[i], cn[j]) < a.cncutoff)
        prop |= 1u << (cn[i] < cn[j] ? i : j);
    }
  }
  if (prop == 0) return;
  uint32_t targets[BIG_ROW], n = 0;
  for (uint32_t k = 0; k < d; k++)
    if (((prop >> k) & 1u) && !vertex_state_marked(g.vstate[nb[k]])) targets[n++] = nb[k];   // :242
  append_proposals(v, n, targets, a.proposals, a.proposals_cap, g.counters);
}

// block per big row; per-block scratch: copy_num[max_deg] f32, mark[max_deg] u8
__global__ void __launch_bounds__(512) k_pairs_big(FilterArgs a) {
  const GraphArgs &g = a.g;
  float *cn = reinterpret_cast<float *>(a.big_scratch + (size_t) blockIdx.x * g.max_deg * BIG_SCRATCH_STRIDE);
  uint8_t *mark = reinterpret_cast<uint8_t *>(cn + 2 * (size_t) g.max_deg);
  for (uint32_t li = blockIdx.x; li < g.n_big_rows; li += gridDim.x) {
    const uint32_t p = g.big_rows[li];
    const uint32_t v = vertex_at(g, p);
    if (vertex_state_marked(g.vstate[v])) continue;
    const uint32_t r0 = g.rs[p], d = g.re[p] - r0;
    for (uint32_t k = threadIdx.x; k < d; k += blockDim.x) {
      cn[k] = g.vattr[g.dst[r0 + k]].copy_num;
      mark[k] = 0;
    }
    __syncthreads();
    for (uint32_t i = 0; i + 1 < d; i++) {
      const int32_t di = g.dist[r0 + i];
      const float si = g.std_dev[r0 + i], ci = cn[i];
      const uint32_t fi = g.flags[r0 + i] & F_SENSE;
      for (uint32_t j = i + 1 + threadIdx.x; j < d; j += blockDim.x) {
        if ((g.flags[r0 + j] & F_SENSE) != fi) continue;
        const float cj = cn[j];
        if (ambiguous_order(di, si, g.dist[r0 + j], g.std_dev[r0 + j], a.ambig) &&
            __fadd_rn(ci, cj) < a.cncutoff)
          mark[ci < cj ? i : j] = 1;
      }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < d; k += blockDim.x) {
      if (mark[k]) {
        const uint32_t p = g.dst[r0 + k];
        if (!vertex_state_marked(g.vstate[p]))
          append_proposals(v, 1, &p, a.proposals, a.proposals_cap, g.counters);
      }
    }
    __syncthreads();
  }
}

void launch_pairs_big(const FilterArgs &a, cudaStream_t s) {
  if (a.g.n_big_rows == 0) return;
  KernelTimer t_("k_pairs_big", s);
  k_pairs_big<<<a.big_blocks, 512, 0, s>>>(a);
}

void launch_filter_pairs(const FilterArgs &a, cudaStream_t s) {
  if (a.g.V == 0) return;
  { KernelTimer t_("k_pairs_small", s);
  k_pairs_small<<<(a.g.V + 127) / 128, 128, 0, s>>>(a); }
  if (a.g.n_big_rows) { KernelTimer t_("k_pairs_big", s); k_pairs_big<<<a.big_blocks, 512, 0, s>>>(a); }
}

// ------------------------------------------------------------------ polyTime fixpoint

// One Jacobi sweep of  polyTime(p) = min{ t : A[t], (t,p) proposed },
// A[t] = !(polyTime(t) < t).  Every dependency points to a smaller index, so
// the sweeps converge to the unique solution in (longest chain) iterations.
__global__ void __launch_bounds__(256) k_poly_reset(const uint2 *__restrict__ proposals, uint32_t n,
                                                     uint32_t *__restrict__ poly_new) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) poly_new[proposals[i].y] = NO_TIME;
}
__global__ void __launch_bounds__(256) k_poly_propose(const uint2 *__restrict__ proposals, uint32_t n,
                                                       const uint32_t *__restrict__ poly_cur,
                                                       uint32_t *__restrict__ poly_new) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2 pr = proposals[i];
  if (!(poly_cur[pr.x] < pr.x)) atomicMin(&poly_new[pr.y], pr.x);
}
__global__ void __launch_bounds__(256) k_poly_commit(const uint2 *__restrict__ proposals, uint32_t n,
                                                      uint32_t *__restrict__ poly_cur,
                                                      const uint32_t *__restrict__ poly_new,
                                                      uint32_t *__restrict__ counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = proposals[i].y;
  const uint32_t nv = poly_new[p];
  if (atomicExch(&poly_cur[p], nv) != nv) counters[CNT_POLY_CHANGED] = 1;
}

void launch_poly_sweep(const FilterArgs &a, uint32_t n, cudaStream_t s) {
  if (n == 0) return;
  const uint32_t blocks = (n + 255) / 256;
  KernelTimer t_("k_poly_sweep(3 kernels)", s);
  k_poly_reset<<<blocks, 256, 0, s>>>(a.proposals, n, a.poly_new);
  k_poly_propose<<<blocks, 256, 0, s>>>(a.proposals, n, a.poly_cur, a.poly_new);
  k_poly_commit<<<blocks, 256, 0, s>>>(a.proposals, n, a.poly_cur, a.poly_new, a.g.counters);
}

// ------------------------------------------------------------------ filter, phase 2

// fstat bits: 0/1 = F[v, antisense/sense], 2/3 = that direction is decided
constexpr uint8_t FS_DECIDED_ALL = 0x0C;

// G[v,s]: max overlap over same-direction pairs that are unmarked when v is
// reached, not yet counting the fires of smaller neighbours (algorithms.c:301-320)
__global__ void __launch_bounds__(128) k_overlap_small(FilterArgs a) {
  const GraphArgs &g = a.g;
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  bool queue = false;
  if (p < g.V) {
    const uint32_t r0 = g.rs[p], d = g.re[p] - r0;
    const uint32_t v = vertex_at(g, p);
    if (d <= BIG_ROW) {
      const bool active = !vertex_state_marked(g.vstate[v]) && !(a.poly_cur[v] < v);
      uint8_t gb = 0;
      if (active && a.ocutoff < 0) {
        gb = 3;            // 0 > ocutoff: both directions fire whatever the pairs (:301-324)
      } else if (active && a.dirty != nullptr && !a.dirty[v]) {
        gb = a.gbits[v];   // k3_pairs' static answer stands: no polymorphic vertex nearby
      } else if (active && d >= 2) {
        int32_t dist[BIG_ROW];
        uint32_t len[BIG_ROW];
        uint32_t sense_mask = 0, ok_mask = 0;
        for (uint32_t k = 0; k < d; k++) {
          const uint32_t w = g.dst[r0 + k];
          dist[k] = g.dist[r0 + k];
          len[k] = g.vattr[w].seq_len;
          if (g.flags[r0 + k] & F_SENSE) sense_mask |= 1u << k;
          if (!edge_state_marked(g.estate[r0 + k]) && !(a.poly_cur[w] <= v)) ok_mask |= 1u << k;
        }
        long long mx[2] = {0, 0};
        for (uint32_t i = 0; i + 1 < d; i++) {
          if (!((ok_mask >> i) & 1u)) continue;
          const uint32_t si = (sense_mask >> i) & 1u;
          for (uint32_t j = i + 1; j < d; j++) {
            if (!((ok_mask >> j) & 1u) || ((sense_mask >> j) & 1u) != si) continue;
            const long long ov = interval_overlap(dist[i], len[i], dist[j], len[j]);
            if (ov > mx[si]) mx[si] = ov;
          }
        }
        gb = (uint8_t) ((mx[0] > a.ocutoff ? 1 : 0) | (mx[1] > a.ocutoff ? 2 : 0));
      }
      a.gbits[v] = gb;
      if (a.ocutoff < 0) {
        a.fstat[v] = FS_DECIDED_ALL | gb;          // no dependence on neighbours
      } else {
        a.fstat[v] = (uint8_t) (FS_DECIDED_ALL & ~(gb << 2));
        queue = gb != 0 && a.dirty == nullptr;     // the dense first round needs no worklist
      }
    }
  }
  warp_append(queue, p, a.work_a, &g.counters[CNT_WORK_A]);
}

__global__ void __launch_bounds__(512) k_overlap_big(FilterArgs a) {
  const GraphArgs &g = a.g;
  __shared__ long long s_mx[2];
  uint32_t *len = reinterpret_cast<uint32_t *>(a.big_scratch + (size_t) blockIdx.x * g.max_deg * BIG_SCRATCH_STRIDE);
  uint8_t *ok = reinterpret_cast<uint8_t *>(len + 2 * (size_t) g.max_deg);
  for (uint32_t li = blockIdx.x; li < g.n_big_rows; li += gridDim.x) {
    const uint32_t p = g.big_rows[li];
    const uint32_t v = vertex_at(g, p);
    const uint32_t r0 = g.rs[p], d = g.re[p] - r0;
    const bool active = !vertex_state_marked(g.vstate[v]) && !(a.poly_cur[v] < v);
    uint8_t gb = 0;
    if (active && a.ocutoff < 0) {
      gb = 3;
    } else if (active) {
      if (threadIdx.x < 2) s_mx[threadIdx.x] = 0;
      for (uint32_t k = threadIdx.x; k < d; k += blockDim.x) {
        const uint32_t w = g.dst[r0 + k];
        len[k] = g.vattr[w].seq_len;
        ok[k] = (!edge_state_marked(g.estate[r0 + k]) && !(a.poly_cur[w] <= v)) ? 1 : 0;
      }
      __syncthreads();
      long long mx[2] = {0, 0};
      for (uint32_t i = 0; i + 1 < d; i++) {
        if (!ok[i]) continue;
        const int32_t di = g.dist[r0 + i];
        const uint32_t li_ = len[i];
        const uint32_t fi = g.flags[r0 + i] & F_SENSE;
        for (uint32_t j = i + 1 + threadIdx.x; j < d; j += blockDim.x) {
          if (!ok[j] || (g.flags[r0 + j] & F_SENSE) != fi) continue;
          const long long ov = interval_overlap(di, li_, g.dist[r0 + j], len[j]);
          if (ov > mx[fi]) mx[fi] = ov;
        }
      }
      if (mx[0] > 0) atomicMax(&s_mx[0], mx[0]);
      if (mx[1] > 0) atomicMax(&s_mx[1], mx[1]);
      __syncthreads();
      gb = (uint8_t) ((s_mx[0] > a.ocutoff ? 1 : 0) | (s_mx[1] > a.ocutoff ? 2 : 0));
    }
    if (threadIdx.x == 0) {
      a.gbits[v] = gb;
      if (a.ocutoff < 0) {
        a.fstat[v] = FS_DECIDED_ALL | gb;
      } else {
        a.fstat[v] = (uint8_t) (FS_DECIDED_ALL & ~(gb << 2));
        if (gb && a.dirty == nullptr) a.work_a[atomicAdd(&g.counters[CNT_WORK_A], 1u)] = p;
      }
    }
    __syncthreads();
  }
}

void launch_filter_overlap(const FilterArgs &a, cudaStream_t s) {
  if (a.g.V == 0) return;
  { KernelTimer t_("k_fire_init", s);
  k_overlap_small<<<(a.g.V + 127) / 128, 128, 0, s>>>(a); }
  if (a.g.n_big_rows) { KernelTimer t_("k_overlap_big", s); k_overlap_big<<<a.big_blocks, 512, 0, s>>>(a); }
}

// ------------------------------------------------------------------ fire fixpoint

// F[v,s] = G[v,s] && no neighbour u < v with F[u, sense(u->v)] and
// twin_dir(u->v) == s.  A direction is decided once every smaller neighbour
// that could fire into it is decided; each round decides at least the smallest
// undecided vertex, and on random orders the depth is logarithmic.
__global__ void __launch_bounds__(128) k_fire_round(FilterArgs a, const uint32_t *__restrict__ work_in,
                                                     uint32_t n_in, uint32_t *__restrict__ work_out,
                                                     uint32_t *__restrict__ n_out) {
  const GraphArgs &g = a.g;
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  bool again = false;
  uint32_t p = 0;
  if (idx < n_in) {
    p = work_in[idx];
    const uint32_t v = vertex_at(g, p);
    const volatile uint8_t *fstat = a.fstat;
    uint8_t st = fstat[v];
    const uint32_t und = (~(uint32_t) st >> 2) & 3u;
    const uint32_t r0 = g.rs[p], d = g.re[p] - r0;
    uint32_t out = 0, pend = 0;
    for (uint32_t k = 0; k < d; k++) {
      const uint32_t u = g.dst[r0 + k];
      if (u >= v) continue;
      const uint32_t f = g.flags[r0 + k];
      const bool rs = (f & F_RSENSE) != 0, rm = (f & F_RSAME) != 0;
      const uint32_t s = twin_dir(rs, rm) ? 1u : 0u;     // direction of v that u's edge hits
      if (!((und >> s) & 1u)) continue;
      const uint32_t su = fstat[u];
      const uint32_t du = rs ? 1u : 0u;                  // direction of u that edge u->v is in
      if ((su >> (2 + du)) & 1u) {
        if ((su >> du) & 1u) out |= 1u << s;
      } else {
        pend |= 1u << s;
      }
    }
    for (uint32_t s = 0; s < 2; s++) {
      if (!((und >> s) & 1u)) continue;
      if ((out >> s) & 1u) st |= (uint8_t) (4u << s);                       // decided, not fired
      else if (!((pend >> s) & 1u)) st |= (uint8_t) ((4u << s) | (1u << s));  // decided, fired
    }
    a.fstat[v] = st;
    again = (st & FS_DECIDED_ALL) != FS_DECIDED_ALL;
  }
  warp_append(again, p, work_out, n_out);
}

void launch_fire_round(const FilterArgs &a, const uint32_t *work_in, uint32_t n_in,
                       uint32_t *work_out, uint32_t *n_out, cudaStream_t s) {
  if (n_in == 0) return;
  KernelTimer t_("k_fire_round", s);
  k_fire_round<<<(n_in + 127) / 128, 128, 0, s>>>(a, work_in, n_in, work_out, n_out);
}

// ------------------------------------------------------------------ final states

__device__ __forceinline__ void finalize_row(const FilterArgs &a, uint32_t v, uint32_t r0, uint32_t d,
                                             uint32_t first, uint32_t step, int pv, uint32_t fv,
                                             const int inc[2]) {
  const GraphArgs &g = a.g;
  for (uint32_t k = first; k < d; k += step) {
    const uint32_t w = g.dst[r0 + k];
    const uint32_t s = (g.flags[r0 + k] & F_SENSE) ? 1u : 0u;
    const uint32_t pw_u = a.poly_cur[w];
    const int pw = pw_u == NO_TIME ? -1 : (int) pw_u;
    const int tp = pv > pw ? pv : pw;
    int ti = inc[s];
    if (((fv >> s) & 1u) && (int) v > ti) ti = (int) v;
    if (tp < 0 && ti < 0) continue;
    g.estate[r0 + k] = ti >= tp ? GIS_INCONSISTENT : GIS_POLYMORPHIC;
  }
}

__global__ void __launch_bounds__(256) k_finalize(FilterArgs a) {
  const GraphArgs &g = a.g;
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t r0 = 0, d = 0, fv = 0, v = 0;
  int pv = -1;
  if (p < g.V) {
    v = vertex_at(g, p);
    r0 = g.rs[p];
    d = g.re[p] - r0;
    const uint32_t pt = a.poly_cur[v];
    if (pt != NO_TIME) {
      pv = (int) pt;
      g.vstate[v] = GIS_POLYMORPHIC;
    }
    fv = a.fstat[v] & 3u;
  }
  const bool big = d > BIG_ROW;
  if (!big && d > 0) {
    int inc[2] = {-1, -1};          // latest neighbour that fired into (v, s)
    for (uint32_t k = 0; k < d; k++) {
      const uint32_t f = g.flags[r0 + k];
      const bool rs = (f & F_RSENSE) != 0, rm = (f & F_RSAME) != 0;
      const uint32_t u = g.dst[r0 + k];
      if ((a.fstat[u] >> (rs ? 1 : 0)) & 1u) {
        const uint32_t s = twin_dir(rs, rm) ? 1u : 0u;
        if ((int) u > inc[s]) inc[s] = (int) u;
      }
    }
    finalize_row(a, v, r0, d, 0, 1, pv, fv, inc);
  }
  unsigned todo = __ballot_sync(0xffffffffu, big);
  while (todo) {
    const int l = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t vv = __shfl_sync(0xffffffffu, v, l), rr = __shfl_sync(0xffffffffu, r0, l),
                   dd = __shfl_sync(0xffffffffu, d, l), ff = __shfl_sync(0xffffffffu, fv, l);
    const int pp = __shfl_sync(0xffffffffu, pv, l);
    int inc[2] = {-1, -1};
    for (uint32_t k = lane_id(); k < dd; k += 32) {
      const uint32_t f = g.flags[rr + k];
      const bool rs = (f & F_RSENSE) != 0, rm = (f & F_RSAME) != 0;
      const uint32_t u = g.dst[rr + k];
      if ((a.fstat[u] >> (rs ? 1 : 0)) & 1u) {
        const uint32_t s = twin_dir(rs, rm) ? 1u : 0u;
        if ((int) u > inc[s]) inc[s] = (int) u;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      inc[0] = max(inc[0], __shfl_xor_sync(0xffffffffu, inc[0], o));
      inc[1] = max(inc[1], __shfl_xor_sync(0xffffffffu, inc[1], o));
    }
    finalize_row(a, vv, rr, dd, lane_id(), 32, pp, ff, inc);
  }
}

void launch_filter_finalize(const FilterArgs &a, cudaStream_t s) {
  if (a.g.V == 0) return;
  KernelTimer t_("k_finalize", s);
  k_finalize<<<(a.g.V + 255) / 256, 256, 0, s>>>(a);
}

}  // namespace gtsb
