// gtsb_filter.cu -- mark_repeats and the polymorphic / inconsistent filter on the
// device-resident graph.
//
// mark_repeats (reference gt_scaffolder_algorithms.c:160-166 with mark_vertex /
// mark_edge, :61-87) has the closed form
//     vstate[v]      = REPEAT  iff pred(v)
//     estate[v -> w] = REPEAT  iff pred(v) || pred(w)          (else untouched)
//
// gt_scaffolder_graph_filter (:261-343) is a sequential sweep over vertices in
// index order whose marks feed later iterations.  It is evaluated through the
// equivalent closed form of SURVEY.md section 8(a) (validated against the
// compiled reference), with "time" = index (id) of the vertex being processed:
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
//
// Mapping to the machine.  Vertices are named by POSITION (gtsb_kernels.h), so
// every per-vertex array streams.  The passes over edges are flat: a warp owns
// a WINDOW of 32 consecutive slots plus the tail of the last row that starts in
// it (rows of at most BIG_ROW slots; longer rows take block/warp-per-row
// kernels), lanes are slots, row boundaries come from one ballot over the srcp
// column, partners of a pair are fetched with shuffles and row-wide facts are
// ballots.  No shared memory, no block barriers; per-neighbour facts are packed
// so that each pass makes ONE gather per slot:
//   vinfo[p] = {copy_num, seq_len | marked-on-entry << 31}          (pairs pass)
//   vres[p]  = polyTime (27 bits) | F[p,antisense] | F[p,sense] | repeat-pred
//                                                                   (final pass)
#include "gtsb_common.cuh"
#include "gtsb_kernels.h"

namespace gtsb {

constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t NONE = 0xffffffffu;
constexpr uint32_t VI_MARKED = 1u << 31;
constexpr uint32_t VR_TIME_MASK = (1u << 27) - 1u;   // all ones = never
constexpr uint32_t VR_F0 = 1u << 27, VR_F1 = 1u << 28, VR_REP = 1u << 29;
// fstat bits: 0/1 = F[v, antisense/sense], 2/3 = that direction is decided
constexpr uint8_t FS_DECIDED_ALL = 0x0C;
constexpr int WARPS = 8;                             // warps per block of the flat passes

__device__ __forceinline__ uint32_t id_at(const GraphArgs &g, uint32_t p) {
  return g.vid != nullptr ? g.vid[p] : p;
}

__device__ __forceinline__ void warp_append2(bool pred, uint2 value, uint2 *list, uint32_t cap,
                                             uint32_t *count, uint32_t *overflow) {
  const unsigned mask = __ballot_sync(FULL, pred);
  if (mask == 0) return;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if ((int) lane_id() == leader) base = atomicAdd(count, (uint32_t) __popc(mask));
  base = __shfl_sync(FULL, base, leader);
  if (pred) {
    const uint32_t at = base + __popc(mask & ((1u << lane_id()) - 1u));
    if (at < cap) list[at] = value; else atomicOr(overflow, 1u);
  }
}

// ------------------------------------------------------------------ windows

// The rows a warp owns: those that START in slots [32w, 32w+32).  Lane l holds
// slot 32w+l ("lo") and, if the last of those rows runs past the window, slot
// 32(w+1)+l of that row ("hi").  Virtual index of lo = l, of hi = 32+l.
struct Window {
  uint32_t s_lo, s_hi;
  uint32_t row, row_last;    // position of the lo slot's row / of the last row (the hi slots' row)
  bool own_lo, own_hi, head, last;
  uint32_t heads;            // lanes whose lo slot starts a row
  uint32_t vb, ve;           // own lo lane: its row spans virtual indices [vb, ve)
  uint32_t vb_last, nhi;
};

__device__ __forceinline__ Window open_window(const GraphArgs &g, uint32_t w) {
  Window W;
  const uint32_t lane = lane_id();
  W.s_lo = w * 32u + lane;
  W.s_hi = W.s_lo + 32u;
  const bool valid_lo = W.s_lo < g.E;
  const bool valid_hi = (uint64_t) W.s_lo + 32u < g.E;
  const uint32_t sp = valid_lo ? g.srcp[W.s_lo] : NONE;
  const uint32_t sph = valid_hi ? g.srcp[W.s_hi] : NONE;
  uint32_t prev = __shfl_up_sync(FULL, sp, 1);
  if (lane == 0) prev = w > 0 ? g.srcp[W.s_lo - 1] : NONE;
  W.head = valid_lo && sp != prev;
  W.heads = __ballot_sync(FULL, W.head);
  const uint32_t upto = W.heads & (FULL >> (31u - lane));
  const uint32_t above = lane == 31u ? 0u : (W.heads & (FULL << (lane + 1u)));
  W.own_lo = valid_lo && upto != 0u && !(sp & S_BIG);
  W.row = sp & S_POS;
  W.vb = upto ? 31u - (uint32_t) __clz(upto) : 0u;
  W.last = above == 0u;
  const uint32_t sp31 = __shfl_sync(FULL, sp, 31);
  W.own_hi = valid_hi && W.heads != 0u && sph == sp31 && !(sp31 & S_BIG);
  W.nhi = (uint32_t) __popc(__ballot_sync(FULL, W.own_hi));
  W.row_last = sp31 & S_POS;
  W.vb_last = W.heads ? 31u - (uint32_t) __clz(W.heads) : 0u;
  W.ve = W.last ? 32u + W.nhi : (uint32_t) __ffs(above) - 1u;
  return W;
}

// OR of a per-slot predicate over the whole row; valid in own lo lanes
__device__ __forceinline__ bool row_any(const Window &W, bool lo, bool hi) {
  const uint32_t mlo = __ballot_sync(FULL, lo), mhi = __ballot_sync(FULL, hi);
  const uint32_t lim = W.ve < 32u ? ((1u << W.ve) - 1u) : FULL;
  return (mlo & lim & (FULL << W.vb)) != 0u || (W.last && mhi != 0u);
}

static uint32_t host_flat_grid(uint32_t E) {
  const uint64_t nwin = ((uint64_t) E + 31u) / 32u;
  const uint64_t blocks = (nwin + WARPS - 1) / WARPS;
  return (uint32_t) (blocks < 1 ? 1 : blocks);
}

// ------------------------------------------------------------------ per-vertex facts

// One pass by position: the repeat predicate of gt_scaffolder_graph_mark_repeats
// (algorithms.c:163-164, float compares) with its vertex marks, and/or the
// packed facts the pairs pass gathers per neighbour.
__global__ void __launch_bounds__(256) k4_vertex_facts(FilterArgs a, int do_repeats, int write_vinfo,
                                                        int fresh, float copy_num_cutoff,
                                                        float astat_cutoff, int use_copy_num) {
  const GraphArgs &g = a.g;
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.V) return;
  const uint32_t v = id_at(g, p);
  const VAttr at = g.vattr[v];
  bool rep = false;
  if (do_repeats) {
    rep = g.astat[v] <= astat_cutoff || (use_copy_num && at.copy_num < copy_num_cutoff);
    a.rep_pred[p] = rep ? 1 : 0;
    if (rep) g.vstate[v] = GIS_REPEAT;
  }
  if (write_vinfo) {
    if (at.seq_len & VI_MARKED) atomicOr(&g.counters[CNT_ERROR], 4u);   // seq_len must fit 31 bits
    const bool marked = rep || (!fresh && vertex_state_marked(g.vstate[v]));
    a.vinfo[p] = make_uint2(__float_as_uint(at.copy_num), at.seq_len | (marked ? VI_MARKED : 0u));
  }
}

void launch_vertex_facts(const FilterArgs &a, int do_repeats, float copy_num_cutoff, float astat_cutoff,
                         int use_copy_num, cudaStream_t s) {
  if (a.g.V == 0) return;
  KernelTimer t_("k4_vertex_facts", s);
  const int write_vinfo = a.vinfo != nullptr ? 1 : 0;
  k4_vertex_facts<<<(a.g.V + 255) / 256, 256, 0, s>>>(a, do_repeats, write_vinfo,
                                                      a.fused_repeats, copy_num_cutoff, astat_cutoff,
                                                      use_copy_num);
}

// edge marks of mark_repeats: estate[v->w] = REPEAT iff pred(v) || pred(w)
__global__ void __launch_bounds__(256) k4_repeat_edges(GraphArgs g, const uint8_t *__restrict__ rep) {
  const uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= g.E) return;
  if (rep[g.srcp[s] & S_POS] || rep[g.dst[s]]) g.estate[s] = GIS_REPEAT;
}

void launch_repeat_edges(const GraphArgs &g, const uint8_t *rep_pred, cudaStream_t s) {
  if (g.E == 0) return;
  KernelTimer t_("k4_repeat_edges", s);
  k4_repeat_edges<<<(uint32_t) (((uint64_t) g.E + 255) / 256), 256, 0, s>>>(g, rep_pred);
}

// srcp column, F_LT flags and the big-row list of a plain CSR (identity positions)
__global__ void __launch_bounds__(256) k4_fill_srcp(GraphArgs g, uint32_t *__restrict__ srcp,
                                                     uint32_t *__restrict__ big_rows) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t r0 = 0, d = 0;
  if (p < g.V) {
    r0 = g.row_ptr[p];
    d = g.row_ptr[p + 1] - r0;
  }
  const bool big = d > BIG_ROW;
  if (!big)
    for (uint32_t k = 0; k < d; k++) {
      srcp[r0 + k] = p;
      g.flags[r0 + k] = (uint8_t) ((g.flags[r0 + k] & 0x0Fu) | (g.dst[r0 + k] < p ? F_LT : 0u));
    }
  if (big_rows != nullptr) {
    if (big) atomicMax(&g.counters[CNT_MAX_DEG], d);
    warp_append(big, p, big_rows, &g.counters[CNT_BIG_ROWS]);
  }
  unsigned todo = __ballot_sync(FULL, big);
  while (todo) {
    const int l = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t pp = __shfl_sync(FULL, p, l), rr = __shfl_sync(FULL, r0, l), dd = __shfl_sync(FULL, d, l);
    for (uint32_t k = lane_id(); k < dd; k += 32) {
      srcp[rr + k] = pp | S_BIG;
      g.flags[rr + k] = (uint8_t) ((g.flags[rr + k] & 0x0Fu) | (g.dst[rr + k] < pp ? F_LT : 0u));
    }
  }
}

void launch_fill_srcp(const GraphArgs &g, uint32_t *srcp, uint32_t *big_rows, cudaStream_t s) {
  if (g.V == 0) return;
  KernelTimer t_("k4_fill_srcp", s);
  k4_fill_srcp<<<(g.V + 255) / 256, 256, 0, s>>>(g, srcp, big_rows);
}

// ------------------------------------------------------------------ phase 1 + static overlap

struct SlotFacts {
  int32_t dist;
  float std_dev, cn;
  uint32_t len, fl;          // fl: bit0 sense, bit1 edge unmarked on entry, bit2 dst marked on entry
};
constexpr uint32_t SF_SENSE = 1u, SF_OK = 2u, SF_WM = 4u;

__device__ __forceinline__ SlotFacts slot_facts(const FilterArgs &a, uint32_t s, uint32_t *dst_out) {
  const GraphArgs &g = a.g;
  SlotFacts f;
  const uint32_t w = g.dst[s];
  const uint2 vi = a.vinfo[w];
  const uint32_t fl = g.flags[s];
  const bool wm = (vi.y & VI_MARKED) != 0;
  const bool ok = a.fused_repeats ? !wm : !edge_state_marked(g.estate[s]);
  f.dist = g.dist[s];
  f.std_dev = g.std_dev[s];
  f.cn = __uint_as_float(vi.x);
  f.len = vi.y & ~VI_MARKED;
  f.fl = (fl & F_SENSE) | (ok ? SF_OK : 0u) | (wm ? SF_WM : 0u);
  *dst_out = w;
  return f;
}

__device__ __forceinline__ SlotFacts shfl_facts(const SlotFacts &f, uint32_t src_lane) {
  SlotFacts r;
  r.dist = __shfl_sync(FULL, f.dist, src_lane);
  r.std_dev = __shfl_sync(FULL, f.std_dev, src_lane);
  r.cn = __shfl_sync(FULL, f.cn, src_lane);
  r.len = __shfl_sync(FULL, f.len, src_lane);
  r.fl = __shfl_sync(FULL, f.fl, src_lane);
  return r;
}

// one same-row pair, e1 = earlier adjacency slot (algorithms.c:283-295, 301-320)
__device__ __forceinline__ void pair_eval(const FilterArgs &a, const SlotFacts &e1, const SlotFacts &e2,
                                          bool &prop1, bool &prop2, bool &fire) {
  if ((e1.fl ^ e2.fl) & SF_SENSE) return;
  // check_mark_polymorphic, algorithms.c:232-238
  if (ambiguous_order(e1.dist, e1.std_dev, e2.dist, e2.std_dev, a.ambig) &&
      __fadd_rn(e1.cn, e2.cn) < a.cncutoff) {
    if (e1.cn < e2.cn) prop1 = true; else prop2 = true;
  }
  if ((e1.fl & e2.fl & SF_OK) && a.ocutoff >= 0)
    fire |= interval_overlap(e1.dist, e1.len, e2.dist, e2.len) > a.ocutoff;
}

// Proposals of check_mark_polymorphic and, in the same sweep over the pairs,
// G0[v,s] = "some same-direction pair of edges that are unmarked on entry
// overlaps by more than ocutoff" (algorithms.c:301-324 before any mark of this
// filter run is taken into account; k4_fire_init repairs the rows next to
// polymorphic vertices).  gbits must be zero on entry.
__global__ void __launch_bounds__(32 * WARPS) k4_pairs(FilterArgs a) {
  const GraphArgs &g = a.g;
  const uint32_t lane = lane_id();
  const uint32_t nwin = (uint32_t) (((uint64_t) g.E + 31u) / 32u);
  for (uint32_t w = blockIdx.x * WARPS + (threadIdx.x >> 5); w < nwin; w += gridDim.x * WARPS) {
    const Window W = open_window(g, w);
    // rows that cannot propose or fire: marked on entry (algorithms.c:279)
    bool act_lo = false;
    if (W.own_lo) act_lo = !(a.vinfo[W.row].y & VI_MARKED) && (W.ve - W.vb) >= 2u;
    if (!__any_sync(FULL, act_lo)) continue;
    const bool act_last = __shfl_sync(FULL, (int) act_lo, W.vb_last) != 0;   // every lane shuffles
    const bool act_hi = W.own_hi && act_last;
    SlotFacts lo = {}, hi = {};
    uint32_t dst_lo = 0, dst_hi = 0;
    if (act_lo) lo = slot_facts(a, W.s_lo, &dst_lo);
    if (act_hi) hi = slot_facts(a, W.s_hi, &dst_hi);
    bool prop_lo = false, prop_hi = false, fire_lo = false, fire_hi = false;
    // lo requesters: partner at virtual index lane + o (lo of lane+o, or hi of lane+o-32)
    const uint32_t maxlen = __reduce_max_sync(FULL, act_lo ? W.ve - lane : 0u);
    for (uint32_t o = 1; o < maxlen; o++) {
      const bool from_lo = lane >= o;
      SlotFacts src;
      src.dist = from_lo ? lo.dist : hi.dist;
      src.std_dev = from_lo ? lo.std_dev : hi.std_dev;
      src.cn = from_lo ? lo.cn : hi.cn;
      src.len = from_lo ? lo.len : hi.len;
      src.fl = from_lo ? lo.fl : hi.fl;
      const SlotFacts p = shfl_facts(src, (lane + o) & 31u);
      bool pm = false, po = false;
      if (act_lo && lane + o < W.ve) pair_eval(a, lo, p, pm, po, fire_lo);
      prop_lo |= pm;
      const bool back = __shfl_sync(FULL, (int) po, (lane - o) & 31u) != 0;
      if (from_lo) prop_lo |= back; else prop_hi |= back;
    }
    // hi requesters pair with hi partners only
    for (uint32_t o = 1; o < W.nhi; o++) {
      const SlotFacts p = shfl_facts(hi, (lane + o) & 31u);
      bool pm = false, po = false;
      if (act_hi && lane + o < W.nhi) pair_eval(a, hi, p, pm, po, fire_hi);
      prop_hi |= pm;
      const bool back = __shfl_sync(FULL, (int) po, (lane - o) & 31u) != 0;
      if (lane >= o) prop_hi |= back;
    }
    // proposals (target must be unmarked, algorithms.c:242)
    warp_append2(act_lo && prop_lo && !(lo.fl & SF_WM), make_uint2(W.row, dst_lo), a.proposals,
                 a.proposals_cap, &g.counters[CNT_PROPOSALS], &g.counters[CNT_OVERFLOW]);
    warp_append2(act_hi && prop_hi && !(hi.fl & SF_WM), make_uint2(W.row_last, dst_hi), a.proposals,
                 a.proposals_cap, &g.counters[CNT_PROPOSALS], &g.counters[CNT_OVERFLOW]);
    const bool g1 = row_any(W, fire_lo && (lo.fl & SF_SENSE), fire_hi && (hi.fl & SF_SENSE));
    const bool g0 = row_any(W, fire_lo && !(lo.fl & SF_SENSE), fire_hi && !(hi.fl & SF_SENSE));
    if (W.head && W.own_lo && (g0 || g1)) a.gbits[W.row] = (uint8_t) ((g0 ? 1u : 0u) | (g1 ? 2u : 0u));
  }
}

// block per big row; per-block scratch: copy_num[max_deg] f32, mark[max_deg] u8
__global__ void __launch_bounds__(512) k4_pairs_big(FilterArgs a) {
  const GraphArgs &g = a.g;
  float *cn = reinterpret_cast<float *>(a.big_scratch + (size_t) blockIdx.x * g.max_deg * BIG_SCRATCH_STRIDE);
  uint8_t *mark = reinterpret_cast<uint8_t *>(cn + 2 * (size_t) g.max_deg);
  for (uint32_t li = blockIdx.x; li < g.n_big_rows; li += gridDim.x) {
    const uint32_t p = g.big_rows[li];
    if (a.vinfo[p].y & VI_MARKED) continue;                        // block-uniform
    const uint32_t r0 = g.row_ptr[p], d = g.row_ptr[p + 1] - r0;
    for (uint32_t k = threadIdx.x; k < d; k += blockDim.x) {
      cn[k] = __uint_as_float(a.vinfo[g.dst[r0 + k]].x);
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
    for (uint32_t k0 = 0; k0 < d; k0 += blockDim.x) {
      const uint32_t k = k0 + threadIdx.x;
      uint32_t t = 0;
      bool emit = false;
      if (k < d && mark[k]) {
        t = g.dst[r0 + k];
        emit = !(a.vinfo[t].y & VI_MARKED);
      }
      warp_append2(emit, make_uint2(p, t), a.proposals, a.proposals_cap, &g.counters[CNT_PROPOSALS],
                   &g.counters[CNT_OVERFLOW]);
    }
    __syncthreads();
  }
}

void launch_pairs(const FilterArgs &a, cudaStream_t s) {
  if (a.g.E == 0) return;
  {
    KernelTimer t_("k4_pairs", s);
    k4_pairs<<<host_flat_grid(a.g.E), 32 * WARPS, 0, s>>>(a);
  }
  if (a.g.n_big_rows) {
    KernelTimer t_("k4_pairs_big", s);
    k4_pairs_big<<<a.big_blocks, 512, 0, s>>>(a);
  }
}

// ------------------------------------------------------------------ polyTime fixpoint

// One Jacobi sweep of  polyTime(p) = min{ t : A[t], (t,p) proposed },
// A[t] = !(polyTime(t) < t).  Every dependency points to a smaller id, so the
// sweeps converge to the unique solution in (longest chain) iterations.
__global__ void __launch_bounds__(256) k_poly_reset(const uint2 *__restrict__ proposals, uint32_t n,
                                                     uint32_t *__restrict__ poly_new) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) poly_new[proposals[i].y] = NO_TIME;
}
__global__ void __launch_bounds__(256) k_poly_propose(GraphArgs g, const uint2 *__restrict__ proposals,
                                                       uint32_t n, const uint32_t *__restrict__ poly_cur,
                                                       uint32_t *__restrict__ poly_new) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2 pr = proposals[i];
  const uint32_t t = id_at(g, pr.x);
  if (!(poly_cur[pr.x] < t)) atomicMin(&poly_new[pr.y], t);
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
  k_poly_propose<<<blocks, 256, 0, s>>>(a.g, a.proposals, n, a.poly_cur, a.poly_new);
  k_poly_commit<<<blocks, 256, 0, s>>>(a.proposals, n, a.poly_cur, a.poly_new, a.g.counters);
}

// rows whose static overlap answer may be stale: every neighbour of a vertex
// that became polymorphic, and that vertex itself
__global__ void __launch_bounds__(256) k4_dirty(FilterArgs a, uint32_t n) {
  const GraphArgs &g = a.g;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2 pr = a.proposals[i];
  if (a.poly_cur[pr.y] != id_at(g, pr.x)) return;             // not the winning proposer
  a.dirty[pr.y] = 1;
  for (uint32_t s = g.row_ptr[pr.y]; s < g.row_ptr[pr.y + 1]; s++) a.dirty[g.dst[s]] = 1;
}

void launch_dirty(const FilterArgs &a, uint32_t n_proposals, cudaStream_t s) {
  if (n_proposals == 0) return;
  KernelTimer t_("k4_dirty", s);
  k4_dirty<<<(n_proposals + 255) / 256, 256, 0, s>>>(a, n_proposals);
}

// ------------------------------------------------------------------ phase 2: fire candidates

__device__ __forceinline__ bool slot_unmarked(const FilterArgs &a, uint32_t s, uint32_t w_info_y,
                                              uint32_t w, uint32_t v_id) {
  const bool ok0 = a.fused_repeats ? !(w_info_y & VI_MARKED) : !edge_state_marked(a.g.estate[s]);
  return ok0 && !(a.poly_cur[w] <= v_id);
}

// G[v,s]: max overlap over same-direction pairs that are unmarked when v is
// reached, not yet counting the fires of smaller neighbours (algorithms.c:301-320).
// Thread per row; only rows next to a polymorphic vertex recompute.
__global__ void __launch_bounds__(128) k4_fire_init(FilterArgs a) {
  const GraphArgs &g = a.g;
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.V) return;
  const uint32_t r0 = g.row_ptr[p], d = g.row_ptr[p + 1] - r0;
  if (d > BIG_ROW) return;
  const uint32_t v_id = id_at(g, p);
  const bool active = !(a.vinfo[p].y & VI_MARKED) && !(a.poly_cur[p] < v_id);
  uint8_t gb = 0;
  if (active && a.ocutoff < 0) {
    gb = 3;            // 0 > ocutoff: both directions fire whatever the pairs (:301-324)
  } else if (active && !a.dirty[p]) {
    gb = a.gbits[p];   // the pairs pass's static answer stands: no polymorphic vertex nearby
  } else if (active && d >= 2) {
    int32_t dist[BIG_ROW];
    uint32_t len[BIG_ROW];
    uint32_t sense_mask = 0, ok_mask = 0;
    for (uint32_t k = 0; k < d; k++) {
      const uint32_t w = g.dst[r0 + k];
      const uint2 vi = a.vinfo[w];
      dist[k] = g.dist[r0 + k];
      len[k] = vi.y & ~VI_MARKED;
      if (g.flags[r0 + k] & F_SENSE) sense_mask |= 1u << k;
      if (slot_unmarked(a, r0 + k, vi.y, w, v_id)) ok_mask |= 1u << k;
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
  a.gbits[p] = gb;
  a.fstat[p] = a.ocutoff < 0 ? (uint8_t) (FS_DECIDED_ALL | gb)       // no dependence on neighbours
                             : (uint8_t) (FS_DECIDED_ALL & ~(gb << 2));
}

// block per big row; undecided big rows go straight to the fire worklist
__global__ void __launch_bounds__(512) k4_fire_init_big(FilterArgs a, uint32_t *__restrict__ work_out,
                                                         uint32_t *__restrict__ n_out) {
  const GraphArgs &g = a.g;
  __shared__ long long s_mx[2];
  uint32_t *len = reinterpret_cast<uint32_t *>(a.big_scratch + (size_t) blockIdx.x * g.max_deg * BIG_SCRATCH_STRIDE);
  uint8_t *ok = reinterpret_cast<uint8_t *>(len + 2 * (size_t) g.max_deg);
  for (uint32_t li = blockIdx.x; li < g.n_big_rows; li += gridDim.x) {
    const uint32_t p = g.big_rows[li];
    const uint32_t v_id = id_at(g, p);
    const uint32_t r0 = g.row_ptr[p], d = g.row_ptr[p + 1] - r0;
    const bool active = !(a.vinfo[p].y & VI_MARKED) && !(a.poly_cur[p] < v_id);
    uint8_t gb = 0;
    if (active && a.ocutoff < 0) {
      gb = 3;
    } else if (active) {
      if (threadIdx.x < 2) s_mx[threadIdx.x] = 0;
      for (uint32_t k = threadIdx.x; k < d; k += blockDim.x) {
        const uint32_t w = g.dst[r0 + k];
        const uint2 vi = a.vinfo[w];
        len[k] = vi.y & ~VI_MARKED;
        ok[k] = slot_unmarked(a, r0 + k, vi.y, w, v_id) ? 1 : 0;
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
      a.gbits[p] = gb;
      if (a.ocutoff < 0) {
        a.fstat[p] = FS_DECIDED_ALL | gb;
      } else {
        a.fstat[p] = (uint8_t) (FS_DECIDED_ALL & ~(gb << 2));
        if (gb) work_out[atomicAdd(n_out, 1u)] = p;
      }
    }
    __syncthreads();
  }
}

void launch_fire_init(const FilterArgs &a, cudaStream_t s) {
  if (a.g.V == 0) return;
  {
    KernelTimer t_("k4_fire_init", s);
    k4_fire_init<<<(a.g.V + 127) / 128, 128, 0, s>>>(a);
  }
  if (a.g.n_big_rows) {
    KernelTimer t_("k4_fire_init_big", s);
    k4_fire_init_big<<<a.big_blocks, 512, 0, s>>>(a, a.work_b, &a.g.counters[CNT_WORK_B]);
  }
}

// ------------------------------------------------------------------ fire fixpoint

// F[v,s] = G[v,s] && no neighbour u < v with F[u, sense(u->v)] and
// twin_dir(u->v) == s.  A direction is decided once every smaller neighbour
// that could fire into it is decided; each round decides at least the smallest
// undecided vertex, and on random orders the depth is logarithmic.

// what slot s says about the undecided directions `und` of its row:
// bit sdir = a smaller neighbour fired into it, bit 2+sdir = such a neighbour is pending
__device__ __forceinline__ uint32_t fire_probe(const GraphArgs &g, const volatile uint8_t *fstat,
                                               uint32_t s, uint32_t und) {
  const uint32_t f = g.flags[s];
  if (!(f & F_LT)) return 0u;                            // only smaller ids matter
  const bool rs = (f & F_RSENSE) != 0, rm = (f & F_RSAME) != 0;
  const uint32_t sdir = twin_dir(rs, rm) ? 1u : 0u;      // direction of v that u's edge hits
  if (!((und >> sdir) & 1u)) return 0u;
  const uint32_t su = fstat[g.dst[s]];
  const uint32_t du = rs ? 1u : 0u;                      // direction of u that edge u->v is in
  if ((su >> (2 + du)) & 1u) return ((su >> du) & 1u) ? (1u << sdir) : 0u;
  return 4u << sdir;
}

__device__ __forceinline__ uint8_t fire_decide(uint8_t st, uint32_t und, uint32_t res) {
  for (uint32_t s = 0; s < 2; s++) {
    if (!((und >> s) & 1u)) continue;
    if ((res >> s) & 1u) st |= (uint8_t) (4u << s);                          // decided, not fired
    else if (!((res >> (2 + s)) & 1u)) st |= (uint8_t) ((4u << s) | (1u << s));  // decided, fired
  }
  return st;
}

// first round over all rows of at most BIG_ROW slots (flat); rows still
// undecided afterwards are appended to the worklist
__global__ void __launch_bounds__(32 * WARPS) k4_fire_dense(FilterArgs a, uint32_t *__restrict__ work_out,
                                                             uint32_t *__restrict__ n_out) {
  const GraphArgs &g = a.g;
  const volatile uint8_t *fstat = a.fstat;
  const uint32_t nwin = (uint32_t) (((uint64_t) g.E + 31u) / 32u);
  for (uint32_t w = blockIdx.x * WARPS + (threadIdx.x >> 5); w < nwin; w += gridDim.x * WARPS) {
    const Window W = open_window(g, w);
    uint32_t st = FS_DECIDED_ALL;
    if (W.own_lo) st = fstat[W.row];
    const uint32_t und_lo = (~st >> 2) & 3u;
    if (!__any_sync(FULL, und_lo != 0u)) continue;
    const uint32_t und_last = __shfl_sync(FULL, und_lo, W.vb_last);          // every lane shuffles
    const uint32_t und_hi = W.own_hi ? und_last : 0u;
    const uint32_t res_lo = und_lo ? fire_probe(g, fstat, W.s_lo, und_lo) : 0u;
    const uint32_t res_hi = und_hi ? fire_probe(g, fstat, W.s_hi, und_hi) : 0u;
    uint32_t res = 0;
#pragma unroll
    for (uint32_t b = 0; b < 4; b++)
      if (row_any(W, (res_lo >> b) & 1u, (res_hi >> b) & 1u)) res |= 1u << b;
    bool again = false;
    if (W.head && W.own_lo && und_lo) {
      const uint8_t nst = fire_decide((uint8_t) st, und_lo, res);
      a.fstat[W.row] = nst;
      again = (nst & FS_DECIDED_ALL) != FS_DECIDED_ALL;
    }
    warp_append(again, W.row, work_out, n_out);
  }
}

void launch_fire_dense(const FilterArgs &a, uint32_t *work_out, uint32_t *n_out, cudaStream_t s) {
  if (a.g.E == 0) return;
  KernelTimer t_("k4_fire_dense", s);
  k4_fire_dense<<<host_flat_grid(a.g.E), 32 * WARPS, 0, s>>>(a, work_out, n_out);
}

// later rounds: thread per listed row (any degree)
__global__ void __launch_bounds__(128) k_fire_round(FilterArgs a, const uint32_t *__restrict__ work_in,
                                                     uint32_t n_in, uint32_t *__restrict__ work_out,
                                                     uint32_t *__restrict__ n_out) {
  const GraphArgs &g = a.g;
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  bool again = false;
  uint32_t p = 0;
  if (idx < n_in) {
    p = work_in[idx];
    const volatile uint8_t *fstat = a.fstat;
    uint8_t st = fstat[p];
    const uint32_t und = (~(uint32_t) st >> 2) & 3u;
    uint32_t res = 0;
    for (uint32_t s = g.row_ptr[p]; s < g.row_ptr[p + 1]; s++) res |= fire_probe(g, fstat, s, und);
    st = fire_decide(st, und, res);
    a.fstat[p] = st;
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

__global__ void __launch_bounds__(256) k4_vres(FilterArgs a) {
  const GraphArgs &g = a.g;
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.V) return;
  const uint32_t t = a.poly_cur[p];
  const uint32_t f = a.fstat[p];
  a.vres[p] = (t == NO_TIME ? VR_TIME_MASK : t) | ((f & 1u) ? VR_F0 : 0u) | ((f & 2u) ? VR_F1 : 0u) |
              ((a.fused_repeats && a.rep_pred[p]) ? VR_REP : 0u);
  if (t != NO_TIME) g.vstate[id_at(g, p)] = GIS_POLYMORPHIC;
}

// last writer wins (INCONSISTENT on ties: phase 3 follows phase 1)
__device__ __forceinline__ void final_state(const FilterArgs &a, uint32_t s, uint32_t f, uint32_t ru,
                                            uint32_t own, uint32_t row, int inc0, int inc1) {
  const GraphArgs &g = a.g;
  const uint32_t sd = (f & F_SENSE) ? 1u : 0u;
  const uint32_t tu = ru & VR_TIME_MASK, tv = own & VR_TIME_MASK;
  const int pw = tu == VR_TIME_MASK ? -1 : (int) tu, pv = tv == VR_TIME_MASK ? -1 : (int) tv;
  const int tp = pv > pw ? pv : pw;
  int ti = sd ? inc1 : inc0;
  if (own & (sd ? VR_F1 : VR_F0)) {
    const int me = (int) id_at(g, row);
    if (me > ti) ti = me;
  }
  if (tp < 0 && ti < 0) {
    if (a.fused_repeats) g.estate[s] = ((own | ru) & VR_REP) ? GIS_REPEAT : GIS_UNVISITED;
  } else {
    g.estate[s] = ti >= tp ? GIS_INCONSISTENT : GIS_POLYMORPHIC;
  }
}

__global__ void __launch_bounds__(32 * WARPS) k4_finalize(FilterArgs a) {
  const GraphArgs &g = a.g;
  const uint32_t nwin = (uint32_t) (((uint64_t) g.E + 31u) / 32u);
  for (uint32_t w = blockIdx.x * WARPS + (threadIdx.x >> 5); w < nwin; w += gridDim.x * WARPS) {
    const Window W = open_window(g, w);
    if (!__any_sync(FULL, W.own_lo)) continue;
    uint32_t own_lo = 0, ru_lo = 0, f_lo = 0, u_lo = 0, ru_hi = 0, f_hi = 0, u_hi = 0;
    if (W.own_lo) {
      own_lo = a.vres[W.row];
      u_lo = g.dst[W.s_lo];
      f_lo = g.flags[W.s_lo];
      ru_lo = a.vres[u_lo];
    }
    const uint32_t own_hi = __shfl_sync(FULL, own_lo, W.vb_last);
    if (W.own_hi) {
      u_hi = g.dst[W.s_hi];
      f_hi = g.flags[W.s_hi];
      ru_hi = a.vres[u_hi];
    }
    // pass 1: latest neighbour that fired into (row, direction)
    int inc_lo0 = -1, inc_lo1 = -1, inc_hi0 = -1, inc_hi1 = -1;
    const bool fin_lo = W.own_lo && (ru_lo & ((f_lo & F_RSENSE) ? VR_F1 : VR_F0));
    const bool fin_hi = W.own_hi && (ru_hi & ((f_hi & F_RSENSE) ? VR_F1 : VR_F0));
    const int id_lo = fin_lo ? (int) id_at(g, u_lo) : -1, id_hi = fin_hi ? (int) id_at(g, u_hi) : -1;
    const uint32_t dir_lo = twin_dir((f_lo & F_RSENSE) != 0, (f_lo & F_RSAME) != 0) ? 1u : 0u;
    const uint32_t dir_hi = twin_dir((f_hi & F_RSENSE) != 0, (f_hi & F_RSAME) != 0) ? 1u : 0u;
    uint32_t m = __ballot_sync(FULL, fin_lo);
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t rvb = __shfl_sync(FULL, W.vb, l), dir = __shfl_sync(FULL, dir_lo, l);
      const int id = __shfl_sync(FULL, id_lo, l);
      const bool rlast = __shfl_sync(FULL, (int) W.last, l) != 0;
      if (W.own_lo && W.vb == rvb) {
        if (dir) inc_lo1 = max(inc_lo1, id); else inc_lo0 = max(inc_lo0, id);
      }
      if (W.own_hi && rlast) {
        if (dir) inc_hi1 = max(inc_hi1, id); else inc_hi0 = max(inc_hi0, id);
      }
    }
    m = __ballot_sync(FULL, fin_hi);
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t dir = __shfl_sync(FULL, dir_hi, l);
      const int id = __shfl_sync(FULL, id_hi, l);
      if (W.own_lo && W.last) {
        if (dir) inc_lo1 = max(inc_lo1, id); else inc_lo0 = max(inc_lo0, id);
      }
      if (W.own_hi) {
        if (dir) inc_hi1 = max(inc_hi1, id); else inc_hi0 = max(inc_hi0, id);
      }
    }
    // pass 2
    if (W.own_lo) final_state(a, W.s_lo, f_lo, ru_lo, own_lo, W.row, inc_lo0, inc_lo1);
    if (W.own_hi) final_state(a, W.s_hi, f_hi, ru_hi, own_hi, W.row_last, inc_hi0, inc_hi1);
  }
}

// warp per big row
__global__ void __launch_bounds__(256) k4_finalize_big(FilterArgs a) {
  const GraphArgs &g = a.g;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t li = warp; li < g.n_big_rows; li += nwarps) {
    const uint32_t p = g.big_rows[li];
    const uint32_t r0 = g.row_ptr[p], d = g.row_ptr[p + 1] - r0;
    const uint32_t own = a.vres[p];
    int inc0 = -1, inc1 = -1;
    for (uint32_t k = lane_id(); k < d; k += 32) {
      const uint32_t f = g.flags[r0 + k];
      const bool rs = (f & F_RSENSE) != 0, rm = (f & F_RSAME) != 0;
      const uint32_t u = g.dst[r0 + k];
      if (a.vres[u] & (rs ? VR_F1 : VR_F0)) {
        const int id = (int) id_at(g, u);
        if (twin_dir(rs, rm)) inc1 = max(inc1, id); else inc0 = max(inc0, id);
      }
    }
    inc0 = __reduce_max_sync(FULL, inc0);
    inc1 = __reduce_max_sync(FULL, inc1);
    for (uint32_t k = lane_id(); k < d; k += 32) {
      const uint32_t u = g.dst[r0 + k];
      final_state(a, r0 + k, g.flags[r0 + k], a.vres[u], own, p, inc0, inc1);
    }
  }
}

void launch_finalize(const FilterArgs &a, cudaStream_t s) {
  if (a.g.V == 0) return;
  {
    KernelTimer t_("k4_vres", s);
    k4_vres<<<(a.g.V + 255) / 256, 256, 0, s>>>(a);
  }
  if (a.g.E == 0) return;
  {
    KernelTimer t_("k4_finalize", s);
    k4_finalize<<<host_flat_grid(a.g.E), 32 * WARPS, 0, s>>>(a);
  }
  if (a.g.n_big_rows) {
    KernelTimer t_("k4_finalize_big", s);
    uint32_t blocks = (a.g.n_big_rows + 7) / 8;
    if (blocks > (uint32_t) a.g.sm_count * 8) blocks = (uint32_t) a.g.sm_count * 8;
    k4_finalize_big<<<blocks, 256, 0, s>>>(a);
  }
}

}  // namespace gtsb
