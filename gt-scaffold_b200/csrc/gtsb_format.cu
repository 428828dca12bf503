// gtsb_format.cu -- `.dot` text from graph arrays on the device (SURVEY.md §8(f) rank 3;
// gt_scaffolder_graph_print_generic / _print_scaffold, graph.c:269-343).  Per call: line
// lengths (one thread per vertex / edge), exclusive scan, the lines themselves; the caller
// gets the bytes in item order and writes them between "digraph {\n" and "}\n".
#include "gtsb_context.h"
#include "gtsb_format_core.h"
#include "gtsb_scan.cuh"

using namespace gtsb;
using namespace gtsbi;
using namespace gtsbf;

namespace gtsbformat {

__global__ void __launch_bounds__(256) kf_vertex_len(uint64_t first, uint32_t count, const uint8_t *__restrict__ vstate,
                                                     const uint64_t *__restrict__ name_off, int scaffold_only,
                                                     uint32_t *__restrict__ len, uint32_t *bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t v = first + i;
  if (vstate[i] >= NOF_STATES) atomicOr(bad, 1u);
  const uint64_t name_len = name_off[v + 1] - name_off[v];
  const uint32_t n = vertex_line_len(v, vstate[i], name_len, scaffold_only != 0);
  len[i] = n;
  // bad[1]: longest line of the call (a header is as long as its FASTA line)
  const uint64_t bound = name_len + 64;
  atomicMax(bad + 1, bound > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t) bound);
}

__global__ void __launch_bounds__(256) kf_vertex_put(uint64_t first, uint32_t count, const uint8_t *__restrict__ vstate,
                                                     const char *__restrict__ names,
                                                     const uint64_t *__restrict__ name_off, int scaffold_only,
                                                     const uint32_t *__restrict__ off, char *__restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t v = first + i;
  put_vertex_line(out + off[i], v, vstate[i], names + name_off[v], name_off[v + 1] - name_off[v],
                  scaffold_only != 0);
}

__global__ void __launch_bounds__(256) kf_edge_len(uint32_t count, const uint32_t *__restrict__ src,
                                                   const uint32_t *__restrict__ dst, const int32_t *__restrict__ dist,
                                                   const uint8_t *__restrict__ estate, const uint8_t *__restrict__ sense,
                                                   int scaffold_only, uint32_t *__restrict__ len, uint32_t *bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  if (estate[i] >= NOF_STATES) atomicOr(bad, 1u);
  if (i == 0) atomicMax(bad + 1, 105u);             // no edge line is longer
  len[i] = edge_line_len(src[i], dst[i], dist[i], estate[i], sense[i] != 0, scaffold_only != 0);
}

__global__ void __launch_bounds__(256) kf_edge_put(uint32_t count, const uint32_t *__restrict__ src,
                                                   const uint32_t *__restrict__ dst, const int32_t *__restrict__ dist,
                                                   const uint8_t *__restrict__ estate, const uint8_t *__restrict__ sense,
                                                   int scaffold_only, const uint32_t *__restrict__ off,
                                                   char *__restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  put_edge_line(out + off[i], src[i], dst[i], dist[i], estate[i], sense[i] != 0, scaffold_only != 0);
}

// ---- `.scaf` records (gt_scaffolder_graph_write_scaffold, algorithms.c:1000-1042): the text of
// record i starts at off_r[i] + off_e[rec_edge_off[i]] (off_r: root headers and newlines of the
// records before it, off_e: edge pieces before its first edge)

__global__ void __launch_bounds__(256) kf_scaf_rec_len(uint32_t n, const uint32_t *__restrict__ root,
                                                       const uint64_t *__restrict__ name_off, uint64_t names_V,
                                                       uint32_t *__restrict__ len, uint32_t *bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = root[i];
  if (v >= names_V) {
    atomicOr(bad, 2u);
    len[i] = 0;
    return;
  }
  const uint64_t name_len = name_off[v + 1] - name_off[v];
  len[i] = (uint32_t) name_len + 1u;
  atomicMax(bad + 1, name_len + 1 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t) (name_len + 1));
}

__global__ void __launch_bounds__(256) kf_scaf_edge_len(uint32_t m, const uint32_t *__restrict__ end,
                                                        const int64_t *__restrict__ dist,
                                                        const uint32_t *__restrict__ std_bits,
                                                        const uint64_t *__restrict__ name_off, uint64_t names_V,
                                                        uint32_t *__restrict__ len, uint32_t *bad) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const uint32_t v = end[j];
  if (v >= names_V) {
    atomicOr(bad, 2u);
    len[j] = 0;
    return;
  }
  const uint64_t name_len = name_off[v + 1] - name_off[v];
  len[j] = scaf_edge_len(name_len, dist[j], std_bits[j]);
  // bad[2]: longest piece (tab, header, 20-character distance, 47-character %f, two flags, five commas)
  const uint64_t bound = name_len + 80;
  atomicMax(bad + 2, bound > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t) bound);
}

__global__ void __launch_bounds__(256) kf_scaf_rec_put(uint32_t n, const uint32_t *__restrict__ root,
                                                       const uint64_t *__restrict__ rec_edge_off,
                                                       const uint32_t *__restrict__ off_r,
                                                       const uint32_t *__restrict__ off_e,
                                                       const char *__restrict__ names,
                                                       const uint64_t *__restrict__ name_off, char *__restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = root[i];
  const uint32_t name_len = (uint32_t) (name_off[v + 1] - name_off[v]);
  put_str(out + off_r[i] + off_e[rec_edge_off[i]], names + name_off[v], name_len);
  out[off_r[i] + name_len + off_e[rec_edge_off[i + 1]]] = '\n';
}

__global__ void __launch_bounds__(256) kf_scaf_edge_put(uint32_t m, uint32_t n, const uint32_t *__restrict__ root,
                                                        const uint64_t *__restrict__ rec_edge_off,
                                                        const uint32_t *__restrict__ end,
                                                        const int64_t *__restrict__ dist,
                                                        const uint32_t *__restrict__ std_bits,
                                                        const uint8_t *__restrict__ flags,
                                                        const uint32_t *__restrict__ off_r,
                                                        const uint32_t *__restrict__ off_e,
                                                        const char *__restrict__ names,
                                                        const uint64_t *__restrict__ name_off, char *__restrict__ out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  // the record of edge j: the last i with rec_edge_off[i] <= j (records without edges are skipped)
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1u) {
    const uint32_t mid = lo + (hi - lo) / 2u;
    if (rec_edge_off[mid] <= j) lo = mid; else hi = mid;
  }
  const uint32_t r = root[lo];
  const uint32_t root_len = (uint32_t) (name_off[r + 1] - name_off[r]);
  const uint32_t v = end[j];
  put_scaf_edge(out + off_r[lo] + root_len + off_e[j], names + name_off[v], name_off[v + 1] - name_off[v], dist[j],
                std_bits[j], (flags[j] & 1u) != 0, (flags[j] & 2u) != 0);
}

// lengths are in c->f_len: offsets, total, room check; returns the total through *bytes
static int offsets(gtsb_context *c, uint64_t count, uint64_t cap, uint64_t *bytes, const char *what) {
  cudaStream_t s = c->stream;
  uint32_t total = 0, bad[2] = {0, 0};
  exclusive_scan<uint32_t>(c->f_len.as<uint32_t>(), count, c->f_off.as<uint32_t>(), c->scan_scratch.as<uint32_t>(), s);
  c->stats.kernel_launches += 3;
  CK(cudaMemcpyAsync(&total, c->f_off.as<uint32_t>() + count, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(bad, c->p_flags.as<uint32_t>(), 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  if (bad[0]) return fail(c, "%s: state outside GraphItemState", what);
  // the offsets are 32-bit sums: longest line * lines must stay below 2^32
  if ((uint64_t) bad[1] * count >= (1ull << 32))
    return fail(c, "%s: too much text for one call, pass fewer items", what);
  if (total > cap) return fail(c, "%s: %u bytes of text, room for %llu", what, total, (unsigned long long) cap);
  *bytes = total;
  return 0;
}

// an edge line is at most 20 + 4 + 20 + 9 + 9 + 9 + 11 + 13 + 6 + 4 = 105 bytes, a vertex line
// 20 + 31 + header; offsets() refuses a call whose lines could sum to 2^32 bytes
constexpr uint64_t MAX_ITEMS = 1ull << 25;

}  // namespace gtsbformat

using namespace gtsbformat;

extern "C" {

int gtsb_dot_vertex_lines_host(gtsb_context *c, int scaffold_only, uint64_t first, uint64_t count,
                               const uint8_t *vstate, char *out, uint64_t cap, uint64_t *bytes) {
  if (c == nullptr) return -1;
  CK(cudaSetDevice(c->device));
  if (bytes == nullptr || (count && (vstate == nullptr || out == nullptr)))
    return fail(c, "gtsb_dot_vertex_lines_host: null argument");
  if (!c->have_names) return fail(c, "gtsb_dot_vertex_lines_host: vertex names not set");
  if (first + count > c->names_V) return fail(c, "gtsb_dot_vertex_lines_host: vertices outside the names set");
  if (count > MAX_ITEMS) return fail(c, "gtsb_dot_vertex_lines_host: more than 2^25 vertices in one call");
  *bytes = 0;
  if (count == 0) return 0;
  ProfScope prof(c);
  cudaStream_t s = c->stream;
  ENSURE(c->f_state, count);
  ENSURE(c->f_len, (count + 1) * 4);
  ENSURE(c->f_off, (count + 2) * 4);
  ENSURE(c->scan_scratch, scan_scratch_elems(count) * 4);
  ENSURE(c->p_flags, 16);
  ENSURE(c->f_out, cap + 16);
  CK(cudaMemsetAsync(c->p_flags.p, 0, 16, s));
  CK(cudaMemcpyAsync(c->f_state.p, vstate, count, cudaMemcpyHostToDevice, s));
  const uint32_t n = (uint32_t) count, blocks = (n + 255) / 256;
  {
    GTSB_TIMED("kf_vertex_len", s);
    kf_vertex_len<<<blocks, 256, 0, s>>>(first, n, c->f_state.as<uint8_t>(), c->p_name_off.as<uint64_t>(),
                                         scaffold_only, c->f_len.as<uint32_t>(), c->p_flags.as<uint32_t>());
  }
  if (offsets(c, count, cap, bytes, "gtsb_dot_vertex_lines_host") != 0) return -1;
  {
    GTSB_TIMED("kf_vertex_put", s);
    kf_vertex_put<<<blocks, 256, 0, s>>>(first, n, c->f_state.as<uint8_t>(), c->p_names.as<char>(),
                                         c->p_name_off.as<uint64_t>(), scaffold_only, c->f_off.as<uint32_t>(),
                                         c->f_out.as<char>());
  }
  c->stats.kernel_launches += 2;
  if (*bytes) CK(cudaMemcpyAsync(out, c->f_out.p, *bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return 0;
}

int gtsb_dot_edge_lines_host(gtsb_context *c, int scaffold_only, uint64_t count, const uint32_t *src,
                             const uint32_t *dst, const int32_t *dist, const uint8_t *estate,
                             const uint8_t *sense, char *out, uint64_t cap, uint64_t *bytes) {
  if (c == nullptr) return -1;
  CK(cudaSetDevice(c->device));
  if (bytes == nullptr || (count && (src == nullptr || dst == nullptr || dist == nullptr || estate == nullptr ||
                                     sense == nullptr || out == nullptr)))
    return fail(c, "gtsb_dot_edge_lines_host: null argument");
  if (count > MAX_ITEMS) return fail(c, "gtsb_dot_edge_lines_host: more than 2^25 edges in one call");
  *bytes = 0;
  if (count == 0) return 0;
  ProfScope prof(c);
  cudaStream_t s = c->stream;
  ENSURE(c->f_src, count * 4);
  ENSURE(c->f_dst, count * 4);
  ENSURE(c->f_dist, count * 4);
  ENSURE(c->f_state, count);
  ENSURE(c->f_sense, count);
  ENSURE(c->f_len, (count + 1) * 4);
  ENSURE(c->f_off, (count + 2) * 4);
  ENSURE(c->scan_scratch, scan_scratch_elems(count) * 4);
  ENSURE(c->p_flags, 16);
  ENSURE(c->f_out, cap + 16);
  CK(cudaMemsetAsync(c->p_flags.p, 0, 16, s));
  CK(cudaMemcpyAsync(c->f_src.p, src, count * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(c->f_dst.p, dst, count * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(c->f_dist.p, dist, count * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(c->f_state.p, estate, count, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(c->f_sense.p, sense, count, cudaMemcpyHostToDevice, s));
  const uint32_t n = (uint32_t) count, blocks = (n + 255) / 256;
  {
    GTSB_TIMED("kf_edge_len", s);
    kf_edge_len<<<blocks, 256, 0, s>>>(n, c->f_src.as<uint32_t>(), c->f_dst.as<uint32_t>(), c->f_dist.as<int32_t>(),
                                       c->f_state.as<uint8_t>(), c->f_sense.as<uint8_t>(), scaffold_only,
                                       c->f_len.as<uint32_t>(), c->p_flags.as<uint32_t>());
  }
  if (offsets(c, count, cap, bytes, "gtsb_dot_edge_lines_host") != 0) return -1;
  {
    GTSB_TIMED("kf_edge_put", s);
    kf_edge_put<<<blocks, 256, 0, s>>>(n, c->f_src.as<uint32_t>(), c->f_dst.as<uint32_t>(), c->f_dist.as<int32_t>(),
                                       c->f_state.as<uint8_t>(), c->f_sense.as<uint8_t>(), scaffold_only,
                                       c->f_off.as<uint32_t>(), c->f_out.as<char>());
  }
  c->stats.kernel_launches += 2;
  if (*bytes) CK(cudaMemcpyAsync(out, c->f_out.p, *bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return 0;
}

int gtsb_scaf_lines_host(gtsb_context *c, uint64_t nof_records, const uint32_t *rec_root,
                         const uint64_t *rec_edge_off, const uint32_t *edge_end, const int64_t *edge_dist,
                         const float *edge_std_dev, const uint8_t *edge_flags, char *out, uint64_t cap,
                         uint64_t *bytes) {
  if (c == nullptr) return -1;
  CK(cudaSetDevice(c->device));
  const uint64_t n = nof_records, m = (n && rec_edge_off != nullptr) ? rec_edge_off[n] : 0;
  if (bytes == nullptr || (n && (rec_root == nullptr || rec_edge_off == nullptr || out == nullptr)) ||
      (m && (edge_end == nullptr || edge_dist == nullptr || edge_std_dev == nullptr || edge_flags == nullptr)))
    return fail(c, "gtsb_scaf_lines_host: null argument");
  if (!c->have_names) return fail(c, "gtsb_scaf_lines_host: vertex names not set");
  if (n > MAX_ITEMS || m > MAX_ITEMS) return fail(c, "gtsb_scaf_lines_host: more than 2^25 records or edges in one call");
  *bytes = 0;
  if (n == 0) return 0;
  if (rec_edge_off[0] != 0) return fail(c, "gtsb_scaf_lines_host: rec_edge_off[0] must be 0");
  for (uint64_t i = 0; i < n; i++)
    if (rec_edge_off[i + 1] < rec_edge_off[i]) return fail(c, "gtsb_scaf_lines_host: rec_edge_off must not decrease");
  ProfScope prof(c);
  cudaStream_t s = c->stream;
  ENSURE(c->s_root, n * 4);
  ENSURE(c->s_recoff, (n + 1) * 8);
  ENSURE(c->s_len_r, (n + 1) * 4);
  ENSURE(c->s_off_r, (n + 2) * 4);
  ENSURE(c->s_end, (m + 1) * 4);
  ENSURE(c->s_dist, (m + 1) * 8);
  ENSURE(c->s_std, (m + 1) * 4);
  ENSURE(c->s_flags, m + 1);
  ENSURE(c->f_len, (m + 1) * 4);
  ENSURE(c->f_off, (m + 2) * 4);
  ENSURE(c->scan_scratch, scan_scratch_elems(n > m ? n : m) * 4);
  ENSURE(c->p_flags, 16);
  ENSURE(c->f_out, cap + 16);
  CK(cudaMemsetAsync(c->p_flags.p, 0, 16, s));
  CK(cudaMemcpyAsync(c->s_root.p, rec_root, n * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(c->s_recoff.p, rec_edge_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
  if (m) {
    CK(cudaMemcpyAsync(c->s_end.p, edge_end, m * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->s_dist.p, edge_dist, m * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->s_std.p, edge_std_dev, m * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->s_flags.p, edge_flags, m, cudaMemcpyHostToDevice, s));
  }
  const uint32_t nn = (uint32_t) n, mm = (uint32_t) m;
  {
    GTSB_TIMED("kf_scaf_len", s);
    kf_scaf_rec_len<<<(nn + 255) / 256, 256, 0, s>>>(nn, c->s_root.as<uint32_t>(), c->p_name_off.as<uint64_t>(),
                                                    c->names_V, c->s_len_r.as<uint32_t>(), c->p_flags.as<uint32_t>());
    if (mm)
      kf_scaf_edge_len<<<(mm + 255) / 256, 256, 0, s>>>(mm, c->s_end.as<uint32_t>(), c->s_dist.as<int64_t>(),
                                                       c->s_std.as<uint32_t>(), c->p_name_off.as<uint64_t>(),
                                                       c->names_V, c->f_len.as<uint32_t>(), c->p_flags.as<uint32_t>());
  }
  exclusive_scan<uint32_t>(c->s_len_r.as<uint32_t>(), n, c->s_off_r.as<uint32_t>(), c->scan_scratch.as<uint32_t>(), s);
  exclusive_scan<uint32_t>(c->f_len.as<uint32_t>(), m, c->f_off.as<uint32_t>(), c->scan_scratch.as<uint32_t>(), s);
  c->stats.kernel_launches += 2 + 6;
  uint32_t total_r = 0, total_e = 0, bad[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(&total_r, c->s_off_r.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&total_e, c->f_off.as<uint32_t>() + m, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(bad, c->p_flags.as<uint32_t>(), 12, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  if (bad[0] & 2u) return fail(c, "gtsb_scaf_lines_host: a vertex id outside the names set");
  // 32-bit offsets: the longest piece times the number of pieces must stay below 2^32
  if ((uint64_t) bad[1] * n + (uint64_t) bad[2] * m >= (1ull << 32))
    return fail(c, "gtsb_scaf_lines_host: too much text for one call, pass fewer records");
  const uint64_t total = (uint64_t) total_r + total_e;
  if (total > cap) return fail(c, "gtsb_scaf_lines_host: %llu bytes of text, room for %llu", (unsigned long long) total,
                               (unsigned long long) cap);
  *bytes = total;
  {
    GTSB_TIMED("kf_scaf_put", s);
    kf_scaf_rec_put<<<(nn + 255) / 256, 256, 0, s>>>(nn, c->s_root.as<uint32_t>(), c->s_recoff.as<uint64_t>(),
                                                    c->s_off_r.as<uint32_t>(), c->f_off.as<uint32_t>(),
                                                    c->p_names.as<char>(), c->p_name_off.as<uint64_t>(),
                                                    c->f_out.as<char>());
    if (mm)
      kf_scaf_edge_put<<<(mm + 255) / 256, 256, 0, s>>>(mm, nn, c->s_root.as<uint32_t>(), c->s_recoff.as<uint64_t>(),
                                                       c->s_end.as<uint32_t>(), c->s_dist.as<int64_t>(),
                                                       c->s_std.as<uint32_t>(), c->s_flags.as<uint8_t>(),
                                                       c->s_off_r.as<uint32_t>(), c->f_off.as<uint32_t>(),
                                                       c->p_names.as<char>(), c->p_name_off.as<uint64_t>(),
                                                       c->f_out.as<char>());
  }
  c->stats.kernel_launches += 2;
  if (total) CK(cudaMemcpyAsync(out, c->f_out.p, total, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"
