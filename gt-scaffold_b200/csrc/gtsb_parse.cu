// gtsb_parse.cu -- `.de` text -> integer records on the device (SURVEY.md §8(f) rank 1;
// replaces the record loop of gt_scaffolder_parser_read_distances, parser.c:323-388, for
// files in the canonical spelling -- gtsb_parse_core.h states the rules and the limits).
//
//   kp_table_insert   headers -> open-addressing table (one thread per contig)
//   kp_chunk_newlines '\n' count per 64-byte chunk                        } line_end[]:
//   exclusive_scan    first line index of every chunk                     } where every
//   kp_line_ends      offsets one past each '\n'                          } physical line stops
//   kp_walk<false>    one thread per physical line: number of records
//   exclusive_scan    first record index of every line
//   kp_walk<true>     the same walk, records written in file order
//
// The walk is a byte loop per thread over its own line; neighbouring threads hold neighbouring
// lines, so a warp's loads fall into one window of a few KB that L1 keeps.
#include "gtsb_context.h"
#include "gtsb_parse_core.h"
#include "gtsb_scan.cuh"

using namespace gtsb;
using namespace gtsbi;
using namespace gtsbp;

namespace gtsbparse {

__global__ void __launch_bounds__(256) kp_table_insert(NameTable t, uint32_t V, uint32_t *irregular) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < V) table_insert(t, v, irregular);
}

__global__ void __launch_bounds__(256) kp_chunk_newlines(const char *__restrict__ text, uint64_t n,
                                                         uint64_t nchunks, uint8_t *__restrict__ cnt,
                                                         uint32_t *irregular) {
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nchunks) cnt[i] = (uint8_t) chunk_newlines(text, n, i, irregular);
}

__global__ void __launch_bounds__(256) kp_line_ends(const char *__restrict__ text, uint64_t n, uint64_t nchunks,
                                                    const uint8_t *__restrict__ cnt,
                                                    const uint32_t *__restrict__ first,
                                                    uint64_t *__restrict__ line_end) {
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nchunks && cnt[i] != 0) chunk_line_ends(text, n, i, first[i], line_end);
}

template <bool EMIT>
__global__ void __launch_bounds__(128) kp_walk(const char *__restrict__ text, const uint64_t *__restrict__ line_end,
                                               uint64_t nlines, NameTable t, Records out,
                                               uint32_t *__restrict__ line_cnt,
                                               const uint32_t *__restrict__ line_off, uint32_t *irregular) {
  const uint64_t l = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nlines) return;
  const uint64_t s = l ? line_end[l - 1] : 0, e = line_end[l];
  if (EMIT) {
    if (line_off[l + 1] != line_off[l]) walk_line<true>(text, s, e, t, out, line_off[l], irregular);
  } else {
    line_cnt[l] = walk_line<false>(text, s, e, t, out, 0, irregular);
  }
}

inline uint32_t blocks_for(uint64_t n, uint32_t threads) { return (uint32_t) ((n + threads - 1) / threads); }

}  // namespace gtsbparse

using namespace gtsbparse;

extern "C" {

int gtsb_set_vertex_names_host(gtsb_context *c, uint64_t V, const char *names, const uint64_t *name_off) {
  if (c == nullptr) return -1;
  CK(cudaSetDevice(c->device));
  c->names_V = 0;
  if (V > GTSB_MAX_VERTICES) return fail(c, "gtsb_set_vertex_names_host: too many contigs");
  if (V && (names == nullptr || name_off == nullptr)) return fail(c, "gtsb_set_vertex_names_host: null input");
  const uint64_t bytes = V ? name_off[V] : 0;
  uint64_t cap = 2;
  while (cap < 2 * V) cap <<= 1;
  ProfScope prof(c);
  cudaStream_t s = c->stream;
  ENSURE(c->p_names, bytes + 8);
  ENSURE(c->p_name_off, (V + 1) * 8);
  ENSURE(c->p_slots, cap * 8);
  ENSURE(c->p_flags, 16);
  if (bytes) CK(cudaMemcpyAsync(c->p_names.p, names, bytes, cudaMemcpyHostToDevice, s));
  if (V) CK(cudaMemcpyAsync(c->p_name_off.p, name_off, (V + 1) * 8, cudaMemcpyHostToDevice, s));
  else CK(cudaMemsetAsync(c->p_name_off.p, 0, 8, s));
  CK(cudaMemsetAsync(c->p_slots.p, 0, cap * 8, s));
  CK(cudaMemsetAsync(c->p_flags.p, 0, 16, s));
  c->names_mask = cap - 1;
  if (V) {
    const NameTable t{c->p_names.as<char>(), c->p_name_off.as<uint64_t>(), c->p_slots.as<uint64_t>(), cap - 1};
    GTSB_TIMED("kp_table_insert", s);
    kp_table_insert<<<blocks_for(V, 256), 256, 0, s>>>(t, (uint32_t) V, c->p_flags.as<uint32_t>());
    c->stats.kernel_launches++;
  }
  uint32_t irregular = 0;
  CK(cudaMemcpyAsync(&irregular, c->p_flags.p, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  if (irregular & IRR_DUP_NAME)
    return fail(c, "gtsb_set_vertex_names_host: two contigs carry the same header");
  c->names_V = V;
  c->have_names = true;
  return 0;
}

int gtsb_parse_de_host(gtsb_context *c, const char *text, uint64_t n, uint64_t *nof_records,
                       uint32_t *irregular_out) {
  if (c == nullptr) return -1;
  CK(cudaSetDevice(c->device));
  if (!c->have_names) return fail(c, "gtsb_parse_de_host: vertex names not set");
  if (nof_records == nullptr || irregular_out == nullptr || (n && text == nullptr))
    return fail(c, "gtsb_parse_de_host: null argument");
  if (c->world > 1) return fail(c, "gtsb_parse_de_host: single-device contexts only");
  // line and record counts are 32-bit sums over the text: every line and every record takes
  // a byte at least, so below 4 GiB none of them can wrap
  if (n >= (1ull << 32)) return fail(c, "gtsb_parse_de_host: text of 4 GiB or more");
  if (await_records(c) != 0) return -1;
  ProfScope prof(c);
  cudaStream_t s = c->stream;
  *nof_records = 0;
  *irregular_out = 0;
  c->have_records = false;
  c->have_graph = false;

  const uint64_t nchunks = (n + CHUNK - 1) / CHUNK;
  ENSURE(c->p_text, n + 8);
  ENSURE(c->p_chunk_cnt, nchunks + 1);
  ENSURE(c->p_chunk_off, (nchunks + 2) * 4);
  ENSURE(c->scan_scratch, scan_scratch_elems(nchunks > c->R ? nchunks : c->R) * 4);
  uint32_t *flags = c->p_flags.as<uint32_t>();
  CK(cudaMemsetAsync(flags, 0, 16, s));
  if (n) CK(cudaMemcpyAsync(c->p_text.p, text, n, cudaMemcpyHostToDevice, s));
  const char *d_text = c->p_text.as<char>();

  // physical lines
  uint32_t newlines = 0;
  if (nchunks) {
    {
      GTSB_TIMED("kp_chunk_newlines", s);
      kp_chunk_newlines<<<blocks_for(nchunks, 256), 256, 0, s>>>(d_text, n, nchunks, c->p_chunk_cnt.as<uint8_t>(),
                                                                flags);
    }
    exclusive_scan<uint8_t>(c->p_chunk_cnt.as<uint8_t>(), nchunks, c->p_chunk_off.as<uint32_t>(),
                            c->scan_scratch.as<uint32_t>(), s);
    c->stats.kernel_launches += 4;
    CK(cudaMemcpyAsync(&newlines, c->p_chunk_off.as<uint32_t>() + nchunks, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  const bool open_end = n != 0 && text[n - 1] != '\n';       // last line without '\n'
  const uint64_t nlines = (uint64_t) newlines + (open_end ? 1 : 0);
  if (nlines >= 0xFFFFFFF0ull) return fail(c, "gtsb_parse_de_host: too many lines");
  ENSURE(c->p_line_end, (nlines + 1) * 8);
  ENSURE(c->p_line_cnt, (nlines + 1) * 4);
  ENSURE(c->p_line_off, (nlines + 2) * 4);
  ENSURE(c->scan_scratch, scan_scratch_elems(nlines > nchunks ? nlines : nchunks) * 4);
  uint64_t *line_end = c->p_line_end.as<uint64_t>();
  if (newlines) {
    GTSB_TIMED("kp_line_ends", s);
    kp_line_ends<<<blocks_for(nchunks, 256), 256, 0, s>>>(d_text, n, nchunks, c->p_chunk_cnt.as<uint8_t>(),
                                                          c->p_chunk_off.as<uint32_t>(), line_end);
    c->stats.kernel_launches++;
  }
  if (open_end) CK(cudaMemcpyAsync(line_end + newlines, &n, 8, cudaMemcpyHostToDevice, s));

  // records per line, then the records
  const NameTable t{c->p_names.as<char>(), c->p_name_off.as<uint64_t>(), c->p_slots.as<uint64_t>(),
                    c->names_mask};
  Records out{};
  uint32_t total = 0, irregular = 0;
  if (nlines) {
    {
      GTSB_TIMED("kp_walk(count)", s);
      kp_walk<false><<<blocks_for(nlines, 128), 128, 0, s>>>(d_text, line_end, nlines, t, out,
                                                            c->p_line_cnt.as<uint32_t>(), nullptr, flags);
    }
    exclusive_scan<uint32_t>(c->p_line_cnt.as<uint32_t>(), nlines, c->p_line_off.as<uint32_t>(),
                             c->scan_scratch.as<uint32_t>(), s);
    c->stats.kernel_launches += 4;
    CK(cudaMemcpyAsync(&total, c->p_line_off.as<uint32_t>() + nlines, 4, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaMemcpyAsync(&irregular, flags, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  if (irregular) {                     // nothing is kept: the caller parses this file on the host
    *irregular_out = irregular;
    return 0;
  }
  const uint64_t R = total;
  DevBuf *bs[] = {&c->root, &c->ctg, &c->dist, &c->std_dev, &c->flags};
  for (DevBuf *b : bs)
    if (!b->owned) *b = DevBuf();
  ENSURE(c->root, R * 4);
  ENSURE(c->ctg, R * 4);
  ENSURE(c->dist, R * 4);
  ENSURE(c->std_dev, R * 4);
  ENSURE(c->flags, R);
  ENSURE(c->num_pairs, R * 4);
  out = Records{c->root.as<uint32_t>(), c->ctg.as<uint32_t>(), c->dist.as<int32_t>(), c->std_dev.as<float>(),
                c->num_pairs.as<uint32_t>(), c->flags.as<uint8_t>()};
  if (R) {
    GTSB_TIMED("kp_walk(emit)", s);
    kp_walk<true><<<blocks_for(nlines, 128), 128, 0, s>>>(d_text, line_end, nlines, t, out, nullptr,
                                                         c->p_line_off.as<uint32_t>(), flags);
    c->stats.kernel_launches++;
    CK(cudaGetLastError());
  }
  c->R = R;
  c->have_records = true;
  c->have_num_pairs = true;
  c->stats.nof_records = R;
  *nof_records = R;
  return 0;
}

int gtsb_get_records(gtsb_context *c, uint32_t *root, uint32_t *ctg, int32_t *dist, float *std_dev,
                     uint8_t *flags, uint32_t *num_pairs) {
  if (c == nullptr) return -1;
  CK(cudaSetDevice(c->device));
  if (!c->have_records) return fail(c, "gtsb_get_records: no records");
  if (num_pairs != nullptr && !c->have_num_pairs)
    return fail(c, "gtsb_get_records: pair counts exist for records parsed on the device only");
  if (await_records(c) != 0) return -1;
  const uint64_t R = c->R;
  cudaStream_t s = c->stream;
  if (R) {
    if (root) CK(cudaMemcpyAsync(root, c->root.p, R * 4, cudaMemcpyDeviceToHost, s));
    if (ctg) CK(cudaMemcpyAsync(ctg, c->ctg.p, R * 4, cudaMemcpyDeviceToHost, s));
    if (dist) CK(cudaMemcpyAsync(dist, c->dist.p, R * 4, cudaMemcpyDeviceToHost, s));
    if (std_dev) CK(cudaMemcpyAsync(std_dev, c->std_dev.p, R * 4, cudaMemcpyDeviceToHost, s));
    if (flags) CK(cudaMemcpyAsync(flags, c->flags.p, R, cudaMemcpyDeviceToHost, s));
    if (num_pairs) CK(cudaMemcpyAsync(num_pairs, c->num_pairs.p, R * 4, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  return 0;
}

}  // extern "C"
