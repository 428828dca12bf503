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
// and, for `.astat` text (algorithms.c:118-149), on the same line index
//   kp_astat<false>   one thread per line: per contig, the offset of the last line naming it
//   kp_astat<true>    that line writes the contig's a-statistic and copy number
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

template <bool APPLY>
__global__ void __launch_bounds__(128) kp_astat(const char *__restrict__ text, const uint64_t *__restrict__ line_end,
                                                uint64_t nlines, NameTable t, uint64_t *last, float *astat,
                                                float *copy_num, uint32_t *irregular) {
  const uint64_t l = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nlines) return;
  walk_astat_line<APPLY>(text, l ? line_end[l - 1] : 0, line_end[l], t, last, astat, copy_num, irregular);
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
  // two contigs with one header: which of them bsearch finds is the C library's business, the
  // texts of this graph are left to the host (every parse call reports GTSB_IRR_DUP_NAME)
  c->names_dup = (irregular & IRR_DUP_NAME) != 0;
  c->names_V = V;
  c->have_names = true;
  return 0;
}

}  // extern "C"

namespace gtsbparse {

// text -> device, physical lines indexed: c->p_line_end[l] = offset one past line l
static int index_lines(gtsb_context *c, const char *text, uint64_t n, uint64_t *nlines_out) {
  cudaStream_t s = c->stream;
  const uint64_t nchunks = (n + CHUNK - 1) / CHUNK;
  ENSURE(c->p_text, n + 8);
  ENSURE(c->p_chunk_cnt, nchunks + 1);
  ENSURE(c->p_chunk_off, (nchunks + 2) * 4);
  ENSURE(c->scan_scratch, scan_scratch_elems(nchunks) * 4);
  uint32_t *flags = c->p_flags.as<uint32_t>();
  CK(cudaMemsetAsync(flags, 0, 16, s));
  if (n) CK(cudaMemcpyAsync(c->p_text.p, text, n, cudaMemcpyHostToDevice, s));
  const char *d_text = c->p_text.as<char>();
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
  ENSURE(c->p_line_end, (nlines + 1) * 8);
  uint64_t *line_end = c->p_line_end.as<uint64_t>();
  if (newlines) {
    GTSB_TIMED("kp_line_ends", s);
    kp_line_ends<<<blocks_for(nchunks, 256), 256, 0, s>>>(d_text, n, nchunks, c->p_chunk_cnt.as<uint8_t>(),
                                                          c->p_chunk_off.as<uint32_t>(), line_end);
    c->stats.kernel_launches++;
  }
  if (open_end) CK(cudaMemcpyAsync(line_end + newlines, &n, 8, cudaMemcpyHostToDevice, s));
  *nlines_out = nlines;
  return 0;
}

static int parse_preamble(gtsb_context *c, const char *what, const char *text, uint64_t n) {
  if (!c->have_names) return fail(c, "%s: vertex names not set", what);
  if (n && text == nullptr) return fail(c, "%s: null argument", what);
  if (c->world > 1) return fail(c, "%s: single-device contexts only", what);
  // line and record counts are 32-bit sums over the text: every line and every record takes
  // a byte at least, so below 4 GiB none of them can wrap
  if (n >= (1ull << 32)) return fail(c, "%s: text of 4 GiB or more", what);
  return 0;
}

}  // namespace gtsbparse

extern "C" {

int gtsb_parse_de_host(gtsb_context *c, const char *text, uint64_t n, uint64_t *nof_records,
                       uint32_t *irregular_out) {
  if (c == nullptr) return -1;
  CK(cudaSetDevice(c->device));
  if (nof_records == nullptr || irregular_out == nullptr) return fail(c, "gtsb_parse_de_host: null argument");
  if (parse_preamble(c, "gtsb_parse_de_host", text, n) != 0) return -1;
  if (await_records(c) != 0) return -1;
  ProfScope prof(c);
  cudaStream_t s = c->stream;
  *nof_records = 0;
  *irregular_out = 0;
  c->have_records = false;
  c->have_graph = false;
  if (c->names_dup) {
    *irregular_out = IRR_DUP_NAME;
    return 0;
  }
  uint64_t nlines = 0;
  if (index_lines(c, text, n, &nlines) != 0) return -1;
  ENSURE(c->p_line_cnt, (nlines + 1) * 4);
  ENSURE(c->p_line_off, (nlines + 2) * 4);
  ENSURE(c->scan_scratch, scan_scratch_elems(nlines) * 4);
  const char *d_text = c->p_text.as<char>();
  const uint64_t *line_end = c->p_line_end.as<uint64_t>();
  uint32_t *flags = c->p_flags.as<uint32_t>();

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
  c->have_root_column = true;
  c->have_lines = false;
  c->have_num_pairs = true;
  c->stats.nof_records = R;
  *nof_records = R;
  return 0;
}

int gtsb_parse_astat_host(gtsb_context *c, const char *text, uint64_t n, float *astat, float *copy_num,
                          uint32_t *irregular_out) {
  if (c == nullptr) return -1;
  CK(cudaSetDevice(c->device));
  if (irregular_out == nullptr) return fail(c, "gtsb_parse_astat_host: null argument");
  if (parse_preamble(c, "gtsb_parse_astat_host", text, n) != 0) return -1;
  const uint64_t V = c->names_V;
  if (V && (astat == nullptr || copy_num == nullptr)) return fail(c, "gtsb_parse_astat_host: null argument");
  ProfScope prof(c);
  cudaStream_t s = c->stream;
  *irregular_out = 0;
  if (c->names_dup) {
    *irregular_out = IRR_DUP_NAME;
    return 0;
  }
  uint64_t nlines = 0;
  if (index_lines(c, text, n, &nlines) != 0) return -1;
  ENSURE(c->p_last, (V + 1) * 8);
  ENSURE(c->p_astat, (V + 1) * 4);
  ENSURE(c->p_copy_num, (V + 1) * 4);
  CK(cudaMemsetAsync(c->p_last.p, 0, (V + 1) * 8, s));
  if (V) {
    CK(cudaMemcpyAsync(c->p_astat.p, astat, V * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->p_copy_num.p, copy_num, V * 4, cudaMemcpyHostToDevice, s));
  }
  const NameTable t{c->p_names.as<char>(), c->p_name_off.as<uint64_t>(), c->p_slots.as<uint64_t>(),
                    c->names_mask};
  uint32_t *flags = c->p_flags.as<uint32_t>();
  uint32_t irregular = 0;
  if (nlines) {
    GTSB_TIMED("kp_astat(last line)", s);
    kp_astat<false><<<blocks_for(nlines, 128), 128, 0, s>>>(c->p_text.as<char>(), c->p_line_end.as<uint64_t>(),
                                                           nlines, t, c->p_last.as<uint64_t>(),
                                                           c->p_astat.as<float>(), c->p_copy_num.as<float>(), flags);
    c->stats.kernel_launches++;
  }
  CK(cudaMemcpyAsync(&irregular, flags, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  if (irregular) {                     // nothing is touched: the caller reads this file on the host
    *irregular_out = irregular;
    return 0;
  }
  if (nlines) {
    GTSB_TIMED("kp_astat(apply)", s);
    kp_astat<true><<<blocks_for(nlines, 128), 128, 0, s>>>(c->p_text.as<char>(), c->p_line_end.as<uint64_t>(),
                                                          nlines, t, c->p_last.as<uint64_t>(),
                                                          c->p_astat.as<float>(), c->p_copy_num.as<float>(), flags);
    c->stats.kernel_launches++;
  }
  if (V) {
    CK(cudaMemcpyAsync(astat, c->p_astat.p, V * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(copy_num, c->p_copy_num.p, V * 4, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
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
  if (root != nullptr && gtsbi::ensure_root_column(c) != 0) return -1;
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
