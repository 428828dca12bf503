// format_emul.cpp -- TEST INFRASTRUCTURE.  The line functions of
// gt-scaffold_b200/csrc/gtsb_format_core.h (the bodies of the kernels in gtsb_format.cu) as
// plain host loops: lengths, prefix sum, lines -- the device driver's order.  Compared byte
// for byte with what the compiled reference prints.  Nothing in the product links this file.
#include <stdint.h>

#include <vector>

#include "../../gt-scaffold_b200/csrc/gtsb_format_core.h"

using namespace gtsbf;

extern "C" {

int emul_dot_vertex_lines(int scaffold_only, uint64_t first, uint64_t count, const uint8_t *vstate,
                          const char *names, const uint64_t *name_off, char *out, uint64_t cap,
                          uint64_t *bytes) {
  std::vector<uint64_t> off(count + 1, 0);
  for (uint64_t i = 0; i < count; i++) {
    const uint64_t v = first + i;
    if (vstate[i] >= NOF_STATES) return -1;
    off[i + 1] = off[i] + vertex_line_len(v, vstate[i], name_off[v + 1] - name_off[v], scaffold_only != 0);
  }
  *bytes = off[count];
  if (off[count] > cap) return -1;
  for (uint64_t k = 0; k < count; k++) {
    const uint64_t i = count - 1 - k, v = first + i;               // any order
    put_vertex_line(out + off[i], v, vstate[i], names + name_off[v], name_off[v + 1] - name_off[v],
                    scaffold_only != 0);
  }
  return 0;
}

int emul_dot_edge_lines(int scaffold_only, uint64_t count, const uint32_t *src, const uint32_t *dst,
                        const int32_t *dist, const uint8_t *estate, const uint8_t *sense, char *out,
                        uint64_t cap, uint64_t *bytes) {
  std::vector<uint64_t> off(count + 1, 0);
  for (uint64_t i = 0; i < count; i++) {
    if (estate[i] >= NOF_STATES) return -1;
    off[i + 1] = off[i] + edge_line_len(src[i], dst[i], dist[i], estate[i], sense[i] != 0, scaffold_only != 0);
  }
  *bytes = off[count];
  if (off[count] > cap) return -1;
  for (uint64_t k = 0; k < count; k++) {
    const uint64_t i = count - 1 - k;
    put_edge_line(out + off[i], src[i], dst[i], dist[i], estate[i], sense[i] != 0, scaffold_only != 0);
  }
  return 0;
}

}  // extern "C"

// ---- `.scaf` (gtsb_scaf_lines_host): the device driver's arithmetic -- piece lengths, two prefix
// sums, pieces in any order

extern "C" {

uint32_t emul_f6(uint32_t bits, char *out) {
  const uint32_t n = f6_len(bits);
  char *e = put_f6(out, bits);
  return (uint32_t) (e - out) == n ? n : 0xFFFFFFFFu;
}

int emul_scaf_lines(uint64_t n, const uint32_t *rec_root, const uint64_t *rec_edge_off, const uint32_t *edge_end,
                    const int64_t *edge_dist, const uint32_t *edge_std_bits, const uint8_t *edge_flags,
                    const char *names, const uint64_t *name_off, uint64_t names_V, char *out, uint64_t cap,
                    uint64_t *bytes) {
  const uint64_t m = n ? rec_edge_off[n] : 0;
  std::vector<uint64_t> off_r(n + 1, 0), off_e(m + 1, 0);
  for (uint64_t i = 0; i < n; i++) {
    if (rec_root[i] >= names_V) return -1;
    off_r[i + 1] = off_r[i] + (name_off[rec_root[i] + 1] - name_off[rec_root[i]]) + 1;
  }
  for (uint64_t j = 0; j < m; j++) {
    if (edge_end[j] >= names_V) return -1;
    off_e[j + 1] = off_e[j] + scaf_edge_len(name_off[edge_end[j] + 1] - name_off[edge_end[j]], edge_dist[j],
                                            edge_std_bits[j]);
  }
  *bytes = off_r[n] + off_e[m];
  if (*bytes > cap) return -1;
  for (uint64_t k = 0; k < n; k++) {
    const uint64_t i = n - 1 - k, v = rec_root[i], len = name_off[v + 1] - name_off[v];
    put_str(out + off_r[i] + off_e[rec_edge_off[i]], names + name_off[v], (uint32_t) len);
    out[off_r[i] + len + off_e[rec_edge_off[i + 1]]] = '\n';
    for (uint64_t j = rec_edge_off[i]; j < rec_edge_off[i + 1]; j++) {
      const uint64_t w = edge_end[j];
      put_scaf_edge(out + off_r[i] + len + off_e[j], names + name_off[w], name_off[w + 1] - name_off[w], edge_dist[j],
                    edge_std_bits[j], (edge_flags[j] & 1u) != 0, (edge_flags[j] & 2u) != 0);
    }
  }
  return 0;
}

}  // extern "C"
