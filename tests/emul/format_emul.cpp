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
