// gtsb_format_core.h -- the `.dot` lines of gt_scaffolder_graph_print_generic and
// gt_scaffolder_graph_print_scaffold (graph.c:269-343), one function per line kind, written
// like gtsb_parse_core.h: the same source is the body of the CUDA kernels (gtsb_format.cu)
// and of a plain C++ loop (tests/emul/format_emul.cpp) that is compared byte for byte with
// what the compiled reference prints.
//
//   vertex, generic :  <id> [color="<colour of state>" label="<header>"];\n
//   vertex, scaffold:  <id> [label="<header>"];\n                         (state == SCAFFOLD only)
//   edge, generic   :  <src> -> <dst> [color="<colour>" label="<dist>" arrowhead="normal|inv"];\n
//   edge, scaffold  :  <src> -> <dst> [label="<dist>" arrowhead="normal|inv"];\n   (SCAFFOLD only)
//
// Only integers and strings are printed (%lu, %ld, %s), so the text is exact by construction.
#pragma once
#include <stdint.h>

#ifndef GTSB_HD
#if defined(__CUDACC__)
#define GTSB_HD __host__ __device__ __forceinline__
#else
#define GTSB_HD inline
#endif
#endif

namespace gtsbf {

constexpr uint32_t STATE_SCAFFOLD = 6;          // GIS_SCAFFOLD, graph.h:29-31
constexpr uint32_t NOF_STATES = 8;

GTSB_HD uint32_t dec_len(uint64_t v) {
  uint32_t n = 1;
  while (v >= 10) {
    v /= 10;
    n++;
  }
  return n;
}

GTSB_HD char *put_dec(char *p, uint64_t v) {
  const uint32_t n = dec_len(v);
  for (uint32_t i = n; i > 0; i--) {
    p[i - 1] = (char) ('0' + v % 10);
    v /= 10;
  }
  return p + n;
}

GTSB_HD uint32_t sdec_len(int64_t v) { return v < 0 ? 1 + dec_len(0 - (uint64_t) v) : dec_len((uint64_t) v); }

GTSB_HD char *put_sdec(char *p, int64_t v) {
  if (v < 0) {
    *p++ = '-';
    return put_dec(p, 0 - (uint64_t) v);
  }
  return put_dec(p, (uint64_t) v);
}

GTSB_HD char *put_str(char *p, const char *s, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) p[i] = s[i];
  return p + n;
}

// color_array of graph.c:277-278, indexed by GraphItemState
GTSB_HD uint32_t colour_len(uint32_t state) {
  switch (state) {
    case 0: return 5;   // black
    case 1: return 6;   // gray80
    case 2: return 9;   // gainsboro
    case 3: return 6;   // ivory3
    case 4: return 3;   // red
    case 5: return 5;   // green
    case 6: return 7;   // magenta
    default: return 4;  // blue
  }
}

GTSB_HD char *put_colour(char *p, uint32_t state) {
  switch (state) {
    case 0: return put_str(p, "black", 5);
    case 1: return put_str(p, "gray80", 6);
    case 2: return put_str(p, "gainsboro", 9);
    case 3: return put_str(p, "ivory3", 6);
    case 4: return put_str(p, "red", 3);
    case 5: return put_str(p, "green", 5);
    case 6: return put_str(p, "magenta", 7);
    default: return put_str(p, "blue", 4);
  }
}

// ---- vertices ---------------------------------------------------------------------------

GTSB_HD uint32_t vertex_line_len(uint64_t id, uint32_t state, uint64_t name_len, bool scaffold_only) {
  if (scaffold_only)
    return state == STATE_SCAFFOLD ? dec_len(id) + 9 + (uint32_t) name_len + 4 : 0;
  return dec_len(id) + 9 + colour_len(state) + 9 + (uint32_t) name_len + 4;
}

GTSB_HD void put_vertex_line(char *p, uint64_t id, uint32_t state, const char *name, uint64_t name_len,
                             bool scaffold_only) {
  if (scaffold_only) {
    if (state != STATE_SCAFFOLD) return;
    p = put_dec(p, id);
    p = put_str(p, " [label=\"", 9);
  } else {
    p = put_dec(p, id);
    p = put_str(p, " [color=\"", 9);
    p = put_colour(p, state);
    p = put_str(p, "\" label=\"", 9);
  }
  p = put_str(p, name, (uint32_t) name_len);
  put_str(p, "\"];\n", 4);
}

// ---- edges ------------------------------------------------------------------------------

GTSB_HD uint32_t edge_line_len(uint64_t src, uint64_t dst, int64_t dist, uint32_t state, bool sense,
                               bool scaffold_only) {
  const uint32_t ends = dec_len(src) + 4 + dec_len(dst);
  const uint32_t tail = sdec_len(dist) + 13 + (sense ? 6u : 3u) + 4;
  if (scaffold_only) return state == STATE_SCAFFOLD ? ends + 9 + tail : 0;
  return ends + 9 + colour_len(state) + 9 + tail;
}

GTSB_HD void put_edge_line(char *p, uint64_t src, uint64_t dst, int64_t dist, uint32_t state, bool sense,
                           bool scaffold_only) {
  if (scaffold_only && state != STATE_SCAFFOLD) return;
  p = put_dec(p, src);
  p = put_str(p, " -> ", 4);
  p = put_dec(p, dst);
  if (scaffold_only) {
    p = put_str(p, " [label=\"", 9);
  } else {
    p = put_str(p, " [color=\"", 9);
    p = put_colour(p, state);
    p = put_str(p, "\" label=\"", 9);
  }
  p = put_sdec(p, dist);
  p = put_str(p, "\" arrowhead=\"", 13);
  p = sense ? put_str(p, "normal", 6) : put_str(p, "inv", 3);
  put_str(p, "\"];\n", 4);
}

}  // namespace gtsbf
