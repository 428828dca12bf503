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
//
// `.scaf` records of gt_scaffolder_graph_write_scaffold (algorithms.c:1000-1042):
//
//   <root header> { \t<end header>,<dist %ld>,<std_dev %f>,<sense %d>,<same %d>, }* \n
//
// "%f" of a float (promoted to double, exactly) is the one formatted number that is not an
// integer: put_f6 produces the correctly rounded six-decimal text from the float's bits with
// integer arithmetic only (round-half-even on the exact value, what glibc's printf does in the
// default rounding mode), "inf" / "nan" with their signs included.
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

// ---- "%f" ---------------------------------------------------------------------------------

// a float = M * 2^E exactly (M < 2^24); returns false for inf / nan
GTSB_HD bool f32_parts(uint32_t bits, uint32_t *M, int *E) {
  const uint32_t ex = (bits >> 23) & 0xFFu, mant = bits & 0x7FFFFFu;
  if (ex == 0xFFu) return false;
  if (ex == 0) {
    *M = mant;
    *E = -149;
  } else {
    *M = mant | 0x800000u;
    *E = (int) ex - 150;
  }
  return true;
}

// value -> integer part (as up to 128 bits: hi, lo) and the six rounded decimals
GTSB_HD void f6_split(uint32_t M, int E, uint64_t *ip_hi, uint64_t *ip_lo, uint32_t *frac6) {
  if (E >= 0) {                                  // an integer, M << E with E <= 104
    if (E >= 64) {
      *ip_hi = (uint64_t) M << (E - 64);
      *ip_lo = 0;
    } else {
      *ip_hi = E == 0 ? 0 : ((uint64_t) M >> (64 - E));
      *ip_lo = (uint64_t) M << E;
    }
    *frac6 = 0;
    return;
  }
  const int sh = -E;                             // 1 .. 149
  uint64_t ip = sh >= 24 ? 0 : (uint64_t) (M >> sh);
  const uint64_t fr = sh >= 24 ? (uint64_t) M : (uint64_t) (M & ((1u << sh) - 1u));
  const uint64_t N = fr * 1000000ull;            // < 2^44
  uint64_t q = 0;
  if (sh < 64) {
    const uint64_t half = 1ull << (sh - 1), rem = N & ((half << 1) - 1ull);
    q = N >> sh;
    if (rem > half || (rem == half && (q & 1ull))) q++;
  }                                              // sh >= 64: N / 2^sh < 1/2, rounds to 0
  if (q >= 1000000ull) {
    q -= 1000000ull;
    ip++;
  }
  *ip_hi = 0;
  *ip_lo = ip;
  *frac6 = (uint32_t) q;
}

// decimal digits of a 128-bit integer (hi, lo) into buf (least significant first); returns the count
GTSB_HD uint32_t u128_digits(uint64_t hi, uint64_t lo, char *buf) {
  uint32_t n = 0;
  if (hi == 0) {
    do {
      buf[n++] = (char) ('0' + lo % 10);
      lo /= 10;
    } while (lo);
    return n;
  }
  // long division by 10 over four 32-bit limbs
  uint32_t w[4] = {(uint32_t) (hi >> 32), (uint32_t) hi, (uint32_t) (lo >> 32), (uint32_t) lo};
  bool nonzero = true;
  while (nonzero) {
    uint64_t rem = 0;
    nonzero = false;
    for (int k = 0; k < 4; k++) {
      const uint64_t cur = (rem << 32) | w[k];
      w[k] = (uint32_t) (cur / 10);
      rem = cur % 10;
      nonzero |= w[k] != 0;
    }
    buf[n++] = (char) ('0' + rem);
  }
  return n;
}

GTSB_HD uint32_t f6_len(uint32_t bits) {
  uint32_t M;
  int E;
  const uint32_t sign = bits >> 31;
  if (!f32_parts(bits, &M, &E)) return sign + 3;                 // [-]inf, [-]nan
  uint64_t hi, lo;
  uint32_t fr;
  f6_split(M, E, &hi, &lo, &fr);
  char tmp[40];
  return sign + u128_digits(hi, lo, tmp) + 7;
}

GTSB_HD char *put_f6(char *p, uint32_t bits) {
  uint32_t M;
  int E;
  if (bits >> 31) *p++ = '-';
  if (!f32_parts(bits, &M, &E)) return (bits & 0x7FFFFFu) ? put_str(p, "nan", 3) : put_str(p, "inf", 3);
  uint64_t hi, lo;
  uint32_t fr;
  f6_split(M, E, &hi, &lo, &fr);
  char tmp[40];
  const uint32_t n = u128_digits(hi, lo, tmp);
  for (uint32_t i = 0; i < n; i++) p[i] = tmp[n - 1 - i];
  p += n;
  *p++ = '.';
  for (int i = 5; i >= 0; i--) {
    p[i] = (char) ('0' + fr % 10);
    fr /= 10;
  }
  return p + 6;
}

// ---- `.scaf` ------------------------------------------------------------------------------

// one edge of a record: \t<end header>,<dist>,<std_dev>,<sense>,<same>,
GTSB_HD uint32_t scaf_edge_len(uint64_t name_len, int64_t dist, uint32_t std_bits) {
  return 1 + (uint32_t) name_len + 1 + sdec_len(dist) + 1 + f6_len(std_bits) + 5;
}

GTSB_HD void put_scaf_edge(char *p, const char *name, uint64_t name_len, int64_t dist, uint32_t std_bits, bool sense,
                           bool same) {
  *p++ = '\t';
  p = put_str(p, name, (uint32_t) name_len);
  *p++ = ',';
  p = put_sdec(p, dist);
  *p++ = ',';
  p = put_f6(p, std_bits);
  *p++ = ',';
  *p++ = sense ? '1' : '0';
  *p++ = ',';
  *p++ = same ? '1' : '0';
  *p++ = ',';
}

}  // namespace gtsbf
