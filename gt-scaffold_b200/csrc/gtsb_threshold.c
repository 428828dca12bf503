/* gtsb_threshold.c -- host side of the ambiguous-order test.

   gt_scaffolder_graph_ambiguousorder (reference gt_scaffolder_algorithms.c:
   174-193) ends in
       prob12  = 0.5 * (1 + erf(interval));      (double math, float result)
       prob21  = 1.0 - prob12;
       p_wrong = 1.0 - MAX(prob12, prob21);
       return p_wrong > cutoff;
   which is a function g of the single float `interval` (and the cutoff).  The
   reference evaluates erf with the host's libm; parity is defined against that
   libm.  Rather than re-implementing erf on the device,
   the host reduces g to two float thresholds per call,
       g(x) <=> 0 <= x <= t_pos    or    -t_neg <= x < 0,
   found by bisection over float bit patterns with this very libm, and checks
   that g really is a step function around them.  The device then only has to
   reproduce `interval` bit-exactly (gtsb_common.cuh: ambiguous_order).

   Compile with FMA contraction off and without -ffast-math. */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "gtsb_threshold.h"

#define GTSB_MAX(a,b) ((a)>(b)?(a):(b))   /* GenomeTools core/minmax.h */

/* the tail of gt_scaffolder_graph_ambiguousorder, conversions as in the C */
int gtsb_ambig_tail(float interval, float cutoff)
{
  float prob12, prob21, p_wrong;
  prob12 = 0.5 * (1 + erf(interval));
  prob21 = 1.0 - prob12;
  p_wrong = 1.0 - GTSB_MAX(prob12, prob21);
  return (p_wrong > cutoff) ? 1 : 0;
}

static float from_bits(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }

/* largest t >= 0 with g(sign*t) true, given g(sign*0) true and g(sign*inf)
   false; *ok is cleared if g is not a clean step around t */
static float bisect_side(float sign, float cutoff, int *ok)
{
  uint32_t lo = 0u, hi = 0x7f800000u, b, k;   /* bits of +0 and +inf */
  while (hi - lo > 1u) {
    uint32_t mid = lo + (hi - lo) / 2u;
    if (gtsb_ambig_tail(sign * from_bits(mid), cutoff)) lo = mid; else hi = mid;
  }
  /* neighbourhood scan: 1<<14 ulps either side */
  for (k = 1; k <= (1u << 14); k++) {
    if (lo >= k && !gtsb_ambig_tail(sign * from_bits(lo - k), cutoff)) *ok = 0;
    b = hi + k - 1u;
    if (b <= 0x7f800000u && gtsb_ambig_tail(sign * from_bits(b), cutoff)) *ok = 0;
  }
  /* coarse scan over the whole positive range: every 2^12-th bit pattern */
  for (b = 0u; b < 0x7f800000u; b += (1u << 12)) {
    int g = gtsb_ambig_tail(sign * from_bits(b), cutoff);
    if ((b <= lo) != (g != 0)) *ok = 0;
  }
  return from_bits(lo);
}

int gtsb_ambig_thresholds(float cutoff, float *t_pos, float *t_neg, int *inf_true)
{
  int ok = 1;
  const float inf = from_bits(0x7f800000u);
  int g_pinf = gtsb_ambig_tail(inf, cutoff), g_ninf = gtsb_ambig_tail(-inf, cutoff);
  *inf_true = g_pinf;
  if (g_pinf != g_ninf) ok = 0;
  if (!gtsb_ambig_tail(0.0f, cutoff)) {
    /* p_wrong is largest (0.5) at interval 0: never ambiguous */
    *t_pos = -1.0f;
    /* verify on a coarse grid that nothing else fires */
    for (uint32_t b = 0u; b <= 0x7f800000u; b += (1u << 12))
      if (gtsb_ambig_tail(from_bits(b), cutoff)) ok = 0;
  }
  else if (g_pinf) {
    *t_pos = inf;
    for (uint32_t b = 0u; b <= 0x7f800000u; b += (1u << 12))
      if (!gtsb_ambig_tail(from_bits(b), cutoff)) ok = 0;
  }
  else
    *t_pos = bisect_side(1.0f, cutoff, &ok);

  if (!gtsb_ambig_tail(-0.0f, cutoff)) {
    *t_neg = -1.0f;
    for (uint32_t b = 0u; b <= 0x7f800000u; b += (1u << 12))
      if (gtsb_ambig_tail(-from_bits(b), cutoff)) ok = 0;
  }
  else if (g_ninf) {
    *t_neg = inf;
    for (uint32_t b = 0u; b <= 0x7f800000u; b += (1u << 12))
      if (!gtsb_ambig_tail(-from_bits(b), cutoff)) ok = 0;
  }
  else
    *t_neg = bisect_side(-1.0f, cutoff, &ok);
  return ok ? 0 : -1;
}
