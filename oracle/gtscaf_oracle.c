/* oracle/gtscaf_oracle.c -- TEST INFRASTRUCTURE, not product code.

   CPU restatement ("port") of the reference's hot path on flat arrays with
   integer vertex ids instead of pointers.  Sequential, single-threaded, in the
   reference's own order; it exists to be OBVIOUSLY equal to the reference, not
   to be fast.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
   load it.

   Parity status: PINNED.  tests/test_oracle.py compares every function here
   with the compiled, unmodified reference (oracle/_ref/libgtscaf_ref.so) on the
   reference's golden testdata (testdata/libPE.*) and on randomized graphs that
   exercise the pairwise filter, which the reference's own goldens do not
   (SURVEY.md section 8c).

   Compile with FMA contraction off (-ffp-contract=off) and no -ffast-math: the
   reference is built for baseline x86-64, src/Makefile:7. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "gtscaf_oracle.h"

static void *xcalloc(size_t n, size_t sz)
{
  void *p = calloc(n ? n : 1, sz);
  if (p == NULL) abort();
  return p;
}

/* algorithms.c:38-47 */
static int vertex_is_marked(const OraGraph *g, uint32_t v)
{
  uint8_t s = g->vstate[v];
  return s == ORA_POLYMORPHIC || s == ORA_REPEAT || s == ORA_CYCLIC;
}

/* algorithms.c:50-58 */
static int edge_is_marked(const OraGraph *g, uint32_t e)
{
  uint8_t s = g->estate[e];
  return s == ORA_INCONSISTENT || s == ORA_POLYMORPHIC || s == ORA_CYCLIC ||
         s == ORA_REPEAT;
}

/* algorithms.c:61-73: the edge, and every edge of its end vertex that points
   back at its start vertex */
static void mark_edge(OraGraph *g, uint32_t e, uint8_t state)
{
  uint32_t end = g->dst[e], start = g->src[e];
  uint64_t k;
  g->estate[e] = state;
  for (k = 0; k < g->nof_vedges[end]; k++)
    if (g->dst[g->vedges[end][k]] == start)
      g->estate[g->vedges[end][k]] = state;
}

/* algorithms.c:76-87 */
static void mark_vertex(OraGraph *g, uint32_t v, uint8_t state)
{
  uint64_t k;
  g->vstate[v] = state;
  for (k = 0; k < g->nof_vedges[v]; k++)
    mark_edge(g, g->vedges[v][k], state);
}

/* graph.c:173-184 */
static int64_t find_edge(const OraGraph *g, uint32_t v1, uint32_t v2)
{
  uint64_t k;
  for (k = 0; k < g->nof_vedges[v1]; k++)
    if (g->dst[g->vedges[v1][k]] == v2)
      return (int64_t) g->vedges[v1][k];
  return -1;
}

/* graph.c:137-170 */
static void add_edge(OraGraph *g, uint32_t vstart, uint32_t vend, int64_t dist,
                     float std_dev, uint64_t num_pairs, int dir, int same,
                     int64_t rec)
{
  uint64_t e = g->nof_edges;
  if (e >= g->max_nof_edges) abort();
  g->src[e] = vstart;
  g->dst[e] = vend;
  g->dist[e] = dist;
  g->std_dev[e] = std_dev;
  g->num_pairs[e] = num_pairs;
  g->flags[e] = (uint8_t) ((dir ? ORA_SENSE : 0) | (same ? ORA_SAME : 0));
  g->estate[e] = ORA_UNVISITED;
  g->win_rec[e] = rec;
  g->vedges[vstart][g->nof_vedges[vstart]++] = (uint32_t) e;
  g->nof_edges++;
}

/* Vertices as given (ids = rank of header, parser.c:172) + the record loop of
   gt_scaffolder_parser_read_distances, parser.c:357-379, with
   ismatepair == false (graph.c:399-400). */
OraGraph *ora_build(uint64_t V, const uint64_t *seq_len, const float *astat,
                    const float *copy_num, uint64_t R, const uint32_t *root,
                    const uint32_t *ctg, const int64_t *dist,
                    const float *std_dev, const uint64_t *num_pairs,
                    const uint8_t *flags)
{
  OraGraph *g = xcalloc(1, sizeof *g);
  uint64_t i, *cap;

  g->nof_vertices = V;
  g->seq_len = xcalloc(V, sizeof *g->seq_len);
  g->astat = xcalloc(V, sizeof *g->astat);
  g->copy_num = xcalloc(V, sizeof *g->copy_num);
  g->vstate = xcalloc(V, 1);                 /* graph.c:129 GIS_UNVISITED */
  g->nof_vedges = xcalloc(V, sizeof *g->nof_vedges);
  g->vedges = xcalloc(V, sizeof *g->vedges);
  memcpy(g->seq_len, seq_len, V * sizeof *seq_len);
  memcpy(g->astat, astat, V * sizeof *astat);
  memcpy(g->copy_num, copy_num, V * sizeof *copy_num);

  /* per-vertex capacity upper bound, parser.c:245-283 */
  cap = xcalloc(V, sizeof *cap);
  for (i = 0; i < R; i++) { cap[root[i]]++; cap[ctg[i]]++; }
  for (i = 0; i < V; i++)
    if (cap[i]) g->vedges[i] = xcalloc(cap[i], sizeof **g->vedges);
  free(cap);

  g->max_nof_edges = 2 * R;
  g->src = xcalloc(2 * R, sizeof *g->src);
  g->dst = xcalloc(2 * R, sizeof *g->dst);
  g->dist = xcalloc(2 * R, sizeof *g->dist);
  g->std_dev = xcalloc(2 * R, sizeof *g->std_dev);
  g->num_pairs = xcalloc(2 * R, sizeof *g->num_pairs);
  g->flags = xcalloc(2 * R, 1);
  g->estate = xcalloc(2 * R, 1);
  g->win_rec = xcalloc(2 * R, sizeof *g->win_rec);

  for (i = 0; i < R; i++) {
    int sense = (flags[i] & ORA_SENSE) != 0, same = (flags[i] & ORA_SAME) != 0;
    uint64_t np = num_pairs ? num_pairs[i] : 0;
    int64_t e = find_edge(g, root[i], ctg[i]);          /* parser.c:359 */
    if (e >= 0) {
      if (g->std_dev[e] < std_dev[i]) {                  /* parser.c:362 */
        g->dist[e] = dist[i];                            /* graph.c:230-234 */
        g->std_dev[e] = std_dev[i];
        g->num_pairs[e] = np;
        g->flags[e] = (uint8_t) ((sense ? ORA_SENSE : 0) |
                                 (same ? ORA_SAME : 0));
        g->win_rec[e] = (int64_t) i;
      }
    }
    else {
      int twin_dir = same ? !sense : sense;              /* parser.c:369-372 */
      add_edge(g, root[i], ctg[i], dist[i], std_dev[i], np, sense, same,
               (int64_t) i);
      add_edge(g, ctg[i], root[i], dist[i], std_dev[i], np, twin_dir, same,
               (int64_t) i);
    }
  }
  return g;
}

/* algorithms.c:160-166; use_copy_num <=> strlen(filename) != 0 */
void ora_mark_repeats(OraGraph *g, int use_copy_num, float copy_num_cutoff,
                      float astat_cutoff)
{
  uint64_t v;
  for (v = 0; v < g->nof_vertices; v++)
    if (g->astat[v] <= astat_cutoff ||
        (use_copy_num && g->copy_num[v] < copy_num_cutoff))
      mark_vertex(g, (uint32_t) v, ORA_REPEAT);
}

/* algorithms.c:174-193, with C's implicit conversions written out:
   float = long - long; float arithmetic for the variance; double for the
   quotient, sqrt and erf; each assignment narrows back to float. */
int ora_ambiguous_interval(float interval, float cutoff)
{
  float prob12, prob21, p_wrong;                     /* algorithms.c:187-192 */
  prob12 = (float) (0.5 * (1 + erf(interval)));
  prob21 = (float) (1.0 - prob12);
  p_wrong = (float) (1.0 - (prob12 > prob21 ? prob12 : prob21));
  return p_wrong > cutoff;
}

/* the same for n intervals (tests of the device's threshold form of this decision) */
void ora_ambiguous_intervals(const float *interval, uint64_t n, float cutoff, uint8_t *out)
{
  uint64_t i;
  for (i = 0; i < n; i++)
    out[i] = (uint8_t) ora_ambiguous_interval(interval[i], cutoff);
}

int ora_ambiguousorder(int64_t dist1, float std1, int64_t dist2, float std2,
                       float cutoff)
{
  float expval, variance, interval;
  expval = (float) (dist1 - dist2);
  variance = 2 * ((std1 * std1) + (std2 * std2));
  interval = (float) ((0 - expval) / sqrt(variance));
  return ora_ambiguous_interval(interval, cutoff);
}

void ora_ambiguousorders(const int64_t *dist1, const float *std1, const int64_t *dist2, const float *std2,
                         uint64_t n, float cutoff, uint8_t *out)
{
  uint64_t i;
  for (i = 0; i < n; i++)
    out[i] = (uint8_t) ora_ambiguousorder(dist1[i], std1[i], dist2[i], std2[i], cutoff);
}

/* algorithms.c:197-220; the mixed GtWord + GtUword sums wrap mod 2^64 */
int64_t ora_overlap(int64_t dist1, uint64_t len1, int64_t dist2, uint64_t len2)
{
  int64_t overlap = 0, start1 = dist1, start2 = dist2;
  int64_t end1 = (int64_t) ((uint64_t) dist1 + len1 - 1);
  int64_t end2 = (int64_t) ((uint64_t) dist2 + len2 - 1);
  if (start2 <= end1 && start1 <= end2) {
    int64_t is = start1 > start2 ? start1 : start2;
    int64_t ie = end1 < end2 ? end1 : end2;
    overlap = ie - is + 1;
  }
  return overlap;
}

/* algorithms.c:223-246 */
static void check_mark_polymorphic(OraGraph *g, uint32_t e1, uint32_t e2,
                                   float pcutoff, float cncutoff)
{
  float cn1 = g->copy_num[g->dst[e1]], cn2 = g->copy_num[g->dst[e2]];
  if (ora_ambiguousorder(g->dist[e1], g->std_dev[e1], g->dist[e2],
                         g->std_dev[e2], pcutoff) &&
      (cn1 + cn2) < cncutoff) {
    uint32_t poly = cn1 < cn2 ? g->dst[e1] : g->dst[e2];
    if (!vertex_is_marked(g, poly))
      mark_vertex(g, poly, ORA_POLYMORPHIC);
  }
}

/* algorithms.c:249-258 */
static void mark_edges_in_twin_dir(OraGraph *g, uint32_t v, int sense)
{
  uint64_t k;
  for (k = 0; k < g->nof_vedges[v]; k++) {
    uint32_t e = g->vedges[v][k];
    if (((g->flags[e] & ORA_SENSE) != 0) == (sense != 0))
      g->estate[e] = ORA_INCONSISTENT;
  }
}

/* algorithms.c:261-343, vertex by vertex in index order */
void ora_filter(OraGraph *g, float pcutoff, float cncutoff, int64_t ocutoff)
{
  uint64_t v, k1, k2;
  for (v = 0; v < g->nof_vertices; v++) {
    const uint32_t *adj = g->vedges[v];
    uint64_t deg = g->nof_vedges[v];
    int64_t sense_max = 0, antisense_max = 0;

    if (vertex_is_marked(g, (uint32_t) v)) continue;          /* :279 */

    for (k1 = 0; k1 < deg; k1++)                              /* :283-295 */
      for (k2 = k1 + 1; k2 < deg; k2++)
        if ((g->flags[adj[k1]] & ORA_SENSE) == (g->flags[adj[k2]] & ORA_SENSE))
          check_mark_polymorphic(g, adj[k1], adj[k2], pcutoff, cncutoff);

    if (vertex_is_marked(g, (uint32_t) v)) continue;          /* :298 */

    for (k1 = 0; k1 < deg; k1++)                              /* :304-320 */
      for (k2 = k1 + 1; k2 < deg; k2++) {
        uint32_t e1 = adj[k1], e2 = adj[k2];
        if ((g->flags[e1] & ORA_SENSE) == (g->flags[e2] & ORA_SENSE) &&
            !edge_is_marked(g, e1) && !edge_is_marked(g, e2)) {
          int64_t ov = ora_overlap(g->dist[e1], g->seq_len[g->dst[e1]],
                                   g->dist[e2], g->seq_len[g->dst[e2]]);
          if ((g->flags[e1] & ORA_SENSE) && ov > sense_max) sense_max = ov;
          if (!(g->flags[e1] & ORA_SENSE) && ov > antisense_max)
            antisense_max = ov;
        }
      }

    if (sense_max > ocutoff || antisense_max > ocutoff)       /* :324-341 */
      for (k1 = 0; k1 < deg; k1++) {
        uint32_t e = adj[k1];
        int sense = (g->flags[e] & ORA_SENSE) != 0;
        int same = (g->flags[e] & ORA_SAME) != 0;
        if (sense_max > ocutoff && sense) {
          g->estate[e] = ORA_INCONSISTENT;
          mark_edges_in_twin_dir(g, g->dst[e], !same);
        }
        if (antisense_max > ocutoff && !sense) {
          g->estate[e] = ORA_INCONSISTENT;
          mark_edges_in_twin_dir(g, g->dst[e], same);
        }
      }
  }
}

void ora_get_adjacency(const OraGraph *g, uint64_t *row_ptr, uint32_t *eids)
{
  uint64_t v, k, pos = 0;
  for (v = 0; v < g->nof_vertices; v++) {
    row_ptr[v] = pos;
    for (k = 0; k < g->nof_vedges[v]; k++) eids[pos++] = g->vedges[v][k];
  }
  row_ptr[g->nof_vertices] = pos;
}

void ora_delete(OraGraph *g)
{
  uint64_t v;
  if (g == NULL) return;
  for (v = 0; v < g->nof_vertices; v++) free(g->vedges[v]);
  free(g->vedges); free(g->nof_vedges); free(g->vstate); free(g->copy_num);
  free(g->astat); free(g->seq_len); free(g->src); free(g->dst); free(g->dist);
  free(g->std_dev); free(g->num_pairs); free(g->flags); free(g->estate);
  free(g->win_rec); free(g);
}
