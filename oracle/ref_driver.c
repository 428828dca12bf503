/* oracle/ref_driver.c -- TEST INFRASTRUCTURE, not product code.

   Array-level driver around the COMPILED, UNMODIFIED reference
   (oracle/_ref/libgtscaf_ref.so links this file with the reference's own
   gt_scaffolder_{graph,parser,algorithms}.o).  It lets the tests and the
   bench's cpu_baseline leg feed integer records to the reference's public
   functions and read the resulting GtScaffolderGraph back as flat arrays.

   The only logic restated here is the insert/dedup loop body of
   gt_scaffolder_parser_read_distances (reference parser.c:357-379), because
   the reference only exposes it behind a text parser whose 1024-byte line
   buffer (parser.c:30,323) cannot hold hub lines.  It is expressed entirely
   through the reference's own gt_scaffolder_graph_find_edge /
   gt_scaffolder_graph_alter_edge / gt_scaffolder_graph_add_edge, and
   tests/test_oracle.py checks it against gt_scaffolder_graph_new_from_file on
   generated text files. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include "core/ma_api.h"
#include "core/str_api.h"
#include "gt_scaffolder_graph.h"
#include "gt_scaffolder_algorithms.h"

/* non-static in graph.c:60 but not declared in graph.h */
GtScaffolderGraph *gt_scaffolder_graph_new(GtUword max_nof_vertices,
                                           GtUword max_nof_edges);

#define REC_SENSE 1u
#define REC_SAME  2u

static double now_sec(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/* Build a graph from integer vertices and file-ordered integer records.
   with_headers: give every vertex the header "c%010lu" (sorts in id order) so
   that .astat files and .dot printing work; costs one GtStr per vertex. */
GtScaffolderGraph *refdrv_build(uint64_t nof_vertices,
                                const uint64_t *seq_len,
                                const float *astat,
                                const float *copy_num,
                                uint64_t nof_records,
                                const uint32_t *root,
                                const uint32_t *ctg,
                                const int64_t *dist,
                                const float *std_dev,
                                const uint64_t *num_pairs,
                                const uint8_t *flags,
                                int with_headers,
                                double *seconds)
{
  GtScaffolderGraph *graph;
  GtScaffolderGraphVertex *v, *root_ctg, *ctg_v;
  GtScaffolderGraphEdge *edge;
  GtUword *capacity;
  uint64_t i;
  char name[32];
  double t0 = now_sec();

  graph = gt_scaffolder_graph_new(nof_vertices ? nof_vertices : 1,
                                  nof_records ? 2 * nof_records : 1);
  for (i = 0; i < nof_vertices; i++) {
    GtStr *hdr = NULL;
    if (with_headers) {
      snprintf(name, sizeof name, "c%010lu", (unsigned long) i);
      hdr = gt_str_new_cstr(name);
    }
    gt_scaffolder_graph_add_vertex(graph, hdr, seq_len[i], astat[i],
                                   copy_num[i]);
    if (!with_headers)
      graph->vertices[i].header_seq = NULL;
  }
  /* per-vertex edge-pointer capacity: upper bound as in parser.c:245-283 */
  capacity = gt_calloc(nof_vertices ? nof_vertices : 1, sizeof (*capacity));
  for (i = 0; i < nof_records; i++) {
    capacity[root[i]]++;
    capacity[ctg[i]]++;
  }
  for (i = 0; i < nof_vertices; i++) {
    v = graph->vertices + i;
    if (capacity[i] != 0)
      v->edges = gt_malloc(sizeof (*v->edges) * capacity[i]);
  }
  gt_free(capacity);

  /* parser.c:357-379, ismatepair == false (graph.c:399-400) */
  for (i = 0; i < nof_records; i++) {
    bool sense = (flags[i] & REC_SENSE) != 0, same = (flags[i] & REC_SAME) != 0;
    bool twin_dir;
    GtUword np = num_pairs ? num_pairs[i] : 0;
    root_ctg = graph->vertices + root[i];
    ctg_v = graph->vertices + ctg[i];
    edge = gt_scaffolder_graph_find_edge(root_ctg, ctg_v);
    if (edge != NULL) {
      if (edge->std_dev < std_dev[i])
        gt_scaffolder_graph_alter_edge(edge, dist[i], std_dev[i], np, sense,
                                       same);
    }
    else {
      twin_dir = same ? !sense : sense;
      gt_scaffolder_graph_add_edge(graph, root_ctg, ctg_v, dist[i],
                                   std_dev[i], np, sense, same);
      gt_scaffolder_graph_add_edge(graph, ctg_v, root_ctg, dist[i],
                                   std_dev[i], np, twin_dir, same);
    }
  }
  if (seconds) *seconds = now_sec() - t0;
  return graph;
}

uint64_t refdrv_nof_vertices(const GtScaffolderGraph *g)
{ return g->nof_vertices; }

uint64_t refdrv_nof_edges(const GtScaffolderGraph *g) { return g->nof_edges; }

/* any output pointer may be NULL */
void refdrv_get_vertices(const GtScaffolderGraph *g, uint64_t *seq_len,
                         float *astat, float *copy_num, uint8_t *state)
{
  GtUword i;
  for (i = 0; i < g->nof_vertices; i++) {
    if (seq_len) seq_len[i] = g->vertices[i].seq_len;
    if (astat) astat[i] = g->vertices[i].astat;
    if (copy_num) copy_num[i] = g->vertices[i].copy_num;
    if (state) state[i] = (uint8_t) g->vertices[i].state;
  }
}

/* edges in graph->edges[] order (= creation order, graph.c:295-303) */
void refdrv_get_edges(const GtScaffolderGraph *g, uint32_t *src, uint32_t *dst,
                      int64_t *dist, float *std_dev, uint64_t *num_pairs,
                      uint8_t *flags, uint8_t *state)
{
  GtUword i;
  for (i = 0; i < g->nof_edges; i++) {
    const GtScaffolderGraphEdge *e = g->edges + i;
    if (src) src[i] = (uint32_t) (e->start - g->vertices);
    if (dst) dst[i] = (uint32_t) (e->end - g->vertices);
    if (dist) dist[i] = e->dist;
    if (std_dev) std_dev[i] = e->std_dev;
    if (num_pairs) num_pairs[i] = e->num_pairs;
    if (flags) flags[i] = (e->sense ? REC_SENSE : 0) | (e->same ? REC_SAME : 0);
    if (state) state[i] = (uint8_t) e->state;
  }
}

/* adjacency order: row_ptr[V+1], eids[E] = index into graph->edges[] */
void refdrv_get_adjacency(const GtScaffolderGraph *g, uint64_t *row_ptr,
                          uint32_t *eids)
{
  GtUword i, k, pos = 0;
  for (i = 0; i < g->nof_vertices; i++) {
    row_ptr[i] = pos;
    for (k = 0; k < g->vertices[i].nof_edges; k++)
      eids[pos++] = (uint32_t) (g->vertices[i].edges[k] - g->edges);
  }
  row_ptr[g->nof_vertices] = pos;
}

void refdrv_set_states(GtScaffolderGraph *g, const uint8_t *vstate,
                       const uint8_t *estate)
{
  GtUword i;
  if (vstate)
    for (i = 0; i < g->nof_vertices; i++)
      g->vertices[i].state = (GraphItemState) vstate[i];
  if (estate)
    for (i = 0; i < g->nof_edges; i++)
      g->edges[i].state = (GraphItemState) estate[i];
}

/* astat_filename "" => A-statistic clause only (algorithms.c:108,164); the
   path of an EMPTY file => both clauses on the in-memory values. */
int refdrv_mark_repeats(GtScaffolderGraph *g, const char *astat_filename,
                        float copy_num_cutoff, float astat_cutoff,
                        double *seconds)
{
  GtError *err = gt_error_new();
  double t0 = now_sec();
  int rc = gt_scaffolder_graph_mark_repeats(astat_filename, g, copy_num_cutoff,
                                            astat_cutoff, err);
  if (seconds) *seconds = now_sec() - t0;
  if (rc != 0) fprintf(stderr, "refdrv: %s\n", gt_error_get(err));
  gt_error_delete(err);
  return rc;
}

void refdrv_filter(GtScaffolderGraph *g, float pcutoff, float cncutoff,
                   int64_t ocutoff, double *seconds)
{
  double t0 = now_sec();
  gt_scaffolder_graph_filter(g, pcutoff, cncutoff, ocutoff);
  if (seconds) *seconds = now_sec() - t0;
}

/* text front door: the reference's own parser + constructor */
GtScaffolderGraph *refdrv_new_from_file(const char *ctg_filename,
                                        uint64_t min_ctg_len,
                                        const char *dist_filename)
{
  GtScaffolderGraph *g = NULL;
  GtError *err = gt_error_new();
  if (gt_scaffolder_graph_new_from_file(&g, ctg_filename, min_ctg_len,
                                        dist_filename, false, err) != 0) {
    fprintf(stderr, "refdrv: %s\n", gt_error_get(err));
    g = NULL;
  }
  gt_error_delete(err);
  return g;
}

int refdrv_print(const GtScaffolderGraph *g, const char *filename)
{
  GtError *err = gt_error_new();
  int rc = gt_scaffolder_graph_print(g, filename, err);
  gt_error_delete(err);
  return rc;
}

/* graph.c:310-343; defined there, declared in no header */
void gt_scaffolder_graph_print_scaffold(const GtScaffolderGraph *g, GtFile *f);

int refdrv_print_scaffold(const GtScaffolderGraph *g, const char *filename)
{
  GtError *err = gt_error_new();
  GtFile *f = gt_file_new(filename, "w", err);
  if (f != NULL) {
    gt_scaffolder_graph_print_scaffold(g, f);
    gt_file_delete(f);
  }
  gt_error_delete(err);
  return f != NULL ? 0 : -1;
}

/* downstream host stages (out of the hot path) for the .dot/.scaf checks */
void refdrv_removecycles(GtScaffolderGraph *g) { gt_scaffolder_removecycles(g); }
void refdrv_makescaffold(GtScaffolderGraph *g) { gt_scaffolder_makescaffold(g); }

int refdrv_write_scaffold(GtScaffolderGraph *g, const char *filename)
{
  GtError *err = gt_error_new();
  GtAssemblyStatsCalculator *stats = gt_assembly_stats_calculator_new();
  GtArray *recs = gt_scaffolder_graph_iterate_scaffolds(g, stats);
  GtUword i;
  int rc = gt_scaffolder_graph_write_scaffold(recs, filename, err);
  for (i = 0; i < gt_array_size(recs); i++)
    gt_scaffolder_graph_record_delete(
      *(GtScaffolderGraphRecord **) gt_array_get(recs, i));
  gt_array_delete(recs);
  gt_assembly_stats_calculator_delete(stats);
  gt_error_delete(err);
  return rc;
}

/* gt_scaffolder_calc_cc_and_terminals (algorithms.c:379-436; not in the header, external linkage):
   the components in the order the reference finds them, as arrays.  cc_off[n_cc + 1] cuts
   terminals[] (vertex ids, in the order the search meets them); returns the number of
   components, or -1 when a capacity is too small.  Leaves the vertex states as the reference
   does (every unmarked vertex GIS_VISITED). */
void gt_scaffolder_calc_cc_and_terminals(const GtScaffolderGraph *graph, GtArray *ccs);

int64_t refdrv_calc_cc(GtScaffolderGraph *g, uint64_t *cc_off, uint64_t cc_cap, uint32_t *terminals,
                       uint64_t term_cap)
{
  GtArray *ccs = gt_array_new(sizeof (GtArray *));
  GtUword i, j;
  uint64_t k = 0;
  int64_t n;
  gt_scaffolder_calc_cc_and_terminals(g, ccs);
  n = (int64_t) gt_array_size(ccs);
  for (i = 0; i < gt_array_size(ccs); i++) {
    GtArray *t = *(GtArray **) gt_array_get(ccs, i);
    if (n >= 0 && i < cc_cap)
      cc_off[i] = k;
    else
      n = -1;
    for (j = 0; j < gt_array_size(t); j++, k++) {
      if (n >= 0 && k < term_cap)
        terminals[k] = (uint32_t) (*(GtScaffolderGraphVertex **) gt_array_get(t, j) - g->vertices);
      else
        n = -1;
    }
    gt_array_delete(t);
  }
  if (n >= 0 && (uint64_t) n < cc_cap)
    cc_off[n] = k;
  else
    n = -1;
  gt_array_delete(ccs);
  return n;
}

void refdrv_delete(GtScaffolderGraph *g)
{
  /* gt_scaffolder_graph_delete calls gt_str_delete on every header;
     gtcompat's accepts NULL (headerless vertices) */
  gt_scaffolder_graph_delete(g);
}
