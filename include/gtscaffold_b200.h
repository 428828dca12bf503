/* gtscaffold_b200.h -- C ABI of the B200-native scaffold-graph hot path.
 *
 * Array-level entry points: plain pointers and sizes, no GenomeTools and no
 * torch types.  They are what the reference's three hot-path functions bind to
 * (the binding itself is integration/gt_scaffolder_b200.c):
 *
 *   gtsb_build         <- the insert/dedup loop of
 *                         gt_scaffolder_parser_read_distances
 *                         (gt_scaffolder_parser.c:357-379) with
 *                         gt_scaffolder_graph_add_edge / find_edge / alter_edge
 *                         (gt_scaffolder_graph.c:137-184, 219-235)
 *   gtsb_mark_repeats  <- the marking loop of gt_scaffolder_graph_mark_repeats
 *                         (gt_scaffolder_algorithms.c:160-166, 61-87)
 *   gtsb_filter        <- gt_scaffolder_graph_filter
 *                         (gt_scaffolder_algorithms.c:261-343, 174-258)
 *
 * and, on either side of that path,
 *
 *   gtsb_parse_de_host     <- the .de record loop (gt_scaffolder_parser.c:323-388)
 *   gtsb_parse_astat_host  <- the .astat line loop (gt_scaffolder_algorithms.c:118-149)
 *   gtsb_dot_*_lines_host  <- gt_scaffolder_graph_print_generic / _print_scaffold
 *                             (gt_scaffolder_graph.c:269-343)
 *
 * All functions return 0 on success and -1 on error (message: gtsb_error), the
 * reference's own convention (gt_scaffolder_algorithms.c:112-113).  There is no
 * CPU fallback: without a CUDA device every compute call fails.
 *
 * Vertex ids are the reference's: rank of the contig header in strcmp order
 * (gt_scaffolder_parser.c:172).  Records are in .de file order.
 */
#ifndef GTSCAFFOLD_B200_H
#define GTSCAFFOLD_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* GraphItemState values (gt_scaffolder_graph.h:29-31) */
enum { GTSB_UNVISITED = 0, GTSB_POLYMORPHIC = 1, GTSB_INCONSISTENT = 2, GTSB_REPEAT = 3,
       GTSB_VISITED = 4, GTSB_PROCESSED = 5, GTSB_SCAFFOLD = 6, GTSB_CYCLIC = 7 };

/* record / edge flag bits */
#define GTSB_SENSE  1u   /* edge->sense  (record: before the ';', parser.c:382) */
#define GTSB_SAME   2u   /* edge->same   (record: '+' suffix, parser.c:347)      */
#define GTSB_RSENSE 4u   /* CSR slots only: sense of the reverse edge dst->src   */
#define GTSB_RSAME  8u   /* CSR slots only: same of the reverse edge             */

#define GTSB_WIN_SEEDED 0x80000000u  /* win_rec: attributes are the twin seed of that record */
#define GTSB_MAX_VERTICES ((1u << 27) - 1u)

typedef struct gtsb_context gtsb_context;

typedef struct gtsb_stats {
  uint64_t nof_vertices, nof_records, nof_edges;
  uint32_t max_degree, big_rows, large_buckets;
  uint32_t proposals, poly_sweeps, fire_rounds;
  uint32_t line_ordered_build;     /* 1: last gtsb_build took the line-ordered fast path */
  uint32_t fallback_reason;        /* why not (bit mask, csrc/gtsb_kernels.h FB_*), else 0 */
  uint64_t kernel_launches;        /* kernels launched by this context so far */
  float ms_build, ms_mark_repeats, ms_filter;   /* device time of the last call of each stage */
} gtsb_stats;

int gtsb_create(gtsb_context **ctx, int device);
void gtsb_destroy(gtsb_context *ctx);
const char *gtsb_error(const gtsb_context *ctx);
/* run on a caller-owned CUDA stream (cudaStream_t) instead of the context's own */
int gtsb_set_stream(gtsb_context *ctx, void *cuda_stream);
/* also record, per edge, which record's attributes it carries (needed to fill
   GtScaffolderGraphEdge.num_pairs); costs 8 B/edge of extra traffic */
int gtsb_want_win_rec(gtsb_context *ctx, int on);
/* testing aid: always take the general (any-input) build path */
int gtsb_force_general_build(gtsb_context *ctx, int on);

/* ---- inputs.  *_host variants copy from host memory (H2D on the context's
   stream or on its copy stream, asynchronously): the host buffers must stay valid and
   unmodified until the next gtsb_build / gtsb_pipeline / gtsb_synchronize /
   gtsb_get_* call has returned (pageable memory is staged before the call returns;
   PINNED memory is read by the DMA engine later).  *_device variants adopt device
   pointers that must stay valid. */
int gtsb_set_vertices_host(gtsb_context *ctx, uint64_t nof_vertices, const uint32_t *seq_len,
                           const float *astat, const float *copy_num);
int gtsb_set_vertices_device(gtsb_context *ctx, uint64_t nof_vertices, const uint32_t *seq_len,
                             const float *astat, const float *copy_num);
/* Partitioned graph (after gtsb_dist_init): this rank uploads only the attributes of the contigs
   [first, first + count) -- the arrays hold `count` values -- and gtsb_pipeline gathers the other
   ranks' slices over NVLink instead of every rank pushing all of them through the host's PCIe.
   The ranks' slices must tile [0, nof_vertices) in rank order. */
int gtsb_set_vertices_slice_host(gtsb_context *ctx, uint64_t nof_vertices, uint64_t first, uint64_t count,
                                 const uint32_t *seq_len, const float *astat, const float *copy_num);
int gtsb_set_records_host(gtsb_context *ctx, uint64_t nof_records, const uint32_t *root,
                          const uint32_t *ctg, const int32_t *dist, const float *std_dev,
                          const uint8_t *flags);
int gtsb_set_records_device(gtsb_context *ctx, uint64_t nof_records, const uint32_t *root,
                            const uint32_t *ctg, const int32_t *dist, const float *std_dev,
                            const uint8_t *flags);
/* the same records in the shape a .de parser has them in (parser.c:323-388: one
   line per root contig): line l holds records [line_start[l], line_start[l+1])
   of root line_root[l].  Saves the 4 B/record of the repeated root on the bus. */
int gtsb_set_record_lines_host(gtsb_context *ctx, uint64_t nof_lines, const uint32_t *line_root,
                               const uint32_t *line_start, uint64_t nof_records, const uint32_t *ctg,
                               const int32_t *dist, const float *std_dev, const uint8_t *flags);
/* the same with device pointers that must stay valid (what gtsb_parse_de_host leaves behind,
   or a caller that tokenises on the GPU itself) */
int gtsb_set_record_lines_device(gtsb_context *ctx, uint64_t nof_lines, const uint32_t *line_root,
                                 const uint32_t *line_start, uint64_t nof_records, const uint32_t *ctg,
                                 const int32_t *dist, const float *std_dev, const uint8_t *flags);
/* an already-built graph (host CSR, flags incl. GTSB_RSENSE/RSAME, states);
   used by the GtScaffolderGraph binding of mark_repeats / filter */
int gtsb_set_graph_host(gtsb_context *ctx, uint64_t nof_vertices, uint64_t nof_edges,
                        const uint32_t *row_ptr, const uint32_t *dst, const int32_t *dist,
                        const float *std_dev, const uint8_t *flags, const uint32_t *seq_len,
                        const float *astat, const float *copy_num, const uint8_t *vstate,
                        const uint8_t *estate);

/* ---- the resident graph between the three calls of the reference's driver (test.c:130-146):
   gtsb_build leaves the graph on the device; mark_repeats and filter then need only what the host
   may have changed in between.  gtsb_update_vertices_host refreshes the per-vertex attributes
   (the .astat values gt_scaffolder_graph_mark_repeats reads, algorithms.c:140-141) and keeps
   the graph and its states; gtsb_set_states_host overwrites the states (vstate[V]; edge states
   in graph->edges[] order) when host code marked something itself.  Single-device graphs built
   by gtsb_build. */
int gtsb_update_vertices_host(gtsb_context *ctx, uint64_t nof_vertices, const uint32_t *seq_len,
                              const float *astat, const float *copy_num);
int gtsb_set_states_host(gtsb_context *ctx, const uint8_t *vstate, const uint8_t *estate_by_eid);

/* ---- .de text on the device: the record loop of gt_scaffolder_parser_read_distances
   (gt_scaffolder_parser.c:323-388) -- 1024-byte fgets pieces, last character dropped,
   ' '-separated tokens, "%[^>,],%ld,%ld,%f" records, ';' switching the direction, unknown
   roots skipping the line and unknown partners the record.  Header -> vertex id is
   gt_scaffolder_graph_get_vertex (gt_scaffolder_graph.c:187-216) as a device hash table
   over the headers handed in once: names = the headers of vertex 0..V-1 back to back
   (no terminators), name_off[V+1] their offsets.
   gtsb_parse_de_host leaves the records in the context as gtsb_set_records_* would
   (gtsb_build follows) and keeps their pair counts for gtsb_get_records.  It accepts
   the canonical spelling of records only (csrc/gtsb_parse_core.h); for any other text
   it returns 0 with *irregular != 0 (GTSB_IRR_* bits) and NO records set -- the caller
   then tokenises on the host with the C library's sscanf, as the reference does. */
#define GTSB_IRR_NUL      1u   /* NUL byte in the text */
#define GTSB_IRR_TOKEN    2u   /* "header,..." token that is not a canonical record */
#define GTSB_IRR_RANGE    4u   /* distance outside int32 / pair count outside uint32 */
#define GTSB_IRR_FLOAT    8u   /* a value the device cannot prove it rounds as strtof does */
#define GTSB_IRR_DUP_NAME 16u  /* two contigs with one header (bsearch's choice is unspecified) */
int gtsb_set_vertex_names_host(gtsb_context *ctx, uint64_t nof_vertices, const char *names,
                               const uint64_t *name_off);
int gtsb_parse_de_host(gtsb_context *ctx, const char *text, uint64_t text_bytes,
                       uint64_t *nof_records, uint32_t *irregular);
/* .astat text: the line loop of gt_scaffolder_graph_mark_repeats
   (gt_scaffolder_algorithms.c:118-149, sscanf "%s\t%ld\t%ld\t%ld\t%f\t%f" == 6 per line).
   astat / copy_num: one value per vertex of the names set, in and out -- the entries of
   contigs the text names are overwritten (last line wins, as the sequential loop has it),
   the others keep their value.  Canonical lines only (single tabs, plain decimals);
   *irregular != 0: nothing was touched, read the file on the host -- that includes every
   text the reference would reject with "Invalid record". */
int gtsb_parse_astat_host(gtsb_context *ctx, const char *text, uint64_t text_bytes, float *astat,
                          float *copy_num, uint32_t *irregular);
/* the records held by the context, file order (NULL pointers are skipped); num_pairs
   only after gtsb_parse_de_host */
int gtsb_get_records(gtsb_context *ctx, uint32_t *root, uint32_t *ctg, int32_t *dist,
                     float *std_dev, uint8_t *flags, uint32_t *num_pairs);

/* ---- .dot text from graph arrays: the lines gt_scaffolder_graph_print_generic and
   gt_scaffolder_graph_print_scaffold write (gt_scaffolder_graph.c:269-343) between
   "digraph {\n" and "}\n", in item order.  scaffold_only != 0: the print_scaffold lines
   (items in state GTSB_SCAFFOLD only, no colour).  Vertices [first, first + count) of the
   names set (gtsb_set_vertex_names_host) with vstate[count]; edges as parallel arrays in
   graph->edges[] order.  At most 2^25 items per call; out must hold the text (bytes <= cap
   is checked; 64 + header bytes per vertex and 105 bytes per edge always suffice). */
int gtsb_dot_vertex_lines_host(gtsb_context *ctx, int scaffold_only, uint64_t first, uint64_t count,
                               const uint8_t *vstate, char *out, uint64_t cap, uint64_t *bytes);
int gtsb_dot_edge_lines_host(gtsb_context *ctx, int scaffold_only, uint64_t count, const uint32_t *src,
                             const uint32_t *dst, const int32_t *dist, const uint8_t *estate,
                             const uint8_t *sense, char *out, uint64_t cap, uint64_t *bytes);

/* ---- `.scaf` text: what gt_scaffolder_graph_write_scaffold prints (gt_scaffolder_algorithms.c:
   1000-1042), one line per scaffold record -- the root's header, then per edge
   "\t<end header>,<dist %ld>,<std_dev %f>,<sense %d>,<same %d>," -- with the "%f" produced by
   integer arithmetic on the float's bits (correctly rounded, as glibc's printf).  Records as
   flat arrays: rec_root[i] = vertex id of the root, its edges are [rec_edge_off[i],
   rec_edge_off[i+1]) of the edge arrays (edge_end = vertex id of edge->end, edge_flags bit 0 =
   sense, bit 1 = same).  Vertex ids index the names set (gtsb_set_vertex_names_host).  At most
   2^25 records and 2^25 edges per call; 80 + header bytes per edge and 1 + header bytes per
   record always suffice for out. */
int gtsb_scaf_lines_host(gtsb_context *ctx, uint64_t nof_records, const uint32_t *rec_root,
                         const uint64_t *rec_edge_off, const uint32_t *edge_end, const int64_t *edge_dist,
                         const float *edge_std_dev, const uint8_t *edge_flags, char *out, uint64_t cap,
                         uint64_t *bytes);

/* ---- the hot path */
int gtsb_build(gtsb_context *ctx);
int gtsb_mark_repeats(gtsb_context *ctx, float copy_num_cutoff, float astat_cutoff,
                      int use_copy_num /* strlen(filename) != 0, algorithms.c:164 */);
int gtsb_filter(gtsb_context *ctx, float pcutoff, float cncutoff, int64_t ocutoff);
/* build + mark_repeats + filter back to back */
int gtsb_pipeline(gtsb_context *ctx, float copy_num_cutoff, float astat_cutoff, int use_copy_num,
                  float pcutoff, float cncutoff, int64_t ocutoff);

/* ---- one graph partitioned over the GPUs of a box (one process per GPU).
   Rank r is given a contiguous chunk of the .de records (whole lines, ranks in
   file order) with gtsb_set_records_*, and the attributes of ALL contigs with
   gtsb_set_vertices_*; gtsb_pipeline then builds and filters the global graph,
   each rank holding the rows of the contigs that head its lines (contigs
   without a line: last rank).  Vertex states come back complete on every rank;
   edges stay with their rank (gtsb_get_edges).  The exchanges are NCCL over
   NVLink; the 128-byte id from rank 0 has to reach every rank by the caller's
   own means (MPI, torch.distributed, a file).
   The partitioned build is the line-ordered one only: a contig heading lines on two ranks, a
   line of more than 64 records, or a link listed only on the later of its two lines fails
   gtsb_pipeline on EVERY rank with the reason mask in the message (a single device rebuilds such
   input with its general sort-based path; the partitioned build has none).  An error on one
   rank -- also an allocation failure -- reaches all ranks at the next agreed exchange. */
int gtsb_dist_unique_id(void *id128);
int gtsb_dist_init(gtsb_context *ctx, int rank, int world, const void *id128);
/* this device's edges with the reference's vertex ids: eid = index into
   graph->edges[] (creation order; within a vertex, adjacency order = eid order) */
int gtsb_get_edges(gtsb_context *ctx, uint64_t *nof_edges, uint32_t *eid, uint32_t *src, uint32_t *dst,
                   int32_t *dist, float *std_dev, uint8_t *flags, uint8_t *estate);

/* ---- distance estimates between contig pairs from read-pair fragments: estimate_dist_using_mle of
   gt_scaffolder_bamparser.c (:385-598) for a batch of contig pairs.  Pair p owns the fragments
   [frag_off[p], frag_off[p+1]) -- (start, end) as calculate_fragment (:600-661) stores them --
   with its FragmentData.ma and the lengths of the two contigs; pmf / pmf_nof / minp are PmfData's
   dist / nof / minp, rf the library orientation, min_dist / max_dist the scan range.  Out: dist[p]
   and pairs_used[p] (the reference's *dist and *nof_pairs).  The likelihood scan runs on the
   device in the reference's summation order; both logarithms are the host C library's
   (csrc/gtsb_mle_core.h).  A negative probability anywhere in the distribution is the reference's
   "negative probability" error. */
int gtsb_mle_host(gtsb_context *ctx, uint64_t nof_pairs, const uint64_t *frag_off, const int64_t *frag_start,
                  const int64_t *frag_end, const uint64_t *ma, const uint64_t *len_ref, const uint64_t *len_mref,
                  const double *pmf, uint64_t pmf_nof, double minp, int rf, int64_t min_dist, int64_t max_dist,
                  int64_t *dist, uint64_t *pairs_used);

/* ---- components and terminal vertices of the current graph and states: the facts
   gt_scaffolder_calc_cc_and_terminals (gt_scaffolder_algorithms.c:379-436) derives by breadth-first
   search before removecycles and makescaffold walk the graph.  label[v] = the smallest vertex id
   from which v is reached along unmarked edges between unmarked vertices (v itself included) --
   the root of the reference's component of v; 0xFFFFFFFF for a marked vertex.  terminal[v] =
   gt_scaffolder_graph_isterminal (:346-373): the unmarked edges of v do not point both ways.
   Single-device graphs. */
int gtsb_components(gtsb_context *ctx, uint32_t *label, uint8_t *terminal);

/* order-independent digests of the result, for comparing runs too large to fetch: out[0] = edges
   on this device, out[1] = sum over them of a 64-bit hash of (eid, src, dst, dist, std_dev bits,
   flags & 15, state), out[2] = sum over ALL vertices of a hash of (id, state).  The ranks' out[0]
   and out[1] of a partitioned graph add up (mod 2^64) to those of the same graph on one device. */
int gtsb_result_digest(gtsb_context *ctx, uint64_t out[3]);

/* ---- results (device-resident until fetched; NULL pointers are skipped) */
uint64_t gtsb_nof_edges(const gtsb_context *ctx);
int gtsb_get_vertex_states(gtsb_context *ctx, uint8_t *vstate);
/* CSR in adjacency order; eid = index into the reference's graph->edges[] */
int gtsb_get_csr(gtsb_context *ctx, uint32_t *row_ptr, uint32_t *dst, int32_t *dist,
                 float *std_dev, uint8_t *flags, uint32_t *eid, uint32_t *win_rec,
                 uint8_t *estate);
/* edge states only, indexed by eid (= index into graph->edges[]): what the binding
   needs to write edge->state back, 1 B/edge on the bus.  Single-device graphs. */
int gtsb_get_edge_states(gtsb_context *ctx, uint8_t *estate_by_eid);
/* device pointers of the resident result (for callers that keep it on the GPU) */
int gtsb_device_pointers(gtsb_context *ctx, const uint32_t **row_ptr, const uint32_t **dst,
                         const uint32_t **eid, const uint8_t **estate, const uint8_t **vstate);
int gtsb_get_stats(gtsb_context *ctx, gtsb_stats *stats);
int gtsb_synchronize(gtsb_context *ctx);
/* per-kernel device time (CUDA events on the launching stream).  set_profile
   resets the totals; get_profile returns the number of distinct kernels and
   fills ';'-joined names, accumulated milliseconds and launch-group counts. */
int gtsb_set_profile(gtsb_context *ctx, int on);
int gtsb_get_profile(gtsb_context *ctx, char *names, uint64_t names_cap, double *ms,
                     uint32_t *calls, uint32_t cap);

/* host helper: thresholds of the ambiguous-order test for a cutoff (exposed
   for tests; see csrc/gtsb_threshold.c) */
int gtsb_ambig_thresholds(float cutoff, float *t_pos, float *t_neg, int *inf_true);

#ifdef __cplusplus
}
#endif
#endif
