/* oracle/gtscaf_oracle.h -- TEST INFRASTRUCTURE, not product code.
   Array-level CPU restatement of the reference's hot path.  See
   gtscaf_oracle.c for the reference file:line each function follows.
   Parity status: PINNED -- checked against the compiled reference
   (oracle/_ref) on the reference's golden testdata and on randomized graphs
   by tests/test_oracle.py. */
#ifndef GTSCAF_ORACLE_H
#define GTSCAF_ORACLE_H
#include <stdint.h>

/* GraphItemState values, reference graph.h:29-31 */
enum { ORA_UNVISITED = 0, ORA_POLYMORPHIC = 1, ORA_INCONSISTENT = 2,
       ORA_REPEAT = 3, ORA_VISITED = 4, ORA_PROCESSED = 5, ORA_SCAFFOLD = 6,
       ORA_CYCLIC = 7 };

#define ORA_SENSE 1u
#define ORA_SAME  2u

typedef struct OraGraph {
  uint64_t nof_vertices;
  uint64_t *seq_len;     /* [V] */
  float *astat;          /* [V] */
  float *copy_num;       /* [V] */
  uint8_t *vstate;       /* [V] */
  uint64_t *nof_vedges;  /* [V] out-degree */
  uint32_t **vedges;     /* [V] -> edge ids in adjacency (insertion) order */
  uint64_t nof_edges, max_nof_edges;
  uint32_t *src, *dst;   /* [E] in creation order */
  int64_t *dist;
  float *std_dev;
  uint64_t *num_pairs;
  uint8_t *flags;        /* ORA_SENSE | ORA_SAME */
  uint8_t *estate;
  int64_t *win_rec;      /* record whose attributes the edge carries */
} OraGraph;

OraGraph *ora_build(uint64_t nof_vertices, const uint64_t *seq_len,
                    const float *astat, const float *copy_num,
                    uint64_t nof_records, const uint32_t *root,
                    const uint32_t *ctg, const int64_t *dist,
                    const float *std_dev, const uint64_t *num_pairs,
                    const uint8_t *flags);
void ora_mark_repeats(OraGraph *g, int use_copy_num, float copy_num_cutoff,
                      float astat_cutoff);
void ora_filter(OraGraph *g, float pcutoff, float cncutoff, int64_t ocutoff);
int ora_ambiguous_interval(float interval, float cutoff);
void ora_ambiguous_intervals(const float *interval, uint64_t n, float cutoff, uint8_t *out);
void ora_ambiguousorders(const int64_t *dist1, const float *std1, const int64_t *dist2, const float *std2,
                         uint64_t n, float cutoff, uint8_t *out);
int ora_ambiguousorder(int64_t dist1, float std1, int64_t dist2, float std2,
                       float cutoff);
int64_t ora_overlap(int64_t dist1, uint64_t len1, int64_t dist2, uint64_t len2);
void ora_get_adjacency(const OraGraph *g, uint64_t *row_ptr, uint32_t *eids);
void ora_delete(OraGraph *g);
#endif
