/* oracle/ref_stubs.c -- TEST INFRASTRUCTURE, not product code.
   Link-time stand-ins for the two reference modules that are OUT of the hot
   path and cannot be built here (they need the htslib-backed / forked
   GenomeTools: SURVEY.md section 2 rows 6-7): the BAM parser and the FASTA
   generator.  test.c references their entry points (test.c:174-176,
   202-222); the `scaffold`, `graph` and `parser` modules never reach them
   when the 4th scaffold argument is "false". */
#include <stdio.h>
#include <stdlib.h>
#include "core/array_api.h"
#include "gt_scaffolder_bamparser.h"
#include "gt_scaffolder_generate_fasta.h"

static void not_built(const char *what)
{
  fprintf(stderr, "oracle: reference module '%s' is not built here\n", what);
  exit(3);
}

DistRecords *gt_scaffolder_bamparser_init_dist_records(void)
{ not_built("bamparser"); return NULL; }

int gt_scaffolder_bamparser_print_dist_records(const DistRecords *dist,
                                               const char *filename,
                                               GtError *err)
{ (void) dist; (void) filename; (void) err; not_built("bamparser"); return -1; }

void gt_scaffolder_bamparser_delete_dist_records(DistRecords *dist)
{ (void) dist; not_built("bamparser"); }

int gt_scaffolder_bamparser_read_paired_information(DistRecords *dist,
                                                    const char *bam_filename,
                                                    GtWord min_dist,
                                                    GtWord max_dist,
                                                    GtUword min_qual,
                                                    GtUword min_nof_pairs,
                                                    GtUword min_ref_length,
                                                    GtUword min_align,
                                                    GtError *err)
{
  (void) dist; (void) bam_filename; (void) min_dist; (void) max_dist;
  (void) min_qual; (void) min_nof_pairs; (void) min_ref_length;
  (void) min_align; (void) err;
  not_built("bamparser");
  return -1;
}

int gt_scaffolder_graph_generate_fasta(char *contig_file, char *spm_file,
                                       char *fasta_file, GtArray *recs,
                                       GtError *err)
{
  (void) contig_file; (void) spm_file; (void) fasta_file; (void) recs;
  (void) err;
  not_built("generate_fasta");
  return -1;
}
