/* oracle/ref_bam_driver.c -- TEST INFRASTRUCTURE, not product code.
   The reference's distance estimator (gt_scaffolder_bamparser.c:385-598: window,
   calculate_fragment_dist, compute_likelihood, maximum_likelihood_estimate,
   estimate_dist_using_mle) consists of static functions; this driver includes the reference's
   source file UNMODIFIED from where it lies and calls them on arrays.  The BAM reader in the
   same file needs the htslib-backed GenomeTools SAM API, which is not here: its entry points are
   declared by gtcompat headers and defined below as stubs that are never reached. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gt_scaffolder_bamparser.c"

static void not_here(const char *what)
{
  fprintf(stderr, "oracle: '%s' needs the GenomeTools SAM API, which is not built here\n", what);
  exit(3);
}
GtAlphabet *gt_alphabet_new_dna(void) { not_here("gt_alphabet_new_dna"); return NULL; }
void gt_alphabet_delete(GtAlphabet *a) { (void) a; }
uint16_t gt_sam_alignment_cigar_length(GtSamAlignment *a) { (void) a; not_here("sam"); return 0; }
unsigned char gt_sam_alignment_cigar_i_operation(GtSamAlignment *a, uint16_t i) { (void) a; (void) i; not_here("sam"); return 0; }
uint32_t gt_sam_alignment_cigar_i_length(GtSamAlignment *a, uint16_t i) { (void) a; (void) i; not_here("sam"); return 0; }
bool gt_sam_alignment_is_reverse(GtSamAlignment *a) { (void) a; not_here("sam"); return false; }
bool gt_sam_alignment_is_unmapped(GtSamAlignment *a) { (void) a; not_here("sam"); return false; }
GtUword gt_sam_alignment_pos(GtSamAlignment *a) { (void) a; not_here("sam"); return 0; }
const char *gt_sam_alignment_identifier(GtSamAlignment *a) { (void) a; not_here("sam"); return NULL; }
int32_t gt_sam_alignment_ref_num(GtSamAlignment *a) { (void) a; not_here("sam"); return 0; }
GtUword gt_sam_alignment_mapping_quality(GtSamAlignment *a) { (void) a; not_here("sam"); return 0; }
GtSamfileIterator *gt_samfile_iterator_new_bam(const char *f, GtAlphabet *a, GtError *e)
{ (void) f; (void) a; (void) e; not_here("samfile_iterator"); return NULL; }
int gt_samfile_iterator_next(GtSamfileIterator *it, GtSamAlignment **aln) { (void) it; (void) aln; not_here("samfile_iterator"); return 0; }
const char *gt_samfile_iterator_reference_name(const GtSamfileIterator *it, int32_t r) { (void) it; (void) r; not_here("samfile_iterator"); return NULL; }
GtUword gt_samfile_iterator_reference_length(const GtSamfileIterator *it, int32_t r) { (void) it; (void) r; not_here("samfile_iterator"); return 0; }
void gt_samfile_iterator_delete(GtSamfileIterator *it) { (void) it; }
GtHashmap *gt_hashmap_new(GtHashType t, GtFree k, GtFree v) { (void) t; (void) k; (void) v; not_here("hashmap"); return NULL; }
void gt_hashmap_add(GtHashmap *hm, void *key, void *value) { (void) hm; (void) key; (void) value; not_here("hashmap"); }
void *gt_hashmap_get(GtHashmap *hm, const void *key) { (void) hm; (void) key; not_here("hashmap"); return NULL; }
void gt_hashmap_delete(GtHashmap *hm) { (void) hm; }

/* estimate_dist_using_mle (bamparser.c:553-598) for one contig pair.  frag_pos = n (start, end)
   pairs as calculate_fragment stores them (it is sorted in place, as the reference does);
   ma = FragmentData.ma; pmf / pmf_nof / minp = PmfData.dist / .nof / .minp. */
int refbam_estimate_dist(int64_t *frag_pos, uint64_t n, uint64_t ma, double *pmf, uint64_t pmf_nof, double minp,
                         int64_t min_dist, int64_t max_dist, uint64_t len_ref, uint64_t len_mref, int rf,
                         int64_t *dist, uint64_t *nof_pairs)
{
  FragmentData fd;
  PmfData pd;
  GtError *err = gt_error_new();
  GtWord d = 0;
  GtUword np = 0;
  int rc;
  memset(&fd, 0, sizeof fd);
  memset(&pd, 0, sizeof pd);
  fd.frag_pos = (GtWord *) frag_pos;
  fd.nof_frag_pos = n;
  fd.size_frag_pos = n;
  fd.ma = ma;
  pd.dist = pmf;
  pd.nof = pmf_nof;
  pd.minp = minp;
  rc = estimate_dist_using_mle(&d, &np, (GtWord) min_dist, (GtWord) max_dist, &fd, pd, len_ref, len_mref,
                               rf != 0, err);
  gt_free(fd.frag_size);
  gt_error_delete(err);
  *dist = d;
  *nof_pairs = np;
  return rc;
}
