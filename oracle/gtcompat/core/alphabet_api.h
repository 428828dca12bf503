/* gtcompat: declarations only -- gt_scaffolder_bamparser.c is compiled for its distance
   estimator (oracle/ref_bam_driver.c); its BAM reader is never called here. */
#ifndef GTCOMPAT_ALPHABET_API_H
#define GTCOMPAT_ALPHABET_API_H
typedef struct GtAlphabet GtAlphabet;
GtAlphabet *gt_alphabet_new_dna(void);
void gt_alphabet_delete(GtAlphabet *a);
#endif
