/* gtcompat: declarations only (see core/alphabet_api.h) */
#ifndef GTCOMPAT_SAMFILE_ITERATOR_H
#define GTCOMPAT_SAMFILE_ITERATOR_H
#include "core/alphabet_api.h"
#include "core/error_api.h"
#include "extended/sam_alignment.h"
typedef struct GtSamfileIterator GtSamfileIterator;
GtSamfileIterator *gt_samfile_iterator_new_bam(const char *filename, GtAlphabet *alphabet, GtError *err);
int gt_samfile_iterator_next(GtSamfileIterator *it, GtSamAlignment **aln);
const char *gt_samfile_iterator_reference_name(const GtSamfileIterator *it, int32_t ref);
GtUword gt_samfile_iterator_reference_length(const GtSamfileIterator *it, int32_t ref);
void gt_samfile_iterator_delete(GtSamfileIterator *it);
#endif
