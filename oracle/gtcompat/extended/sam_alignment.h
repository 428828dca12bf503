/* gtcompat: declarations only (see core/alphabet_api.h) */
#ifndef GTCOMPAT_SAM_ALIGNMENT_H
#define GTCOMPAT_SAM_ALIGNMENT_H
#include <stdbool.h>
#include <stdint.h>
#include "core/types_api.h"
typedef struct GtSamAlignment GtSamAlignment;
uint16_t gt_sam_alignment_cigar_length(GtSamAlignment *a);
unsigned char gt_sam_alignment_cigar_i_operation(GtSamAlignment *a, uint16_t i);
uint32_t gt_sam_alignment_cigar_i_length(GtSamAlignment *a, uint16_t i);
bool gt_sam_alignment_is_reverse(GtSamAlignment *a);
bool gt_sam_alignment_is_unmapped(GtSamAlignment *a);
GtUword gt_sam_alignment_pos(GtSamAlignment *a);
const char *gt_sam_alignment_identifier(GtSamAlignment *a);
int32_t gt_sam_alignment_ref_num(GtSamAlignment *a);
GtUword gt_sam_alignment_mapping_quality(GtSamAlignment *a);
#endif
