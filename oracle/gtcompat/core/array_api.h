/* gtcompat: GtArray of fixed-size elements (GenomeTools core/array_api.h). */
#ifndef GTCOMPAT_ARRAY_API_H
#define GTCOMPAT_ARRAY_API_H
#include "core/types_api.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GtArray GtArray;
GtArray *gt_array_new(size_t size_of_elem);
void gt_array_add_elem(GtArray *a, void *elem, size_t size_of_elem);
#define gt_array_add(a, elem) gt_array_add_elem(a, &(elem), sizeof (elem))
void *gt_array_get(const GtArray *a, GtUword idx);
void *gt_array_pop(GtArray *a);
GtUword gt_array_size(const GtArray *a);
void gt_array_reset(GtArray *a);
void gt_array_delete(GtArray *a);
#ifdef __cplusplus
}
#endif
#endif
