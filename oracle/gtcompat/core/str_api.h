/* gtcompat: GtStr growable string (GenomeTools core/str_api.h surface). */
#ifndef GTCOMPAT_STR_API_H
#define GTCOMPAT_STR_API_H
#include "core/types_api.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GtStr GtStr;
GtStr *gt_str_new(void);
GtStr *gt_str_new_cstr(const char *cstr);
GtStr *gt_str_clone(const GtStr *str);
void gt_str_set(GtStr *str, const char *cstr);
void gt_str_append_cstr(GtStr *str, const char *cstr);
char *gt_str_get(const GtStr *str);
GtUword gt_str_length(const GtStr *str);
/* byte-wise strcmp order; defines the vertex ids (parser.c:45-52,172) */
int gt_str_cmp(const GtStr *a, const GtStr *b);
void gt_str_delete(GtStr *str);
#ifdef __cplusplus
}
#endif
#endif
