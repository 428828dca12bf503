/* gtcompat: GtError object (GenomeTools core/error_api.h surface). */
#ifndef GTCOMPAT_ERROR_API_H
#define GTCOMPAT_ERROR_API_H
#include "core/types_api.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GtError GtError;
GtError *gt_error_new(void);
void gt_error_set(GtError *err, const char *format, ...)
  __attribute__((format(printf, 2, 3)));
bool gt_error_is_set(const GtError *err);
void gt_error_unset(GtError *err);
const char *gt_error_get(const GtError *err);
void gt_error_delete(GtError *err);
#ifdef __cplusplus
}
#endif
#endif
