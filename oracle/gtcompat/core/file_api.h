/* gtcompat: GtFile output stream (GenomeTools core/file_api.h surface). */
#ifndef GTCOMPAT_FILE_API_H
#define GTCOMPAT_FILE_API_H
#include "core/error.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GtFile GtFile;
GtFile *gt_file_new(const char *path, const char *mode, GtError *err);
void gt_file_xprintf(GtFile *file, const char *format, ...)
  __attribute__((format(printf, 2, 3)));
void gt_file_delete(GtFile *file);
#ifdef __cplusplus
}
#endif
#endif
