/* gtcompat: memory allocation wrappers (GenomeTools core/ma_api.h surface). */
#ifndef GTCOMPAT_MA_API_H
#define GTCOMPAT_MA_API_H
#include "core/types_api.h"
#ifdef __cplusplus
extern "C" {
#endif
void *gt_malloc_mem(size_t size, const char *file, int line);
void *gt_calloc_mem(size_t nmemb, size_t size, const char *file, int line);
void *gt_realloc_mem(void *ptr, size_t size, const char *file, int line);
void gt_free_mem(void *ptr);
#ifdef __cplusplus
}
#endif
#define gt_malloc(size) gt_malloc_mem(size, __FILE__, __LINE__)
#define gt_calloc(nmemb, size) gt_calloc_mem(nmemb, size, __FILE__, __LINE__)
#define gt_realloc(ptr, size) gt_realloc_mem(ptr, size, __FILE__, __LINE__)
#define gt_free(ptr) gt_free_mem(ptr)
#endif
