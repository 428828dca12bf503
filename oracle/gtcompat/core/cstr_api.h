#ifndef GTCOMPAT_CSTR_API_H
#define GTCOMPAT_CSTR_API_H
#include "core/types_api.h"
#ifdef __cplusplus
extern "C" {
#endif
char *gt_cstr_dup(const char *cstr);
#ifdef __cplusplus
}
#endif
#endif
