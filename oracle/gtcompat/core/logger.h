#ifndef GTCOMPAT_LOGGER_H
#define GTCOMPAT_LOGGER_H
#include "core/types_api.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GtLogger GtLogger;
GtLogger *gt_logger_new(bool enabled, const char *prefix, FILE *target);
void gt_logger_log(GtLogger *logger, const char *format, ...)
  __attribute__((format(printf, 2, 3)));
void gt_logger_delete(GtLogger *logger);
#ifdef __cplusplus
}
#endif
#endif
