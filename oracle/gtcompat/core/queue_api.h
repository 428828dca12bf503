/* gtcompat: FIFO queue of pointers (GenomeTools core/queue_api.h). */
#ifndef GTCOMPAT_QUEUE_API_H
#define GTCOMPAT_QUEUE_API_H
#include "core/types_api.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GtQueue GtQueue;
GtQueue *gt_queue_new(void);
void gt_queue_add(GtQueue *q, void *elem);
void *gt_queue_get(GtQueue *q);
GtUword gt_queue_size(const GtQueue *q);
void gt_queue_delete(GtQueue *q);
#ifdef __cplusplus
}
#endif
#endif
