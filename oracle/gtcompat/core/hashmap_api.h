/* gtcompat: the scaffolder graph/parser/algorithms sources include this
   header but use nothing from it. */
#ifndef GTCOMPAT_HASHMAP_API_H
#define GTCOMPAT_HASHMAP_API_H
#include "core/types_api.h"
typedef struct GtHashmap GtHashmap;
/* declarations for gt_scaffolder_bamparser.c (its histogram code; never called here) */
typedef enum { GT_HASH_DIRECT, GT_HASH_STRING } GtHashType;
typedef void (*GtFree)(void *);
GtHashmap *gt_hashmap_new(GtHashType keyhashtype, GtFree keyfree, GtFree valuefree);
void gt_hashmap_add(GtHashmap *hm, void *key, void *value);
void *gt_hashmap_get(GtHashmap *hm, const void *key);
void gt_hashmap_delete(GtHashmap *hm);
#endif
