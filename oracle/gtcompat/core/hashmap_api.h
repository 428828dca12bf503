/* gtcompat: the scaffolder graph/parser/algorithms sources include this
   header but use nothing from it. */
#ifndef GTCOMPAT_HASHMAP_API_H
#define GTCOMPAT_HASHMAP_API_H
#include "core/types_api.h"
typedef struct GtHashmap GtHashmap;
#endif
