#ifndef GTCOMPAT_ASSERT_API_H
#define GTCOMPAT_ASSERT_API_H
#include "core/types_api.h"
#endif
