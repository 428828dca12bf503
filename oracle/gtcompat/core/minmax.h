#ifndef GTCOMPAT_MINMAX_H
#define GTCOMPAT_MINMAX_H
#ifndef MAX
#define MAX(a,b) ((a)>(b)?(a):(b))
#endif
#ifndef MIN
#define MIN(a,b) ((a)<(b)?(a):(b))
#endif
#endif
