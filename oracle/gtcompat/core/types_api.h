/* gtcompat: minimal stand-in for GenomeTools core/types_api.h (LP64 build).
   GenomeTools is an external, un-vendored dependency of the reference
   (README.md:30-33); this header provides only the names the scaffolder
   sources use. Written from the public GenomeTools API, not copied. */
#ifndef GTCOMPAT_TYPES_API_H
#define GTCOMPAT_TYPES_API_H
#include <limits.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef unsigned long GtUword;
typedef long GtWord;
#define GT_WU "%lu"
#define GT_WD "%ld"
#define GT_WORD_MAX LONG_MAX
#define GT_WORD_MIN LONG_MIN
typedef unsigned char GtUchar;
#define GT_UWORD_MAX ULONG_MAX
#define GT_UNUSED __attribute__((unused))

/* GenomeTools aborts with exit code 2 on a failed assertion
   (the reference's testsuite depends on it: scaffolder_include.rb:19-23). */
#define GT_EXIT_PROGRAMMING_ERROR 2
#define gt_assert(expr)                                                      \
  do {                                                                       \
    if (!(expr)) {                                                           \
      fprintf(stderr, "Assertion failed: (%s), function %s, file %s, "       \
              "line %d.\n", #expr, __func__, __FILE__, __LINE__);            \
      exit(GT_EXIT_PROGRAMMING_ERROR);                                       \
    }                                                                        \
  } while (0)
#endif
