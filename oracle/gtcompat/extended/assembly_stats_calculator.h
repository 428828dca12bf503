/* gtcompat: N50-style length statistics, reported on stderr only
   (test.c:165-192, algorithms.c:993); never part of a parity observable. */
#ifndef GTCOMPAT_ASSEMBLY_STATS_CALCULATOR_H
#define GTCOMPAT_ASSEMBLY_STATS_CALCULATOR_H
#include "core/logger.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GtAssemblyStatsCalculator GtAssemblyStatsCalculator;
GtAssemblyStatsCalculator *gt_assembly_stats_calculator_new(void);
void gt_assembly_stats_calculator_add(GtAssemblyStatsCalculator *calc,
                                      GtUword length);
void gt_assembly_stats_calculator_nstat(GtAssemblyStatsCalculator *calc,
                                        GtUword n);
void gt_assembly_stats_calculator_show(GtAssemblyStatsCalculator *calc,
                                       GtLogger *logger);
void gt_assembly_stats_calculator_delete(GtAssemblyStatsCalculator *calc);
#ifdef __cplusplus
}
#endif
#endif
