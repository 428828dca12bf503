// cusim/cooperative_groups.h -- TEST INFRASTRUCTURE.  grid.sync() of a cooperative launch
// (cusim::run_grid_coop keeps every block of the grid alive at once).
#pragma once
#include "cuda_runtime.h"
namespace cooperative_groups {
struct grid_group {
  void sync() const { cusim::grid_barrier(); }
};
static inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
