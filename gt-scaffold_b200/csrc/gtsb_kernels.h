// gtsb_kernels.h -- launch interface between the pipeline (gtsb_api.cu) and the
// kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gtsb_common.cuh"

namespace gtsb {

// Optional per-kernel device timing (CUDA events on the launching stream),
// switched on by gtsb_set_profile; a no-op otherwise.
struct KernelTimer {
  const char *name;
  cudaStream_t stream;
  int slot;
  explicit KernelTimer(const char *n, cudaStream_t s);
  ~KernelTimer();
};
#define GTSB_TIMED(name, stream) ::gtsb::KernelTimer gtsb_timer_##__LINE__(name, stream)

constexpr uint32_t BIG_ROW = 32;    // rows above this take the block-per-row path

// device counter block (uint32 each)
enum {
  CNT_ERROR = 0,          // bit0: vertex id out of range, bit1: self link
  CNT_LARGE_BUCKETS,      // buckets queued for k_resolve_large
  CNT_LARGE_PAD,          // scratch entries they need
  CNT_BIG_ROWS,           // rows with degree > BIG_ROW
  CNT_PROPOSALS,          // (proposer, target) pairs found by phase 1
  CNT_POLY_CHANGED,       // Jacobi sweep changed a polyTime
  CNT_WORK_A,             // fire worklists (ping-pong)
  CNT_WORK_B,
  CNT_OVERFLOW,           // a list ran out of capacity
  CNT_MAX_DEG,
  CNT_NUM
};

struct BuildArgs {
  uint64_t R;
  uint32_t V;
  int sm_count;
  // records
  const uint32_t *root, *ctg;
  const int32_t *dist;
  const float *std_dev;
  const uint8_t *flags;
  // work arrays
  uint32_t *cnt, *bptr, *cursor, *deg, *krank, *scan_scratch;
  uint4 *entries;
  uint32_t *bwin;           // optional (nullptr): winning record per bucket slot
  uint8_t *creator_flag;
  uint2 *large_list;
  uint32_t *big_rows;
  uint32_t *counters;
  // CSR out
  uint32_t *row_ptr, *dst, *eid, *win_rec;
  int32_t *edist;
  float *estd;
  uint8_t *eflags;
};

void launch_build_count(const BuildArgs &a, cudaStream_t s);
void launch_build_scatter_resolve(const BuildArgs &a, cudaStream_t s);
void launch_build_resolve_large(const BuildArgs &a, uint4 *scratch, uint32_t *scratch_tag,
                                uint32_t nlarge, cudaStream_t s);
void launch_build_emit(const BuildArgs &a, cudaStream_t s);

struct GraphArgs {          // a device-resident CSR graph + vertex attributes
  uint32_t V;
  int sm_count;
  const uint32_t *row_ptr, *dst;
  const int32_t *dist;
  const float *std_dev;
  const uint8_t *flags;
  const VAttr *vattr;
  const float *astat;
  uint8_t *vstate, *estate;
  const uint32_t *big_rows;
  uint32_t n_big_rows, max_deg;
  uint32_t *counters;
};

struct FilterArgs {
  GraphArgs g;
  AmbigParams ambig;
  float cncutoff;
  long long ocutoff;
  // work arrays
  uint2 *proposals;
  uint32_t proposals_cap;
  uint32_t *poly_cur, *poly_new;
  uint8_t *gbits, *fstat;
  uint32_t *work_a, *work_b;
  uint8_t *big_scratch;      // per block: max_deg * BIG_SCRATCH_STRIDE bytes
  uint32_t big_blocks;
};
constexpr uint32_t BIG_SCRATCH_STRIDE = 12;   // cn f32, len u32, u8 marks (padded)

void launch_mark_repeats(const GraphArgs &g, uint8_t *rep_pred, float copy_num_cutoff,
                         float astat_cutoff, int use_copy_num, cudaStream_t s);
void launch_filter_pairs(const FilterArgs &a, cudaStream_t s);
void launch_poly_sweep(const FilterArgs &a, uint32_t n_proposals, cudaStream_t s);
void launch_filter_overlap(const FilterArgs &a, cudaStream_t s);
void launch_fire_round(const FilterArgs &a, const uint32_t *work_in, uint32_t n_in,
                       uint32_t *work_out, uint32_t *n_out, cudaStream_t s);
void launch_filter_finalize(const FilterArgs &a, cudaStream_t s);

}  // namespace gtsb
