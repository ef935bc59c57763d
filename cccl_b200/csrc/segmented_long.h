// Whole-grid passes for the long segments of a segmented sort (segmented_long.cu), called from segmented.cu.
#pragma once

#include "common.cuh"

namespace b200rs
{

// segments longer than this take the whole-grid path (shorter ones: one CTA per segment, segmented.cu); default of the
// run-time threshold (b200rs_set_segmented_long_min)
constexpr uint32_t SEG_LONG_MIN_DEFAULT = 1; // raised to one tile of the per-segment kernel (measured best, r4i sweep)

struct SegLong
{
  unsigned long long begin; // first item of the segment
  uint32_t len;             // < 2^30
  uint32_t first_tile;      // global id of the segment's first tile
};

struct SegLongCtl
{
  uint32_t n_long;
  uint32_t total_tiles;
  uint32_t overflow; // != 0: the table could not take every long segment -- segmented.cu sorts all of them
  uint32_t pad;
  uint32_t tile_counter[8]; // dynamic tile ids, one counter per pass
};

struct SegLongPlan
{
  const void* keys_in;
  void* keys_out;
  void* keys_tmp;
  const void* vals_in;
  void* vals_out;
  void* vals_tmp;
  const void* begin_offsets;
  const void* end_offsets;
  long long num_segments;
  int offset_bytes;
  int begin_bit;
  int end_bit;
  int passes;
  KeyXform xf;
  SegLongCtl* ctl;
  SegLong* table;
  unsigned long long* bins; // [max_long][passes][256]
  uint32_t* lookback[2];    // [tiles_bound][256] each; [0] directly behind the bins
  size_t zero_bytes;        // bins and lookback[0] (contiguous): cleared by one memset per call
  uint32_t max_long;
  uint32_t tiles_bound;
  uint32_t long_min; // segments longer than this many items are the long ones
  int sms;
};

bool seg_long_supported(int key_bytes, int value_bytes);
uint32_t seg_long_tile_items(int key_bytes, int value_bytes);
// phase 0 (on the caller's stream): build the table of the long segments; phase 1 (any stream ordered after phase 0):
// histograms, scans and the passes
cudaError_t seg_long_sort(const SegLongPlan& plan, int phase, int key_bytes, int value_bytes, cudaStream_t stream);

} // namespace b200rs
