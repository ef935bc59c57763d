// Host-visible table of compiled onesweep tile configurations (one table per key width x value width).
#pragma once

#include "common.cuh"

namespace b200rs
{

typedef cudaError_t (*onesweep_launch_fn)(const PassArgs& args, unsigned grid, cudaStream_t stream);

struct OnesweepConfig
{
  int threads;
  int items_per_thread;
  int rank_algo;
  int min_blocks;
  int tile_items;
  size_t smem_bytes;
  onesweep_launch_fn launch;
  int bulk_store; // 1: TMA bulk-store kernel (onesweep_tma.cuh); needs 16-byte aligned output pointers
  int opt;        // OnesweepOpt bits of the classic kernel
  onesweep_launch_fn launch_bucket; // the same configuration in bucket mode (OPT_BUCKET), or nullptr
};

// index 0 is the default for the (key_bytes, value_bytes) combination
const OnesweepConfig* onesweep_configs_k1(int value_bytes, int* count);
const OnesweepConfig* onesweep_configs_k2(int value_bytes, int* count);
// 4- and 8-byte keys: one translation unit per value-width group so that the library builds in parallel
const OnesweepConfig* onesweep_configs_k4_v0(int* count);
const OnesweepConfig* onesweep_configs_k4_v4(int* count);
const OnesweepConfig* onesweep_configs_k4_vx(int value_bytes, int* count);
const OnesweepConfig* onesweep_configs_k8_v0(int* count);
const OnesweepConfig* onesweep_configs_k8_v4(int* count);
const OnesweepConfig* onesweep_configs_k8_vx(int value_bytes, int* count);

cudaError_t launch_histogram(
  const void* keys, unsigned long long n, int key_bytes, unsigned long long* bins, int passes, int begin_bit,
  int end_bit, const KeyXform& xf, int sm_count, cudaStream_t stream, uint32_t* zero_flag = nullptr);
cudaError_t launch_scan_bins(unsigned long long* bins, int passes, cudaStream_t stream);

// whole sort in one CTA (single_tile.cu); capacity in items for the given key / value widths
unsigned long long single_tile_capacity(int key_bytes, int value_bytes);
cudaError_t launch_single_tile(
  const void* keys_in, void* keys_out, const void* vals_in, void* vals_out, unsigned long long n, int key_bytes,
  int value_bytes, int begin_bit, int end_bit, const KeyXform& xf, cudaStream_t stream);

// the whole sort of a mid-size input in one cooperative launch (small.cu); `passes` is what the general path would launch
bool small_sort_supported(int key_bytes, int value_bytes);
unsigned long long small_sort_tile_items();
cudaError_t launch_small_sort(const PassArgs* passes, int num_passes, unsigned long long* bins, unsigned tiles,
                              int key_bytes, int value_bytes, int sms, cudaStream_t stream);

cudaError_t launch_splitter_ranks(
  const void* sorted_keys, unsigned long long n, int key_bytes, const KeyXform& xf, const void* splitters,
  int num_splitters, unsigned long long* lt, unsigned long long* eq, cudaStream_t stream);

} // namespace b200rs
