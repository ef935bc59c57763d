/*
 * b200rs.h -- C ABI of the B200-native (sm_100a) LSD radix sort: the drop-in boundary for the
 * reference's cub::DeviceRadixSort hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and returns a cudaError_t value
 * as int (0 == cudaSuccess).  The model is the reference's own C ABI for this path,
 *   cccl_device_radix_sort(build, d_temp_storage, temp_storage_bytes, keys_in, keys_out, values_in,
 *                          values_out, decomposer, num_items, begin_bit, end_bit, is_overwrite_okay,
 *                          selector, stream)
 *   (/root/reference/c/parallel/include/cccl/c/radix_sort.h:110-124, impl c/parallel/src/radix_sort.cu:582-661),
 * minus the NVRTC "build" object (kernels here are compiled ahead of time) and with the iterator structs
 * replaced by (pointer, key kind, key bytes, value bytes).  The C++ templates cub::DeviceRadixSort /
 * cub::DoubleBuffer (include/cub/device/device_radix_sort.cuh in this repo) and the thrust::sort shim
 * (include/thrust/sort.h) are header-only wrappers that type-erase to these calls.
 *
 * Conventions (same as cub/cub/device/device_radix_sort.cuh:300-317, dispatch_radix_sort.cuh:1739-1743,
 * :1950-1977, util_temporary_storage.cuh:75-78 in the reference):
 *   - two-phase: d_temp_storage == NULL  => only *temp_storage_bytes is written, no work is enqueued;
 *     otherwise *temp_storage_bytes is the size of the blob provided, too small => cudaErrorInvalidValue (1);
 *   - all work is enqueued on `stream` and is NOT synchronised; no allocation, no host sync; legal under
 *     CUDA stream capture;
 *   - the caller owns every buffer; keys/values ranges must not overlap each other or the temp blob;
 *   - never throws, never prints.
 */
#ifndef B200RS_H_
#define B200RS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#  define B200RS_API __declspec(dllexport)
#else
#  define B200RS_API __attribute__((visibility("default")))
#endif

/* Opaque CUDA stream (cudaStream_t / CUstream); NULL is the legacy default stream. */
typedef struct CUstream_st* b200rs_stream_t;

/* How the key bits are mapped to an order-preserving unsigned integer ("twiddle"):
 * cub::Traits<T>::TwiddleIn/Out, /root/reference/cub/cub/util_type.cuh:857-865 (unsigned),
 * :906-914 (signed), :953-963 (floating point). */
typedef enum b200rs_key_kind
{
  B200RS_KEY_UINT  = 0, /* unsigned integers, bool, char: identity            */
  B200RS_KEY_INT   = 1, /* two's complement signed: flip the sign bit         */
  B200RS_KEY_FLOAT = 2  /* IEEE-754: negative -> ~x, else flip the sign bit;  */
                        /* -0.0 and +0.0 compare equal (stable), NaNs by bits. */
                        /* key_bytes 2 (__half or __nv_bfloat16: same sign-     */
                        /* magnitude layout, util_type.cuh:1017-1095), 4 or 8   */
} b200rs_key_kind;

#define B200RS_VERSION 100 /* 0.1.0 */

/* Library version (B200RS_VERSION of the build). */
B200RS_API int b200rs_version(void);

/*
 * Device-wide stable LSD radix sort of `num_items` keys (and, if value_bytes > 0, the values that travel
 * with them) on bits [begin_bit, end_bit) of the transformed key.
 *
 * Replaces: cub::DeviceRadixSort::{SortKeys,SortKeysDescending,SortPairs,SortPairsDescending}, pointer and
 * DoubleBuffer overloads (/root/reference/cub/cub/device/device_radix_sort.cuh:412,1127,1776,2295,3034,3644,
 * 4211,4669) == detail::radix_sort::dispatch (dispatch/dispatch_radix_sort.cuh:2016-2065).
 *
 *   d_keys_in / d_keys_out      "buffer 0" / "buffer 1" of the key DoubleBuffer (device pointers)
 *   d_values_in / d_values_out  same for values; ignored when value_bytes == 0
 *   key_kind                    b200rs_key_kind
 *   key_bytes                   1, 2, 4 or 8
 *   value_bytes                 0 (keys only), 1, 2, 4, 8 or 16; pointers must be aligned to
 *                               min(value_bytes, 8)
 *   begin_bit, end_bit          0 <= begin_bit <= end_bit <= 8*key_bytes
 *   descending                  0 ascending, 1 descending (equal keys keep INPUT order in both)
 *   is_overwrite_okay           0: pointer API -- buffer 0 is never written, the result is in buffer 1,
 *                                  *selector = 1, temp blob holds ~N*(key_bytes+value_bytes) extra;
 *                               1: DoubleBuffer API -- both buffers are clobbered, *selector (0/1) tells
 *                                  which one holds the result, temp blob is O(N/tile)
 *   selector                    host pointer, written by every successful EXECUTE call (untouched by the size
 *                               query; may be NULL)
 *
 * Returns 0 (cudaSuccess), 1 (cudaErrorInvalidValue: bad arguments or temp blob too small),
 * 801 (cudaErrorNotSupported: unsupported key/value width) or the CUDA error of a failed launch.
 */
B200RS_API int b200rs_sort(
  void* d_temp_storage,
  size_t* temp_storage_bytes,
  const void* d_keys_in,
  void* d_keys_out,
  const void* d_values_in,
  void* d_values_out,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int begin_bit,
  int end_bit,
  int descending,
  int is_overwrite_okay,
  int* selector,
  b200rs_stream_t stream);

/*
 * b200rs_sort with PER-CALL tuning -- the reference's `cuda::execution::tune(policy)` carried by the environment of the
 * env overloads (/root/reference/cub/cub/device/device_radix_sort.cuh:200-202, detail/env_dispatch.cuh:39-77): which
 * compiled tile configuration runs the digit passes (index into b200rs_describe_config's table), up to which size the
 * one-launch cooperative kernel is used, whether inputs of one tile take the single-CTA kernel.  A negative field
 * keeps the library's own choice.  tuning == NULL is b200rs_sort.  The temp-storage size may depend on the tuning: query
 * and run with the same one.  Thread-safe (nothing process-wide is touched).
 */
typedef struct b200rs_tuning
{
  int config_index;          /* onesweep tile configuration, or -1 */
  long long small_max_items; /* largest input for the one-launch kernel (0 = never), or -1 */
  int single_tile;           /* 1 / 0: use / do not use the single-CTA kernel, or -1 */
} b200rs_tuning;

B200RS_API int b200rs_sort_tuned(
  void* d_temp_storage,
  size_t* temp_storage_bytes,
  const void* d_keys_in,
  void* d_keys_out,
  const void* d_values_in,
  void* d_values_out,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int begin_bit,
  int end_bit,
  int descending,
  int is_overwrite_okay,
  int* selector,
  b200rs_stream_t stream,
  const b200rs_tuning* tuning);

/*
 * Segmented sort: every segment [begin_offsets[s], end_offsets[s]) of the arrays is sorted independently, with the
 * order, bit window and stability of b200rs_sort; items outside every segment are neither read nor written.
 *
 * Replaces cub::DeviceSegmentedRadixSort::SortKeys / SortPairs [Descending], pointer form
 * (/root/reference/cub/cub/device/device_segmented_radix_sort.cuh:234,887,1531,2139; dispatch_segmented_radix_sort.cuh;
 * kernel kernel_segmented_radix_sort.cuh:113-300).  d_begin_offsets / d_end_offsets: DEVICE arrays of num_segments
 * integers of offset_bytes (4 or 8) bytes each; segments may be empty, leave gaps, and must not overlap; a segment holds
 * fewer than 2^31 items.  Two-phase temp-storage query like b200rs_sort (the blob holds a scratch copy of the arrays
 * for segments longer than one tile).  The inputs are never written; results are always in d_keys_out / d_values_out.
 * ONE kernel launch: one CTA per segment, all digit passes inside it.
 */
B200RS_API int b200rs_segmented_sort(
  void* d_temp_storage,
  size_t* temp_storage_bytes,
  const void* d_keys_in,
  void* d_keys_out,
  const void* d_values_in,
  void* d_values_out,
  uint64_t num_items,
  uint64_t num_segments,
  const void* d_begin_offsets,
  const void* d_end_offsets,
  int offset_bytes,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int begin_bit,
  int end_bit,
  int descending,
  b200rs_stream_t stream);

/*
 * Sort of user-defined keys described as a tuple of arithmetic members, most significant first (the reference's
 * "decomposer" keys), and of 128-bit integers (two 64-bit members).
 *
 * Replaces the decomposer overloads of cub::DeviceRadixSort
 * (/root/reference/cub/cub/device/device_radix_sort.cuh:671,905,1359,1565 and their Descending / SortKeys twins;
 * member-wise digit extraction cub/cub/block/radix_rank_sort_operations.cuh:436-527).  d_keys_in / d_keys_out: arrays of
 * key_stride_bytes-byte items; fields[f] = {byte offset inside the item, member width 1/2/4/8, b200rs_key_kind}; members
 * compare like b200rs_sort keys (-0.0 == +0.0 per member).  [begin_bit, end_bit) counts from the least significant bit of
 * the LAST member over the concatenated members.  Values: any item size (value_bytes == 0: keys only).  Stable;
 * inputs are never written; fewer than 2^32 items.  Two-phase temp-storage query like b200rs_sort.
 */
typedef struct b200rs_key_field
{
  uint32_t offset;
  uint32_t bytes;
  int32_t kind;
} b200rs_key_field;

B200RS_API int b200rs_sort_fields(
  void* d_temp_storage,
  size_t* temp_storage_bytes,
  const void* d_keys_in,
  void* d_keys_out,
  int key_stride_bytes,
  const b200rs_key_field* fields,
  int num_fields,
  const void* d_values_in,
  void* d_values_out,
  int value_bytes,
  uint64_t num_items,
  int begin_bit,
  int end_bit,
  int descending,
  b200rs_stream_t stream);

/*
 * Top-K selection: the k best keys (largest != 0: the k largest, else the k smallest; -0.0 == +0.0, NaNs by bits as in
 * b200rs_sort) and their values are written to d_keys_out / d_values_out[0 .. min(k, num_items)) in NO particular order;
 * which keys tied with the k-th one are returned is unspecified.
 *
 * Replaces cub::DeviceTopK::{Max,Min}{Keys,Pairs} (/root/reference/cub/cub/device/device_topk.cuh:297,775,1238;
 * dispatch/dispatch_topk.cuh).  Exact MSD radix select of the k-th key (the splitter selection of the multi-GPU sort with
 * one target) + one filter pass: three reads of the keys, no sort, no host wait.  Two-phase temp-storage query.
 */
B200RS_API int b200rs_topk(
  void* d_temp_storage,
  size_t* temp_storage_bytes,
  const void* d_keys_in,
  void* d_keys_out,
  const void* d_values_in,
  void* d_values_out,
  uint64_t num_items,
  uint64_t k,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int largest,
  b200rs_stream_t stream);

/*
 * In-place sort of device memory: the thrust::sort / thrust::sort_by_key front door (full key width).
 *
 * Replaces thrust::cuda_cub::__radix_sort::radix_sort + the tail of __smart_sort::smart_sort
 * (/root/reference/thrust/thrust/system/cuda/detail/sort.h:219-266, :288-339): obtains scratch for the alternate
 * buffers and the temp blob (stream-ordered pool of the current device instead of the reference's cudaMalloc /
 * cudaFree per call), runs the DoubleBuffer sort, copies back iff the result landed in the scratch half, releases the
 * scratch; `synchronize` != 0 additionally waits for the stream (the reference's synchronize_optional).
 * d_values is ignored when value_bytes == 0.  thrust::less => descending = 0, thrust::greater => descending = 1.
 */
B200RS_API int b200rs_sort_inplace(
  void* d_keys,
  void* d_values,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int descending,
  int synchronize,
  b200rs_stream_t stream);

/*
 * The upsweep on its own: one read of the keys produces the digit histogram of every 8-bit pass.
 * d_bins is a device array of uint64[ceil((end_bit-begin_bit)/8) * 256], overwritten.
 *
 * Replaces: DeviceRadixSortHistogramKernel / AgentRadixSortHistogram::Process
 * (/root/reference/cub/cub/device/dispatch/kernels/kernel_radix_sort.cuh:447-473,
 *  cub/cub/agent/agent_radix_sort_histogram.cuh:248-279).  Exposed so the kernel can be checked and timed alone.
 */
B200RS_API int b200rs_digit_histogram(
  const void* d_keys_in,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int begin_bit,
  int end_bit,
  int descending,
  uint64_t* d_bins,
  b200rs_stream_t stream);

/*
 * Multi-GPU support kernel: for a locally SORTED key array, count for each of
 * `num_splitters` splitter keys how many local keys order strictly before it (d_lt[i]) and how many are
 * equal to it (d_eq[i]) under the same transform/order as b200rs_sort with the full bit range.
 * d_splitters: device array of `num_splitters` keys (same type as the keys).  d_lt/d_eq: device uint64 arrays.
 *
 * Replaces the per-probe local counts of the reference's multi-GPU sort
 * (/root/reference/cudax/include/cuda/experimental/__multi_gpu/algorithm/sort/hss/histogramming.h:522-610).
 */
B200RS_API int b200rs_splitter_ranks(
  const void* d_sorted_keys,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int descending,
  const void* d_splitters,
  int num_splitters,
  uint64_t* d_lt,
  uint64_t* d_eq,
  b200rs_stream_t stream);

/*
 * Multi-GPU support kernels of the partition-first protocol (cccl_b200/multi_gpu.py).
 * Both work on an UNSORTED shard, in the sort's own key domain (bit-ordered value after the key transform, descending
 * included, -0.0 viewed as +0.0), i.e. "equal" means what b200rs_sort treats as equal.
 *
 * b200rs_select_histogram: one round of the exact MSD radix select.  d_prefixes (DEVICE array, num_prefixes <= 15, so a
 * round can consume the previous round's choice without a host round trip) holds the high digits chosen so far, as
 * values of `round` digits (round 0: all zero); duplicates are allowed.  Optional candidate compaction: with
 * d_candidates_out / d_candidate_state_out (uint64[1026]: overflow flag and per-CTA counts) every key that carries one of the
 * prefixes is also appended to d_candidates_out (capacity in keys; on overflow the flag is set); with d_candidates_in /
 * d_candidate_state_in a later round scans that buffer instead of all keys (or all keys if the flag is set).  d_hist[p * 256 + b] receives the
 * number of local keys whose top `round` digits equal d_prefixes[p] and whose next 8-bit digit is b.  Overwritten.
 *
 * b200rs_bucket_ids: h_splitters (HOST array, strictly increasing bit-ordered values, num_splitters <= 15);
 * d_ids[i] = 2 * #{j : splitter_j < key_i} + [key_i == some splitter_j]  (uint8).
 *
 * Replace the counting rounds and the destination computation of the reference's multi-GPU sort
 * (/root/reference/cudax/include/cuda/experimental/__multi_gpu/algorithm/sort/hss/histogramming.h:522-610,
 *  hss/data_exchange.h:380-450).
 */
B200RS_API int b200rs_select_histogram(
  const void* d_keys,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int descending,
  const uint64_t* d_prefixes,
  int num_prefixes,
  int round,
  uint64_t* d_hist,
  const void* d_candidates_in,
  const uint64_t* d_candidate_state_in,
  void* d_candidates_out,
  uint64_t* d_candidate_state_out,
  uint64_t candidate_capacity,
  b200rs_stream_t stream);

B200RS_API int b200rs_bucket_ids(
  const void* d_keys,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int descending,
  const uint64_t* h_splitters,
  int num_splitters,
  uint8_t* d_ids,
  b200rs_stream_t stream);

/*
 * The partition pass of the multi-GPU protocol as ONE onesweep launch: keys (and values) are moved, stably, into
 * destination-bucket order, the bucket of a key being 2 * #{splitters below it} + [it equals a splitter] -- what
 * b200rs_bucket_ids computes, here evaluated inside the kernel so no id array is written or read.
 * h_splitters: HOST array, strictly increasing bit-ordered values (num_splitters <= 15).  h_bucket_offsets: HOST array
 * of 2 * num_splitters + 1 exclusive output offsets (the caller knows the bucket sizes from the select rounds).
 * Two-phase temp-storage query like b200rs_sort.  Compiled for 4- and 8-byte keys with 0-, 4- or 8-byte values and
 * fewer than 2^30 items; anything else returns cudaErrorNotSupported (callers fall back to b200rs_bucket_ids +
 * b200rs_sort on the ids).  Input arrays are not modified.
 */
B200RS_API int b200rs_partition_by_splitters(
  void* d_temp_storage,
  size_t* temp_storage_bytes,
  const void* d_keys_in,
  void* d_keys_out,
  const void* d_values_in,
  void* d_values_out,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int descending,
  const uint64_t* h_splitters,
  int num_splitters,
  const uint64_t* h_bucket_offsets,
  b200rs_stream_t stream);

/*
 * The partition pass FUSED with the exchange: the same launch as b200rs_partition_by_splitters, but the partitioned
 * order is cut into num_dests contiguous segments (h_segment_ends[r] = first partitioned index that does not go to
 * ranks <= r; num_dests - 1 entries) and segment r is stored straight into rank r's receive buffer over NVLink.
 * h_rank_dst_keys[r] / h_rank_dst_vals[r] (HOST arrays of device addresses valid on THIS GPU, e.g. peer-mapped
 * symmetric memory) are biased: the item with partitioned index idx is stored at address + idx * item size.
 * The caller orders the launch between two cross-GPU barriers (receive buffers free / all stores landed).
 * One kernel does the compute step and the collective that follows it (SURVEY.md 8e steps 3 + 5).
 */
B200RS_API int b200rs_partition_to_peers(
  void* d_temp_storage,
  size_t* temp_storage_bytes,
  const void* d_keys_in,
  const void* d_values_in,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int descending,
  const uint64_t* h_splitters,
  int num_splitters,
  const uint64_t* h_bucket_offsets,
  int num_dests,
  const uint64_t* h_segment_ends,
  const uint64_t* h_rank_dst_keys,
  const uint64_t* h_rank_dst_vals,
  b200rs_stream_t stream);

/*
 * One-box multi-GPU stable sort (one process per GPU), SURVEY.md 8e / 8(b)-2.
 *
 * Replaces cudax::sort over a communicator
 * (/root/reference/cudax/include/cuda/experimental/__multi_gpu/algorithm/sort/sort.h:96-127; protocol hss/execute.h:56-129),
 * with the C-ABI shape of /root/reference/c/parallel/include/cccl/c/radix_sort.h:110-124.  Semantics: the concatenation
 * of the per-rank outputs in rank order equals ONE stable cub::DeviceRadixSort::SortPairs[Descending] of the
 * concatenation of the per-rank inputs in rank order; rank r's output has as many items as its input; inputs are not
 * modified.  (The reference's multi-GPU sort is keys-only and unstable; this one is stable and takes values.)
 *
 * b200rs_multi_comm_create is COLLECTIVE: every rank calls it with the same `world` and `receive_bytes` (>= the largest
 * shard's num_items * max(key_bytes, value_bytes) of any later sort) and an all-gather over its own transport
 * (MPI, torch.distributed, sockets ...): allgather(ctx, send, recv, n) must place the n bytes `send` of rank r at
 * recv + r * n on every rank and return 0.  It is only used here (CUDA IPC handles); a sort itself makes no host-side
 * communication, no NCCL call and no host wait: it enqueues a fixed sequence of kernels on `stream`, the small
 * all-reduces of the splitter selection being done by those kernels over peer-mapped memory (NVLink) and the key/value
 * exchange being fused into the partition kernel.
 *
 * b200rs_sort_multi is COLLECTIVE over the communicator (same key/value widths and order on every rank, any
 * num_items per rank, zero included); calls on one communicator must be stream-ordered on each rank.  Two-phase
 * temp-storage query like b200rs_sort.  4- and 8-byte keys with 0-, 4- or 8-byte values, < 2^30 items per rank, up to 16
 * ranks; other shapes return cudaErrorNotSupported (801).  world == 1 is a plain b200rs_sort.
 * b200rs_multi_status waits for the device and returns (and clears) the sticky device-side status: 0 ok, 1 a peer
 * never signalled (timeout), 2 some rank's shard exceeds receive_bytes, 4 internal inconsistency.
 */
typedef struct b200rs_multi_comm b200rs_multi_comm;
typedef int (*b200rs_allgather_fn)(void* ctx, const void* send, void* recv, size_t bytes_per_rank);

B200RS_API int b200rs_multi_comm_create(
  b200rs_multi_comm** comm, int rank, int world, size_t receive_bytes, b200rs_allgather_fn allgather, void* allgather_ctx);
B200RS_API int b200rs_multi_comm_destroy(b200rs_multi_comm* comm);
B200RS_API int b200rs_multi_status(b200rs_multi_comm* comm, int* status);
B200RS_API int b200rs_multi_last_launch_count(b200rs_multi_comm* comm);
/* Per-phase device time of the next sorts on this communicator (bench.py): while enabled b200rs_sort_multi records an
 * event on `stream` at every phase boundary; b200rs_multi_timing_read waits for the last sort and writes
 * ms4 = {splitter selection, fused partition + exchange, cross-GPU barrier, final local sort}. */
B200RS_API int b200rs_multi_timing_enable(b200rs_multi_comm* comm, int on);
B200RS_API int b200rs_multi_timing_read(b200rs_multi_comm* comm, float* ms4);
B200RS_API int b200rs_sort_multi(
  b200rs_multi_comm* comm,
  void* d_temp_storage,
  size_t* temp_storage_bytes,
  const void* d_keys_in,
  void* d_keys_out,
  const void* d_values_in,
  void* d_values_out,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int descending,
  b200rs_stream_t stream);

/* Number of kernel launches / async ops the last b200rs_sort call on this host thread enqueued
 * (bench.py's `gpu_launches`).  Thread-local. */
B200RS_API int b200rs_last_launch_count(void);

/* Per-op device timing for the calling host thread (bench.py's roofline leg).  While enabled, b200rs_sort records a
 * CUDA event on `stream` before each op it enqueues and one after the last.  b200rs_timing_read waits for the last
 * sort call of this thread and returns how many ops it wrote: kinds[i] in {0 memset, 1 upsweep histogram, 2 bin scan,
 * 3 onesweep pass, 4 copy}, ms[i] = device milliseconds between consecutive events.  Do not enable under capture. */
B200RS_API int b200rs_timing_enable(int on);
B200RS_API int b200rs_timing_read(int* kinds, float* ms, int capacity);

/* Tuning/diagnostic override: force the onesweep tile configuration index for subsequent calls from this
 * process (-1 = automatic).  Used only by tools/sweep and tests; not part of the drop-in surface. */
B200RS_API int b200rs_set_config(int config_index);

/* Diagnostic: cap the number of items one onesweep launch handles (0 = default, < 2^30) so tests can exercise the
 * multi-portion path (reference: portion_size, dispatch_radix_sort.cuh:1710-1716) at small N. */
B200RS_API int b200rs_set_portion_items(unsigned long long items);

/* Diagnostic: force the 64-bit-offset kernel variants (normally selected only for arrays of >= 2^32 items, the
 * reference's 64-bit OffsetT case, detail/choose_offset.cuh:35-52) so tests can exercise them at small N. */
B200RS_API int b200rs_set_force_big(int on);

/* Diagnostic: 0 routes inputs of at most one tile (which normally take the single-CTA kernel that replaces
 * DeviceRadixSortSingleTileKernel, kernel_radix_sort.cuh:330-434) through the general multi-kernel path as well, so
 * tests can compare the two; 1 (default) restores the single-CTA kernel. */
B200RS_API int b200rs_set_single_tile(int on);

/* Tuning/diagnostic: inputs of at most `items` items (4- or 8-byte keys with 0-, 4- or 8-byte values) are sorted by ONE
 * cooperative launch that runs every phase of the general path between grid-wide barriers (the latency path; the
 * reference uses programmatic dependent launch for the same regime, dispatch_radix_sort.cuh:1755-1756).  Default
 * (items == ~0): up to 4 MiB of keys (2^20 4-byte / 2^19 8-byte keys, the measured break-even on B200); 0 sends
 * everything above one tile through the general multi-kernel path. */
B200RS_API int b200rs_set_small_max(unsigned long long items);
/* b200rs_segmented_sort: segments longer than this many items are sorted by whole-grid passes over all long segments at
 * once, shorter ones by one CTA per segment (0 = always one CTA per segment).  Tuning / test hook, process-wide; the
 * temp-storage query and the sort must see the same value. */
B200RS_API int b200rs_set_segmented_long_min(unsigned long long items);
/* b200rs_segmented_sort: segments of at most this many items (at most 256) are sorted by one warp each instead of one
 * CTA each; 0 = never.  Tuning / test hook, process-wide. */
B200RS_API int b200rs_set_segmented_tiny_max(unsigned long long items);
/* b200rs_topk: inputs of at most this many BYTES of keys are selected by one cooperative launch (the rounds of the radix
 * select between grid-wide barriers) instead of one launch per round; default 32 MiB (the measured crossover on B200), 0 = never.  Tuning / test hook. */
B200RS_API int b200rs_set_topk_small_max(unsigned long long key_bytes_total);

/* Human-readable description of configuration `config_index` for (key_bytes, value_bytes); returns the number of
 * configurations available when config_index < 0.  buf may be NULL. */
B200RS_API int b200rs_describe_config(int key_bytes, int value_bytes, int config_index, char* buf, size_t buf_len);

#ifdef __cplusplus
}
#endif
#endif /* B200RS_H_ */
