// Host orchestration + C ABI (include/b200rs.h).
//
// Replaces detail::radix_sort::dispatch / dispatch_impl::invoke / invoke_onesweep / invoke_copy
// (/root/reference/cub/cub/device/dispatch/dispatch_radix_sort.cuh:2016-2065, :1950-2005, :1698-1948, :1364-1410)
// and the temp-blob rule of detail::alias_temporaries (cub/cub/util_temporary_storage.cuh:49-88).
//
// Stream ops of one sort call:  1 memset (bins + tile counters + first look-back array)
//                               1 upsweep histogram kernel (reads the keys once, all passes)
//                               1 bin-scan kernel
//                               passes x portions onesweep kernels (each zeroes the NEXT launch's look-back array)
// No allocation, no host synchronisation, legal under stream capture.
#include "../../include/b200rs.h"
#include "configs.h"

#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdlib>
#include <cstdio>
#include <cstring>

namespace b200rs
{

static std::atomic<int> g_config_override{-1};
static std::atomic<unsigned long long> g_portion_override{0};
static std::atomic<bool> g_force_big{false};
static std::atomic<bool> g_no_single_tile{false};
// inputs of at most this many KEY BYTES take the one-launch cooperative kernel (small.cu); measured on B200 with
// tools/ubench/small_crossover.py: 2.0-2.3x faster than the general path up to 2^16 items, level at 4 MiB of keys
// (2^20 4-byte keys / 2^19 8-byte keys), behind from there on.  b200rs_set_small_max overrides it in items.
static std::atomic<unsigned long long> g_small_max{~0ull};
constexpr unsigned long long SMALL_MAX_KEY_BYTES = 4ull << 20;
static thread_local int t_last_launches = 0;

// Optional per-op device timing (bench.py's roofline leg): when enabled, an event is recorded on the stream before
// every op of b200rs_sort and after the last one.  Off by default; never enabled under stream capture by callers.
constexpr int MAX_TIMED_OPS = 96;
static thread_local bool t_timing_on = false;
static thread_local cudaEvent_t t_events[MAX_TIMED_OPS + 1];
static thread_local int t_event_kinds[MAX_TIMED_OPS];
static thread_local int t_events_made = 0;
static thread_local int t_events_used = 0;

enum OpKind
{
  OP_MEMSET    = 0,
  OP_HISTOGRAM = 1,
  OP_SCAN      = 2,
  OP_ONESWEEP  = 3,
  OP_COPY      = 4,
  OP_SINGLE    = 5,
  OP_SMALL     = 6
};

// B200RS_DEBUG_SYNC=1 in the environment (the counterpart of the reference's CUB_DEBUG_SYNC / CubDebug logging,
// /root/reference/cub/cub/util_debug.cuh:37-72): every stream operation of b200rs_sort is followed by a stream
// synchronisation and its status is logged to stderr.  Ignored while the stream is being captured.
static int debug_sync_level()
{
  static const int level = [] {
    const char* v = getenv("B200RS_DEBUG_SYNC");
    return v != nullptr ? atoi(v) : 0;
  }();
  return level;
}
static thread_local int t_prev_kind = -1;
static void debug_sync(cudaStream_t stream)
{
  if (debug_sync_level() <= 0 || t_prev_kind < 0)
  {
    return;
  }
  static const char* const names[] = {"memset", "histogram", "scan", "onesweep", "copy", "single_tile", "small_sort"};
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone)
  {
    const cudaError_t e = cudaStreamSynchronize(stream);
    fprintf(stderr, "b200rs_sort: op %d (%s) on stream %p: %s\n", t_last_launches, names[t_prev_kind], (void*) stream,
            cudaGetErrorString(e));
  }
  t_prev_kind = -1;
}

static void mark_op(cudaStream_t stream, int kind)
{
  debug_sync(stream);
  t_prev_kind = kind;
  t_last_launches++;
  if (!t_timing_on || t_events_used >= MAX_TIMED_OPS)
  {
    return;
  }
  while (t_events_made <= t_events_used + 1 && t_events_made <= MAX_TIMED_OPS)
  {
    if (cudaEventCreate(&t_events[t_events_made]) != cudaSuccess)
    {
      return;
    }
    t_events_made++;
  }
  t_event_kinds[t_events_used] = kind;
  cudaEventRecord(t_events[t_events_used], stream);
  t_events_used++;
}

static void mark_end(cudaStream_t stream)
{
  debug_sync(stream);
  if (t_timing_on && t_events_used > 0 && t_events_used < t_events_made)
  {
    cudaEventRecord(t_events[t_events_used], stream);
  }
}

static const OnesweepConfig* configs_for(int key_bytes, int value_bytes, int* count)
{
  switch (key_bytes)
  {
    case 1:
      return onesweep_configs_k1(value_bytes, count);
    case 2:
      return onesweep_configs_k2(value_bytes, count);
    case 4:
      return value_bytes == 0   ? onesweep_configs_k4_v0(count)
             : value_bytes == 4 ? onesweep_configs_k4_v4(count)
                                : onesweep_configs_k4_vx(value_bytes, count);
    case 8:
      return value_bytes == 0   ? onesweep_configs_k8_v0(count)
             : value_bytes == 4 ? onesweep_configs_k8_v4(count)
                                : onesweep_configs_k8_vx(value_bytes, count);
    default:
      *count = 0;
      return nullptr;
  }
}

// Per-call tuning (b200rs_sort_tuned): consulted before the process-wide diagnostic switches; only set for the
// duration of one call on the calling thread.
static thread_local const b200rs_tuning* t_tuning = nullptr;

static int effective_config_override()
{
  return (t_tuning != nullptr && t_tuning->config_index >= 0) ? t_tuning->config_index
                                                              : g_config_override.load(std::memory_order_relaxed);
}
static unsigned long long effective_small_max()
{
  return (t_tuning != nullptr && t_tuning->small_max_items >= 0) ? (unsigned long long) t_tuning->small_max_items
                                                                 : g_small_max.load(std::memory_order_relaxed);
}
static bool effective_no_single_tile()
{
  return (t_tuning != nullptr && t_tuning->single_tile >= 0) ? t_tuning->single_tile == 0
                                                             : g_no_single_tile.load(std::memory_order_relaxed);
}

static const OnesweepConfig* pick_config(int key_bytes, int value_bytes)
{
  int count                 = 0;
  const OnesweepConfig* tab = configs_for(key_bytes, value_bytes, &count);
  if (tab == nullptr || count == 0)
  {
    return nullptr;
  }
  int idx = effective_config_override();
  if (idx < 0 || idx >= count)
  {
    idx = 0;
  }
  return tab + idx;
}

// the classic kernel used when a bulk-store configuration meets output pointers that are not 16-byte aligned
static const OnesweepConfig* fallback_config(int key_bytes, int value_bytes)
{
  int count                 = 0;
  const OnesweepConfig* tab = configs_for(key_bytes, value_bytes, &count);
  for (int i = 0; tab != nullptr && i < count; ++i)
  {
    if (!tab[i].bulk_store)
    {
      return tab + i;
    }
  }
  return nullptr;
}

struct PortionPlan
{
  uint64_t tile, portion_items, portions, max_tiles;
};

static PortionPlan plan_portions(uint64_t num_items, uint64_t tile)
{
  PortionPlan p;
  p.tile                    = tile;
  uint64_t portion_cap      = (uint64_t(1) << 30) - 1;
  const uint64_t portion_ov = g_portion_override.load(std::memory_order_relaxed);
  if (portion_ov != 0 && portion_ov < portion_cap)
  {
    portion_cap = portion_ov;
  }
  p.portion_items = (portion_cap / tile > 0 ? portion_cap / tile : 1) * tile;
  p.portions      = (num_items + p.portion_items - 1) / p.portion_items;
  p.max_tiles     = ((num_items < p.portion_items ? num_items : p.portion_items) + tile - 1) / tile;
  return p;
}

static int sm_count_of_current_device(int* out)
{
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
  {
    return int(e);
  }
  if (dev >= 0 && dev < 64)
  {
    int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0)
    {
      *out = c;
      return 0;
    }
  }
  int sms = 0;
  e       = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess)
  {
    return int(e);
  }
  if (dev >= 0 && dev < 64)
  {
    cache[dev].store(sms, std::memory_order_relaxed);
  }
  *out = sms;
  return 0;
}

static inline size_t align_up(size_t x, size_t a)
{
  return (x + a - 1) / a * a;
}

struct TempLayout
{
  size_t off_bins, off_ctrs, off_lb0, off_lb1, off_keys, off_vals, control_bytes, total;
};

} // namespace b200rs

using namespace b200rs;

extern "C" {

int b200rs_version(void)
{
  return B200RS_VERSION;
}

int b200rs_last_launch_count(void)
{
  return t_last_launches;
}

int b200rs_sort_tuned(
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
  const b200rs_tuning* tuning)
{
  const b200rs_tuning* saved = t_tuning;
  t_tuning                   = tuning;
  const int rc = b200rs_sort(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items,
                             key_kind, key_bytes, value_bytes, begin_bit, end_bit, descending, is_overwrite_okay, selector,
                             stream);
  t_tuning     = saved;
  return rc;
}

int b200rs_set_config(int config_index)
{
  g_config_override.store(config_index, std::memory_order_relaxed);
  return 0;
}

// diagnostic: cap the number of items one onesweep launch handles (0 = default, < 2^30) so tests can
// exercise the multi-portion path (reference: portion_size, dispatch_radix_sort.cuh:1710-1716) at small N
B200RS_API int b200rs_set_portion_items(unsigned long long items)
{
  g_portion_override.store(items, std::memory_order_relaxed);
  return 0;
}

// diagnostic: force the 64-bit-offset kernels (normally only used for arrays of >= 2^32 items) so tests can
// exercise them at small N
B200RS_API int b200rs_set_force_big(int on)
{
  g_force_big.store(on != 0, std::memory_order_relaxed);
  return 0;
}

// diagnostic: route inputs of at most one tile through the general path as well (tests compare both)
B200RS_API int b200rs_set_small_max(unsigned long long items)
{
  g_small_max.store(items, std::memory_order_relaxed);
  return 0;
}

int b200rs_set_single_tile(int on)
{
  g_no_single_tile.store(on == 0, std::memory_order_relaxed);
  return 0;
}

int b200rs_describe_config(int key_bytes, int value_bytes, int config_index, char* buf, size_t buf_len)
{
  int count                 = 0;
  const OnesweepConfig* tab = configs_for(key_bytes, value_bytes, &count);
  if (config_index < 0 || tab == nullptr)
  {
    return count;
  }
  if (config_index >= count)
  {
    return -1;
  }
  if (buf != nullptr && buf_len > 0)
  {
    const OnesweepConfig& c = tab[config_index];
    snprintf(buf, buf_len, "k%dv%d threads=%d items=%d minb=%d tile=%d smem=%zu opt=%d%s", key_bytes, value_bytes,
             c.threads, c.items_per_thread, c.min_blocks, c.tile_items, c.smem_bytes, c.opt, c.bulk_store ? " tma" : "");
  }
  return count;
}

int b200rs_digit_histogram(
  const void* d_keys_in,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int begin_bit,
  int end_bit,
  int descending,
  uint64_t* d_bins,
  b200rs_stream_t stream_)
{
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (key_kind < 0 || key_kind > 2 || (key_bytes != 1 && key_bytes != 2 && key_bytes != 4 && key_bytes != 8)
      || begin_bit < 0 || end_bit < begin_bit || end_bit > key_bytes * 8 || d_bins == nullptr
      || (key_kind == 2 && key_bytes < 2))
  {
    return int(cudaErrorInvalidValue);
  }
  const int passes = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
  if (passes == 0)
  {
    return 0;
  }
  int sms = 0;
  if (int e = sm_count_of_current_device(&sms))
  {
    return e;
  }
  cudaError_t e = cudaMemsetAsync(d_bins, 0, size_t(passes) * RADIX * sizeof(uint64_t), stream);
  if (e != cudaSuccess)
  {
    return int(e);
  }
  if (num_items == 0)
  {
    return 0;
  }
  const KeyXform xf = make_xform(key_kind, key_bytes, descending);
  return int(launch_histogram(d_keys_in, num_items, key_bytes, reinterpret_cast<unsigned long long*>(d_bins), passes,
                              begin_bit, end_bit, xf, sms, stream));
}

int b200rs_splitter_ranks(
  const void* d_sorted_keys,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int descending,
  const void* d_splitters,
  int num_splitters,
  uint64_t* d_lt,
  uint64_t* d_eq,
  b200rs_stream_t stream_)
{
  if (key_kind < 0 || key_kind > 2 || num_splitters < 0 || (key_kind == 2 && key_bytes < 2))
  {
    return int(cudaErrorInvalidValue);
  }
  const KeyXform xf = make_xform(key_kind, key_bytes, descending);
  return int(launch_splitter_ranks(d_sorted_keys, num_items, key_bytes, xf, d_splitters, num_splitters,
                                   reinterpret_cast<unsigned long long*>(d_lt),
                                   reinterpret_cast<unsigned long long*>(d_eq),
                                   reinterpret_cast<cudaStream_t>(stream_)));
}

int b200rs_sort(
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
  b200rs_stream_t stream_)
{
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  t_last_launches     = 0;
  t_events_used       = 0;
  t_prev_kind         = -1;
  // an NVTX range around the execute call only, as the reference (_CCCL_NVTX_RANGE_SCOPE_IF(d_temp_storage, ...),
  // device_radix_sort.cuh:424); costs nothing while no tool is attached
  struct nvtx_scope
  {
    bool on;
    explicit nvtx_scope(bool enable)
        : on(enable)
    {
      if (on)
      {
        nvtxRangePushA("b200rs_sort");
      }
    }
    ~nvtx_scope()
    {
      if (on)
      {
        nvtxRangePop();
      }
    }
  } nvtx_range(d_temp_storage != nullptr);
  if (temp_storage_bytes == nullptr || key_kind < 0 || key_kind > 2)
  {
    return int(cudaErrorInvalidValue);
  }
  if ((key_bytes != 1 && key_bytes != 2 && key_bytes != 4 && key_bytes != 8)
      || (value_bytes != 0 && value_bytes != 1 && value_bytes != 2 && value_bytes != 4 && value_bytes != 8
          && value_bytes != 16)
      || (key_kind == 2 && key_bytes < 2))
  {
    return int(cudaErrorNotSupported);
  }
  if (begin_bit < 0 || end_bit < begin_bit || end_bit > key_bytes * 8)
  {
    return int(cudaErrorInvalidValue);
  }
  const bool overwrite = is_overwrite_okay != 0;
  const bool query     = d_temp_storage == nullptr;

  // empty problem, or nothing to sort on and both buffers are the caller's scratch anyway
  // (dispatch_radix_sort.cuh:1956-1963)
  if (num_items == 0 || (begin_bit == end_bit && overwrite))
  {
    if (query)
    {
      *temp_storage_bytes = 1;
    }
    else if (selector != nullptr)
    {
      *selector = 0;
    }
    return 0;
  }
  if (!query)
  {
    if (d_keys_in == nullptr || d_keys_out == nullptr
        || (value_bytes > 0 && (d_values_in == nullptr || d_values_out == nullptr)))
    {
      return int(cudaErrorInvalidValue);
    }
    const size_t valign = value_bytes > 8 ? 8 : size_t(value_bytes);
    if ((reinterpret_cast<size_t>(d_keys_in) | reinterpret_cast<size_t>(d_keys_out)) % size_t(key_bytes) != 0
        || (value_bytes > 0
            && (reinterpret_cast<size_t>(d_values_in) | reinterpret_cast<size_t>(d_values_out)) % valign != 0))
    {
      return int(cudaErrorMisalignedAddress);
    }
  }

  // pointer API with an empty bit range: plain copy (dispatch_radix_sort.cuh:1966-1977, :1364-1410)
  if (begin_bit == end_bit)
  {
    if (query)
    {
      *temp_storage_bytes = 1;
      return 0;
    }
    mark_op(stream, OP_COPY);
    cudaError_t e = cudaMemcpyAsync(d_keys_out, d_keys_in, size_t(num_items) * key_bytes, cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess && value_bytes > 0)
    {
      mark_op(stream, OP_COPY);
      e = cudaMemcpyAsync(d_values_out, d_values_in, size_t(num_items) * value_bytes, cudaMemcpyDeviceToDevice, stream);
    }
    mark_end(stream);
    if (e == cudaSuccess && selector != nullptr)
    {
      *selector = 1;
    }
    return int(e);
  }

  // at most one tile: the whole sort is one launch of one CTA, no temp storage
  // (dispatch_radix_sort.cuh:1980, kernel_radix_sort.cuh:330-434)
  if (num_items <= single_tile_capacity(key_bytes, value_bytes) && !effective_no_single_tile())
  {
    if (query)
    {
      *temp_storage_bytes = 1;
      return 0;
    }
    const KeyXform sxf =
      make_xform(key_kind, key_bytes, descending, num_items <= reference_single_tile_items(key_bytes, value_bytes));
    mark_op(stream, OP_SINGLE);
    const cudaError_t se = launch_single_tile(d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, key_bytes,
                                              value_bytes, begin_bit, end_bit, sxf, stream);
    mark_end(stream);
    if (se == cudaSuccess && selector != nullptr)
    {
      *selector = 1;
    }
    return int(se);
  }

  const OnesweepConfig* cfg = pick_config(key_bytes, value_bytes);
  if (cfg == nullptr)
  {
    return int(cudaErrorNotSupported);
  }
  const int passes = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
  // A bulk-store configuration needs 16-byte aligned destinations; otherwise the classic kernel runs.  The temp blob
  // is sized for whichever of the two needs more, so the size query does not depend on the pointers.
  const OnesweepConfig* fb = cfg->bulk_store ? fallback_config(key_bytes, value_bytes) : nullptr;
  if (cfg->bulk_store && fb == nullptr)
  {
    return int(cudaErrorNotSupported);
  }
  PortionPlan plan       = plan_portions(num_items, uint64_t(cfg->tile_items));
  uint64_t size_portions = plan.portions;
  uint64_t size_tiles    = plan.max_tiles;
  // mid-size inputs: every phase of the sort in ONE cooperative launch (small.cu).  Not with a forced configuration,
  // forced 64-bit offsets or a portion override: those diagnostics are about the general path.
  const bool small = small_sort_supported(key_bytes, value_bytes) && passes <= 8
                  && (effective_small_max() == ~0ull ? num_items * uint64_t(key_bytes) <= SMALL_MAX_KEY_BYTES
                                                     : num_items <= effective_small_max())
                  && plan.portions == 1 && effective_config_override() < 0
                  && !g_force_big.load(std::memory_order_relaxed)
                  && g_portion_override.load(std::memory_order_relaxed) == 0;
  if (small)
  {
    fb            = nullptr;
    plan          = plan_portions(num_items, small_sort_tile_items());
    size_portions = plan.portions;
    size_tiles    = plan.max_tiles;
  }
  if (fb != nullptr)
  {
    const PortionPlan alt = plan_portions(num_items, uint64_t(fb->tile_items));
    size_portions         = alt.portions > size_portions ? alt.portions : size_portions;
    size_tiles            = alt.max_tiles > size_tiles ? alt.max_tiles : size_tiles;
    if (!query)
    {
      size_t bits = reinterpret_cast<size_t>(d_keys_out) | reinterpret_cast<size_t>(d_values_out);
      if (overwrite)
      {
        bits |= reinterpret_cast<size_t>(d_keys_in) | reinterpret_cast<size_t>(d_values_in);
      }
      if (bits % 16 != 0)
      {
        cfg  = fb;
        plan = alt;
      }
    }
  }
  const uint64_t tile          = plan.tile;
  const uint64_t portion_items = plan.portion_items;
  const uint64_t portions      = plan.portions;
  const bool need_tmp          = !overwrite && passes > 1;

  // temp blob: every sub-allocation 256-byte aligned, +255 so an unaligned blob can be aligned up
  // (same rule as util_temporary_storage.cuh:49-88)
  TempLayout L;
  size_t off = 0;
  L.off_bins = off;
  off += align_up(size_t(size_portions) * passes * RADIX * sizeof(unsigned long long), 256);
  L.off_ctrs = off;
  off += align_up((size_t(size_portions) * passes + 1) * sizeof(uint32_t), 256); // + the upsweep's zero flag
  L.off_lb0 = off;
  off += align_up(size_t(size_tiles) * RADIX * sizeof(uint32_t), 256);
  L.control_bytes = off;
  L.off_lb1       = off;
  off += align_up(size_t(size_tiles) * RADIX * sizeof(uint32_t), 256);
  L.off_keys = off;
  off += need_tmp ? align_up(size_t(num_items) * key_bytes, 256) : 0;
  L.off_vals = off;
  off += (need_tmp && value_bytes > 0) ? align_up(size_t(num_items) * value_bytes, 256) : 0;
  L.total = off + 255;

  if (query)
  {
    *temp_storage_bytes = L.total;
    return 0;
  }
  if (*temp_storage_bytes < L.total)
  {
    return int(cudaErrorInvalidValue);
  }

  unsigned char* base = reinterpret_cast<unsigned char*>(align_up(reinterpret_cast<size_t>(d_temp_storage), 256));
  unsigned long long* bins = reinterpret_cast<unsigned long long*>(base + L.off_bins);
  uint32_t* ctrs           = reinterpret_cast<uint32_t*>(base + L.off_ctrs);
  // floating-point keys: set by the upsweep when a key with the aliased zero pattern exists (PassArgs::zero_flag)
  uint32_t* zero_flag      = key_kind == 2 ? ctrs + size_t(size_portions) * passes : nullptr;
  uint32_t* lb[2]          = {reinterpret_cast<uint32_t*>(base + L.off_lb0), reinterpret_cast<uint32_t*>(base + L.off_lb1)};
  void* keys_tmp           = need_tmp ? base + L.off_keys : nullptr;
  void* vals_tmp           = (need_tmp && value_bytes > 0) ? base + L.off_vals : nullptr;

  int sms = 0;
  if (int e = sm_count_of_current_device(&sms))
  {
    return e;
  }
  const KeyXform xf =
    make_xform(key_kind, key_bytes, descending, num_items <= reference_single_tile_items(key_bytes, value_bytes));

  cudaError_t e = cudaSuccess;
  PassArgs small_pass[8];
  if (!small)
  {
    mark_op(stream, OP_MEMSET);
    e = cudaMemsetAsync(base, 0, L.control_bytes, stream);
    if (e != cudaSuccess)
    {
      return int(e);
    }
    // upsweep over the whole input: counts land in portion 0's bins, then become exclusive offsets
    mark_op(stream, OP_HISTOGRAM);
    e = launch_histogram(d_keys_in, num_items, key_bytes, bins, passes, begin_bit, end_bit, xf, sms, stream, zero_flag);
    if (e != cudaSuccess)
    {
      return int(e);
    }
    mark_op(stream, OP_SCAN);
    e = launch_scan_bins(bins, passes, stream);
    if (e != cudaSuccess)
    {
      return int(e);
    }
  }

  const uint64_t launches = uint64_t(passes) * portions;
  uint64_t launch_idx     = 0;
  const void* src_k       = d_keys_in;
  const void* src_v       = d_values_in;
  for (int pass = 0; pass < passes; ++pass)
  {
    void* dst_k;
    void* dst_v;
    if (overwrite)
    {
      // ping-pong between the caller's two buffers
      dst_k = (pass % 2 == 0) ? d_keys_out : const_cast<void*>(d_keys_in);
      dst_v = (pass % 2 == 0) ? d_values_out : const_cast<void*>(d_values_in);
    }
    else
    {
      // never write buffer 0; the last pass must land in buffer 1 (dispatch_radix_sort.cuh:1850-1857, :1935-1942)
      const bool to_out = ((passes - 1 - pass) % 2) == 0;
      dst_k             = to_out ? d_keys_out : keys_tmp;
      dst_v             = to_out ? d_values_out : vals_tmp;
    }
    const int bit   = begin_bit + pass * RADIX_BITS;
    const int nbits = (end_bit - bit) < RADIX_BITS ? (end_bit - bit) : RADIX_BITS;
    for (uint64_t portion = 0; portion < portions; ++portion, ++launch_idx)
    {
      const uint64_t first = portion * portion_items;
      const uint64_t count = (num_items - first) < portion_items ? (num_items - first) : portion_items;
      const unsigned tiles = unsigned((count + tile - 1) / tile);
      PassArgs a;
      a.keys_in       = static_cast<const unsigned char*>(src_k) + first * key_bytes;
      a.keys_out      = dst_k;
      a.vals_in       = value_bytes > 0 ? static_cast<const unsigned char*>(src_v) + first * value_bytes : nullptr;
      a.vals_out      = value_bytes > 0 ? dst_v : nullptr;
      a.lookback      = lb[launch_idx & 1];
      a.lookback_next = nullptr;
      a.lookback_next_tiles = 0;
      if (launch_idx + 1 < launches)
      {
        const uint64_t nportion = (portion + 1 < portions) ? portion + 1 : 0;
        const uint64_t nfirst   = nportion * portion_items;
        const uint64_t ncount   = (num_items - nfirst) < portion_items ? (num_items - nfirst) : portion_items;
        a.lookback_next         = lb[(launch_idx + 1) & 1];
        a.lookback_next_tiles   = uint32_t((ncount + tile - 1) / tile);
      }
      a.tile_counter = ctrs + (portion * passes + pass);
      a.bins         = bins + (portion * passes + pass) * RADIX;
      a.bins_next    = (portion + 1 < portions) ? bins + ((portion + 1) * passes + pass) * RADIX : nullptr;
      a.num_items    = uint32_t(count);
      a.num_tiles    = tiles;
      a.all_ones     = 0xffffffffu;
      a.shift        = bit;
      a.mask         = (1u << nbits) - 1u;
      // unsigned ascending keys: the transform is the identity, the kernels skip it altogether
      const bool identity = xf.float_mask == 0 && xf.sign_mask == 0 && xf.desc_mask == 0;
      a.first_pass   = pass == 0 && !identity;
      a.last_pass    = pass == passes - 1 && !identity;
      // 64-bit output offsets from 2^32 - 2^16 items on: the folded scatter biases its 32-bit offsets by one tile
      a.big          = ((num_items + 65536) >> 32) != 0 || g_force_big.load(std::memory_order_relaxed);
      a.xf           = xf;
      a.num_splitters = 0;
      a.peer          = nullptr;
      a.plan          = nullptr;
      a.sm_count      = sms;
      a.zero_flag     = small ? nullptr : zero_flag;
      if (small)
      {
        small_pass[pass] = a; // (one portion) launched together below
        continue;
      }
      mark_op(stream, OP_ONESWEEP);
      e = cfg->launch(a, tiles, stream);
      if (e != cudaSuccess)
      {
        return int(e);
      }
    }
    src_k = dst_k;
    src_v = dst_v;
  }
  if (small)
  {
    mark_op(stream, OP_SMALL);
    e = launch_small_sort(small_pass, passes, bins, unsigned(plan.max_tiles), key_bytes, value_bytes, sms, stream);
    if (e != cudaSuccess)
    {
      return int(e);
    }
  }
  mark_end(stream);
  if (selector != nullptr)
  {
    *selector = overwrite ? (passes & 1) : 1;
  }
  return 0;
}

} // extern "C"

// Host-built tables of the partition pass travel as KERNEL ARGUMENTS (copied by value at launch, so the call stays legal
// under stream capture and never reads pageable host memory asynchronously).
struct PartitionTables
{
  PeerTable peer;
  unsigned long long bucket_offsets[31];
  uint32_t num_buckets;
  uint32_t write_peer;
};
__global__ void write_partition_tables_kernel(const PartitionTables t, PeerTable* peer_out, unsigned long long* bins_out)
{
  if (t.write_peer != 0)
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&t.peer);
    uint32_t* dst       = reinterpret_cast<uint32_t*>(peer_out);
    for (uint32_t w = threadIdx.x; w < uint32_t(sizeof(PeerTable) / 4); w += blockDim.x)
    {
      dst[w] = src[w];
    }
  }
  if (threadIdx.x < t.num_buckets)
  {
    bins_out[threadIdx.x] = t.bucket_offsets[threadIdx.x];
  }
}

// The partition pass (one onesweep launch in bucket mode).  num_dests == 0: results go to d_keys_out / d_values_out;
// otherwise the partitioned order is cut at h_segment_ends and segment r is stored through the (biased) pointers
// h_rank_dst_keys[r] / h_rank_dst_vals[r] -- peer-mapped receive buffers of the destination GPUs.
static int partition_impl(
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
  int num_dests,
  const uint64_t* h_segment_ends,
  const uint64_t* h_rank_dst_keys,
  const uint64_t* h_rank_dst_vals,
  b200rs_stream_t stream_,
  const PartitionPlan* d_plan = nullptr) // device-side plan (multi.cu): replaces every h_* argument
{
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  t_last_launches     = 0;
  t_events_used       = 0;
  if (temp_storage_bytes == nullptr || key_kind < 0 || key_kind > 2 || num_splitters < 0 || num_splitters > 15
      || (key_kind == 2 && key_bytes < 2) || num_dests < 0 || num_dests > 16)
  {
    return int(cudaErrorInvalidValue);
  }
  const OnesweepConfig* cfg = nullptr;
  {
    int count                 = 0;
    const OnesweepConfig* tab = configs_for(key_bytes, value_bytes, &count);
    if (tab != nullptr && count > 0 && tab[0].launch_bucket != nullptr)
    {
      cfg = tab; // bucket mode is compiled for the default configuration only
    }
  }
  // one launch: the chained scan carries 30-bit counts
  if (cfg == nullptr || num_items >= (uint64_t(1) << 30))
  {
    return int(cudaErrorNotSupported);
  }
  const bool query     = d_temp_storage == nullptr;
  const uint64_t tile  = uint64_t(cfg->tile_items);
  const uint64_t tiles = (num_items + tile - 1) / tile;
  size_t off           = 0;
  const size_t off_bins = off;
  off += align_up(RADIX * sizeof(unsigned long long), 256);
  const size_t off_ctr = off;
  off += 256;
  const size_t off_lb = off;
  off += align_up(size_t(tiles > 0 ? tiles : 1) * RADIX * sizeof(uint32_t), 256);
  const size_t control = off; // everything up to here is zeroed
  const size_t off_peer = off;
  off += align_up(sizeof(PeerTable), 256);
  const size_t total = off + 255;
  if (query)
  {
    *temp_storage_bytes = total;
    return 0;
  }
  if (*temp_storage_bytes < total)
  {
    return int(cudaErrorInvalidValue);
  }
  if (num_items == 0)
  {
    return 0;
  }
  const bool remote = num_dests > 0;
  if (d_plan != nullptr)
  {
    if (d_keys_in == nullptr || (value_bytes > 0 && d_values_in == nullptr) || !remote)
    {
      return int(cudaErrorInvalidValue);
    }
  }
  else if (d_keys_in == nullptr || (value_bytes > 0 && d_values_in == nullptr) || (num_splitters > 0 && h_splitters == nullptr)
      || h_bucket_offsets == nullptr
      || (!remote && (d_keys_out == nullptr || (value_bytes > 0 && d_values_out == nullptr)))
      || (remote && (h_rank_dst_keys == nullptr || (value_bytes > 0 && h_rank_dst_vals == nullptr)
                     || (num_dests > 1 && h_segment_ends == nullptr))))
  {
    return int(cudaErrorInvalidValue);
  }
  unsigned char* base = reinterpret_cast<unsigned char*>(align_up(reinterpret_cast<size_t>(d_temp_storage), 256));
  int sms             = 0;
  if (int e = sm_count_of_current_device(&sms))
  {
    return e;
  }
  mark_op(stream, OP_MEMSET);
  cudaError_t e = cudaMemsetAsync(base, 0, control, stream);
  if (e != cudaSuccess)
  {
    return int(e);
  }
  PartitionTables tables;
  memset(&tables, 0, sizeof(tables));
  if (remote && d_plan == nullptr)
  {
    // which rank each bucket goes to (0 = a segment boundary falls inside it: resolved per item in the kernel)
    PeerTable& pt = tables.peer;
    tables.write_peer = 1;
    pt.num_dests = uint32_t(num_dests);
    for (int r = 0; r < num_dests; ++r)
    {
      pt.rank_dst_keys[r] = h_rank_dst_keys[r];
      pt.rank_dst_vals[r] = value_bytes > 0 ? h_rank_dst_vals[r] : 0;
      pt.seg_end[r]       = r + 1 < num_dests ? uint32_t(h_segment_ends[r]) : 0xffffffffu;
    }
    const int nb = 2 * num_splitters + 1;
    for (int b = 0; b < nb; ++b)
    {
      const uint64_t lo = h_bucket_offsets[b];
      const uint64_t hi = b + 1 < nb ? h_bucket_offsets[b + 1] : num_items;
      int r_lo = 0, r_hi = 0;
      for (int r = 0; r + 1 < num_dests; ++r)
      {
        r_lo += h_segment_ends[r] <= lo ? 1 : 0;
        r_hi += (hi > lo && h_segment_ends[r] < hi) ? 1 : 0;
      }
      const bool one_rank     = hi <= lo || r_lo == r_hi;
      pt.bucket_dst_keys[b]   = one_rank ? pt.rank_dst_keys[r_lo] : 0;
      pt.bucket_dst_vals[b]   = one_rank ? pt.rank_dst_vals[r_lo] : 0;
    }
  }
  // exclusive bucket offsets (2 * num_splitters + 1 of them) -> the pass's per-digit output offsets
  if (d_plan == nullptr)
  {
    tables.num_buckets = uint32_t(2 * num_splitters + 1);
    for (uint32_t b = 0; b < tables.num_buckets; ++b)
    {
      tables.bucket_offsets[b] = h_bucket_offsets[b];
    }
    write_partition_tables_kernel<<<1, 256, 0, stream>>>(
      tables, reinterpret_cast<PeerTable*>(base + off_peer), reinterpret_cast<unsigned long long*>(base + off_bins));
    e = cudaPeekAtLastError();
    if (e != cudaSuccess)
    {
      return int(e);
    }
  }
  PassArgs a;
  a.keys_in             = d_keys_in;
  a.keys_out            = d_keys_out;
  a.vals_in             = value_bytes > 0 ? d_values_in : nullptr;
  a.vals_out            = value_bytes > 0 ? d_values_out : nullptr;
  a.lookback            = reinterpret_cast<uint32_t*>(base + off_lb);
  a.lookback_next       = nullptr;
  a.lookback_next_tiles = 0;
  a.tile_counter        = reinterpret_cast<uint32_t*>(base + off_ctr);
  a.bins                = d_plan != nullptr ? d_plan->bins : reinterpret_cast<unsigned long long*>(base + off_bins);
  a.bins_next           = nullptr;
  a.num_items           = uint32_t(num_items);
  a.num_tiles           = unsigned(tiles);
  a.all_ones            = 0xffffffffu;
  a.shift               = 0;
  a.mask                = 0xffu;
  a.first_pass          = 1;
  a.last_pass           = 1;
  a.big                 = 0;
  a.xf                  = make_xform(key_kind, key_bytes, descending);
  a.num_splitters       = d_plan != nullptr ? 0 : num_splitters;
  for (int i = 0; i < num_splitters && d_plan == nullptr; ++i)
  {
    a.splitters[i] = h_splitters[i];
  }
  a.peer     = d_plan != nullptr ? &d_plan->peer
                                 : (remote ? reinterpret_cast<const PeerTable*>(base + off_peer) : nullptr);
  a.plan     = d_plan;
  a.sm_count = sms;
  mark_op(stream, OP_ONESWEEP);
  e = cfg->launch_bucket(a, unsigned(tiles), stream);
  mark_end(stream);
  return int(e);
}

namespace b200rs
{
// multi.cu: the fused partition + exchange pass driven by a PartitionPlan in device memory
int partition_with_plan(void* d_temp_storage, size_t* temp_storage_bytes, const void* d_keys_in, const void* d_values_in,
                        uint64_t num_items, int key_kind, int key_bytes, int value_bytes, int descending, int num_dests,
                        const PartitionPlan* d_plan, cudaStream_t stream)
{
  return partition_impl(d_temp_storage, temp_storage_bytes, d_keys_in, nullptr, d_values_in, nullptr, num_items, key_kind,
                        key_bytes, value_bytes, descending, nullptr, 0, nullptr, num_dests, nullptr, nullptr, nullptr,
                        reinterpret_cast<b200rs_stream_t>(stream), d_plan);
}
} // namespace b200rs

extern "C" {

int b200rs_partition_by_splitters(
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
  b200rs_stream_t stream)
{
  return partition_impl(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items,
                        key_kind, key_bytes, value_bytes, descending, h_splitters, num_splitters, h_bucket_offsets, 0,
                        nullptr, nullptr, nullptr, stream);
}

int b200rs_partition_to_peers(
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
  b200rs_stream_t stream)
{
  if (num_dests < 1)
  {
    return int(cudaErrorInvalidValue);
  }
  return partition_impl(d_temp_storage, temp_storage_bytes, d_keys_in, nullptr, d_values_in, nullptr, num_items, key_kind,
                        key_bytes, value_bytes, descending, h_splitters, num_splitters, h_bucket_offsets, num_dests,
                        h_segment_ends, h_rank_dst_keys, h_rank_dst_vals, stream);
}

int b200rs_timing_enable(int on)
{
  t_timing_on   = on != 0;
  t_events_used = 0;
  return 0;
}

int b200rs_timing_read(int* kinds, float* ms, int capacity)
{
  const int n = t_events_used;
  if (n == 0)
  {
    return 0;
  }
  cudaError_t e = cudaEventSynchronize(t_events[n]);
  if (e != cudaSuccess)
  {
    return -int(e);
  }
  int written = 0;
  for (int i = 0; i < n && written < capacity; ++i, ++written)
  {
    float t = 0.f;
    cudaEventElapsedTime(&t, t_events[i], t_events[i + 1]);
    kinds[written] = t_event_kinds[i];
    ms[written]    = t;
  }
  return written;
}

} // extern "C"
