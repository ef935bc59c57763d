// In-place front door of the C ABI: what thrust::sort / thrust::sort_by_key on device memory do around the
// DoubleBuffer radix sort.
//
// Replaces thrust::cuda_cub::__radix_sort::radix_sort and the tail of __smart_sort::smart_sort
// (/root/reference/thrust/thrust/system/cuda/detail/sort.h:219-266, :288-339): query the temp size, obtain scratch
// for the alternate key/value buffers plus the temp blob in ONE allocation (keys and values rounded up to 128 bytes,
// :236-239), run the DoubleBuffer sort, copy the result back iff it landed in the scratch half (:252-264), release
// the scratch, optionally synchronise (:338).
//
// Difference from the reference: it goes to cudaMalloc/cudaFree on every call
// (thrust/system/cuda/detail/malloc_and_free.h:70,99), which synchronises the device twice per sort; here scratch
// comes from the device's stream-ordered pool (cudaMallocAsync/cudaFreeAsync on the caller's stream) with the pool's
// release threshold raised once, so repeated sorts reuse the same physical memory without a device synchronisation.
#include "../../include/b200rs.h"
#include "common.cuh"

#include <atomic>

namespace b200rs
{

static inline size_t round_up(size_t x, size_t a)
{
  return (x + a - 1) / a * a;
}

// Keep freed scratch cached in the device's default pool (otherwise the pool trims at every synchronisation and
// each call pays the page mapping again).  Once per device per process.
static cudaError_t keep_pool_warm(int dev)
{
  static std::atomic<unsigned long long> done{0};
  if (dev < 0 || dev >= 64 || (done.load(std::memory_order_relaxed) >> dev) & 1ull)
  {
    return cudaSuccess;
  }
  cudaMemPool_t pool;
  cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, dev);
  if (e != cudaSuccess)
  {
    return e;
  }
  unsigned long long threshold = ~0ull;
  e                            = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  if (e == cudaSuccess)
  {
    done.fetch_or(1ull << dev, std::memory_order_relaxed);
  }
  return e;
}

} // namespace b200rs

using namespace b200rs;

extern "C" B200RS_API int b200rs_sort_inplace(
  void* d_keys,
  void* d_values,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int descending,
  int synchronize,
  b200rs_stream_t stream_)
{
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (key_bytes <= 0 || value_bytes < 0)
  {
    return int(cudaErrorInvalidValue);
  }
  if (num_items == 0)
  {
    return 0;
  }
  if (d_keys == nullptr || (value_bytes > 0 && d_values == nullptr))
  {
    return int(cudaErrorInvalidValue);
  }
  const int end_bit = key_bytes * 8;
  size_t temp_bytes = 0;
  int rc = b200rs_sort(nullptr, &temp_bytes, d_keys, d_keys, d_values, d_values, num_items, key_kind, key_bytes,
                       value_bytes, 0, end_bit, descending, 1, nullptr, stream_);
  if (rc != 0)
  {
    return rc;
  }
  if (temp_bytes == 1)
  {
    // At most one tile (the only non-empty full-width sort with a 1-byte temp): the single-CTA kernel reads every item
    // before it writes any, so it sorts IN PLACE -- no scratch allocation, no copy back.
    int sel = 0;
    rc      = b200rs_sort(d_keys /* never dereferenced */, &temp_bytes, d_keys, d_keys, d_values, d_values, num_items,
                          key_kind, key_bytes, value_bytes, 0, end_bit, descending, 1, &sel, stream_);
    if (rc == 0 && synchronize)
    {
      rc = int(cudaStreamSynchronize(stream));
    }
    return rc;
  }
  const size_t keys_scratch = round_up(size_t(num_items) * key_bytes, 128);
  const size_t vals_scratch = round_up(size_t(num_items) * value_bytes, 128);
  const size_t total        = keys_scratch + vals_scratch + temp_bytes;

  int dev       = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess)
  {
    e = keep_pool_warm(dev);
  }
  unsigned char* scratch = nullptr;
  if (e == cudaSuccess)
  {
    e = cudaMallocAsync(reinterpret_cast<void**>(&scratch), total, stream);
  }
  if (e != cudaSuccess)
  {
    return int(e);
  }
  void* keys_alt = scratch;
  void* vals_alt = value_bytes > 0 ? scratch + keys_scratch : nullptr;
  void* temp     = scratch + keys_scratch + vals_scratch;
  int selector   = 0;
  rc = b200rs_sort(temp, &temp_bytes, d_keys, keys_alt, d_values, vals_alt, num_items, key_kind, key_bytes, value_bytes,
                   0, end_bit, descending, 1, &selector, stream_);
  if (rc == 0 && selector != 0)
  {
    e = cudaMemcpyAsync(d_keys, keys_alt, size_t(num_items) * key_bytes, cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess && value_bytes > 0)
    {
      e = cudaMemcpyAsync(d_values, vals_alt, size_t(num_items) * value_bytes, cudaMemcpyDeviceToDevice, stream);
    }
    rc = int(e);
  }
  const cudaError_t fe = cudaFreeAsync(scratch, stream);
  if (rc == 0)
  {
    rc = int(fe);
  }
  if (rc == 0 && synchronize)
  {
    rc = int(cudaStreamSynchronize(stream));
  }
  return rc;
}
