// Radix sort of user-defined keys described as a tuple of arithmetic members ("decomposer" keys) and of 128-bit
// integers.
//
// Replaces the decomposer overloads of cub::DeviceRadixSort
// (/root/reference/cub/cub/device/device_radix_sort.cuh:671,905,1359,1565 ..., radix_sort_with_decomposer :223-300;
// digit extraction over a tuple of members: cub/cub/block/radix_rank_sort_operations.cuh:436-527; 128-bit keys:
// cub/cub/util_type.cuh).  The reference runs its digit passes over the concatenated bits of the members, least
// significant member first.  LSD radix sort is a chain of stable sorts, so the same order is obtained here by chaining
// the existing passes member by member: a 32-bit permutation is sorted as the VALUE of each member's keys
// (b200rs_sort, pairs), least significant member first, and the key structs / user values are gathered once at the
// end.  Per member: one gather of that member through the current permutation + one (member, index) pair sort; the
// caller's arrays are read twice and written once whatever their item sizes (any key stride, any value size).
// The bit window [begin_bit, end_bit) counts from the least significant bit of the LAST member, as in the reference.
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/b200rs.h"
#include "common.cuh"

namespace b200rs
{

constexpr int FIELD_THREADS = 256;

// out[i] = member at byte offset `off` of item perm[i] (perm == nullptr: item i)
template <class U>
__global__ void __launch_bounds__(FIELD_THREADS)
gather_member_kernel(const unsigned char* items, uint32_t stride, uint32_t off, const uint32_t* perm, U* out, uint32_t n)
{
  for (uint32_t i = blockIdx.x * FIELD_THREADS + threadIdx.x; i < n; i += gridDim.x * FIELD_THREADS)
  {
    const size_t src = perm != nullptr ? perm[i] : i;
    out[i]           = *reinterpret_cast<const U*>(items + src * stride + off);
  }
}

__global__ void __launch_bounds__(FIELD_THREADS) iota_kernel(uint32_t* perm, uint32_t n)
{
  for (uint32_t i = blockIdx.x * FIELD_THREADS + threadIdx.x; i < n; i += gridDim.x * FIELD_THREADS)
  {
    perm[i] = i;
  }
}

// out[i] = in[perm[i]] for items of W-byte words (W = 16, 8, 4, 2 or 1 chosen from the item size and alignment)
template <class W>
__global__ void __launch_bounds__(FIELD_THREADS)
gather_items_kernel(const W* in, W* out, const uint32_t* perm, uint32_t n, uint32_t words_per_item)
{
  const unsigned long long total = (unsigned long long) n * words_per_item;
  for (unsigned long long j = blockIdx.x * (unsigned long long) FIELD_THREADS + threadIdx.x; j < total;
       j += (unsigned long long) gridDim.x * FIELD_THREADS)
  {
    const uint32_t i = uint32_t(j / words_per_item), w = uint32_t(j % words_per_item);
    out[j]           = in[size_t(perm[i]) * words_per_item + w];
  }
}

static cudaError_t gather_items(const void* in, void* out, const uint32_t* perm, uint32_t n, uint32_t item_bytes,
                                unsigned grid, cudaStream_t stream)
{
  const size_t bits = reinterpret_cast<size_t>(in) | reinterpret_cast<size_t>(out) | item_bytes;
  if (bits % 16 == 0)
  {
    gather_items_kernel<uint4><<<grid, FIELD_THREADS, 0, stream>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out),
                                                                  perm, n, item_bytes / 16);
  }
  else if (bits % 8 == 0)
  {
    gather_items_kernel<uint64_t><<<grid, FIELD_THREADS, 0, stream>>>(static_cast<const uint64_t*>(in),
                                                                     static_cast<uint64_t*>(out), perm, n, item_bytes / 8);
  }
  else if (bits % 4 == 0)
  {
    gather_items_kernel<uint32_t><<<grid, FIELD_THREADS, 0, stream>>>(static_cast<const uint32_t*>(in),
                                                                     static_cast<uint32_t*>(out), perm, n, item_bytes / 4);
  }
  else if (bits % 2 == 0)
  {
    gather_items_kernel<uint16_t><<<grid, FIELD_THREADS, 0, stream>>>(static_cast<const uint16_t*>(in),
                                                                     static_cast<uint16_t*>(out), perm, n, item_bytes / 2);
  }
  else
  {
    gather_items_kernel<uint8_t><<<grid, FIELD_THREADS, 0, stream>>>(static_cast<const uint8_t*>(in),
                                                                    static_cast<uint8_t*>(out), perm, n, item_bytes);
  }
  return cudaPeekAtLastError();
}

} // namespace b200rs

using namespace b200rs;

static size_t f_align(size_t x)
{
  return (x + 255) / 256 * 256;
}

extern "C" int b200rs_sort_fields(
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
  b200rs_stream_t stream_)
{
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (temp_storage_bytes == nullptr || fields == nullptr || num_fields < 1 || num_fields > 16 || key_stride_bytes < 1
      || value_bytes < 0)
  {
    return int(cudaErrorInvalidValue);
  }
  int total_bits = 0;
  for (int f = 0; f < num_fields; ++f)
  {
    const b200rs_key_field& m = fields[f];
    if ((m.bytes != 1 && m.bytes != 2 && m.bytes != 4 && m.bytes != 8) || m.kind < 0 || m.kind > 2
        || (m.kind == 2 && m.bytes < 2) || m.offset % m.bytes != 0 || m.offset + m.bytes > uint32_t(key_stride_bytes))
    {
      return int(cudaErrorInvalidValue);
    }
    total_bits += int(m.bytes) * 8;
  }
  if (begin_bit < 0 || end_bit < begin_bit || end_bit > total_bits)
  {
    return int(cudaErrorInvalidValue);
  }
  if (num_items >= (uint64_t(1) << 32))
  {
    return int(cudaErrorNotSupported); // the permutation is 32-bit
  }
  // temp blob: two member-key buffers, two permutation buffers, the pair sort's own temp storage (sized for 8-byte keys)
  size_t sort_bytes = 0;
  if (int rc = b200rs_sort(nullptr, &sort_bytes, nullptr, nullptr, nullptr, nullptr, num_items, 0, 8, 4, 0, 64, 0, 0,
                           nullptr, stream_))
  {
    return rc;
  }
  const size_t kbytes = f_align(size_t(num_items) * 8), pbytes = f_align(size_t(num_items) * 4);
  const size_t total  = 2 * kbytes + 2 * pbytes + sort_bytes + 255;
  if (d_temp_storage == nullptr)
  {
    *temp_storage_bytes = num_items == 0 ? 1 : total;
    return 0;
  }
  if (num_items == 0)
  {
    return 0;
  }
  if (*temp_storage_bytes < total)
  {
    return int(cudaErrorInvalidValue);
  }
  if (d_keys_in == nullptr || d_keys_out == nullptr || (value_bytes > 0 && (d_values_in == nullptr || d_values_out == nullptr)))
  {
    return int(cudaErrorInvalidValue);
  }
  unsigned char* base = reinterpret_cast<unsigned char*>(f_align(reinterpret_cast<size_t>(d_temp_storage)));
  void* mk[2]         = {base, base + kbytes};
  uint32_t* perm[2]   = {reinterpret_cast<uint32_t*>(base + 2 * kbytes), reinterpret_cast<uint32_t*>(base + 2 * kbytes + pbytes)};
  void* sort_temp     = base + 2 * kbytes + 2 * pbytes;
  const uint32_t n    = uint32_t(num_items);
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess)
  {
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (e != cudaSuccess)
  {
    return int(e);
  }
  const unsigned long long want = (num_items + FIELD_THREADS - 1) / FIELD_THREADS;
  const unsigned grid           = unsigned(want < (unsigned long long) sms * 16 ? want : (unsigned long long) sms * 16);
  const unsigned char* items    = static_cast<const unsigned char*>(d_keys_in);

  const uint32_t* cur = nullptr; // identity until the first member has been sorted
  int which           = 0;
  int low             = 0; // bit index (in the concatenated key) of the current member's least significant bit
  for (int f = num_fields - 1; f >= 0; --f)
  {
    const b200rs_key_field& m = fields[f];
    const int bits            = int(m.bytes) * 8;
    const int lo              = begin_bit > low ? begin_bit - low : 0;
    const int hi              = end_bit < low + bits ? end_bit - low : bits;
    low += bits;
    if (hi <= lo)
    {
      continue; // the window does not touch this member
    }
    switch (m.bytes)
    {
      case 1: gather_member_kernel<uint8_t><<<grid, FIELD_THREADS, 0, stream>>>(items, uint32_t(key_stride_bytes), m.offset, cur, static_cast<uint8_t*>(mk[0]), n); break;
      case 2: gather_member_kernel<uint16_t><<<grid, FIELD_THREADS, 0, stream>>>(items, uint32_t(key_stride_bytes), m.offset, cur, static_cast<uint16_t*>(mk[0]), n); break;
      case 4: gather_member_kernel<uint32_t><<<grid, FIELD_THREADS, 0, stream>>>(items, uint32_t(key_stride_bytes), m.offset, cur, static_cast<uint32_t*>(mk[0]), n); break;
      default: gather_member_kernel<uint64_t><<<grid, FIELD_THREADS, 0, stream>>>(items, uint32_t(key_stride_bytes), m.offset, cur, static_cast<uint64_t*>(mk[0]), n); break;
    }
    if ((e = cudaPeekAtLastError()) != cudaSuccess)
    {
      return int(e);
    }
    if (cur == nullptr)
    {
      iota_kernel<<<grid, FIELD_THREADS, 0, stream>>>(perm[which], n);
      if ((e = cudaPeekAtLastError()) != cudaSuccess)
      {
        return int(e);
      }
      cur = perm[which];
    }
    size_t sb = sort_bytes;
    if (int rc = b200rs_sort(sort_temp, &sb, mk[0], mk[1], cur, perm[which ^ 1], num_items, m.kind, int(m.bytes), 4, lo, hi,
                             descending, 0, nullptr, stream_))
    {
      return rc;
    }
    which ^= 1;
    cur = perm[which];
  }
  if (cur == nullptr) // empty window: the items are copied (dispatch_radix_sort.cuh:1966-1977)
  {
    e = cudaMemcpyAsync(d_keys_out, d_keys_in, size_t(num_items) * key_stride_bytes, cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess && value_bytes > 0)
    {
      e = cudaMemcpyAsync(d_values_out, d_values_in, size_t(num_items) * value_bytes, cudaMemcpyDeviceToDevice, stream);
    }
    return int(e);
  }
  e = gather_items(d_keys_in, d_keys_out, cur, n, uint32_t(key_stride_bytes), grid, stream);
  if (e == cudaSuccess && value_bytes > 0)
  {
    e = gather_items(d_values_in, d_values_out, cur, n, uint32_t(value_bytes), grid, stream);
  }
  return int(e);
}
