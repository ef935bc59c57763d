// cub::DeviceSegmentedRadixSort on top of libb200rs.so -- header-only drop-in for the arithmetic-key surface of
// /root/reference/cub/cub/device/device_segmented_radix_sort.cuh (cub 3.6.0).
//
//   reference overload (device_segmented_radix_sort.cuh)   pointer   DoubleBuffer
//   SortPairs                                               :234      :417
//   SortPairsDescending                                     :887      :1073
//   SortKeys                                                :1531     :1702
//   SortKeysDescending                                      :2139     :2308
// Same names, parameter order and defaults; every overload is ONE call of b200rs_segmented_sort (include/b200rs.h).
// Offsets: device pointers to 32- or 64-bit integers (the reference takes any random-access iterator; fancy iterators
// are a compile error here).  Semantics kept: d_temp_storage == nullptr => size query only; items outside every segment
// are neither read nor written (:59); the pointer overloads never write their inputs.  The DoubleBuffer overloads leave
// the result in Alternate() and flip the selectors (the reference flips them by the parity of its digit passes,
// dispatch_segmented_radix_sort.cuh; any buffer is allowed by its contract, :301-306) -- Current() after the call is
// the sorted data in both.
#pragma once

#include "device_radix_sort.cuh"

namespace cub
{
namespace detail
{
template <class OffsetIt>
struct b200rs_offset_bytes
{
  static_assert(std::is_pointer<OffsetIt>::value, "segment offsets must be device pointers to 32- or 64-bit integers");
  using T = std::remove_cv_t<std::remove_pointer_t<OffsetIt>>;
  static_assert(std::is_integral<T>::value && (sizeof(T) == 4 || sizeof(T) == 8),
                "segment offsets must be 32- or 64-bit integers");
  static constexpr int value = int(sizeof(T));
};

template <class KeyT, class ValueT, class BeginIt, class EndIt>
inline cudaError_t b200rs_segmented(
  void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out, const ValueT* d_values_in,
  ValueT* d_values_out, long long num_items, long long num_segments, BeginIt d_begin_offsets, EndIt d_end_offsets,
  int begin_bit, int end_bit, bool descending, cudaStream_t stream)
{
  static_assert(b200rs_offset_bytes<BeginIt>::value == b200rs_offset_bytes<EndIt>::value,
                "begin and end offsets must have the same width");
  return static_cast<cudaError_t>(b200rs_segmented_sort(
    d_temp_storage, &temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out,
    static_cast<unsigned long long>(num_items), static_cast<unsigned long long>(num_segments), d_begin_offsets,
    d_end_offsets, b200rs_offset_bytes<BeginIt>::value, b200rs_key_kind_of<KeyT>(), int(sizeof(KeyT)),
    b200rs_value_bytes_of<ValueT>(), begin_bit, end_bit, descending ? 1 : 0, reinterpret_cast<b200rs_stream_t>(stream)));
}

template <class KeyT, class ValueT, class BeginIt, class EndIt>
inline cudaError_t b200rs_segmented_db(
  void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, DoubleBuffer<ValueT>* d_values,
  long long num_items, long long num_segments, BeginIt d_begin_offsets, EndIt d_end_offsets, int begin_bit, int end_bit,
  bool descending, cudaStream_t stream)
{
  const cudaError_t e = b200rs_segmented<KeyT, ValueT>(
    d_temp_storage, temp_storage_bytes, d_keys.Current(), d_keys.Alternate(),
    d_values != nullptr ? d_values->Current() : nullptr, d_values != nullptr ? d_values->Alternate() : nullptr, num_items,
    num_segments, d_begin_offsets, d_end_offsets, begin_bit, end_bit, descending, stream);
  if (e == cudaSuccess && d_temp_storage != nullptr)
  {
    // items outside every segment must read the same through Current() after the call: they were never moved, so
    // the caller keeps seeing its own data for them only in the OLD buffer.  The reference has the same property
    // (its passes never touch them either); callers index the result by segment.
    d_keys.selector ^= 1;
    if (d_values != nullptr)
    {
      d_values->selector ^= 1;
    }
  }
  return e;
}
} // namespace detail

struct DeviceSegmentedRadixSort
{
#define B200RS_SEG_POINTER(NAME, DESC)                                                                               \
  template <typename KeyT, typename ValueT, typename BeginOffsetIteratorT, typename EndOffsetIteratorT>             \
  static cudaError_t NAME(                                                                                          \
    void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,                      \
    const ValueT* d_values_in, ValueT* d_values_out, long long num_items, long long num_segments,                   \
    BeginOffsetIteratorT d_begin_offsets, EndOffsetIteratorT d_end_offsets, int begin_bit = 0,                      \
    int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = nullptr)                                                  \
  {                                                                                                                 \
    return detail::b200rs_segmented<KeyT, ValueT>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out,        \
                                                  d_values_in, d_values_out, num_items, num_segments,               \
                                                  d_begin_offsets, d_end_offsets, begin_bit, end_bit, DESC, stream); \
  }                                                                                                                 \
  template <typename KeyT, typename ValueT, typename BeginOffsetIteratorT, typename EndOffsetIteratorT>             \
  static cudaError_t NAME(                                                                                          \
    void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, DoubleBuffer<ValueT>& d_values,   \
    long long num_items, long long num_segments, BeginOffsetIteratorT d_begin_offsets,                              \
    EndOffsetIteratorT d_end_offsets, int begin_bit = 0, int end_bit = sizeof(KeyT) * 8,                            \
    cudaStream_t stream = nullptr)                                                                                  \
  {                                                                                                                 \
    return detail::b200rs_segmented_db<KeyT, ValueT>(d_temp_storage, temp_storage_bytes, d_keys, &d_values,         \
                                                     num_items, num_segments, d_begin_offsets, d_end_offsets,       \
                                                     begin_bit, end_bit, DESC, stream);                             \
  }
#define B200RS_SEG_KEYS(NAME, DESC)                                                                                  \
  template <typename KeyT, typename BeginOffsetIteratorT, typename EndOffsetIteratorT>                              \
  static cudaError_t NAME(                                                                                          \
    void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out, long long num_items, \
    long long num_segments, BeginOffsetIteratorT d_begin_offsets, EndOffsetIteratorT d_end_offsets,                 \
    int begin_bit = 0, int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = nullptr)                               \
  {                                                                                                                 \
    return detail::b200rs_segmented<KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out,      \
                                                    nullptr, nullptr, num_items, num_segments, d_begin_offsets,     \
                                                    d_end_offsets, begin_bit, end_bit, DESC, stream);               \
  }                                                                                                                 \
  template <typename KeyT, typename BeginOffsetIteratorT, typename EndOffsetIteratorT>                              \
  static cudaError_t NAME(                                                                                          \
    void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, long long num_items,              \
    long long num_segments, BeginOffsetIteratorT d_begin_offsets, EndOffsetIteratorT d_end_offsets,                 \
    int begin_bit = 0, int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = nullptr)                               \
  {                                                                                                                 \
    return detail::b200rs_segmented_db<KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys, nullptr,         \
                                                       num_items, num_segments, d_begin_offsets, d_end_offsets,     \
                                                       begin_bit, end_bit, DESC, stream);                           \
  }
  B200RS_SEG_POINTER(SortPairs, false)
  B200RS_SEG_POINTER(SortPairsDescending, true)
  B200RS_SEG_KEYS(SortKeys, false)
  B200RS_SEG_KEYS(SortKeysDescending, true)
#undef B200RS_SEG_POINTER
#undef B200RS_SEG_KEYS
};

} // namespace cub
