// thrust::sort / stable_sort / sort_by_key / stable_sort_by_key for arithmetic keys in device memory, on top of
// libb200rs.so -- the cases the reference's CUDA backend routes to radix sort
// (/root/reference/thrust/thrust/system/cuda/detail/sort.h:288-339, __smart_sort::can_use_primitive_sort: arithmetic
// key + thrust::less / thrust::greater [or cuda::std::less/greater]).  Public entry points being mirrored:
// /root/reference/thrust/thrust/sort.h (sort :212, stable_sort :404, sort_by_key :589, stable_sort_by_key :804 and the
// execution-policy overloads next to them).
//
// Every call is ONE b200rs_sort_inplace (include/b200rs.h): scratch from the stream-ordered pool, DoubleBuffer sort,
// copy-back iff needed, synchronise (the reference synchronises unless the policy is par_nosync,
// sort.h:260 `synchronize_optional`).  A non-zero return throws thrust::system_error like the reference
// (sort.h:260-262).  Comparators other than less / greater take cub::DeviceMergeSort (include/cub/device/
// device_merge_sort.cuh), as the reference's __smart_sort does.  No host backend exists here.
#pragma once

#include <cuda_runtime.h>

#include <functional>
#include <type_traits>

#if defined(__has_include)
#  if __has_include(<cuda/std/functional>)
#    include <cuda/std/functional>
#    define B200RS_THRUST_SHIM_HAS_CUDA_STD 1
#  endif
#endif

#include "../b200rs.h"
#include "device_vector.h"
#ifdef __CUDACC__
#  include "../cub/device/device_merge_sort.cuh"
#endif

namespace thrust
{

template <class T = void>
struct less
{};
template <class T = void>
struct greater
{};

// Execution policies: thrust::device, thrust::cuda::par.on(stream), thrust::cuda::par_nosync.on(stream)
namespace cuda
{
struct execute_on_stream
{
  cudaStream_t stream = nullptr;
  bool sync           = true;
};
struct par_t : execute_on_stream
{
  execute_on_stream on(cudaStream_t s) const
  {
    return execute_on_stream{s, sync};
  }
};
static const par_t par{};
static const par_t par_nosync{{nullptr, false}};
} // namespace cuda
static const cuda::par_t device{};

namespace detail
{
template <class K>
constexpr int key_kind_of()
{
  static_assert(std::is_arithmetic<K>::value, "thrust::sort shim: arithmetic keys only (radix-sort path)");
  return std::is_floating_point<K>::value ? B200RS_KEY_FLOAT
       : (std::is_signed<K>::value && !std::is_same<K, bool>::value) ? B200RS_KEY_INT
                                                                     : B200RS_KEY_UINT;
}
template <class T>
T* unwrap(device_ptr<T> it)
{
  return it.get();
}
template <class T>
T* unwrap(T* it)
{
  return it;
}
// Comparators the radix-sort path accepts (reference: __smart_sort::can_use_primitive_sort, sort.h:288-301): the
// less / greater function objects of thrust, std and cuda::std.  0 = not a radix comparator, 1 = ascending,
// 2 = descending.  Anything else would need the reference's merge sort, which this path does not have: it is a compile
// error here, never a silent ascending sort.
template <class C>
struct radix_order : std::integral_constant<int, 0>
{};
template <class T>
struct radix_order<less<T>> : std::integral_constant<int, 1>
{};
template <class T>
struct radix_order<greater<T>> : std::integral_constant<int, 2>
{};
template <class T>
struct radix_order<std::less<T>> : std::integral_constant<int, 1>
{};
template <class T>
struct radix_order<std::greater<T>> : std::integral_constant<int, 2>
{};
#ifdef B200RS_THRUST_SHIM_HAS_CUDA_STD
template <class T>
struct radix_order<::cuda::std::less<T>> : std::integral_constant<int, 1>
{};
template <class T>
struct radix_order<::cuda::std::greater<T>> : std::integral_constant<int, 2>
{};
#endif
template <class C>
struct is_descending : std::integral_constant<bool, radix_order<C>::value == 2>
{
  static_assert(radix_order<C>::value != 0,
                "thrust::sort shim: only less<T> / greater<T> (thrust, std or cuda::std) select the radix-sort path; "
                "other comparators need the reference's merge sort");
};

template <class KeyIt, class Compare>
void radix_sort_keys(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last, Compare)
{
  auto* k = unwrap(first);
  using K = std::remove_pointer_t<decltype(k)>;
  const int rc = b200rs_sort_inplace(k, nullptr, static_cast<uint64_t>(unwrap(last) - k), key_kind_of<K>(),
                                     int(sizeof(K)), 0, is_descending<Compare>::value ? 1 : 0, pol.sync ? 1 : 0,
                                     reinterpret_cast<b200rs_stream_t>(pol.stream));
  throw_on_error(static_cast<cudaError_t>(rc), "radix_sort: failed on 2nd step"); // sort.h:261
}
template <class KeyIt, class ValIt, class Compare>
void radix_sort_pairs(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last, ValIt values, Compare)
{
  auto* k = unwrap(first);
  auto* v = unwrap(values);
  using K = std::remove_pointer_t<decltype(k)>;
  using V = std::remove_pointer_t<decltype(v)>;
  static_assert(std::is_trivially_copyable<V>::value, "values are moved as opaque blobs");
  static_assert(sizeof(V) == 1 || sizeof(V) == 2 || sizeof(V) == 4 || sizeof(V) == 8 || sizeof(V) == 16,
                "value width must be 1/2/4/8/16 bytes");
  const int rc = b200rs_sort_inplace(k, v, static_cast<uint64_t>(unwrap(last) - k), key_kind_of<K>(), int(sizeof(K)),
                                     int(sizeof(V)), is_descending<Compare>::value ? 1 : 0, pol.sync ? 1 : 0,
                                     reinterpret_cast<b200rs_stream_t>(pol.stream));
  throw_on_error(static_cast<cudaError_t>(rc), "radix_sort: failed on 2nd step");
}

// Any other comparator (a user functor, a lambda, less / greater on a key type without a bit-ordered image) takes the
// comparison sort, as in the reference (sort.h:288-301 -> __merge_sort): cub::DeviceMergeSort of this repo, whose
// kernels are templates instantiated here by the caller's nvcc (a comparator cannot cross the C ABI).
#ifdef __CUDACC__
template <class K, class V, class Compare>
void merge_sort_impl(const cuda::execute_on_stream& pol, K* k, V* v, unsigned long long n, Compare comp)
{
  size_t bytes = 0;
  auto call    = [&](void* t) {
    if constexpr (std::is_same<V, ::cub::NullType>::value)
    {
      return ::cub::DeviceMergeSort::StableSortKeys(t, bytes, k, n, comp, pol.stream);
    }
    else
    {
      return ::cub::DeviceMergeSort::StableSortPairs(t, bytes, k, v, n, comp, pol.stream);
    }
  };
  throw_on_error(call(nullptr), "merge_sort: failed on 1st step");
  void* temp = nullptr;
  throw_on_error(cudaMallocAsync(&temp, bytes, pol.stream), "merge_sort: failed to get memory buffer");
  const cudaError_t e = call(temp);
  const cudaError_t f = cudaFreeAsync(temp, pol.stream);
  throw_on_error(e, "merge_sort: failed on 2nd step");
  throw_on_error(f, "merge_sort: failed to free memory buffer");
  if (pol.sync)
  {
    throw_on_error(cudaStreamSynchronize(pol.stream), "merge_sort: failed to synchronize");
  }
}
#endif

template <class KeyIt, class Compare>
void sort_keys(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last, Compare comp)
{
  using K = std::remove_pointer_t<decltype(unwrap(first))>;
  if constexpr (radix_order<Compare>::value != 0 && std::is_arithmetic<K>::value)
  {
    radix_sort_keys(pol, first, last, comp);
  }
  else
  {
#ifdef __CUDACC__
    static_assert(radix_order<Compare>::value == 0 || std::is_void<K>::value,
                  "less<T> / greater<T> of a non-arithmetic key type: pass a comparator with operator()");
    merge_sort_impl(pol, unwrap(first), static_cast<::cub::NullType*>(nullptr),
                    static_cast<unsigned long long>(unwrap(last) - unwrap(first)), comp);
#else
    static_assert(radix_order<Compare>::value != 0, "user comparators need nvcc (the merge-sort kernels are templates)");
#endif
  }
}
template <class KeyIt, class ValIt, class Compare>
void sort_pairs(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last, ValIt values, Compare comp)
{
  using K = std::remove_pointer_t<decltype(unwrap(first))>;
  if constexpr (radix_order<Compare>::value != 0 && std::is_arithmetic<K>::value)
  {
    radix_sort_pairs(pol, first, last, values, comp);
  }
  else
  {
#ifdef __CUDACC__
    static_assert(radix_order<Compare>::value == 0 || std::is_void<K>::value,
                  "less<T> / greater<T> of a non-arithmetic key type: pass a comparator with operator()");
    merge_sort_impl(pol, unwrap(first), unwrap(values), static_cast<unsigned long long>(unwrap(last) - unwrap(first)),
                    comp);
#else
    static_assert(radix_order<Compare>::value != 0, "user comparators need nvcc (the merge-sort kernels are templates)");
#endif
  }
}
} // namespace detail

// ---- sort / stable_sort (radix sort is stable, so both names are the same call, as in the reference)
template <class KeyIt>
void sort(KeyIt first, KeyIt last)
{
  detail::radix_sort_keys(device, first, last, less<>{});
}
template <class KeyIt, class Compare>
void sort(KeyIt first, KeyIt last, Compare comp)
{
  detail::sort_keys(device, first, last, comp);
}
template <class KeyIt>
void sort(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last)
{
  detail::radix_sort_keys(pol, first, last, less<>{});
}
template <class KeyIt, class Compare>
void sort(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last, Compare comp)
{
  detail::sort_keys(pol, first, last, comp);
}
template <class KeyIt>
void stable_sort(KeyIt first, KeyIt last)
{
  detail::radix_sort_keys(device, first, last, less<>{});
}
template <class KeyIt, class Compare>
void stable_sort(KeyIt first, KeyIt last, Compare comp)
{
  detail::sort_keys(device, first, last, comp);
}
template <class KeyIt>
void stable_sort(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last)
{
  detail::radix_sort_keys(pol, first, last, less<>{});
}
template <class KeyIt, class Compare>
void stable_sort(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last, Compare comp)
{
  detail::sort_keys(pol, first, last, comp);
}

// ---- sort_by_key / stable_sort_by_key
template <class KeyIt, class ValIt>
void sort_by_key(KeyIt first, KeyIt last, ValIt values)
{
  detail::radix_sort_pairs(device, first, last, values, less<>{});
}
template <class KeyIt, class ValIt, class Compare>
void sort_by_key(KeyIt first, KeyIt last, ValIt values, Compare comp)
{
  detail::sort_pairs(device, first, last, values, comp);
}
template <class KeyIt, class ValIt>
void sort_by_key(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last, ValIt values)
{
  detail::radix_sort_pairs(pol, first, last, values, less<>{});
}
template <class KeyIt, class ValIt, class Compare>
void sort_by_key(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last, ValIt values, Compare comp)
{
  detail::sort_pairs(pol, first, last, values, comp);
}
template <class KeyIt, class ValIt>
void stable_sort_by_key(KeyIt first, KeyIt last, ValIt values)
{
  detail::radix_sort_pairs(device, first, last, values, less<>{});
}
template <class KeyIt, class ValIt, class Compare>
void stable_sort_by_key(KeyIt first, KeyIt last, ValIt values, Compare comp)
{
  detail::sort_pairs(device, first, last, values, comp);
}
template <class KeyIt, class ValIt>
void stable_sort_by_key(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last, ValIt values)
{
  detail::radix_sort_pairs(pol, first, last, values, less<>{});
}
template <class KeyIt, class ValIt, class Compare>
void stable_sort_by_key(const cuda::execute_on_stream& pol, KeyIt first, KeyIt last, ValIt values, Compare comp)
{
  detail::sort_pairs(pol, first, last, values, comp);
}

} // namespace thrust
