// cub::DeviceRadixSort on top of libb200rs.so -- a header-only drop-in for the arithmetic-key surface of
// /root/reference/cub/cub/device/device_radix_sort.cuh (cub 3.6.0).
//
// Put this repo's include/ ahead of the CCCL include path and link -lb200rs.  Every overload below keeps the
// reference's name, template parameters that matter for deduction, parameter order and defaults, and returns
// cudaError_t; each one type-erases to ONE call of the C ABI (include/b200rs.h, b200rs_sort).  Nothing here
// contains a kernel and nothing falls back to another implementation: an unsupported key type is a compile error,
// an unsupported size a cudaErrorNotSupported from the library.
//
//   reference overload (device_radix_sort.cuh)            here
//   SortPairs            pointer :412   DoubleBuffer :1127   env :532 / :1230
//   SortPairsDescending  pointer :1776  DoubleBuffer :2295   env :1894 / :2398
//   SortKeys             pointer :3034  DoubleBuffer :3644   env :3138 / :3743
//   SortKeysDescending   pointer :4211  DoubleBuffer :4669   env :4310 / :4768
// Decomposer overloads for user-defined key types (:671, :905, :1359, :1565 and the Descending / SortKeys twins, with and
// without a bit window, temp-storage and env forms) and 128-bit integer keys go through b200rs_sort_fields: the members the decomposer returns are
// sorted as a chain of stable passes, least significant member first (see cccl_b200/csrc/fields.cu).
//
// Semantics kept from the reference:
//   * d_temp_storage == nullptr  => only temp_storage_bytes is written (device_radix_sort.cuh:300-317);
//   * pointer overloads never write the input and need ~N extra temp; DoubleBuffer overloads may clobber both
//     buffers, need O(N / tile) temp, and flip d_keys.selector / d_values.selector to the buffer holding the
//     result (dispatch_radix_sort.cuh:1943-1944);
//   * keys are ordered by bits [begin_bit, end_bit) of Traits<KeyT>::TwiddleIn(key) (util_type.cuh:857-963),
//     -0.0 == +0.0, stable in both directions;
//   * work is enqueued on `stream` without synchronisation; legal under stream capture.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cuda/std/tuple>

#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <utility>

// cuda::get_stream (CCCL 3.x): lets the env overloads take any environment that carries a stream
#if defined(__has_include)
#  if __has_include(<cuda/__stream/get_stream.h>)
#    include <cuda/__stream/get_stream.h>
#    define B200RS_HAS_GET_STREAM 1
#  endif
#endif

#include "../../b200rs.h"

namespace cub
{

/// Same layout and members as cub::DoubleBuffer (/root/reference/cub/cub/util_type.cuh:749-779).
template <typename T>
struct DoubleBuffer
{
  T* d_buffers[2]{};
  int selector = 0;

  DoubleBuffer() = default;
  __host__ __device__ DoubleBuffer(T* d_current, T* d_alternate)
      : d_buffers{d_current, d_alternate}
  {}
  __host__ __device__ T* Current()
  {
    return d_buffers[selector];
  }
  __host__ __device__ T* Alternate()
  {
    return d_buffers[selector ^ 1];
  }
};

/// Placeholder value type of keys-only sorts (/root/reference/cub/cub/util_type.cuh NullType).
struct NullType
{};

enum class SortOrder
{
  Ascending,
  Descending
};

namespace detail
{
// Traits<T>::CATEGORY of the reference (util_type.cuh:857-963) reduced to what the transform needs.
// 16-bit floating-point keys (reference: __can_use_radix_sort, device_radix_sort.cuh:51-57; NumericTraits<__half>,
// NumericTraits<__nv_bfloat16>, util_type.cuh:1017-1095): same sign-magnitude transform on a 16-bit pattern.
template <class K>
struct b200rs_is_float16 : std::integral_constant<bool, std::is_same<K, __half>::value || std::is_same<K, __nv_bfloat16>::value>
{};

template <class KeyT>
constexpr int b200rs_key_kind_of()
{
  using K = std::remove_cv_t<KeyT>;
  static_assert(std::is_arithmetic<K>::value || b200rs_is_float16<K>::value,
                "arithmetic keys (or pass a decomposer for a user-defined key type)");
  static_assert(sizeof(K) == 1 || sizeof(K) == 2 || sizeof(K) == 4 || sizeof(K) == 8, "key width must be 1/2/4/8");
  return (std::is_floating_point<K>::value || b200rs_is_float16<K>::value) ? B200RS_KEY_FLOAT
       : (std::is_signed<K>::value && !std::is_same<K, bool>::value) ? B200RS_KEY_INT
                                                                     : B200RS_KEY_UINT;
}

template <class ValueT>
constexpr int b200rs_value_bytes_of();

// 128-bit integer keys (reference: util_type.cuh __int128 traits): two 64-bit members, the high half first
template <class KeyT>
struct b200rs_is_int128
    : std::integral_constant<bool, std::is_same<std::remove_cv_t<KeyT>, __int128>::value
                                     || std::is_same<std::remove_cv_t<KeyT>, unsigned __int128>::value>
{};

struct b200rs_field_list
{
  b200rs_key_field f[16];
  int n = 0;
  void add(size_t offset, size_t bytes, int kind)
  {
    if (n < 16)
    {
      f[n].offset = uint32_t(offset);
      f[n].bytes  = uint32_t(bytes);
      f[n].kind   = kind;
    }
    ++n;
  }
};

// members of KeyT the decomposer exposes: tuple of references, leftmost = most significant (device_radix_sort.cuh:620-626)
template <class KeyT, class DecomposerT>
inline b200rs_field_list b200rs_fields_of(DecomposerT decomposer)
{
  alignas(KeyT) unsigned char raw[sizeof(KeyT)] = {};
  KeyT& probe                                   = *reinterpret_cast<KeyT*>(raw); // never read, only addressed
  b200rs_field_list out;
  auto members = decomposer(probe);
  ::cuda::std::apply(
    [&](auto&... m) {
      (out.add(size_t(reinterpret_cast<const unsigned char*>(&m) - raw), sizeof(m),
               b200rs_key_kind_of<std::remove_reference_t<decltype(m)>>()),
       ...);
    },
    members);
  return out;
}
template <class KeyT>
inline b200rs_field_list b200rs_fields_of_int128()
{
  b200rs_field_list out;
  out.add(8, 8, std::is_same<std::remove_cv_t<KeyT>, __int128>::value ? B200RS_KEY_INT : B200RS_KEY_UINT);
  out.add(0, 8, B200RS_KEY_UINT);
  return out;
}

template <class KeyT, class ValueT>
inline cudaError_t b200rs_fields_sort(
  void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out, const ValueT* d_values_in,
  ValueT* d_values_out, unsigned long long num_items, const b200rs_field_list& fields, int begin_bit, int end_bit,
  bool descending, cudaStream_t stream)
{
  if (fields.n > 16)
  {
    return cudaErrorNotSupported;
  }
  int total_bits = 0;
  for (int i = 0; i < fields.n; ++i)
  {
    total_bits += int(fields.f[i].bytes) * 8;
  }
  return static_cast<cudaError_t>(b200rs_sort_fields(
    d_temp_storage, &temp_storage_bytes, d_keys_in, d_keys_out, int(sizeof(KeyT)), fields.f, fields.n, d_values_in,
    d_values_out, b200rs_value_bytes_of<ValueT>(), num_items, begin_bit, end_bit < 0 ? total_bits : end_bit,
    descending ? 1 : 0, reinterpret_cast<b200rs_stream_t>(stream)));
}

template <class KeyT, class ValueT>
inline cudaError_t b200rs_fields_sort_db(
  void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, DoubleBuffer<ValueT>* d_values,
  unsigned long long num_items, const b200rs_field_list& fields, int begin_bit, int end_bit, bool descending,
  cudaStream_t stream)
{
  const cudaError_t e = b200rs_fields_sort<KeyT, ValueT>(
    d_temp_storage, temp_storage_bytes, d_keys.Current(), d_keys.Alternate(),
    d_values != nullptr ? d_values->Current() : nullptr, d_values != nullptr ? d_values->Alternate() : nullptr, num_items,
    fields, begin_bit, end_bit, descending, stream);
  if (e == cudaSuccess && d_temp_storage != nullptr)
  {
    d_keys.selector ^= 1; // the result is gathered into the alternate buffers
    if (d_values != nullptr)
    {
      d_values->selector ^= 1;
    }
  }
  return e;
}
template <class ValueT>
constexpr int b200rs_value_bytes_of()
{
  // the library moves values as opaque blobs of these sizes; 16-byte values must be 8-byte aligned (they are loaded
  // and stored as two 64-bit words).  Anything else would be a cudaErrorNotSupported / misaligned access at run time.
  static_assert(std::is_same<std::remove_cv_t<ValueT>, NullType>::value || sizeof(ValueT) == 1 || sizeof(ValueT) == 2
                  || sizeof(ValueT) == 4 || sizeof(ValueT) == 8 || sizeof(ValueT) == 16,
                "value types of 1, 2, 4, 8 or 16 bytes (wrap other sizes in an index sort + gather)");
  static_assert(std::is_same<std::remove_cv_t<ValueT>, NullType>::value || sizeof(ValueT) != 16 || alignof(ValueT) >= 8,
                "16-byte value types must be at least 8-byte aligned");
  return std::is_same<std::remove_cv_t<ValueT>, NullType>::value ? 0 : int(sizeof(ValueT));
}

// tuning of the env overload being executed on this thread (nullptr outside one): see stream_env below
inline const b200rs_tuning*& b200rs_current_tuning()
{
  static thread_local const b200rs_tuning* cur = nullptr;
  return cur;
}

template <class KeyT, class ValueT>
inline cudaError_t b200rs_pointer_sort(
  void* d_temp_storage,
  size_t& temp_storage_bytes,
  const KeyT* d_keys_in,
  KeyT* d_keys_out,
  const ValueT* d_values_in,
  ValueT* d_values_out,
  unsigned long long num_items,
  int begin_bit,
  int end_bit,
  bool descending,
  cudaStream_t stream)
{
  if constexpr (b200rs_is_int128<KeyT>::value)
  {
    return b200rs_fields_sort<KeyT, ValueT>(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in,
                                            d_values_out, num_items, b200rs_fields_of_int128<KeyT>(), begin_bit, end_bit,
                                            descending, stream);
  }
  else
  return static_cast<cudaError_t>(b200rs_sort_tuned(
    d_temp_storage,
    &temp_storage_bytes,
    d_keys_in,
    d_keys_out,
    d_values_in,
    d_values_out,
    num_items,
    b200rs_key_kind_of<KeyT>(),
    int(sizeof(KeyT)),
    b200rs_value_bytes_of<ValueT>(),
    begin_bit,
    end_bit,
    descending ? 1 : 0,
    /*is_overwrite_okay=*/0,
    nullptr,
    reinterpret_cast<b200rs_stream_t>(stream),
    b200rs_current_tuning()));
}

template <class KeyT, class ValueT>
inline cudaError_t b200rs_double_buffer_sort(
  void* d_temp_storage,
  size_t& temp_storage_bytes,
  DoubleBuffer<KeyT>& d_keys,
  DoubleBuffer<ValueT>* d_values,
  unsigned long long num_items,
  int begin_bit,
  int end_bit,
  bool descending,
  cudaStream_t stream)
{
  if constexpr (b200rs_is_int128<KeyT>::value)
  {
    return b200rs_fields_sort_db<KeyT, ValueT>(d_temp_storage, temp_storage_bytes, d_keys, d_values, num_items,
                                               b200rs_fields_of_int128<KeyT>(), begin_bit, end_bit, descending, stream);
  }
  else
  {
  int selector = 0;
  const int rc = b200rs_sort_tuned(
    d_temp_storage,
    &temp_storage_bytes,
    d_keys.Current(),
    d_keys.Alternate(),
    d_values ? static_cast<const void*>(d_values->Current()) : nullptr,
    d_values ? static_cast<void*>(d_values->Alternate()) : nullptr,
    num_items,
    b200rs_key_kind_of<KeyT>(),
    int(sizeof(KeyT)),
    d_values ? b200rs_value_bytes_of<ValueT>() : 0,
    begin_bit,
    end_bit,
    descending ? 1 : 0,
    /*is_overwrite_okay=*/1,
    &selector,
    reinterpret_cast<b200rs_stream_t>(stream),
    b200rs_current_tuning());
  if (rc == 0 && d_temp_storage != nullptr)
  {
    d_keys.selector ^= selector; // dispatch_radix_sort.cuh:1943-1944
    if (d_values)
    {
      d_values->selector ^= selector;
    }
  }
  return static_cast<cudaError_t>(rc);
  }
}
} // namespace detail

/// Execution environment of the env overloads: the reference takes a cuda::std::execution::env carrying a stream
/// and a memory resource (device_radix_sort.cuh:532, detail::dispatch_with_env); this stand-alone shim carries the
/// stream and obtains the temporary storage from the device's stream-ordered pool (cudaMallocAsync / cudaFreeAsync).
/// The reference's env may also carry `cuda::execution::tune(policy)` (device_radix_sort.cuh:200-202): here the tuning is
/// the C ABI's b200rs_tuning (tile configuration index, one-launch threshold, single-CTA switch), applied to THIS call only.
struct stream_env
{
  cudaStream_t stream = nullptr;
  b200rs_tuning tuning{-1, -1, -1};
  bool tuned = false;
  stream_env() = default;
  stream_env(cudaStream_t s)
      : stream(s)
  {}
  stream_env(cudaStream_t s, const b200rs_tuning& t)
      : stream(s)
      , tuning(t)
      , tuned(true)
  {}
};

/// Property an execution environment may answer to carry a per-call tuning (the counterpart of the reference's
/// `cuda::execution::tune`, device_radix_sort.cuh:200-202): `env.query(cub::b200rs_get_tuning_t{})` -> b200rs_tuning.
struct b200rs_get_tuning_t
{};

namespace detail
{
// What the env overloads accept as EnvT (reference: any environment queryable with cuda::get_stream,
// detail/env_dispatch.cuh:39-77): cub::stream_env, a raw cudaStream_t, cuda::stream_ref / anything whose .get() yields
// a cudaStream_t, and any environment with .query(cuda::get_stream_t) (cuda::std::execution::env of CCCL 3.x carrying a
// stream) when those headers are on the include path.  Other properties of such an environment (memory resource,
// requirements) are not consulted: temporary storage always comes from the stream-ordered pool of the device.
template <class E, class = void>
struct b200rs_has_get : std::false_type
{};
template <class E>
struct b200rs_has_get<E, std::enable_if_t<std::is_convertible<decltype(std::declval<const E&>().get()), cudaStream_t>::value>>
    : std::true_type
{};
template <class E, class = void>
struct b200rs_has_stream_query : std::false_type
{};
#ifdef B200RS_HAS_GET_STREAM
template <class E>
struct b200rs_has_stream_query<E, std::void_t<decltype(::cuda::get_stream(std::declval<const E&>()))>> : std::true_type
{};
#endif
template <class E, class = void>
struct b200rs_has_tuning_query : std::false_type
{};
template <class E>
struct b200rs_has_tuning_query<
  E,
  std::enable_if_t<std::is_convertible<decltype(std::declval<const E&>().query(b200rs_get_tuning_t{})), b200rs_tuning>::value>>
    : std::true_type
{};
template <class E>
struct b200rs_is_env
    : std::integral_constant<bool,
                             std::is_same<E, stream_env>::value || std::is_same<E, cudaStream_t>::value
                               || b200rs_has_get<E>::value || b200rs_has_stream_query<E>::value>
{};

template <class E>
inline stream_env b200rs_env_view(const E& env)
{
  stream_env out;
  if constexpr (std::is_same<E, stream_env>::value)
  {
    out = env;
  }
  else if constexpr (std::is_same<E, cudaStream_t>::value)
  {
    out.stream = env;
  }
  else if constexpr (b200rs_has_get<E>::value)
  {
    out.stream = env.get();
  }
#ifdef B200RS_HAS_GET_STREAM
  else
  {
    out.stream = ::cuda::get_stream(env).get();
  }
#endif
  if constexpr (!std::is_same<E, stream_env>::value && b200rs_has_tuning_query<E>::value)
  {
    out.tuning = env.query(b200rs_get_tuning_t{});
    out.tuned  = true;
  }
  return out;
}

template <class F>
inline cudaError_t b200rs_with_env(const stream_env& env, F&& call)
{
  struct scoped_tuning // the env's tuning applies to the calls made below, on this thread, and to nothing else
  {
    const b200rs_tuning* saved;
    explicit scoped_tuning(const b200rs_tuning* t)
        : saved(b200rs_current_tuning())
    {
      b200rs_current_tuning() = t;
    }
    ~scoped_tuning()
    {
      b200rs_current_tuning() = saved;
    }
  } guard(env.tuned ? &env.tuning : nullptr);
  size_t bytes    = 0;
  cudaError_t err = call(nullptr, bytes, env.stream);
  if (err != cudaSuccess)
  {
    return err;
  }
  void* d_temp = nullptr;
  err          = cudaMallocAsync(&d_temp, bytes, env.stream);
  if (err != cudaSuccess)
  {
    return err;
  }
  err                   = call(d_temp, bytes, env.stream);
  const cudaError_t fre = cudaFreeAsync(d_temp, env.stream);
  return err != cudaSuccess ? err : fre;
}
} // namespace detail

struct DeviceRadixSort
{
  // ------------------------------------------------------------------ SortPairs
  template <typename KeyT, typename ValueT, typename NumItemsT>
  static cudaError_t SortPairs(
    void* d_temp_storage,
    size_t& temp_storage_bytes,
    const KeyT* d_keys_in,
    KeyT* d_keys_out,
    const ValueT* d_values_in,
    ValueT* d_values_out,
    NumItemsT num_items,
    int begin_bit       = 0,
    int end_bit         = sizeof(KeyT) * 8,
    cudaStream_t stream = nullptr)
  {
    return detail::b200rs_pointer_sort(
      d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out,
      static_cast<unsigned long long>(num_items), begin_bit, end_bit, false, stream);
  }

  template <typename KeyT, typename ValueT, typename NumItemsT>
  static cudaError_t SortPairs(
    void* d_temp_storage,
    size_t& temp_storage_bytes,
    DoubleBuffer<KeyT>& d_keys,
    DoubleBuffer<ValueT>& d_values,
    NumItemsT num_items,
    int begin_bit       = 0,
    int end_bit         = sizeof(KeyT) * 8,
    cudaStream_t stream = nullptr)
  {
    return detail::b200rs_double_buffer_sort(
      d_temp_storage, temp_storage_bytes, d_keys, &d_values, static_cast<unsigned long long>(num_items), begin_bit,
      end_bit, false, stream);
  }

  template <typename KeyT, typename ValueT, typename NumItemsT, typename EnvT = stream_env,
            std::enable_if_t<std::is_integral<NumItemsT>::value && detail::b200rs_is_env<EnvT>::value, int> = 0>
  [[nodiscard]] static cudaError_t SortPairs(
    const KeyT* d_keys_in,
    KeyT* d_keys_out,
    const ValueT* d_values_in,
    ValueT* d_values_out,
    NumItemsT num_items,
    int begin_bit         = 0,
    int end_bit           = sizeof(KeyT) * 8,
    const EnvT& env = {})
  {
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {
      return SortPairs(t, b, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, begin_bit, end_bit, s);
    });
  }

  template <typename KeyT, typename ValueT, typename NumItemsT, typename EnvT = stream_env,
            std::enable_if_t<std::is_integral<NumItemsT>::value && detail::b200rs_is_env<EnvT>::value, int> = 0>
  [[nodiscard]] static cudaError_t SortPairs(
    DoubleBuffer<KeyT>& d_keys,
    DoubleBuffer<ValueT>& d_values,
    NumItemsT num_items,
    int begin_bit         = 0,
    int end_bit           = sizeof(KeyT) * 8,
    const EnvT& env = {})
  {
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {
      return SortPairs(t, b, d_keys, d_values, num_items, begin_bit, end_bit, s);
    });
  }

  // ------------------------------------------------------------------ SortPairsDescending
  template <typename KeyT, typename ValueT, typename NumItemsT>
  static cudaError_t SortPairsDescending(
    void* d_temp_storage,
    size_t& temp_storage_bytes,
    const KeyT* d_keys_in,
    KeyT* d_keys_out,
    const ValueT* d_values_in,
    ValueT* d_values_out,
    NumItemsT num_items,
    int begin_bit       = 0,
    int end_bit         = sizeof(KeyT) * 8,
    cudaStream_t stream = nullptr)
  {
    return detail::b200rs_pointer_sort(
      d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out,
      static_cast<unsigned long long>(num_items), begin_bit, end_bit, true, stream);
  }

  template <typename KeyT, typename ValueT, typename NumItemsT>
  static cudaError_t SortPairsDescending(
    void* d_temp_storage,
    size_t& temp_storage_bytes,
    DoubleBuffer<KeyT>& d_keys,
    DoubleBuffer<ValueT>& d_values,
    NumItemsT num_items,
    int begin_bit       = 0,
    int end_bit         = sizeof(KeyT) * 8,
    cudaStream_t stream = nullptr)
  {
    return detail::b200rs_double_buffer_sort(
      d_temp_storage, temp_storage_bytes, d_keys, &d_values, static_cast<unsigned long long>(num_items), begin_bit,
      end_bit, true, stream);
  }

  template <typename KeyT, typename ValueT, typename NumItemsT, typename EnvT = stream_env,
            std::enable_if_t<std::is_integral<NumItemsT>::value && detail::b200rs_is_env<EnvT>::value, int> = 0>
  [[nodiscard]] static cudaError_t SortPairsDescending(
    const KeyT* d_keys_in,
    KeyT* d_keys_out,
    const ValueT* d_values_in,
    ValueT* d_values_out,
    NumItemsT num_items,
    int begin_bit         = 0,
    int end_bit           = sizeof(KeyT) * 8,
    const EnvT& env = {})
  {
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {
      return SortPairsDescending(t, b, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, begin_bit,
                                 end_bit, s);
    });
  }

  template <typename KeyT, typename ValueT, typename NumItemsT, typename EnvT = stream_env,
            std::enable_if_t<std::is_integral<NumItemsT>::value && detail::b200rs_is_env<EnvT>::value, int> = 0>
  [[nodiscard]] static cudaError_t SortPairsDescending(
    DoubleBuffer<KeyT>& d_keys,
    DoubleBuffer<ValueT>& d_values,
    NumItemsT num_items,
    int begin_bit         = 0,
    int end_bit           = sizeof(KeyT) * 8,
    const EnvT& env = {})
  {
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {
      return SortPairsDescending(t, b, d_keys, d_values, num_items, begin_bit, end_bit, s);
    });
  }

  // ------------------------------------------------------------------ SortKeys
  template <typename KeyT, typename NumItemsT>
  static cudaError_t SortKeys(
    void* d_temp_storage,
    size_t& temp_storage_bytes,
    const KeyT* d_keys_in,
    KeyT* d_keys_out,
    NumItemsT num_items,
    int begin_bit       = 0,
    int end_bit         = sizeof(KeyT) * 8,
    cudaStream_t stream = nullptr)
  {
    return detail::b200rs_pointer_sort<KeyT, NullType>(
      d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, nullptr, nullptr,
      static_cast<unsigned long long>(num_items), begin_bit, end_bit, false, stream);
  }

  template <typename KeyT, typename NumItemsT>
  static cudaError_t SortKeys(
    void* d_temp_storage,
    size_t& temp_storage_bytes,
    DoubleBuffer<KeyT>& d_keys,
    NumItemsT num_items,
    int begin_bit       = 0,
    int end_bit         = sizeof(KeyT) * 8,
    cudaStream_t stream = nullptr)
  {
    return detail::b200rs_double_buffer_sort<KeyT, NullType>(
      d_temp_storage, temp_storage_bytes, d_keys, nullptr, static_cast<unsigned long long>(num_items), begin_bit,
      end_bit, false, stream);
  }

  template <typename KeyT, typename NumItemsT, typename EnvT = stream_env,
            std::enable_if_t<std::is_integral<NumItemsT>::value && detail::b200rs_is_env<EnvT>::value, int> = 0>
  [[nodiscard]] static cudaError_t SortKeys(
    const KeyT* d_keys_in,
    KeyT* d_keys_out,
    NumItemsT num_items,
    int begin_bit         = 0,
    int end_bit           = sizeof(KeyT) * 8,
    const EnvT& env = {})
  {
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {
      return SortKeys(t, b, d_keys_in, d_keys_out, num_items, begin_bit, end_bit, s);
    });
  }

  template <typename KeyT, typename NumItemsT, typename EnvT = stream_env,
            std::enable_if_t<std::is_integral<NumItemsT>::value && detail::b200rs_is_env<EnvT>::value, int> = 0>
  [[nodiscard]] static cudaError_t SortKeys(
    DoubleBuffer<KeyT>& d_keys,
    NumItemsT num_items,
    int begin_bit         = 0,
    int end_bit           = sizeof(KeyT) * 8,
    const EnvT& env = {})
  {
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {
      return SortKeys(t, b, d_keys, num_items, begin_bit, end_bit, s);
    });
  }

  // ------------------------------------------------------------------ SortKeysDescending
  template <typename KeyT, typename NumItemsT>
  static cudaError_t SortKeysDescending(
    void* d_temp_storage,
    size_t& temp_storage_bytes,
    const KeyT* d_keys_in,
    KeyT* d_keys_out,
    NumItemsT num_items,
    int begin_bit       = 0,
    int end_bit         = sizeof(KeyT) * 8,
    cudaStream_t stream = nullptr)
  {
    return detail::b200rs_pointer_sort<KeyT, NullType>(
      d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, nullptr, nullptr,
      static_cast<unsigned long long>(num_items), begin_bit, end_bit, true, stream);
  }

  template <typename KeyT, typename NumItemsT>
  static cudaError_t SortKeysDescending(
    void* d_temp_storage,
    size_t& temp_storage_bytes,
    DoubleBuffer<KeyT>& d_keys,
    NumItemsT num_items,
    int begin_bit       = 0,
    int end_bit         = sizeof(KeyT) * 8,
    cudaStream_t stream = nullptr)
  {
    return detail::b200rs_double_buffer_sort<KeyT, NullType>(
      d_temp_storage, temp_storage_bytes, d_keys, nullptr, static_cast<unsigned long long>(num_items), begin_bit,
      end_bit, true, stream);
  }

  template <typename KeyT, typename NumItemsT, typename EnvT = stream_env,
            std::enable_if_t<std::is_integral<NumItemsT>::value && detail::b200rs_is_env<EnvT>::value, int> = 0>
  [[nodiscard]] static cudaError_t SortKeysDescending(
    const KeyT* d_keys_in,
    KeyT* d_keys_out,
    NumItemsT num_items,
    int begin_bit         = 0,
    int end_bit           = sizeof(KeyT) * 8,
    const EnvT& env = {})
  {
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {
      return SortKeysDescending(t, b, d_keys_in, d_keys_out, num_items, begin_bit, end_bit, s);
    });
  }

  template <typename KeyT, typename NumItemsT, typename EnvT = stream_env,
            std::enable_if_t<std::is_integral<NumItemsT>::value && detail::b200rs_is_env<EnvT>::value, int> = 0>
  [[nodiscard]] static cudaError_t SortKeysDescending(
    DoubleBuffer<KeyT>& d_keys,
    NumItemsT num_items,
    int begin_bit         = 0,
    int end_bit           = sizeof(KeyT) * 8,
    const EnvT& env = {})
  {
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {
      return SortKeysDescending(t, b, d_keys, num_items, begin_bit, end_bit, s);
    });
  }

  // ------------------------------------------------------------------ decomposer overloads (user-defined key types)
  // reference: device_radix_sort.cuh:671 / :905 (SortPairs), :1359 / :1565 (DoubleBuffer) and the Descending / SortKeys
  // twins; DecomposerT: cuda::std::tuple<ArithmeticTs&...> operator()(KeyT&), leftmost member most significant
#define B200RS_DECOMPOSER_OVERLOADS(PAIRS_NAME, KEYS_NAME, DESC)                                                        \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT,                                  \
            std::enable_if_t<!std::is_convertible<DecomposerT, int>::value, int> = 0>                                  \
  static cudaError_t PAIRS_NAME(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in,               \
                                KeyT* d_keys_out, const ValueT* d_values_in, ValueT* d_values_out, NumItemsT num_items, \
                                DecomposerT decomposer, cudaStream_t stream = nullptr)                                 \
  {                                                                                                                    \
    return detail::b200rs_fields_sort<KeyT, ValueT>(                                                                   \
      d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out,                            \
      static_cast<unsigned long long>(num_items), detail::b200rs_fields_of<KeyT>(decomposer), 0, -1, DESC, stream);     \
  }                                                                                                                    \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT,                                  \
            std::enable_if_t<!std::is_convertible<DecomposerT, int>::value, int> = 0>                                  \
  static cudaError_t PAIRS_NAME(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in,               \
                                KeyT* d_keys_out, const ValueT* d_values_in, ValueT* d_values_out, NumItemsT num_items, \
                                DecomposerT decomposer, int begin_bit, int end_bit, cudaStream_t stream = nullptr)     \
  {                                                                                                                    \
    return detail::b200rs_fields_sort<KeyT, ValueT>(                                                                   \
      d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out,                            \
      static_cast<unsigned long long>(num_items), detail::b200rs_fields_of<KeyT>(decomposer), begin_bit, end_bit, DESC, \
      stream);                                                                                                         \
  }                                                                                                                    \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT,                                  \
            std::enable_if_t<!std::is_convertible<DecomposerT, int>::value, int> = 0>                                  \
  static cudaError_t PAIRS_NAME(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys,          \
                                DoubleBuffer<ValueT>& d_values, NumItemsT num_items, DecomposerT decomposer,           \
                                cudaStream_t stream = nullptr)                                                         \
  {                                                                                                                    \
    return detail::b200rs_fields_sort_db<KeyT, ValueT>(d_temp_storage, temp_storage_bytes, d_keys, &d_values,          \
                                                       static_cast<unsigned long long>(num_items),                     \
                                                       detail::b200rs_fields_of<KeyT>(decomposer), 0, -1, DESC, stream); \
  }                                                                                                                    \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT,                                  \
            std::enable_if_t<!std::is_convertible<DecomposerT, int>::value, int> = 0>                                  \
  static cudaError_t PAIRS_NAME(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys,          \
                                DoubleBuffer<ValueT>& d_values, NumItemsT num_items, DecomposerT decomposer,           \
                                int begin_bit, int end_bit, cudaStream_t stream = nullptr)                             \
  {                                                                                                                    \
    return detail::b200rs_fields_sort_db<KeyT, ValueT>(                                                                \
      d_temp_storage, temp_storage_bytes, d_keys, &d_values, static_cast<unsigned long long>(num_items),               \
      detail::b200rs_fields_of<KeyT>(decomposer), begin_bit, end_bit, DESC, stream);                                   \
  }                                                                                                                    \
  template <typename KeyT, typename NumItemsT, typename DecomposerT,                                                   \
            std::enable_if_t<!std::is_convertible<DecomposerT, int>::value, int> = 0>                                  \
  static cudaError_t KEYS_NAME(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in,                \
                               KeyT* d_keys_out, NumItemsT num_items, DecomposerT decomposer,                          \
                               cudaStream_t stream = nullptr)                                                          \
  {                                                                                                                    \
    return detail::b200rs_fields_sort<KeyT, NullType>(                                                                 \
      d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, nullptr, nullptr,                                     \
      static_cast<unsigned long long>(num_items), detail::b200rs_fields_of<KeyT>(decomposer), 0, -1, DESC, stream);     \
  }                                                                                                                    \
  template <typename KeyT, typename NumItemsT, typename DecomposerT,                                                   \
            std::enable_if_t<!std::is_convertible<DecomposerT, int>::value, int> = 0>                                  \
  static cudaError_t KEYS_NAME(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in,                \
                               KeyT* d_keys_out, NumItemsT num_items, DecomposerT decomposer, int begin_bit,           \
                               int end_bit, cudaStream_t stream = nullptr)                                             \
  {                                                                                                                    \
    return detail::b200rs_fields_sort<KeyT, NullType>(                                                                 \
      d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, nullptr, nullptr,                                     \
      static_cast<unsigned long long>(num_items), detail::b200rs_fields_of<KeyT>(decomposer), begin_bit, end_bit, DESC, \
      stream);                                                                                                         \
  }                                                                                                                    \
  template <typename KeyT, typename NumItemsT, typename DecomposerT,                                                   \
            std::enable_if_t<!std::is_convertible<DecomposerT, int>::value, int> = 0>                                  \
  static cudaError_t KEYS_NAME(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys,           \
                               NumItemsT num_items, DecomposerT decomposer, cudaStream_t stream = nullptr)             \
  {                                                                                                                    \
    return detail::b200rs_fields_sort_db<KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys, nullptr,          \
                                                         static_cast<unsigned long long>(num_items),                   \
                                                         detail::b200rs_fields_of<KeyT>(decomposer), 0, -1, DESC,      \
                                                         stream);                                                      \
  }                                                                                                                    \
  template <typename KeyT, typename NumItemsT, typename DecomposerT,                                                   \
            std::enable_if_t<!std::is_convertible<DecomposerT, int>::value, int> = 0>                                  \
  static cudaError_t KEYS_NAME(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys,           \
                               NumItemsT num_items, DecomposerT decomposer, int begin_bit, int end_bit,                \
                               cudaStream_t stream = nullptr)                                                          \
  {                                                                                                                    \
    return detail::b200rs_fields_sort_db<KeyT, NullType>(                                                              \
      d_temp_storage, temp_storage_bytes, d_keys, nullptr, static_cast<unsigned long long>(num_items),                 \
      detail::b200rs_fields_of<KeyT>(decomposer), begin_bit, end_bit, DESC, stream);                                   \
  }
  B200RS_DECOMPOSER_OVERLOADS(SortPairs, SortKeys, false)
  B200RS_DECOMPOSER_OVERLOADS(SortPairsDescending, SortKeysDescending, true)
#undef B200RS_DECOMPOSER_OVERLOADS

  // ------------------------------------------------------------------ decomposer + environment overloads
  // reference: the env twins of the decomposer overloads (catch2_test_device_radix_sort_env_api.cu:227-400 passes a
  // decomposer and a cuda::stream_ref, with and without a bit window, pointer and DoubleBuffer forms): temporary storage
  // from the stream-ordered pool, everything else as the overloads above
#define B200RS_DECOMPOSER_ENV_GUARD                                                                                    \
  std::enable_if_t<!std::is_convertible<DecomposerT, int>::value && !detail::b200rs_is_env<DecomposerT>::value         \
                     && detail::b200rs_is_env<EnvT>::value && std::is_integral<NumItemsT>::value,                      \
                   int> = 0
#define B200RS_DECOMPOSER_ENV_OVERLOADS(PAIRS_NAME, KEYS_NAME)                                                         \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT, typename EnvT = stream_env,       \
            B200RS_DECOMPOSER_ENV_GUARD>                                                                               \
  [[nodiscard]] static cudaError_t PAIRS_NAME(const KeyT* d_keys_in, KeyT* d_keys_out, const ValueT* d_values_in,      \
                                              ValueT* d_values_out, NumItemsT num_items, DecomposerT decomposer,      \
                                              const EnvT& env = {})                                                    \
  {                                                                                                                    \
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {             \
      return PAIRS_NAME(t, b, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, decomposer, s);             \
    });                                                                                                                \
  }                                                                                                                    \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT, typename EnvT = stream_env,       \
            B200RS_DECOMPOSER_ENV_GUARD>                                                                               \
  [[nodiscard]] static cudaError_t PAIRS_NAME(const KeyT* d_keys_in, KeyT* d_keys_out, const ValueT* d_values_in,      \
                                              ValueT* d_values_out, NumItemsT num_items, DecomposerT decomposer,      \
                                              int begin_bit, int end_bit, const EnvT& env = {})                        \
  {                                                                                                                    \
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {             \
      return PAIRS_NAME(t, b, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, decomposer, begin_bit,      \
                        end_bit, s);                                                                                   \
    });                                                                                                                \
  }                                                                                                                    \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT, typename EnvT = stream_env,       \
            B200RS_DECOMPOSER_ENV_GUARD>                                                                               \
  [[nodiscard]] static cudaError_t PAIRS_NAME(DoubleBuffer<KeyT>& d_keys, DoubleBuffer<ValueT>& d_values,              \
                                              NumItemsT num_items, DecomposerT decomposer, const EnvT& env = {})       \
  {                                                                                                                    \
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {             \
      return PAIRS_NAME(t, b, d_keys, d_values, num_items, decomposer, s);                                             \
    });                                                                                                                \
  }                                                                                                                    \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename DecomposerT, typename EnvT = stream_env,       \
            B200RS_DECOMPOSER_ENV_GUARD>                                                                               \
  [[nodiscard]] static cudaError_t PAIRS_NAME(DoubleBuffer<KeyT>& d_keys, DoubleBuffer<ValueT>& d_values,              \
                                              NumItemsT num_items, DecomposerT decomposer, int begin_bit, int end_bit, \
                                              const EnvT& env = {})                                                    \
  {                                                                                                                    \
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {             \
      return PAIRS_NAME(t, b, d_keys, d_values, num_items, decomposer, begin_bit, end_bit, s);                         \
    });                                                                                                                \
  }                                                                                                                    \
  template <typename KeyT, typename NumItemsT, typename DecomposerT, typename EnvT = stream_env,                        \
            B200RS_DECOMPOSER_ENV_GUARD>                                                                               \
  [[nodiscard]] static cudaError_t KEYS_NAME(const KeyT* d_keys_in, KeyT* d_keys_out, NumItemsT num_items,             \
                                             DecomposerT decomposer, const EnvT& env = {})                             \
  {                                                                                                                    \
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {             \
      return KEYS_NAME(t, b, d_keys_in, d_keys_out, num_items, decomposer, s);                                         \
    });                                                                                                                \
  }                                                                                                                    \
  template <typename KeyT, typename NumItemsT, typename DecomposerT, typename EnvT = stream_env,                        \
            B200RS_DECOMPOSER_ENV_GUARD>                                                                               \
  [[nodiscard]] static cudaError_t KEYS_NAME(const KeyT* d_keys_in, KeyT* d_keys_out, NumItemsT num_items,             \
                                             DecomposerT decomposer, int begin_bit, int end_bit, const EnvT& env = {}) \
  {                                                                                                                    \
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {             \
      return KEYS_NAME(t, b, d_keys_in, d_keys_out, num_items, decomposer, begin_bit, end_bit, s);                     \
    });                                                                                                                \
  }                                                                                                                    \
  template <typename KeyT, typename NumItemsT, typename DecomposerT, typename EnvT = stream_env,                        \
            B200RS_DECOMPOSER_ENV_GUARD>                                                                               \
  [[nodiscard]] static cudaError_t KEYS_NAME(DoubleBuffer<KeyT>& d_keys, NumItemsT num_items, DecomposerT decomposer,  \
                                             const EnvT& env = {})                                                     \
  {                                                                                                                    \
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {             \
      return KEYS_NAME(t, b, d_keys, num_items, decomposer, s);                                                        \
    });                                                                                                                \
  }                                                                                                                    \
  template <typename KeyT, typename NumItemsT, typename DecomposerT, typename EnvT = stream_env,                        \
            B200RS_DECOMPOSER_ENV_GUARD>                                                                               \
  [[nodiscard]] static cudaError_t KEYS_NAME(DoubleBuffer<KeyT>& d_keys, NumItemsT num_items, DecomposerT decomposer,  \
                                             int begin_bit, int end_bit, const EnvT& env = {})                         \
  {                                                                                                                    \
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {             \
      return KEYS_NAME(t, b, d_keys, num_items, decomposer, begin_bit, end_bit, s);                                    \
    });                                                                                                                \
  }
  B200RS_DECOMPOSER_ENV_OVERLOADS(SortPairs, SortKeys)
  B200RS_DECOMPOSER_ENV_OVERLOADS(SortPairsDescending, SortKeysDescending)
#undef B200RS_DECOMPOSER_ENV_OVERLOADS
#undef B200RS_DECOMPOSER_ENV_GUARD
};

} // namespace cub
