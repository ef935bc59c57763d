// cub::DeviceTopK on top of libb200rs.so -- a header-only drop-in for the arithmetic-key surface of
// /root/reference/cub/cub/device/device_topk.cuh (cub 3.6.0): MaxPairs :297 / :406, MinPairs :775 / :884,
// MaxKeys :1238, MinKeys (same file, the twin of MaxKeys).
//
// Each overload keeps the reference's name, parameter order and defaults and type-erases to ONE call of the C ABI
// (include/b200rs.h, b200rs_topk).  Contract kept from the reference: the K selected items are written to
// d_keys_out / d_values_out[0 .. min(k, num_items)) in NO particular order, which of several keys tied with the K-th one
// are returned is unspecified (the reference demands `cuda::execution::require(determinism::not_guaranteed,
// output_ordering::unsorted)` in the env, device_topk.cuh:179-200 -- the only combination it implements); -0.0 == +0.0;
// d_temp_storage == nullptr => only temp_storage_bytes is written; work is enqueued on the env's stream.
// The environment is whatever the radix-sort shim accepts (cub::stream_env, cudaStream_t, cuda::stream_ref, a CCCL 3.x
// env carrying a stream); its requirements are not inspected.
#pragma once

#include "device_radix_sort.cuh"

namespace cub
{
namespace detail
{
template <class KeyT, class ValueT, class EnvT>
inline cudaError_t b200rs_topk_call(
  void* d_temp_storage,
  size_t& temp_storage_bytes,
  const KeyT* d_keys_in,
  KeyT* d_keys_out,
  const ValueT* d_values_in,
  ValueT* d_values_out,
  unsigned long long num_items,
  unsigned long long k,
  bool largest,
  const EnvT& env)
{
  static_assert(b200rs_is_env<EnvT>::value, "the last argument must be an execution environment carrying a stream");
  return static_cast<cudaError_t>(b200rs_topk(
    d_temp_storage,
    &temp_storage_bytes,
    d_keys_in,
    d_keys_out,
    d_values_in,
    d_values_out,
    num_items,
    k,
    b200rs_key_kind_of<KeyT>(),
    int(sizeof(KeyT)),
    b200rs_value_bytes_of<ValueT>(),
    largest ? 1 : 0,
    reinterpret_cast<b200rs_stream_t>(b200rs_env_view(env).stream)));
}
} // namespace detail

struct DeviceTopK
{
#define B200RS_TOPK_OVERLOADS(NAME, LARGEST)                                                                           \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename NumOutItemsT, typename EnvT = stream_env>     \
  static cudaError_t NAME##Pairs(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in,              \
                                 KeyT* d_keys_out, const ValueT* d_values_in, ValueT* d_values_out,                    \
                                 NumItemsT num_items, NumOutItemsT k, const EnvT& env = {})                            \
  {                                                                                                                    \
    return detail::b200rs_topk_call(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in,            \
                                    d_values_out, static_cast<unsigned long long>(num_items),                          \
                                    static_cast<unsigned long long>(k), LARGEST, env);                                 \
  }                                                                                                                    \
  template <typename KeyT, typename ValueT, typename NumItemsT, typename NumOutItemsT, typename EnvT = stream_env,     \
            std::enable_if_t<std::is_integral<NumItemsT>::value && std::is_integral<NumOutItemsT>::value, int> = 0>    \
  [[nodiscard]] static cudaError_t NAME##Pairs(const KeyT* d_keys_in, KeyT* d_keys_out, const ValueT* d_values_in,     \
                                               ValueT* d_values_out, NumItemsT num_items, NumOutItemsT k,              \
                                               const EnvT& env = {})                                                   \
  {                                                                                                                    \
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {             \
      return NAME##Pairs(t, b, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, k, s);                     \
    });                                                                                                                \
  }                                                                                                                    \
  template <typename KeyT, typename NumItemsT, typename NumOutItemsT, typename EnvT = stream_env>                      \
  static cudaError_t NAME##Keys(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in,               \
                                KeyT* d_keys_out, NumItemsT num_items, NumOutItemsT k, const EnvT& env = {})           \
  {                                                                                                                    \
    return detail::b200rs_topk_call(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out,                         \
                                    static_cast<const NullType*>(nullptr), static_cast<NullType*>(nullptr),            \
                                    static_cast<unsigned long long>(num_items), static_cast<unsigned long long>(k),    \
                                    LARGEST, env);                                                                     \
  }                                                                                                                    \
  template <typename KeyT, typename NumItemsT, typename NumOutItemsT, typename EnvT = stream_env,                      \
            std::enable_if_t<std::is_integral<NumItemsT>::value && std::is_integral<NumOutItemsT>::value, int> = 0>    \
  [[nodiscard]] static cudaError_t NAME##Keys(const KeyT* d_keys_in, KeyT* d_keys_out, NumItemsT num_items,            \
                                              NumOutItemsT k, const EnvT& env = {})                                    \
  {                                                                                                                    \
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {             \
      return NAME##Keys(t, b, d_keys_in, d_keys_out, num_items, k, s);                                                 \
    });                                                                                                                \
  }

  B200RS_TOPK_OVERLOADS(Max, true)
  B200RS_TOPK_OVERLOADS(Min, false)
#undef B200RS_TOPK_OVERLOADS
};
} // namespace cub
