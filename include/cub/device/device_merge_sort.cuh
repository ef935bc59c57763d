// cub::DeviceMergeSort for sm_100a -- the comparison-sort fallback the reference takes whenever a sort cannot use the
// radix path: user comparators and key types without a bit-ordered image
// (/root/reference/cub/cub/device/device_merge_sort.cuh:250 SortPairs, :467 SortPairsCopy, :712 SortKeys, :926
//  SortKeysCopy, :1126 StableSortPairs, :1313 StableSortKeys, :1502 StableSortKeysCopy; dispatch/dispatch_merge_sort.cuh;
//  thrust front door: thrust/system/cuda/detail/sort.h:288-301 picks it when `can_use_primitive_sort` is false).
//
// Unlike the radix path this one cannot sit behind the C ABI: the comparator is a C++ type that has to be compiled into
// the kernels, so -- exactly as in the reference -- the kernels are templates in this header, instantiated by the
// caller's nvcc.  Same names, parameter order, two-phase temp-storage protocol, in-place semantics (the Copy variants
// leave the input untouched), stream-ordered, no allocation; every variant is STABLE (the reference's SortKeys /
// SortPairs make no promise, its StableSort* do -- one implementation serves both).
//
// Algorithm (a different arrangement from the reference's block sort + partition + merge agents, same asymptotics):
//   1. tile sort: one CTA of 256 threads sorts a tile of 256 * IPT items -- per-thread odd-even transposition network
//      in registers, then log2(256) rounds of merge-path merges through shared memory;
//   2. log2(#tiles) merge passes, ping-pong between the caller's array and the temp copy: one partition kernel finds
//      the merge-path split of every output tile (binary search in global memory, one thread per tile boundary), one
//      merge kernel per pass stages each output tile's two input ranges in shared memory, merges them (merge path per
//      thread + serial merge of IPT items) and writes the tile coalesced.
// Items past the end of a ragged tile never get a sentinel (there is none for an arbitrary comparator): every range is
// clipped to the valid length instead.  Ties always take the left run: stable.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <iterator>
#include <type_traits>

#include "device_radix_sort.cuh" // NullType, stream_env and the env plumbing

namespace cub
{
namespace detail
{
namespace b200ms
{
constexpr int NT = 256;

template <class K, class V>
constexpr int items_per_thread()
{
  constexpr size_t item = sizeof(K) + (std::is_same<V, NullType>::value ? 0 : sizeof(V));
  return item <= 16 ? 8 : item <= 32 ? 4 : item <= 80 ? 2 : 1;
}

template <class V>
struct has_values : std::integral_constant<bool, !std::is_same<V, NullType>::value>
{};

// storage for the values of a tile (empty for keys-only sorts)
template <class V, int N, bool = has_values<V>::value>
struct value_tile
{
  V v[N];
};
template <class V, int N>
struct value_tile<V, N, false>
{
  unsigned char v[1];
};

// merge path: how many of the first `diag` outputs of merge(A[0,la), B[0,lb)) come from A (ties take A)
template <class K, class Cmp>
__device__ __forceinline__ uint32_t merge_path(const K* a, uint32_t la, const K* b, uint32_t lb, uint32_t diag, Cmp& cmp)
{
  uint32_t lo = diag > lb ? diag - lb : 0u, hi = diag < la ? diag : la;
  while (lo < hi)
  {
    const uint32_t mid = (lo + hi) >> 1;
    if (!cmp(b[diag - 1 - mid], a[mid]))
    {
      lo = mid + 1; // a[mid] <= b[...]: a[mid] is among the first diag outputs
    }
    else
    {
      hi = mid;
    }
  }
  return lo;
}
template <class K, class Cmp>
__device__ __forceinline__ unsigned long long merge_path_global(
  const K* a, unsigned long long la, const K* b, unsigned long long lb, unsigned long long diag, Cmp& cmp)
{
  unsigned long long lo = diag > lb ? diag - lb : 0ull, hi = diag < la ? diag : la;
  while (lo < hi)
  {
    const unsigned long long mid = (lo + hi) >> 1;
    if (!cmp(b[diag - 1 - mid], a[mid]))
    {
      lo = mid + 1;
    }
    else
    {
      hi = mid;
    }
  }
  return lo;
}

// serial merge of IPT outputs starting at (ai, bi) of two runs living in one shared array: keys into out[], the
// shared-memory index each one came from into src[]
template <int IPT, class K, class Cmp>
__device__ __forceinline__ void serial_merge(
  const K* s, uint32_t ai, uint32_t a_end, uint32_t bi, uint32_t b_end, K (&out)[IPT], uint32_t (&src)[IPT], Cmp& cmp)
{
  K ka = ai < a_end ? s[ai] : K(), kb = bi < b_end ? s[bi] : K();
#pragma unroll
  for (int j = 0; j < IPT; ++j)
  {
    const bool take_b = bi < b_end && (ai >= a_end || cmp(kb, ka));
    out[j]            = take_b ? kb : ka;
    src[j]            = take_b ? bi : ai;
    if (take_b)
    {
      ++bi;
      kb = bi < b_end ? s[bi] : kb;
    }
    else
    {
      ++ai;
      ka = ai < a_end ? s[ai] : ka;
    }
  }
}

template <class K, class V, int IPT, class KeyInIt, class ValInIt, class Cmp>
__global__ void __launch_bounds__(NT)
tile_sort_kernel(KeyInIt keys_in, ValInIt vals_in, K* keys_out, V* vals_out, unsigned long long n, Cmp cmp)
{
  constexpr int TILE    = NT * IPT;
  constexpr bool VALUES = has_values<V>::value;
  // raw storage: key / value types with user-provided constructors cannot be __shared__ objects
  __shared__ __align__(16) unsigned char raw_k[sizeof(K) * TILE];
  __shared__ __align__(16) unsigned char raw_v[sizeof(value_tile<V, TILE>)];
  K* const sk = reinterpret_cast<K*>(raw_k);
  auto& sv    = *reinterpret_cast<value_tile<V, TILE>*>(raw_v);
  const unsigned long long base = (unsigned long long) blockIdx.x * TILE;
  const uint32_t valid          = uint32_t(n - base < (unsigned long long) TILE ? n - base : TILE);
  const uint32_t tid            = threadIdx.x;
  for (uint32_t p = tid; p < valid; p += NT)
  {
    sk[p] = keys_in[base + p];
    if constexpr (VALUES)
    {
      sv.v[p] = vals_in[base + p];
    }
  }
  __syncthreads();
  // ---- per-thread stable sort of IPT consecutive items (odd-even transposition; items past `valid` never move)
  K k[IPT];
  uint32_t src[IPT];
  const uint32_t first = tid * IPT;
#pragma unroll
  for (int j = 0; j < IPT; ++j)
  {
    src[j] = first + j;
    if (first + j < valid)
    {
      k[j] = sk[first + j];
    }
  }
#pragma unroll
  for (int round = 0; round < IPT; ++round)
  {
#pragma unroll
    for (int j = round & 1; j + 1 < IPT; j += 2)
    {
      if (first + j + 1 < valid && cmp(k[j + 1], k[j]))
      {
        const K t = k[j];
        k[j]      = k[j + 1];
        k[j + 1]  = t;
        const uint32_t u = src[j];
        src[j]           = src[j + 1];
        src[j + 1]       = u;
      }
    }
  }
  V v[VALUES ? IPT : 1];
  auto gather_and_store = [&]() {
    if constexpr (VALUES)
    {
#pragma unroll
      for (int j = 0; j < IPT; ++j)
      {
        if (first + j < valid)
        {
          v[j] = sv.v[src[j]];
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IPT; ++j)
    {
      if (first + j < valid)
      {
        sk[first + j] = k[j];
        if constexpr (VALUES)
        {
          sv.v[first + j] = v[j];
        }
      }
    }
    __syncthreads();
  };
  gather_and_store();
  // ---- merge rounds: runs of L items -> runs of 2L
  for (uint32_t L = IPT; L < uint32_t(TILE); L *= 2)
  {
    const uint32_t pair = first / (2 * L) * (2 * L);
    const uint32_t a0 = pair < valid ? pair : valid, a1 = pair + L < valid ? pair + L : valid;
    const uint32_t b1 = pair + 2 * L < valid ? pair + 2 * L : valid;
    const uint32_t la = a1 - a0, lb = b1 - a1;
    const uint32_t diag = first - pair < la + lb ? first - pair : la + lb;
    const uint32_t ai   = merge_path(sk + a0, la, sk + a1, lb, diag, cmp);
    serial_merge<IPT>(sk, a0 + ai, a1, a1 + (diag - ai), b1, k, src, cmp);
    gather_and_store();
  }
  for (uint32_t p = tid; p < valid; p += NT)
  {
    keys_out[base + p] = sk[p];
    if constexpr (VALUES)
    {
      vals_out[base + p] = sv.v[p];
    }
  }
}

// splits[t] = number of items of run A among the outputs before output tile t of its pair (t = 0 .. tiles)
template <class K, int IPT, class Cmp>
__global__ void merge_partition_kernel(
  const K* keys, unsigned long long n, unsigned long long width, unsigned long long tiles, unsigned long long* splits,
  Cmp cmp)
{
  constexpr unsigned long long TILE = NT * IPT;
  const unsigned long long t        = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t > tiles)
  {
    return;
  }
  const unsigned long long out0 = t * TILE < n ? t * TILE : n;
  const unsigned long long pair = out0 / (2 * width) * (2 * width);
  const unsigned long long a1   = pair + width < n ? pair + width : n;
  const unsigned long long b1   = pair + 2 * width < n ? pair + 2 * width : n;
  splits[t] = merge_path_global(keys + pair, a1 - pair, keys + a1, b1 - a1, out0 - pair, cmp);
}

template <class K, class V, int IPT, class Cmp>
__global__ void __launch_bounds__(NT) merge_pass_kernel(
  const K* keys_in, const V* vals_in, K* keys_out, V* vals_out, unsigned long long n, unsigned long long width,
  const unsigned long long* splits, Cmp cmp)
{
  constexpr int TILE    = NT * IPT;
  constexpr bool VALUES = has_values<V>::value;
  __shared__ __align__(16) unsigned char raw_k[sizeof(K) * TILE];
  __shared__ __align__(16) unsigned char raw_v[sizeof(value_tile<V, TILE>)];
  K* const sk = reinterpret_cast<K*>(raw_k);
  auto& sv    = *reinterpret_cast<value_tile<V, TILE>*>(raw_v);
  const uint32_t tid            = threadIdx.x;
  const unsigned long long out0 = (unsigned long long) blockIdx.x * TILE;
  const unsigned long long out1 = out0 + TILE < n ? out0 + TILE : n;
  const unsigned long long pair = out0 / (2 * width) * (2 * width);
  const unsigned long long a_hi = pair + width < n ? pair + width : n;   // end of run A == start of run B
  const unsigned long long b_hi = pair + 2 * width < n ? pair + 2 * width : n;
  // A-side split at both ends of this tile; the next tile may belong to the next pair, where its split restarts at 0
  const unsigned long long sa0 = splits[blockIdx.x];
  const bool last_of_pair      = out1 >= b_hi;
  const unsigned long long sa1 = last_of_pair ? a_hi - pair : splits[blockIdx.x + 1];
  const unsigned long long sb0 = (out0 - pair) - sa0, sb1 = (out1 - pair) - sa1;
  const uint32_t la = uint32_t(sa1 - sa0), lb = uint32_t(sb1 - sb0);
  const K* ga = keys_in + pair + sa0;
  const K* gb = keys_in + a_hi + sb0;
  for (uint32_t p = tid; p < la + lb; p += NT)
  {
    sk[p] = p < la ? ga[p] : gb[p - la];
    if constexpr (VALUES)
    {
      sv.v[p] = p < la ? vals_in[pair + sa0 + p] : vals_in[a_hi + sb0 + (p - la)];
    }
  }
  __syncthreads();
  const uint32_t total = la + lb;
  const uint32_t first = tid * IPT;
  const uint32_t diag  = first < total ? first : total;
  const uint32_t ai    = merge_path(sk, la, sk + la, lb, diag, cmp);
  K k[IPT];
  uint32_t src[IPT];
  serial_merge<IPT>(sk, ai, la, la + (diag - ai), total, k, src, cmp);
  V v[VALUES ? IPT : 1];
  if constexpr (VALUES)
  {
#pragma unroll
    for (int j = 0; j < IPT; ++j)
    {
      if (first + j < total)
      {
        v[j] = sv.v[src[j]];
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < IPT; ++j)
  {
    if (first + j < total)
    {
      sk[first + j] = k[j];
      if constexpr (VALUES)
      {
        sv.v[first + j] = v[j];
      }
    }
  }
  __syncthreads();
  for (uint32_t p = tid; p < total; p += NT)
  {
    keys_out[out0 + p] = sk[p];
    if constexpr (VALUES)
    {
      vals_out[out0 + p] = sv.v[p];
    }
  }
}

inline size_t align256(size_t x)
{
  return (x + 255) / 256 * 256;
}

// keys_in/vals_in -> keys_out/vals_out (may be the same arrays: in-place), temp = one more copy of both + the splits
template <class KeyInIt, class ValInIt, class K, class V, class Cmp>
cudaError_t merge_sort(
  void* d_temp_storage, size_t& temp_storage_bytes, KeyInIt keys_in, ValInIt vals_in, K* keys_out, V* vals_out,
  unsigned long long n, Cmp cmp, cudaStream_t stream)
{
  static_assert(std::is_trivially_copyable<K>::value, "keys are moved bitwise");
  constexpr int IPT                 = items_per_thread<K, V>();
  constexpr unsigned long long TILE = NT * IPT;
  constexpr bool VALUES             = has_values<V>::value;
  const unsigned long long tiles    = (n + TILE - 1) / TILE;
  const size_t keys_bytes = align256(size_t(n) * sizeof(K)), vals_bytes = VALUES ? align256(size_t(n) * sizeof(V)) : 0;
  const size_t split_bytes = align256(size_t(tiles + 1) * sizeof(unsigned long long));
  const size_t total       = tiles <= 1 ? 1 : keys_bytes + vals_bytes + split_bytes + 255;
  if (d_temp_storage == nullptr)
  {
    temp_storage_bytes = total;
    return cudaSuccess;
  }
  if (temp_storage_bytes < total)
  {
    return cudaErrorInvalidValue;
  }
  if (n == 0)
  {
    return cudaSuccess;
  }
  unsigned char* base = reinterpret_cast<unsigned char*>(align256(reinterpret_cast<size_t>(d_temp_storage)));
  K* keys_tmp         = reinterpret_cast<K*>(base);
  V* vals_tmp         = reinterpret_cast<V*>(base + keys_bytes);
  auto* splits        = reinterpret_cast<unsigned long long*>(base + keys_bytes + vals_bytes);
  int passes          = 0;
  for (unsigned long long w = TILE; w < n; w *= 2)
  {
    ++passes;
  }
  // the tile sort lands where an even number of merge passes is left to go, so the last pass writes the output arrays
  K* cur_k = (passes % 2 == 0) ? keys_out : keys_tmp;
  V* cur_v = (passes % 2 == 0) ? vals_out : vals_tmp;
  K* alt_k = (passes % 2 == 0) ? keys_tmp : keys_out;
  V* alt_v = (passes % 2 == 0) ? vals_tmp : vals_out;
  tile_sort_kernel<K, V, IPT><<<unsigned(tiles), NT, 0, stream>>>(keys_in, vals_in, cur_k, cur_v, n, cmp);
  cudaError_t e = cudaPeekAtLastError();
  for (unsigned long long w = TILE; w < n && e == cudaSuccess; w *= 2)
  {
    merge_partition_kernel<K, IPT><<<unsigned((tiles + 1 + 255) / 256), 256, 0, stream>>>(cur_k, n, w, tiles, splits, cmp);
    merge_pass_kernel<K, V, IPT><<<unsigned(tiles), NT, 0, stream>>>(cur_k, cur_v, alt_k, alt_v, n, w, splits, cmp);
    e = cudaPeekAtLastError();
    K* tk = cur_k;
    cur_k = alt_k;
    alt_k = tk;
    V* tv = cur_v;
    cur_v = alt_v;
    alt_v = tv;
  }
  return e;
}
} // namespace b200ms
} // namespace detail

struct DeviceMergeSort
{
  // ---- in place (device_merge_sort.cuh:250, :712, :1126, :1313)
  template <typename KeyT, typename ValueT, typename OffsetT, typename CompareOpT>
  static cudaError_t SortPairs(void* d_temp_storage, size_t& temp_storage_bytes, KeyT* d_keys, ValueT* d_items,
                               OffsetT num_items, CompareOpT compare_op, cudaStream_t stream = nullptr)
  {
    return detail::b200ms::merge_sort(d_temp_storage, temp_storage_bytes, static_cast<const KeyT*>(d_keys),
                                      static_cast<const ValueT*>(d_items), d_keys, d_items,
                                      static_cast<unsigned long long>(num_items), compare_op, stream);
  }
  template <typename KeyT, typename OffsetT, typename CompareOpT>
  static cudaError_t SortKeys(void* d_temp_storage, size_t& temp_storage_bytes, KeyT* d_keys, OffsetT num_items,
                              CompareOpT compare_op, cudaStream_t stream = nullptr)
  {
    return detail::b200ms::merge_sort(d_temp_storage, temp_storage_bytes, static_cast<const KeyT*>(d_keys),
                                      static_cast<const NullType*>(nullptr), d_keys, static_cast<NullType*>(nullptr),
                                      static_cast<unsigned long long>(num_items), compare_op, stream);
  }
  template <typename KeyT, typename ValueT, typename OffsetT, typename CompareOpT>
  static cudaError_t StableSortPairs(void* d_temp_storage, size_t& temp_storage_bytes, KeyT* d_keys, ValueT* d_items,
                                     OffsetT num_items, CompareOpT compare_op, cudaStream_t stream = nullptr)
  {
    return SortPairs(d_temp_storage, temp_storage_bytes, d_keys, d_items, num_items, compare_op, stream);
  }
  template <typename KeyT, typename OffsetT, typename CompareOpT>
  static cudaError_t StableSortKeys(void* d_temp_storage, size_t& temp_storage_bytes, KeyT* d_keys, OffsetT num_items,
                                    CompareOpT compare_op, cudaStream_t stream = nullptr)
  {
    return SortKeys(d_temp_storage, temp_storage_bytes, d_keys, num_items, compare_op, stream);
  }

  // ---- copy: the input ranges are left untouched (device_merge_sort.cuh:467, :926, :1502)
  template <typename KeyInputIteratorT, typename ValueInputIteratorT, typename KeyT, typename ValueT, typename OffsetT,
            typename CompareOpT>
  static cudaError_t SortPairsCopy(void* d_temp_storage, size_t& temp_storage_bytes, KeyInputIteratorT d_input_keys,
                                   ValueInputIteratorT d_input_items, KeyT* d_output_keys, ValueT* d_output_items,
                                   OffsetT num_items, CompareOpT compare_op, cudaStream_t stream = nullptr)
  {
    return detail::b200ms::merge_sort(d_temp_storage, temp_storage_bytes, d_input_keys, d_input_items, d_output_keys,
                                      d_output_items, static_cast<unsigned long long>(num_items), compare_op, stream);
  }
  template <typename KeyInputIteratorT, typename KeyT, typename OffsetT, typename CompareOpT>
  static cudaError_t SortKeysCopy(void* d_temp_storage, size_t& temp_storage_bytes, KeyInputIteratorT d_input_keys,
                                  KeyT* d_output_keys, OffsetT num_items, CompareOpT compare_op,
                                  cudaStream_t stream = nullptr)
  {
    return detail::b200ms::merge_sort(d_temp_storage, temp_storage_bytes, d_input_keys,
                                      static_cast<const NullType*>(nullptr), d_output_keys,
                                      static_cast<NullType*>(nullptr), static_cast<unsigned long long>(num_items),
                                      compare_op, stream);
  }
  template <typename KeyInputIteratorT, typename KeyT, typename OffsetT, typename CompareOpT>
  static cudaError_t StableSortKeysCopy(void* d_temp_storage, size_t& temp_storage_bytes,
                                        KeyInputIteratorT d_input_keys, KeyT* d_output_keys, OffsetT num_items,
                                        CompareOpT compare_op, cudaStream_t stream = nullptr)
  {
    return SortKeysCopy(d_temp_storage, temp_storage_bytes, d_input_keys, d_output_keys, num_items, compare_op, stream);
  }

  // ---- env overloads owning their temporary storage (device_merge_sort.cuh:775, :1378 and the pairs twins)
  template <typename KeyT, typename OffsetT, typename CompareOpT, typename EnvT = stream_env,
            std::enable_if_t<detail::b200rs_is_env<EnvT>::value && std::is_integral<OffsetT>::value, int> = 0>
  [[nodiscard]] static cudaError_t SortKeys(KeyT* d_keys, OffsetT num_items, CompareOpT compare_op, const EnvT& env = {})
  {
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {
      return SortKeys(t, b, d_keys, num_items, compare_op, s);
    });
  }
  template <typename KeyT, typename OffsetT, typename CompareOpT, typename EnvT = stream_env,
            std::enable_if_t<detail::b200rs_is_env<EnvT>::value && std::is_integral<OffsetT>::value, int> = 0>
  [[nodiscard]] static cudaError_t
  StableSortKeys(KeyT* d_keys, OffsetT num_items, CompareOpT compare_op, const EnvT& env = {})
  {
    return SortKeys(d_keys, num_items, compare_op, env);
  }
  template <typename KeyT, typename ValueT, typename OffsetT, typename CompareOpT, typename EnvT = stream_env,
            std::enable_if_t<detail::b200rs_is_env<EnvT>::value && std::is_integral<OffsetT>::value, int> = 0>
  [[nodiscard]] static cudaError_t
  SortPairs(KeyT* d_keys, ValueT* d_items, OffsetT num_items, CompareOpT compare_op, const EnvT& env = {})
  {
    return detail::b200rs_with_env(detail::b200rs_env_view(env), [&](void* t, size_t& b, cudaStream_t s) {
      return SortPairs(t, b, d_keys, d_items, num_items, compare_op, s);
    });
  }
  template <typename KeyT, typename ValueT, typename OffsetT, typename CompareOpT, typename EnvT = stream_env,
            std::enable_if_t<detail::b200rs_is_env<EnvT>::value && std::is_integral<OffsetT>::value, int> = 0>
  [[nodiscard]] static cudaError_t
  StableSortPairs(KeyT* d_keys, ValueT* d_items, OffsetT num_items, CompareOpT compare_op, const EnvT& env = {})
  {
    return SortPairs(d_keys, d_items, num_items, compare_op, env);
  }
};
} // namespace cub
