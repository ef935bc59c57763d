// TEST INFRASTRUCTURE ONLY -- thin C-ABI wrapper around the UNMODIFIED reference's own CPU sort:
// thrust::sort / thrust::sort_by_key with THRUST_DEVICE_SYSTEM=OMP (or CPP), compiled straight from the
// headers under /root/reference (thrust/thrust/system/omp/detail/sort.h:79-243 ->
// thrust/thrust/system/detail/sequential/stable_radix_sort.h:224-329).  No reference source is copied
// into this repo: this file only #includes the reference headers where they lie; the recipe is
// oracle/Makefile target `_ref` and the output goes to oracle/_ref/ (git-ignored, travels to the GPU box).
//
// Used (a) to pin oracle/radix_sort_oracle.cpp (integer keys; floats without +-0 ties, see
// SURVEY.md 8c) and (b) as the CPU baseline of config #1 (`cpu_baseline.kind = "reference"`).

#include <thrust/device_vector.h>
#include <thrust/functional.h>
#include <thrust/sort.h>

#include <chrono>
#include <cstdint>
#include <cstring>

#include <omp.h>

namespace
{

template <class K>
double sort_keys(void* keys, uint64_t n, int descending)
{
  K* k = static_cast<K*>(keys);
  thrust::device_vector<K> d(k, k + n); // "device" == host memory under the OMP/CPP backends
  auto t0 = std::chrono::steady_clock::now();
  if (descending)
  {
    thrust::sort(d.begin(), d.end(), thrust::greater<K>());
  }
  else
  {
    thrust::sort(d.begin(), d.end());
  }
  auto t1 = std::chrono::steady_clock::now();
  thrust::copy(d.begin(), d.end(), k);
  return std::chrono::duration<double>(t1 - t0).count();
}

template <class K, class V>
double sort_pairs(void* keys, void* vals, uint64_t n, int descending)
{
  K* k = static_cast<K*>(keys);
  V* v = static_cast<V*>(vals);
  thrust::device_vector<K> dk(k, k + n);
  thrust::device_vector<V> dv(v, v + n);
  auto t0 = std::chrono::steady_clock::now();
  if (descending)
  {
    thrust::sort_by_key(dk.begin(), dk.end(), dv.begin(), thrust::greater<K>());
  }
  else
  {
    thrust::sort_by_key(dk.begin(), dk.end(), dv.begin());
  }
  auto t1 = std::chrono::steady_clock::now();
  thrust::copy(dk.begin(), dk.end(), k);
  thrust::copy(dv.begin(), dv.end(), v);
  return std::chrono::duration<double>(t1 - t0).count();
}

template <class K>
double dispatch_values(void* keys, void* vals, uint64_t n, int value_bytes, int descending)
{
  switch (value_bytes)
  {
    case 0:
      return sort_keys<K>(keys, n, descending);
    case 1:
      return sort_pairs<K, uint8_t>(keys, vals, n, descending);
    case 2:
      return sort_pairs<K, uint16_t>(keys, vals, n, descending);
    case 4:
      return sort_pairs<K, uint32_t>(keys, vals, n, descending);
    case 8:
      return sort_pairs<K, uint64_t>(keys, vals, n, descending);
    default:
      return -1.0;
  }
}

} // namespace

extern "C" {

int ref_thrust_max_threads()
{
  return omp_get_max_threads();
}

void ref_thrust_set_threads(int t)
{
  omp_set_num_threads(t);
}

// In-place thrust::sort / sort_by_key on host arrays.  key_kind: 0 uint, 1 int, 2 float.
// Returns the wall-clock seconds of the sort call itself (copies excluded), or < 0 on bad arguments.
double ref_thrust_sort(void* keys, void* vals, uint64_t n, int key_kind, int key_bytes, int value_bytes, int descending)
{
  switch (key_kind * 16 + key_bytes)
  {
    case 0 * 16 + 1:
      return dispatch_values<uint8_t>(keys, vals, n, value_bytes, descending);
    case 0 * 16 + 2:
      return dispatch_values<uint16_t>(keys, vals, n, value_bytes, descending);
    case 0 * 16 + 4:
      return dispatch_values<uint32_t>(keys, vals, n, value_bytes, descending);
    case 0 * 16 + 8:
      return dispatch_values<uint64_t>(keys, vals, n, value_bytes, descending);
    case 1 * 16 + 1:
      return dispatch_values<int8_t>(keys, vals, n, value_bytes, descending);
    case 1 * 16 + 2:
      return dispatch_values<int16_t>(keys, vals, n, value_bytes, descending);
    case 1 * 16 + 4:
      return dispatch_values<int32_t>(keys, vals, n, value_bytes, descending);
    case 1 * 16 + 8:
      return dispatch_values<int64_t>(keys, vals, n, value_bytes, descending);
    case 2 * 16 + 4:
      return dispatch_values<float>(keys, vals, n, value_bytes, descending);
    case 2 * 16 + 8:
      return dispatch_values<double>(keys, vals, n, value_bytes, descending);
    default:
      return -1.0;
  }
}

} // extern "C"
