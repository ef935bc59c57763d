// GPU test program for the `cccl.c.parallel` radix-sort entry points exported by libb200rs.so
// (include/b200rs_cccl_c.h).  Shape of the reference's own test, /root/reference/c/parallel/test/test_radix_sort.cpp:
// build -> size query -> sort -> cleanup for keys and pairs, both orders, bit windows, is_overwrite_okay; plus the
// serialize / deserialize round trip (:332-380).  Host check: std::stable_sort on the bit-ordered key.
#include <b200rs_cccl_c.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

static int g_failed = 0;
#define REQUIRE(cond)                                          \
  do                                                           \
  {                                                            \
    if (!(cond))                                               \
    {                                                          \
      printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      ++g_failed;                                              \
    }                                                          \
  } while (0)

template <class T>
cccl_type_enum type_enum();
template <> cccl_type_enum type_enum<int32_t>() { return CCCL_INT32; }
template <> cccl_type_enum type_enum<uint32_t>() { return CCCL_UINT32; }
template <> cccl_type_enum type_enum<uint64_t>() { return CCCL_UINT64; }
template <> cccl_type_enum type_enum<int64_t>() { return CCCL_INT64; }
template <> cccl_type_enum type_enum<float>() { return CCCL_FLOAT32; }
template <> cccl_type_enum type_enum<double>() { return CCCL_FLOAT64; }
template <> cccl_type_enum type_enum<uint16_t>() { return CCCL_UINT16; }

template <class T>
cccl_iterator_t pointer_it(T* p)
{
  cccl_iterator_t it;
  memset(&it, 0, sizeof(it));
  it.size       = sizeof(T*);
  it.alignment  = alignof(T*);
  it.type       = CCCL_POINTER;
  it.value_type = cccl_type_info{sizeof(T), alignof(T), type_enum<T>()};
  it.state      = p;
  return it;
}

template <class K>
uint64_t ordered(K k, bool desc)
{
  uint64_t u = 0;
  memcpy(&u, &k, sizeof(K));
  const uint64_t high = 1ull << (sizeof(K) * 8 - 1), all = sizeof(K) == 8 ? ~0ull : (1ull << (sizeof(K) * 8)) - 1;
  if (std::is_floating_point<K>::value)
  {
    if ((u & ~high & all) == 0) { u = 0; } // -0.0 ranks as +0.0
    u ^= (u & high) ? all : high;
  }
  else if (std::is_signed<K>::value)
  {
    u ^= high;
  }
  return desc ? (~u & all) : u;
}

template <class K, class V>
void run_case(size_t n, bool desc, bool pairs, int begin_bit, int end_bit, bool overwrite)
{
  std::mt19937_64 rng(n * 7 + desc + 2 * pairs);
  std::vector<K> hk(n);
  std::vector<V> hv(n);
  for (size_t i = 0; i < n; ++i)
  {
    uint64_t r = rng() & rng(); // some duplicates
    memcpy(&hk[i], &r, sizeof(K));
    // keep NaNs and exact zeros out of the host check (the reference's N-dependent +-0.0 rule is pinned against the real
    // cub::DeviceRadixSort in tests/test_vs_reference_gpu.py and tests/golden/cub_*.npz)
    if (std::is_floating_point<K>::value && (hk[i] != hk[i] || hk[i] == K(0))) { hk[i] = K(i % 7) + K(1); }
    hv[i] = V(i);
  }
  K *dk0, *dk1;
  V *dv0 = nullptr, *dv1 = nullptr;
  cudaMalloc(&dk0, std::max<size_t>(n, 1) * sizeof(K));
  cudaMalloc(&dk1, std::max<size_t>(n, 1) * sizeof(K));
  cudaMemcpy(dk0, hk.data(), n * sizeof(K), cudaMemcpyHostToDevice);
  if (pairs)
  {
    cudaMalloc(&dv0, std::max<size_t>(n, 1) * sizeof(V));
    cudaMalloc(&dv1, std::max<size_t>(n, 1) * sizeof(V));
    cudaMemcpy(dv0, hv.data(), n * sizeof(V), cudaMemcpyHostToDevice);
  }
  cccl_iterator_t kin = pointer_it(dk0), kout = pointer_it(dk1), vin = pointer_it(dv0), vout = pointer_it(dv1);
  cccl_op_t decomposer;
  memset(&decomposer, 0, sizeof(decomposer));
  cccl_device_radix_sort_build_result_t build;
  REQUIRE(cccl_device_radix_sort_build(&build, desc ? CCCL_DESCENDING : CCCL_ASCENDING, kin, vin, decomposer, "", 10, 0,
                                       nullptr, nullptr, nullptr, nullptr)
          == CUDA_SUCCESS);
  // serialization round trip gives an equivalent build result
  void* blob  = nullptr;
  size_t blen = 0;
  REQUIRE(cccl_device_radix_sort_serialize(&build, &blob, &blen) == CUDA_SUCCESS);
  cccl_device_radix_sort_build_result_t build2;
  REQUIRE(cccl_device_radix_sort_deserialize(&build2, blob, blen) == CUDA_SUCCESS);
  REQUIRE(cccl_device_radix_sort_load(&build2) == CUDA_SUCCESS);
  cccl_serialization_buffer_free(blob);
  REQUIRE(build2.key_type.size == sizeof(K) && build2.order == build.order);
  size_t bytes = 0;
  int selector = -1;
  REQUIRE(cccl_device_radix_sort(build2, nullptr, &bytes, kin, kout, vin, vout, decomposer, n, begin_bit, end_bit,
                                 overwrite, &selector, nullptr)
          == CUDA_SUCCESS);
  REQUIRE(selector == -1); // the size query does not touch the selector
  void* temp = nullptr;
  cudaMalloc(&temp, std::max<size_t>(bytes, 1));
  REQUIRE(cccl_device_radix_sort(build2, temp, &bytes, kin, kout, vin, vout, decomposer, n, begin_bit, end_bit, overwrite,
                                 &selector, nullptr)
          == CUDA_SUCCESS);
  REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
  REQUIRE(selector == 0 || selector == 1);
  if (!overwrite) { REQUIRE(selector == 1 || n == 0); }
  std::vector<K> gk(n);
  std::vector<V> gv(n);
  cudaMemcpy(gk.data(), selector == 1 ? dk1 : dk0, n * sizeof(K), cudaMemcpyDeviceToHost);
  if (pairs) { cudaMemcpy(gv.data(), selector == 1 ? dv1 : dv0, n * sizeof(V), cudaMemcpyDeviceToHost); }
  std::vector<size_t> order(n);
  std::iota(order.begin(), order.end(), size_t(0));
  const uint64_t mask = (end_bit - begin_bit) == 64 ? ~0ull : ((1ull << (end_bit - begin_bit)) - 1);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    return ((ordered(hk[a], desc) >> begin_bit) & mask) < ((ordered(hk[b], desc) >> begin_bit) & mask);
  });
  bool same = true;
  for (size_t i = 0; i < n; ++i)
  {
    const bool ok = memcmp(&gk[i], &hk[order[i]], sizeof(K)) == 0 && (!pairs || gv[i] == hv[order[i]]);
    if (!ok && same)
    {
      printf("first mismatch: key bytes %zu value bytes %zu n %zu desc %d pairs %d bits [%d,%d) overwrite %d selector %d at %zu\n",
             sizeof(K), sizeof(V), n, int(desc), int(pairs), begin_bit, end_bit, int(overwrite), selector, i);
    }
    same = same && ok;
  }
  REQUIRE(same);
  REQUIRE(cccl_device_radix_sort_cleanup(&build) == CUDA_SUCCESS);
  cudaFree(dk0); cudaFree(dk1); cudaFree(dv0); cudaFree(dv1); cudaFree(temp);
}

int main()
{
  for (size_t n : {size_t(0), size_t(1), size_t(3000), size_t(70001), size_t(1) << 21})
  {
    run_case<uint32_t, uint32_t>(n, false, false, 0, 32, false);
    run_case<int32_t, uint32_t>(n, true, true, 0, 32, true);
    run_case<uint64_t, uint32_t>(n, false, true, 0, 64, false);
    run_case<float, uint64_t>(n, true, true, 8, 24, false);
    run_case<double, uint32_t>(n, false, false, 0, 64, true);
    run_case<int64_t, uint16_t>(n, true, true, 16, 48, false);
  }
  { // what this build cannot do says so: user-defined (CCCL_STORAGE) keys, non-pointer iterators
    uint32_t* p = nullptr;
    cccl_iterator_t k = pointer_it(p), v = pointer_it(p);
    cccl_op_t decomposer;
    memset(&decomposer, 0, sizeof(decomposer));
    cccl_device_radix_sort_build_result_t build;
    k.value_type.type = CCCL_STORAGE;
    REQUIRE(cccl_device_radix_sort_build(&build, CCCL_ASCENDING, k, v, decomposer, "", 10, 0, nullptr, nullptr, nullptr,
                                         nullptr) == CUDA_ERROR_NOT_SUPPORTED);
    k = pointer_it(p);
    k.type = CCCL_ITERATOR;
    REQUIRE(cccl_device_radix_sort_build(&build, CCCL_ASCENDING, k, v, decomposer, "", 10, 0, nullptr, nullptr, nullptr,
                                         nullptr) == CUDA_ERROR_NOT_SUPPORTED);
  }
  if (g_failed == 0)
  {
    printf("test_cccl_c: all checks passed\n");
  }
  return g_failed == 0 ? 0 : 1;
}
