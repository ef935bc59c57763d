// GPU test program for the thrust::sort / sort_by_key shim (include/thrust/sort.h, include/thrust/device_vector.h).
// Golden vectors: /root/reference/thrust/testing/sort.cu:40-48 (TestSortSimple), sort_by_key.cu:46-53
// (InitializeSimpleKeyValueSortTest), plus the descending / stable / policy variants of
// thrust/testing/{sort,stable_sort,stable_sort_by_key}.cu checked against std::stable_sort on the host.
// Run by tests/test_cxx_shims.py under `pytest -m gpu`; exit code 0 == pass.
#include <thrust/device_vector.h>
#include <thrust/sort.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <numeric>
#include <random>
#include <cstdlib>
#include <vector>

static int g_failed = 0;
#define ASSERT_TRUE(cond)                                                 \
  do                                                                      \
  {                                                                       \
    if (!(cond))                                                          \
    {                                                                     \
      printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);            \
      ++g_failed;                                                         \
    }                                                                     \
  } while (0)

struct by_last_digit
{
  __host__ __device__ bool operator()(int a, int b) const
  {
    return a % 10 < b % 10;
  }
};

template <class K>
void random_case(size_t n, bool desc)
{
  std::mt19937_64 rng(n * 2 + desc);
  std::vector<K> hk(n);
  for (auto& k : hk)
  {
    k = K(int64_t(rng() % 2001) - 1000); // many duplicates: stability of the by-key variants is observable
  }
  std::vector<uint32_t> hv(n);
  std::iota(hv.begin(), hv.end(), 0u);
  std::vector<uint32_t> perm = hv;
  if (desc)
  {
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return hk[a] > hk[b]; });
  }
  else
  {
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return hk[a] < hk[b]; });
  }
  std::vector<K> ek(n);
  for (size_t i = 0; i < n; ++i)
  {
    ek[i] = hk[perm[i]];
  }

  thrust::device_vector<K> dk(hk);
  if (desc)
  {
    thrust::sort(dk.begin(), dk.end(), thrust::greater<K>());
  }
  else
  {
    thrust::sort(dk.begin(), dk.end());
  }
  ASSERT_TRUE(dk.to_host() == ek);

  thrust::device_vector<K> dk2(hk);
  thrust::device_vector<uint32_t> dv2(hv);
  if (desc)
  {
    thrust::stable_sort_by_key(dk2.begin(), dk2.end(), dv2.begin(), thrust::greater<K>());
  }
  else
  {
    thrust::sort_by_key(thrust::device, dk2.begin(), dk2.end(), dv2.begin());
  }
  ASSERT_TRUE(dk2.to_host() == ek);
  ASSERT_TRUE(dv2.to_host() == perm);
}

int main()
{
  { // TestSortSimple (thrust/testing/sort.cu:40-48)
    thrust::device_vector<int> v(std::vector<int>{1, 3, 6, 5, 2, 0, 4});
    thrust::sort(v.begin(), v.end());
    ASSERT_TRUE((v.to_host() == std::vector<int>{0, 1, 2, 3, 4, 5, 6}));
    thrust::sort(v.begin(), v.end(), thrust::greater<int>());
    ASSERT_TRUE((v.to_host() == std::vector<int>{6, 5, 4, 3, 2, 1, 0}));
    // the std / cuda::std function objects select the same radix path (sort.h:288-301); any other comparator is a
    // compile error in this shim (checked by tests/test_cxx_shims.py), never a silent ascending sort
    thrust::sort(v.begin(), v.end(), std::less<int>());
    ASSERT_TRUE((v.to_host() == std::vector<int>{0, 1, 2, 3, 4, 5, 6}));
    thrust::sort(v.begin(), v.end(), std::greater<int>());
    ASSERT_TRUE((v.to_host() == std::vector<int>{6, 5, 4, 3, 2, 1, 0}));
    thrust::stable_sort(v.begin(), v.end(), ::cuda::std::less<int>());
    ASSERT_TRUE((v.to_host() == std::vector<int>{0, 1, 2, 3, 4, 5, 6}));
    thrust::stable_sort(v.begin(), v.end(), ::cuda::std::greater<int>());
    ASSERT_TRUE((v.to_host() == std::vector<int>{6, 5, 4, 3, 2, 1, 0}));
  }
  { // user comparators take the comparison sort (sort.h:288-301 -> merge sort), stable
    std::vector<int> h(50001);
    std::mt19937 rng(7);
    for (auto& x : h)
    {
      x = int(rng() % 100000);
    }
    std::vector<int> idx(h.size());
    std::iota(idx.begin(), idx.end(), 0);
    thrust::device_vector<int> k(h), v(idx);
    thrust::sort_by_key(k.begin(), k.end(), v.begin(), by_last_digit());
    std::vector<int> p(idx);
    std::stable_sort(p.begin(), p.end(), [&](int a, int b) { return h[a] % 10 < h[b] % 10; });
    ASSERT_TRUE(v.to_host() == p);
    thrust::device_vector<int> k2(h);
    thrust::stable_sort(k2.begin(), k2.end(), by_last_digit());
    std::vector<int> e(h);
    std::stable_sort(e.begin(), e.end(), [](int a, int b) { return a % 10 < b % 10; });
    ASSERT_TRUE(k2.to_host() == e);
  }
  { // TestSortByKeySimple (thrust/testing/sort_by_key.cu:46-53)
    thrust::device_vector<int> k(std::vector<int>{1, 3, 6, 5, 2, 0, 4});
    thrust::device_vector<int> v(std::vector<int>{0, 1, 2, 3, 4, 5, 6});
    thrust::sort_by_key(k.begin(), k.end(), v.begin());
    ASSERT_TRUE((k.to_host() == std::vector<int>{0, 1, 2, 3, 4, 5, 6}));
    ASSERT_TRUE((v.to_host() == std::vector<int>{5, 0, 4, 1, 6, 3, 2}));
  }
  { // raw device pointers + a stream policy without synchronisation
    cudaStream_t s;
    cudaStreamCreate(&s);
    std::vector<float> h{3.5f, -0.0f, 0.0f, -7.25f, 1e30f, -1e30f, 2.0f};
    float* d;
    cudaMalloc(&d, h.size() * sizeof(float));
    cudaMemcpyAsync(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, s);
    thrust::stable_sort(thrust::cuda::par_nosync.on(s), d, d + h.size());
    std::vector<float> g(h.size());
    cudaMemcpyAsync(g.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    const std::vector<float> e{-1e30f, -7.25f, -0.0f, 0.0f, 2.0f, 3.5f, 1e30f};
    ASSERT_TRUE(std::equal(g.begin(), g.end(), e.begin(),
                           [](float a, float b) { return a == b && std::signbit(a) == std::signbit(b); }));
    cudaFree(d);
    cudaStreamDestroy(s);
  }
  { // empty and one-element ranges
    thrust::device_vector<uint64_t> e0;
    thrust::sort(e0.begin(), e0.end());
    thrust::device_vector<uint64_t> e1(std::vector<uint64_t>{42});
    thrust::sort(e1.begin(), e1.end());
    ASSERT_TRUE((e1.to_host() == std::vector<uint64_t>{42}));
  }
  const char* cap_env = getenv("B200RS_TEST_MAX_N"); // compute-sanitizer runs cap the size (tools/sanitize.sh)
  const size_t max_n  = cap_env != nullptr ? size_t(atoll(cap_env)) : ~size_t(0);
  for (size_t n : {size_t(17), size_t(5000), size_t(1) << 20})
  {
    if (n > max_n)
    {
      continue;
    }
    random_case<int8_t>(n, false);
    random_case<int16_t>(n, true);
    random_case<int32_t>(n, false);
    random_case<int64_t>(n, true);
    random_case<float>(n, true);
    random_case<double>(n, false);
  }
  if (g_failed == 0)
  {
    printf("test_thrust_shim: all checks passed\n");
  }
  return g_failed == 0 ? 0 : 1;
}
