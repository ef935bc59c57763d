// GPU test program for the cub::DeviceRadixSort header shim (include/cub/device/device_radix_sort.cuh).
// Mirrors the reference's own API tests: the env-API golden vectors of
// /root/reference/cub/test/catch2_test_device_radix_sort_env_api.cu:84-136 and the shape of
// catch2_test_device_radix_sort_{keys,pairs}.cu (pointer + DoubleBuffer, descending, bit windows, stream, stability),
// checked on the host against std::stable_sort over the bit-ordered key (the helper the reference tests use,
// catch2_radix_sort_helper.cuh:174-312).  Run by tests/test_cxx_shims.py under `pytest -m gpu`; exit code 0 == pass.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>
#include <cub/device/device_topk.cuh>
#include <cub/device/device_merge_sort.cuh>

#include <cuda/stream_ref>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

static int g_failed = 0;
#define REQUIRE(cond)                                                     \
  do                                                                      \
  {                                                                       \
    if (!(cond))                                                          \
    {                                                                     \
      printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);            \
      ++g_failed;                                                         \
    }                                                                     \
  } while (0)

template <class T>
struct dev
{
  T* p = nullptr;
  size_t n;
  explicit dev(size_t n_)
      : n(n_)
  {
    cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
  }
  explicit dev(const std::vector<T>& h)
      : dev(h.size())
  {
    cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice);
  }
  ~dev()
  {
    cudaFree(p);
  }
  std::vector<T> host() const
  {
    std::vector<T> h(n);
    cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost);
    return h;
  }
};

// order-preserving unsigned image of a key (Traits<T>::TwiddleIn, util_type.cuh:857-963)
template <class K>
uint64_t ordered_bits(K k)
{
  using U = std::conditional_t<sizeof(K) == 1, uint8_t,
            std::conditional_t<sizeof(K) == 2, uint16_t, std::conditional_t<sizeof(K) == 4, uint32_t, uint64_t>>>;
  U u;
  std::memcpy(&u, &k, sizeof(K));
  const U high = U(1) << (sizeof(K) * 8 - 1);
  if (std::is_floating_point<K>::value)
  {
    u = (u & high) ? U(~u) : U(u ^ high);
  }
  else if (std::is_signed<K>::value)
  {
    u ^= high;
  }
  return u;
}

// Largest N the reference sorts with its single-CTA kernel on sm_100 (dispatch_radix_sort.cuh:1980, policy
// tuning_radix_sort.cuh:1747,1833-1841 scaled by util_arch.cuh:128-138).  Only the float-zero rule depends on it.
inline size_t reference_single_tile_items(int key_bytes, int value_bytes)
{
  int dom     = std::max(std::max(key_bytes, value_bytes), 4);
  int items   = std::max(19 * 4 / dom, 1);
  int threads = std::min((48 * 1024 / (dom * items) + 31) / 32 * 32, 256);
  return size_t(threads) * size_t(items);
}

template <class K>
std::vector<uint32_t>
expected_permutation(const std::vector<K>& keys, bool desc, int begin_bit, int end_bit, int value_bytes = 0)
{
  std::vector<uint32_t> perm(keys.size());
  std::iota(perm.begin(), perm.end(), 0u);
  const int kbits     = int(sizeof(K)) * 8;
  const uint64_t all  = kbits == 64 ? ~0ull : ((1ull << kbits) - 1);
  const uint64_t high = 1ull << (kbits - 1);
  const int nbits     = end_bit - begin_bit;
  const uint64_t mask = nbits >= 64 ? ~0ull : ((1ull << nbits) - 1);
  const bool single   = keys.size() <= reference_single_tile_items(int(sizeof(K)), value_bytes);
  auto digit          = [&](uint32_t i) {
    uint64_t b = ordered_bits(keys[i]);
    if (desc)
    {
      b = ~b & all; // descending == ascending on the inverted bit image (radix_rank_sort_operations.cuh:533-573)
    }
    if (std::is_floating_point<K>::value)
    {
      // -0.0 == +0.0 (radix_rank_sort_operations.cuh:44-82): the onesweep path replaces the pattern 0x7f..f by
      // 0x80..0 in the (possibly inverted) key domain; the single-tile kernel applies the rule before inverting.
      if (desc && single)
      {
        b = b == high ? (~high & all) : b;
      }
      else
      {
        b = b == (~high & all) ? high : b;
      }
    }
    return (b >> begin_bit) & mask;
  };
  std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return digit(a) < digit(b); });
  return perm;
}

template <class K>
std::vector<K> random_keys(size_t n, unsigned seed, int and_rounds = 1)
{
  std::mt19937_64 rng(seed);
  std::vector<K> k(n);
  for (auto& x : k)
  {
    uint64_t r = rng();
    for (int i = 1; i < and_rounds; ++i)
    {
      r &= rng();
    }
    if (std::is_floating_point<K>::value)
    {
      // finite values, some zeros of both signs
      const int e = int(r % 41) - 20;
      double v    = double(int64_t(r >> 20) % 2000001 - 1000000) * std::ldexp(1.0, e);
      x           = K(v);
      if (r % 97 == 0) x = K(0.0);
      if (r % 89 == 0) x = K(-0.0);
    }
    else
    {
      std::memcpy(&x, &r, sizeof(K));
    }
  }
  return k;
}

template <class K>
bool same_bits(const std::vector<K>& a, const std::vector<K>& b)
{
  return a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(K)) == 0);
}

template <class K, class V>
void check_pairs(size_t n, bool desc, int begin_bit, int end_bit, bool double_buffer, cudaStream_t stream, int and_rounds)
{
  auto hk = random_keys<K>(n, 1234 + unsigned(n), and_rounds);
  std::vector<V> hv(n);
  for (size_t i = 0; i < n; ++i)
  {
    hv[i] = V(i);
  }
  auto perm = expected_permutation(hk, desc, begin_bit, end_bit, int(sizeof(V)));
  std::vector<K> ek(n);
  std::vector<V> ev(n);
  for (size_t i = 0; i < n; ++i)
  {
    ek[i] = hk[perm[i]];
    ev[i] = hv[perm[i]];
  }
  dev<K> k0(hk), k1(n);
  dev<V> v0(hv), v1(n);
  size_t bytes = 0;
  void* tmp    = nullptr;
  cudaError_t e;
  if (double_buffer)
  {
    cub::DoubleBuffer<K> dk(k0.p, k1.p);
    cub::DoubleBuffer<V> dv(v0.p, v1.p);
    e = desc ? cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, dk, dv, n, begin_bit, end_bit, stream)
             : cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, n, begin_bit, end_bit, stream);
    REQUIRE(e == cudaSuccess);
    REQUIRE(dk.selector == 0); // the size query must not touch the selector
    cudaMalloc(&tmp, bytes);
    e = desc ? cub::DeviceRadixSort::SortPairsDescending(tmp, bytes, dk, dv, n, begin_bit, end_bit, stream)
             : cub::DeviceRadixSort::SortPairs(tmp, bytes, dk, dv, n, begin_bit, end_bit, stream);
    REQUIRE(e == cudaSuccess);
    cudaStreamSynchronize(stream);
    REQUIRE(dk.selector == dv.selector);
    std::vector<K> gk(n);
    std::vector<V> gv(n);
    cudaMemcpy(gk.data(), dk.Current(), n * sizeof(K), cudaMemcpyDeviceToHost);
    cudaMemcpy(gv.data(), dv.Current(), n * sizeof(V), cudaMemcpyDeviceToHost);
    REQUIRE(same_bits(gk, ek));
    REQUIRE(same_bits(gv, ev));
  }
  else
  {
    e = desc ? cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, k0.p, k1.p, v0.p, v1.p, n, begin_bit, end_bit,
                                                         stream)
             : cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0.p, k1.p, v0.p, v1.p, n, begin_bit, end_bit, stream);
    REQUIRE(e == cudaSuccess);
    cudaMalloc(&tmp, bytes);
    e = desc ? cub::DeviceRadixSort::SortPairsDescending(tmp, bytes, k0.p, k1.p, v0.p, v1.p, n, begin_bit, end_bit,
                                                         stream)
             : cub::DeviceRadixSort::SortPairs(tmp, bytes, k0.p, k1.p, v0.p, v1.p, n, begin_bit, end_bit, stream);
    REQUIRE(e == cudaSuccess);
    cudaStreamSynchronize(stream);
    REQUIRE(same_bits(k1.host(), ek));
    REQUIRE(same_bits(v1.host(), ev));
    REQUIRE(same_bits(k0.host(), hk)); // the pointer API never writes its input (device_radix_sort.cuh:310)
    REQUIRE(same_bits(v0.host(), hv));
  }
  cudaFree(tmp);
}

template <class K>
void check_keys(size_t n, bool desc, int begin_bit, int end_bit, bool double_buffer, cudaStream_t stream)
{
  auto hk   = random_keys<K>(n, 99 + unsigned(n));
  auto perm = expected_permutation(hk, desc, begin_bit, end_bit);
  std::vector<K> ek(n);
  for (size_t i = 0; i < n; ++i)
  {
    ek[i] = hk[perm[i]];
  }
  dev<K> k0(hk), k1(n);
  size_t bytes = 0;
  void* tmp    = nullptr;
  cudaError_t e;
  if (double_buffer)
  {
    cub::DoubleBuffer<K> dk(k0.p, k1.p);
    e = desc ? cub::DeviceRadixSort::SortKeysDescending(nullptr, bytes, dk, n, begin_bit, end_bit, stream)
             : cub::DeviceRadixSort::SortKeys(nullptr, bytes, dk, n, begin_bit, end_bit, stream);
    REQUIRE(e == cudaSuccess);
    cudaMalloc(&tmp, bytes);
    e = desc ? cub::DeviceRadixSort::SortKeysDescending(tmp, bytes, dk, n, begin_bit, end_bit, stream)
             : cub::DeviceRadixSort::SortKeys(tmp, bytes, dk, n, begin_bit, end_bit, stream);
    REQUIRE(e == cudaSuccess);
    cudaStreamSynchronize(stream);
    std::vector<K> gk(n);
    cudaMemcpy(gk.data(), dk.Current(), n * sizeof(K), cudaMemcpyDeviceToHost);
    REQUIRE(same_bits(gk, ek));
  }
  else
  {
    e = desc ? cub::DeviceRadixSort::SortKeysDescending(nullptr, bytes, k0.p, k1.p, n, begin_bit, end_bit, stream)
             : cub::DeviceRadixSort::SortKeys(nullptr, bytes, k0.p, k1.p, n, begin_bit, end_bit, stream);
    REQUIRE(e == cudaSuccess);
    cudaMalloc(&tmp, bytes);
    e = desc ? cub::DeviceRadixSort::SortKeysDescending(tmp, bytes, k0.p, k1.p, n, begin_bit, end_bit, stream)
             : cub::DeviceRadixSort::SortKeys(tmp, bytes, k0.p, k1.p, n, begin_bit, end_bit, stream);
    REQUIRE(e == cudaSuccess);
    cudaStreamSynchronize(stream);
    REQUIRE(same_bits(k1.host(), ek));
    REQUIRE(same_bits(k0.host(), hk));
  }
  cudaFree(tmp);
}

// the reference's documented examples (catch2_test_device_radix_sort_env_api.cu:84-136 and the keys variants)
void env_api_goldens()
{
  const std::vector<int> keys{8, 6, 7, 5, 3, 0, 9}, vals{0, 1, 2, 3, 4, 5, 6};
  {
    dev<int> ki(keys), ko(7), vi(vals), vo(7);
    auto error = cub::DeviceRadixSort::SortPairs(ki.p, ko.p, vi.p, vo.p, static_cast<int>(keys.size()));
    cudaDeviceSynchronize();
    REQUIRE(error == cudaSuccess);
    REQUIRE((ko.host() == std::vector<int>{0, 3, 5, 6, 7, 8, 9}));
    REQUIRE((vo.host() == std::vector<int>{5, 4, 3, 1, 2, 0, 6}));
  }
  {
    dev<int> ki(keys), ko(7), vi(vals), vo(7);
    auto error = cub::DeviceRadixSort::SortPairsDescending(ki.p, ko.p, vi.p, vo.p, static_cast<int>(keys.size()));
    cudaDeviceSynchronize();
    REQUIRE(error == cudaSuccess);
    REQUIRE((ko.host() == std::vector<int>{9, 8, 7, 6, 5, 3, 0}));
    REQUIRE((vo.host() == std::vector<int>{6, 0, 2, 1, 3, 4, 5}));
  }
  {
    dev<int> ki(keys), ko(7);
    auto error = cub::DeviceRadixSort::SortKeys(ki.p, ko.p, static_cast<int>(keys.size()));
    cudaDeviceSynchronize();
    REQUIRE(error == cudaSuccess);
    REQUIRE((ko.host() == std::vector<int>{0, 3, 5, 6, 7, 8, 9}));
  }
  {
    dev<int> ki(keys), ko(7);
    cudaStream_t s;
    cudaStreamCreate(&s);
    auto error =
      cub::DeviceRadixSort::SortKeysDescending(ki.p, ko.p, static_cast<int>(keys.size()), 0, 32, cub::stream_env{s});
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    REQUIRE(error == cudaSuccess);
    REQUIRE((ko.host() == std::vector<int>{9, 8, 7, 6, 5, 3, 0}));
  }
  {
    // the reference's own env examples pass a cuda::stream_ref (catch2_test_device_radix_sort_env_api.cu:236-247)
    dev<int> ki(keys), ko(7), vi(vals), vo(7);
    cudaStream_t s;
    cudaStreamCreate(&s);
    cuda::stream_ref stream_ref{s};
    auto error = cub::DeviceRadixSort::SortPairs(ki.p, ko.p, vi.p, vo.p, static_cast<int>(keys.size()), 0, 32, stream_ref);
    cudaStreamSynchronize(s);
    REQUIRE(error == cudaSuccess);
    REQUIRE((ko.host() == std::vector<int>{0, 3, 5, 6, 7, 8, 9}));
    REQUIRE((vo.host() == std::vector<int>{5, 4, 3, 1, 2, 0, 6}));
    // a raw stream, and a user environment answering a stream (.get()) and a tuning query
    dev<int> k2(keys), o2(7);
    error = cub::DeviceRadixSort::SortKeysDescending(k2.p, o2.p, 7, 0, 32, s);
    cudaStreamSynchronize(s);
    REQUIRE(error == cudaSuccess);
    REQUIRE((o2.host() == std::vector<int>{9, 8, 7, 6, 5, 3, 0}));
    struct my_env
    {
      cudaStream_t s;
      cudaStream_t get() const
      {
        return s;
      }
      b200rs_tuning query(cub::b200rs_get_tuning_t) const
      {
        return b200rs_tuning{-1, 0, 0};
      }
    };
    cub::DoubleBuffer<int> dk(k2.p, o2.p);
    error = cub::DeviceRadixSort::SortKeys(dk, 7, 0, 32, my_env{s});
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    REQUIRE(error == cudaSuccess);
    std::vector<int> got(7);
    cudaMemcpy(got.data(), dk.Current(), 7 * sizeof(int), cudaMemcpyDeviceToHost);
    REQUIRE((got == std::vector<int>{0, 3, 5, 6, 7, 8, 9}));
  }
  {
    dev<int> a(keys), b(7), va(vals), vb(7);
    cub::DoubleBuffer<int> dk(a.p, b.p), dv(va.p, vb.p);
    auto error = cub::DeviceRadixSort::SortPairs(dk, dv, 7);
    cudaDeviceSynchronize();
    REQUIRE(error == cudaSuccess);
    std::vector<int> gk(7), gv(7);
    cudaMemcpy(gk.data(), dk.Current(), 28, cudaMemcpyDeviceToHost);
    cudaMemcpy(gv.data(), dv.Current(), 28, cudaMemcpyDeviceToHost);
    REQUIRE((gk == std::vector<int>{0, 3, 5, 6, 7, 8, 9}));
    REQUIRE((gv == std::vector<int>{5, 4, 3, 1, 2, 0, 6}));
  }
}

void edge_cases()
{
  // empty input: success, temp size >= 1, nothing touched (dispatch_radix_sort.cuh:1950-1977)
  size_t bytes = 0;
  auto e       = cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const uint32_t*) nullptr, (uint32_t*) nullptr, 0);
  REQUIRE(e == cudaSuccess);
  REQUIRE(bytes >= 1);
  // temp blob too small -> cudaErrorInvalidValue (util_temporary_storage.cuh:75-78)
  dev<uint32_t> a(1 << 16), b(1 << 16);
  bytes = 0;
  e     = cub::DeviceRadixSort::SortKeys(nullptr, bytes, a.p, b.p, 1 << 16);
  REQUIRE(e == cudaSuccess);
  size_t too_small = bytes / 2;
  void* tmp;
  cudaMalloc(&tmp, bytes);
  e = cub::DeviceRadixSort::SortKeys(tmp, too_small, a.p, b.p, 1 << 16);
  REQUIRE(e == cudaErrorInvalidValue);
  cudaFree(tmp);
  // begin_bit == end_bit: a copy (dispatch_radix_sort.cuh:1364-1410)
  check_keys<uint32_t>(10000, false, 7, 7, false, nullptr);
  check_keys<uint32_t>(10000, true, 7, 7, true, nullptr);
}

// 16-bit floating-point keys through the shim (reference: device_radix_sort.cuh:51-57, util_type.cuh:1017-1095):
// compared with a host stable sort on the float values; -0.0 / +0.0 tie and keep input order.
template <class H>
void half_keys(size_t n, bool descending)
{
  std::vector<H> keys(n);
  std::vector<uint32_t> vals(n);
  std::mt19937 rng(77);
  for (size_t i = 0; i < n; ++i)
  {
    const float f = float(int(rng() % 4001) - 2000) / 16.0f; // exactly representable in half and bfloat16
    keys[i]       = H(i % 13 == 0 ? -0.0f : (i % 17 == 0 ? 0.0f : f));
    vals[i]       = uint32_t(i);
  }
  dev<H> ki(keys), ko(n);
  dev<uint32_t> vi(vals), vo(n);
  size_t bytes = 0;
  auto e = descending ? cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, ki.p, ko.p, vi.p, vo.p, n)
                      : cub::DeviceRadixSort::SortPairs(nullptr, bytes, ki.p, ko.p, vi.p, vo.p, n);
  REQUIRE(e == cudaSuccess);
  void* tmp;
  cudaMalloc(&tmp, bytes);
  e = descending ? cub::DeviceRadixSort::SortPairsDescending(tmp, bytes, ki.p, ko.p, vi.p, vo.p, n)
                 : cub::DeviceRadixSort::SortPairs(tmp, bytes, ki.p, ko.p, vi.p, vo.p, n);
  cudaDeviceSynchronize();
  REQUIRE(e == cudaSuccess);
  cudaFree(tmp);
  std::vector<uint32_t> order(n);
  for (size_t i = 0; i < n; ++i)
  {
    order[i] = uint32_t(i);
  }
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    const float fa = float(keys[a]), fb = float(keys[b]);
    return descending ? fa > fb : fa < fb;
  });
  REQUIRE(vo.host() == order);
  const auto got = ko.host();
  bool same      = true;
  for (size_t i = 0; i < n; ++i)
  {
    same = same && memcmp(&got[i], &keys[order[i]], sizeof(H)) == 0;
  }
  REQUIRE(same);
}

// The multi-GPU partition pass through the C ABI (b200rs_partition_by_splitters): stable partition of (u32 key, u32
// value) pairs into destination-bucket order, bucket = 2 * #{splitters below} + [equals a splitter]; host check.
void partition_case(size_t n, int and_rounds)
{
  std::mt19937_64 rng(99 + n);
  std::vector<uint32_t> hk(n), hv(n);
  for (size_t i = 0; i < n; ++i)
  {
    uint32_t k = uint32_t(rng());
    for (int r = 1; r < and_rounds; ++r)
    {
      k &= uint32_t(rng());
    }
    hk[i] = k;
    hv[i] = uint32_t(i);
  }
  std::vector<uint32_t> sorted(hk);
  std::sort(sorted.begin(), sorted.end());
  std::vector<uint64_t> sp;
  for (int q = 1; q < 4 && n > 0; ++q)
  {
    const uint64_t v = sorted[n * q / 4];
    if (sp.empty() || sp.back() < v)
    {
      sp.push_back(v);
    }
  }
  const int ns = int(sp.size()), nb = 2 * ns + 1;
  auto bucket = [&](uint32_t k) {
    int id = 0;
    for (int j = 0; j < ns; ++j)
    {
      id += (k > sp[j] ? 1 : 0) + (k >= sp[j] ? 1 : 0);
    }
    return id;
  };
  std::vector<uint64_t> sizes(nb, 0), offs(nb, 0);
  for (size_t i = 0; i < n; ++i)
  {
    ++sizes[bucket(hk[i])];
  }
  for (int b = 1; b < nb; ++b)
  {
    offs[b] = offs[b - 1] + sizes[b - 1];
  }
  std::vector<uint32_t> ek(n), ev(n);
  {
    std::vector<uint64_t> cur(offs);
    for (size_t i = 0; i < n; ++i)
    {
      const uint64_t at = cur[bucket(hk[i])]++;
      ek[at]            = hk[i];
      ev[at]            = hv[i];
    }
  }
  dev<uint32_t> dk(hk), dv(hv), ok(n), ov(n);
  size_t bytes = 0;
  int rc = b200rs_partition_by_splitters(nullptr, &bytes, nullptr, nullptr, nullptr, nullptr, n, B200RS_KEY_UINT, 4, 4, 0,
                                         sp.data(), ns, offs.data(), nullptr);
  REQUIRE(rc == 0);
  dev<unsigned char> temp(bytes);
  rc = b200rs_partition_by_splitters(temp.p, &bytes, dk.p, ok.p, dv.p, ov.p, n, B200RS_KEY_UINT, 4, 4, 0, sp.data(), ns,
                                     offs.data(), nullptr);
  REQUIRE(rc == 0);
  REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
  REQUIRE(ok.host() == ek);
  REQUIRE(ov.host() == ev);
}

// Decomposer overloads (user-defined keys) and 128-bit keys.  Golden vectors: the reference's documentation examples,
// /root/reference/cub/test/catch2_test_device_radix_sort_custom.cu:479-700 (custom_t {float f; int unused; long long lli}).
struct custom_t
{
  float f;
  int unused;
  long long int lli;
  custom_t() = default;
  custom_t(float f_, long long int lli_)
      : f(f_)
      , unused(42)
      , lli(lli_)
  {}
};
struct decomposer_t
{
  __host__ __device__ cuda::std::tuple<float&, long long int&> operator()(custom_t& key) const
  {
    return {key.f, key.lli};
  }
};
static bool same_custom(const std::vector<custom_t>& a, const std::vector<custom_t>& b)
{
  bool ok = a.size() == b.size();
  for (size_t i = 0; ok && i < a.size(); ++i)
  {
    ok = memcmp(&a[i].f, &b[i].f, 4) == 0 && a[i].lli == b[i].lli && a[i].unused == b[i].unused;
  }
  return ok;
}

void decomposer_cases()
{
  const std::vector<custom_t> in{{+2.5f, 4}, {-2.5f, 0}, {+1.1f, 3}, {+0.0f, 1}, {-0.0f, 2}, {+3.7f, 5}};
  { // "Keys" (:513-556) and "KeysDescending" (:560-600)
    dev<custom_t> k(in), o(in.size());
    size_t bytes = 0;
    REQUIRE(cub::DeviceRadixSort::SortKeys(nullptr, bytes, k.p, o.p, 6, decomposer_t{}) == cudaSuccess);
    dev<unsigned char> temp(bytes);
    REQUIRE(cub::DeviceRadixSort::SortKeys(temp.p, bytes, k.p, o.p, 6, decomposer_t{}) == cudaSuccess);
    REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
    REQUIRE(same_custom(o.host(), {{-2.5f, 0}, {+0.0f, 1}, {-0.0f, 2}, {+1.1f, 3}, {+2.5f, 4}, {+3.7f, 5}}));
    const std::vector<custom_t> in2{{+1.1f, 2}, {+2.5f, 1}, {-0.0f, 4}, {+0.0f, 3}, {-2.5f, 5}, {+3.7f, 0}};
    dev<custom_t> k2(in2);
    REQUIRE(cub::DeviceRadixSort::SortKeysDescending(temp.p, bytes, k2.p, o.p, 6, decomposer_t{}) == cudaSuccess);
    REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
    REQUIRE(same_custom(o.host(), {{+3.7f, 0}, {+2.5f, 1}, {+1.1f, 2}, {-0.0f, 4}, {+0.0f, 3}, {-2.5f, 5}}));
  }
  { // "Pairs" (:602-650)
    dev<custom_t> k(in), o(in.size());
    dev<int> v(std::vector<int>{4, 0, 3, 1, 2, 5}), vo(6);
    size_t bytes = 0;
    REQUIRE(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k.p, o.p, v.p, vo.p, 6, decomposer_t{}) == cudaSuccess);
    dev<unsigned char> temp(bytes);
    REQUIRE(cub::DeviceRadixSort::SortPairs(temp.p, bytes, k.p, o.p, v.p, vo.p, 6, decomposer_t{}) == cudaSuccess);
    REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
    REQUIRE(same_custom(o.host(), {{-2.5f, 0}, {+0.0f, 1}, {-0.0f, 2}, {+1.1f, 3}, {+2.5f, 4}, {+3.7f, 5}}));
    REQUIRE((vo.host() == std::vector<int>{0, 1, 2, 3, 4, 5}));
  }
  { // decomposer + environment (catch2_test_device_radix_sort_env_api.cu:227-400): cuda::stream_ref as the env, with and
    // without a bit window, pointer and DoubleBuffer forms, keys and pairs
    cudaStream_t stream;
    cudaStreamCreate(&stream);
    cuda::stream_ref stream_ref{stream};
    const std::vector<custom_t> sorted{{-2.5f, 0}, {+0.0f, 1}, {-0.0f, 2}, {+1.1f, 3}, {+2.5f, 4}, {+3.7f, 5}};
    dev<custom_t> k(in), o(in.size());
    REQUIRE(cub::DeviceRadixSort::SortKeys(k.p, o.p, 6, decomposer_t{}, stream_ref) == cudaSuccess);
    cudaStreamSynchronize(stream);
    REQUIRE(same_custom(o.host(), sorted));
    dev<custom_t> o2(in.size());
    REQUIRE(cub::DeviceRadixSort::SortKeys(k.p, o2.p, 6, decomposer_t{}, 0, 96, stream_ref) == cudaSuccess);
    cudaStreamSynchronize(stream);
    REQUIRE(same_custom(o2.host(), sorted));
    dev<int> v(std::vector<int>{4, 0, 3, 1, 2, 5}), vo(6);
    dev<custom_t> o3(in.size());
    REQUIRE(cub::DeviceRadixSort::SortPairs(k.p, o3.p, v.p, vo.p, 6, decomposer_t{}, stream_ref) == cudaSuccess);
    cudaStreamSynchronize(stream);
    REQUIRE(same_custom(o3.host(), sorted));
    REQUIRE((vo.host() == std::vector<int>{0, 1, 2, 3, 4, 5}));
    dev<custom_t> ka(in), kb(in.size());
    dev<int> va(std::vector<int>{4, 0, 3, 1, 2, 5}), vb(6);
    cub::DoubleBuffer<custom_t> dk(ka.p, kb.p);
    cub::DoubleBuffer<int> dv(va.p, vb.p);
    REQUIRE(cub::DeviceRadixSort::SortPairsDescending(dk, dv, 6, decomposer_t{}, 0, 96, stream_ref) == cudaSuccess);
    cudaStreamSynchronize(stream);
    std::vector<custom_t> got(6);
    std::vector<int> gotv(6);
    cudaMemcpy(got.data(), dk.Current(), 6 * sizeof(custom_t), cudaMemcpyDeviceToHost);
    cudaMemcpy(gotv.data(), dv.Current(), 6 * sizeof(int), cudaMemcpyDeviceToHost);
    // -0.0 == +0.0 in the float member, so the two zeros are ordered by the second member (descending: 2 before 1)
    REQUIRE(same_custom(got, {{+3.7f, 5}, {+2.5f, 4}, {+1.1f, 3}, {-0.0f, 2}, {+0.0f, 1}, {-2.5f, 0}}));
    REQUIRE((gotv == std::vector<int>{5, 4, 3, 2, 1, 0}));
    cub::DoubleBuffer<custom_t> dk2(ka.p, kb.p);
    cudaMemcpy(ka.p, in.data(), 6 * sizeof(custom_t), cudaMemcpyHostToDevice);
    REQUIRE(cub::DeviceRadixSort::SortKeys(dk2, 6, decomposer_t{}, stream_ref) == cudaSuccess);
    cudaStreamSynchronize(stream);
    cudaMemcpy(got.data(), dk2.Current(), 6 * sizeof(custom_t), cudaMemcpyDeviceToHost);
    REQUIRE(same_custom(got, sorted));
    cudaStreamDestroy(stream);
  }
  // randomized: (float, long long) keys with many ties against a host stable sort of the tuple order; a bit window that
  // cuts into both members; DoubleBuffer form
  std::mt19937_64 rng(2024);
  for (size_t n : {size_t(1000), size_t(200003)})
  {
    std::vector<custom_t> hk(n);
    std::vector<uint64_t> hv(n);
    for (size_t i = 0; i < n; ++i)
    {
      hk[i] = custom_t(float(int(rng() % 13) - 6) * 0.5f, (long long) (rng() % 9) - 4);
      hv[i] = i;
    }
    auto ordered_f = [](float f) {
      uint32_t u;
      memcpy(&u, &f, 4);
      if ((u & 0x7fffffffu) == 0) { u = 0; }
      return (u & 0x80000000u) ? ~u : (u ^ 0x80000000u);
    };
    auto ordered_l = [](long long v) { return uint64_t(v) ^ 0x8000000000000000ull; };
    for (int variant = 0; variant < 3; ++variant)
    {
      const bool desc = variant == 1;
      const int b = variant == 2 ? 60 : 0, e = variant == 2 ? 70 : 96; // window: top 4 bits of lli + low 6 bits of f
      auto keybits = [&](const custom_t& c) { // 96-bit concatenation f:lli restricted to [b, e)
        unsigned __int128 x = ((unsigned __int128) ordered_f(c.f) << 64) | ordered_l(c.lli);
        x >>= b;
        const unsigned __int128 mask = (e - b) == 128 ? ~(unsigned __int128) 0 : (((unsigned __int128) 1 << (e - b)) - 1);
        return x & mask;
      };
      std::vector<size_t> order(n);
      std::iota(order.begin(), order.end(), size_t(0));
      std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t c) {
        return desc ? keybits(hk[a]) > keybits(hk[c]) : keybits(hk[a]) < keybits(hk[c]);
      });
      std::vector<custom_t> ek(n);
      std::vector<uint64_t> ev(n);
      for (size_t i = 0; i < n; ++i)
      {
        ek[i] = hk[order[i]];
        ev[i] = hv[order[i]];
      }
      dev<custom_t> k0(hk), k1(n);
      dev<uint64_t> v0(hv), v1(n);
      cub::DoubleBuffer<custom_t> dk(k0.p, k1.p);
      cub::DoubleBuffer<uint64_t> dv(v0.p, v1.p);
      size_t bytes = 0;
      auto call = [&](void* t) {
        return desc ? cub::DeviceRadixSort::SortPairsDescending(t, bytes, dk, dv, n, decomposer_t{})
             : variant == 2 ? cub::DeviceRadixSort::SortPairs(t, bytes, dk, dv, n, decomposer_t{}, b, e)
                            : cub::DeviceRadixSort::SortPairs(t, bytes, dk, dv, n, decomposer_t{});
      };
      REQUIRE(call(nullptr) == cudaSuccess);
      dev<unsigned char> temp(bytes);
      REQUIRE(call(temp.p) == cudaSuccess);
      REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
      std::vector<custom_t> gk(n);
      std::vector<uint64_t> gv(n);
      cudaMemcpy(gk.data(), dk.Current(), n * sizeof(custom_t), cudaMemcpyDeviceToHost);
      cudaMemcpy(gv.data(), dv.Current(), n * sizeof(uint64_t), cudaMemcpyDeviceToHost);
      REQUIRE(same_custom(gk, ek));
      REQUIRE(gv == ev);
    }
    // 128-bit integer keys through the ordinary overloads
    std::vector<__int128> h128(n);
    for (size_t i = 0; i < n; ++i)
    {
      h128[i] = ((__int128) (long long) (rng() % 7 - 3) << 64) | (unsigned long long) (rng() & rng());
    }
    std::vector<__int128> e128(h128);
    std::stable_sort(e128.begin(), e128.end());
    dev<__int128> k(h128), o(n);
    size_t bytes = 0;
    REQUIRE(cub::DeviceRadixSort::SortKeys(nullptr, bytes, k.p, o.p, n) == cudaSuccess);
    dev<unsigned char> temp(bytes);
    REQUIRE(cub::DeviceRadixSort::SortKeys(temp.p, bytes, k.p, o.p, n) == cudaSuccess);
    REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
    REQUIRE(o.host() == e128);
    std::reverse(e128.begin(), e128.end());
    REQUIRE(cub::DeviceRadixSort::SortKeysDescending(temp.p, bytes, k.p, o.p, n) == cudaSuccess);
    REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
    REQUIRE(o.host() == e128);
  }
}

// The call-site contract of cuda::std::sort(policy, first, last, comp) for arithmetic keys
// (/root/reference/libcudacxx/include/cuda/std/__pstl/cuda/sort.h:67-135): it takes the ADDRESS of
// cub::DeviceRadixSort::SortKeys[Descending] as cudaError_t(*)(void*, size_t&, DoubleBuffer<T>&, size_t, int, int,
// cudaStream_t), queries, sorts a DoubleBuffer{first, scratch} and copies back iff the selector moved.
template <class T>
void pstl_callsite_contract(bool descending)
{
  using radix_fn = cudaError_t (*)(void*, size_t&, cub::DoubleBuffer<T>&, size_t, int, int, cudaStream_t);
  radix_fn fn    = descending ? static_cast<radix_fn>(cub::DeviceRadixSort::SortKeysDescending)
                              : static_cast<radix_fn>(cub::DeviceRadixSort::SortKeys);
  std::mt19937_64 rng(5);
  const size_t n = 123457;
  std::vector<T> h(n);
  for (auto& x : h) { x = T(rng()); }
  dev<T> first(h), scratch(n);
  cub::DoubleBuffer<T> buffer{first.p, scratch.p};
  size_t bytes = 0;
  REQUIRE(fn(nullptr, bytes, buffer, n, 0, int(sizeof(T) * 8), nullptr) == cudaSuccess);
  dev<unsigned char> temp(bytes);
  REQUIRE(fn(temp.p, bytes, buffer, n, 0, int(sizeof(T) * 8), nullptr) == cudaSuccess);
  if (buffer.selector != 0)
  {
    cudaMemcpy(first.p, buffer.Current(), n * sizeof(T), cudaMemcpyDeviceToDevice);
  }
  REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
  std::sort(h.begin(), h.end());
  if (descending) { std::reverse(h.begin(), h.end()); }
  REQUIRE(first.host() == h);
}

// cub::DeviceSegmentedRadixSort shim: pointer + DoubleBuffer forms, keys and pairs, 32- and 64-bit offsets, gaps and
// empty segments (documentation example of device_segmented_radix_sort.cuh:140-170 first).
void segmented_cases()
{
  { // the reference's documentation example
    std::vector<int> hk{8, 6, 7, 5, 3, 0, 9}, hv{0, 1, 2, 3, 4, 5, 6}, off{0, 3, 3, 7};
    dev<int> k(hk), ko(hk.size()), v(hv), vo(hv.size()), o(off);
    size_t bytes = 0;
    REQUIRE(cub::DeviceSegmentedRadixSort::SortPairs(nullptr, bytes, k.p, ko.p, v.p, vo.p, 7, 3, o.p, o.p + 1) == cudaSuccess);
    dev<unsigned char> temp(bytes);
    REQUIRE(cub::DeviceSegmentedRadixSort::SortPairs(temp.p, bytes, k.p, ko.p, v.p, vo.p, 7, 3, o.p, o.p + 1) == cudaSuccess);
    REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
    REQUIRE((ko.host() == std::vector<int>{6, 7, 8, 0, 3, 5, 9}));
    REQUIRE((vo.host() == std::vector<int>{1, 2, 0, 5, 4, 3, 6}));
  }
  std::mt19937_64 rng(31);
  const std::vector<long long> lens{0, 1, 17, 5000, 5200, 0, 40000, 3, 9000};
  std::vector<long long> begins, ends;
  long long pos = 0;
  for (long long l : lens)
  {
    pos += 2; // a gap before every segment
    begins.push_back(pos);
    ends.push_back(pos + l);
    pos += l;
  }
  const size_t n = size_t(pos + 5);
  std::vector<uint64_t> hk(n);
  std::vector<uint32_t> hv(n);
  for (size_t i = 0; i < n; ++i)
  {
    hk[i] = rng() & rng() & 0xffffffffffull;
    hv[i] = uint32_t(i);
  }
  for (int desc = 0; desc < 2; ++desc)
  {
    std::vector<uint64_t> ek(hk);
    std::vector<uint32_t> ev(hv);
    for (size_t s = 0; s < begins.size(); ++s)
    {
      std::vector<size_t> idx(size_t(ends[s] - begins[s]));
      std::iota(idx.begin(), idx.end(), size_t(begins[s]));
      std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return desc ? hk[a] > hk[b] : hk[a] < hk[b]; });
      for (size_t j = 0; j < idx.size(); ++j)
      {
        ek[size_t(begins[s]) + j] = hk[idx[j]];
        ev[size_t(begins[s]) + j] = hv[idx[j]];
      }
    }
    dev<long long> db(begins), de(ends);
    { // pointer form; the output starts as the input so that untouched gaps compare equal
      dev<uint64_t> k(hk), ko(hk);
      dev<uint32_t> v(hv), vo(hv);
      size_t bytes = 0;
      auto call = [&](void* t) {
        return desc ? cub::DeviceSegmentedRadixSort::SortPairsDescending(t, bytes, k.p, ko.p, v.p, vo.p, (long long) n,
                                                                         (long long) begins.size(), db.p, de.p)
                    : cub::DeviceSegmentedRadixSort::SortPairs(t, bytes, k.p, ko.p, v.p, vo.p, (long long) n,
                                                               (long long) begins.size(), db.p, de.p);
      };
      REQUIRE(call(nullptr) == cudaSuccess);
      dev<unsigned char> temp(bytes);
      REQUIRE(call(temp.p) == cudaSuccess);
      REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
      REQUIRE(ko.host() == ek);
      REQUIRE(vo.host() == ev);
      REQUIRE(k.host() == hk);
    }
    { // DoubleBuffer form, keys only, 32-bit offsets
      std::vector<int> b32(begins.begin(), begins.end()), e32(ends.begin(), ends.end());
      dev<int> db32(b32), de32(e32);
      dev<uint64_t> k0(hk), k1(hk);
      cub::DoubleBuffer<uint64_t> dk(k0.p, k1.p);
      size_t bytes = 0;
      auto call = [&](void* t) {
        return desc ? cub::DeviceSegmentedRadixSort::SortKeysDescending(t, bytes, dk, (long long) n,
                                                                        (long long) begins.size(), db32.p, de32.p)
                    : cub::DeviceSegmentedRadixSort::SortKeys(t, bytes, dk, (long long) n, (long long) begins.size(),
                                                              db32.p, de32.p);
      };
      REQUIRE(call(nullptr) == cudaSuccess);
      REQUIRE(dk.selector == 0);
      dev<unsigned char> temp(bytes);
      REQUIRE(call(temp.p) == cudaSuccess);
      REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
      std::vector<uint64_t> got(n);
      cudaMemcpy(got.data(), dk.Current(), n * sizeof(uint64_t), cudaMemcpyDeviceToHost);
      REQUIRE(got == ek);
    }
  }
}

// Per-call tuning through the env overloads (reference: cuda::execution::tune in the env, device_radix_sort.cuh:200-202):
// every compiled tile configuration, the one-launch kernel on / off and the single-CTA kernel on / off give the same
// result, and the launch structure follows the tuning of THAT call only.
void tuned_env_cases()
{
  std::mt19937_64 rng(77);
  for (size_t n : {size_t(3000), size_t(150000)})
  {
    std::vector<uint32_t> h(n);
    for (auto& x : h) { x = uint32_t(rng()); }
    std::vector<uint32_t> e(h);
    std::sort(e.begin(), e.end());
    const int configs = b200rs_describe_config(4, 0, -1, nullptr, 0);
    REQUIRE(configs >= 2);
    for (int cfg = -1; cfg < configs; ++cfg)
    {
      for (int fused = 0; fused < 2; ++fused)
      {
        dev<uint32_t> k(h), o(n);
        const b200rs_tuning t{cfg, fused ? -1 : 0, fused ? -1 : 0};
        REQUIRE(cub::DeviceRadixSort::SortKeys(k.p, o.p, n, 0, 32, cub::stream_env(nullptr, t)) == cudaSuccess);
        const int ops = b200rs_last_launch_count();
        REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
        REQUIRE(o.host() == e);
        // default tuning: one launch (single-CTA kernel or the cooperative kernel); switched off, or with a forced tile
        // configuration: memset + histogram + scan + 4 passes
        REQUIRE((ops == 1) == (fused == 1 && (cfg < 0 || n <= 5120)));
      }
    }
    // an untuned call right after is unaffected
    dev<uint32_t> k(h), o(n);
    REQUIRE(cub::DeviceRadixSort::SortKeys(k.p, o.p, n) == cudaSuccess);
    REQUIRE(b200rs_last_launch_count() == 1);
    REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
    REQUIRE(o.host() == e);
  }
}

// cub::DeviceTopK (device_topk.cuh:297,775,1238): the K best keys in any order, any subset of the ties of the K-th key;
// checked as the reference's own test does (catch2_test_device_topk_*.cu): sorted output == first K of the sorted input,
// and for pairs every value (an input index) still carries its key.
template <class K>
void topk_case(size_t n, size_t k, bool largest, int and_rounds, bool pairs, cudaStream_t stream)
{
  std::mt19937_64 rng(n * 31 + k * 7 + (largest ? 1 : 0) + and_rounds);
  std::vector<K> keys(n);
  for (auto& x : keys)
  {
    uint64_t bits = rng();
    for (int r = 1; r < and_rounds; ++r)
    {
      bits &= rng();
    }
    using U = std::conditional_t<sizeof(K) == 1, uint8_t,
              std::conditional_t<sizeof(K) == 2, uint16_t, std::conditional_t<sizeof(K) == 4, uint32_t, uint64_t>>>;
    U u = U(bits);
    std::memcpy(&x, &u, sizeof(K));
  }
  if (std::is_floating_point<K>::value && n > 8)
  {
    keys[1] = K(-0.0);
    keys[5] = K(0.0);
  }
  std::vector<uint32_t> vals(n);
  std::iota(vals.begin(), vals.end(), 0u);
  const size_t kk = std::min(k, n);
  dev<K> ki(keys), ko(std::max<size_t>(kk, 1));
  dev<uint32_t> vi(vals), vo(std::max<size_t>(kk, 1));
  size_t bytes = 0;
  cudaError_t e;
  auto call = [&](void* t) {
    if (pairs)
    {
      return largest ? cub::DeviceTopK::MaxPairs(t, bytes, ki.p, ko.p, vi.p, vo.p, n, k, stream)
                     : cub::DeviceTopK::MinPairs(t, bytes, ki.p, ko.p, vi.p, vo.p, n, k, stream);
    }
    return largest ? cub::DeviceTopK::MaxKeys(t, bytes, ki.p, ko.p, n, k, stream)
                   : cub::DeviceTopK::MinKeys(t, bytes, ki.p, ko.p, n, k, stream);
  };
  e = call(nullptr);
  REQUIRE(e == cudaSuccess);
  dev<unsigned char> temp(bytes);
  e = call(temp.p);
  REQUIRE(e == cudaSuccess);
  REQUIRE(cudaStreamSynchronize(stream) == cudaSuccess);
  // order-preserving image with -0.0 == +0.0
  auto img = [&](K x) {
    uint64_t b = ordered_bits(x);
    if (std::is_floating_point<K>::value && x == K(0))
    {
      b = ordered_bits(K(0.0));
    }
    return b;
  };
  std::vector<uint64_t> want(n);
  for (size_t i = 0; i < n; ++i)
  {
    want[i] = img(keys[i]);
  }
  if (largest)
  {
    std::sort(want.begin(), want.end(), std::greater<uint64_t>());
  }
  else
  {
    std::sort(want.begin(), want.end());
  }
  want.resize(kk);
  auto got_k = ko.host();
  auto got_v = vo.host();
  std::vector<uint64_t> got(kk);
  for (size_t i = 0; i < kk; ++i)
  {
    got[i] = img(got_k[i]);
  }
  if (largest)
  {
    std::sort(got.begin(), got.end(), std::greater<uint64_t>());
  }
  else
  {
    std::sort(got.begin(), got.end());
  }
  REQUIRE(got == want);
  if (pairs)
  {
    std::vector<char> seen(n, 0);
    bool ok = true;
    for (size_t i = 0; i < kk; ++i)
    {
      ok = ok && got_v[i] < n && !seen[got_v[i]] && std::memcmp(&keys[got_v[i]], &got_k[i], sizeof(K)) == 0;
      if (got_v[i] < n)
      {
        seen[got_v[i]] = 1;
      }
    }
    REQUIRE(ok);
  }
}

void topk_cases(cudaStream_t stream, size_t max_n)
{
  for (size_t n : {size_t(1), size_t(7), size_t(1000), size_t(70001), size_t((1u << 21) + 5)})
  {
    if (n > max_n)
    {
      continue;
    }
    for (size_t k : {size_t(1), size_t(5), size_t(100), n / 2 + 1, n, n + 3})
    {
      topk_case<uint32_t>(n, k, true, 1, false, stream);
      topk_case<int32_t>(n, k, false, 1, true, stream);
      topk_case<float>(n, k, true, 1, true, stream);
      topk_case<uint32_t>(n, k, false, 5, true, stream); // few distinct keys: the K-th key has many ties
      topk_case<uint64_t>(n, k, true, 1, true, stream);
      topk_case<double>(n, k, false, 1, false, stream);
      topk_case<int16_t>(n, k, true, 1, true, stream);
      topk_case<uint8_t>(n, k, false, 1, false, stream);
    }
  }
  // env overload owning its temp storage, k == 0 and empty input
  std::vector<int> keys{8, 6, 7, 5, 3, 0, 9};
  dev<int> ki(keys), ko(3);
  REQUIRE(cub::DeviceTopK::MaxKeys(ki.p, ko.p, 7, 3, cuda::stream_ref{stream}) == cudaSuccess);
  cudaStreamSynchronize(stream);
  auto h = ko.host();
  std::sort(h.begin(), h.end());
  REQUIRE((h == std::vector<int>{7, 8, 9}));
  size_t bytes = 0;
  REQUIRE(cub::DeviceTopK::MinKeys(nullptr, bytes, ki.p, ko.p, 7, 0) == cudaSuccess);
  REQUIRE(bytes >= 1);
  REQUIRE(cub::DeviceTopK::MinKeys(ki.p, ko.p, 0, 3) == cudaSuccess);
}

// cub::DeviceMergeSort (device_merge_sort.cuh:250,467,712,926,1126,1313,1502): arbitrary comparators and key types, every
// variant stable; checked against std::stable_sort with the same comparator (as catch2_test_device_merge_sort.cu does).
struct ms_item
{
  int group;
  float weight;
  unsigned short tag;
};
struct ms_item_less
{
  __host__ __device__ bool operator()(const ms_item& a, const ms_item& b) const
  {
    return a.group != b.group ? a.group < b.group : a.weight > b.weight;
  }
};
struct ms_mod_less
{
  unsigned m;
  __host__ __device__ bool operator()(unsigned a, unsigned b) const
  {
    return a % m < b % m;
  }
};

void merge_sort_case(size_t n, unsigned mod, cudaStream_t stream)
{
  std::mt19937 rng(unsigned(n * 13 + mod));
  std::vector<unsigned> keys(n);
  std::vector<unsigned long long> vals(n);
  for (size_t i = 0; i < n; ++i)
  {
    keys[i] = rng();
    vals[i] = i;
  }
  const ms_mod_less cmp{mod};
  // expected: stable sort of the indices
  std::vector<unsigned long long> perm(vals);
  std::stable_sort(perm.begin(), perm.end(), [&](unsigned long long a, unsigned long long b) { return cmp(keys[a], keys[b]); });
  std::vector<unsigned> want(n);
  for (size_t i = 0; i < n; ++i)
  {
    want[i] = keys[perm[i]];
  }
  { // in place, pairs
    dev<unsigned> k(keys);
    dev<unsigned long long> v(vals);
    size_t bytes = 0;
    REQUIRE(cub::DeviceMergeSort::SortPairs(nullptr, bytes, k.p, v.p, n, cmp, stream) == cudaSuccess);
    dev<unsigned char> temp(bytes);
    REQUIRE(cub::DeviceMergeSort::SortPairs(temp.p, bytes, k.p, v.p, n, cmp, stream) == cudaSuccess);
    REQUIRE(cudaStreamSynchronize(stream) == cudaSuccess);
    REQUIRE(k.host() == want);
    REQUIRE(v.host() == perm);
  }
  { // copy, keys only: the input stays untouched
    dev<unsigned> k(keys), o(n);
    size_t bytes = 0;
    REQUIRE(cub::DeviceMergeSort::StableSortKeysCopy(nullptr, bytes, k.p, o.p, n, cmp, stream) == cudaSuccess);
    dev<unsigned char> temp(bytes);
    REQUIRE(cub::DeviceMergeSort::StableSortKeysCopy(temp.p, bytes, k.p, o.p, n, cmp, stream) == cudaSuccess);
    REQUIRE(cudaStreamSynchronize(stream) == cudaSuccess);
    REQUIRE(o.host() == want);
    REQUIRE(k.host() == keys);
  }
  { // struct keys with 1-byte values, env overload owning its temporary storage
    std::vector<ms_item> items(n);
    std::vector<unsigned char> tags(n);
    for (size_t i = 0; i < n; ++i)
    {
      items[i] = ms_item{int(rng() % 7), float(rng() % 5), (unsigned short) i};
      tags[i]  = (unsigned char) (i * 7);
    }
    std::vector<size_t> p2(n);
    std::iota(p2.begin(), p2.end(), size_t(0));
    std::stable_sort(p2.begin(), p2.end(), [&](size_t a, size_t b) { return ms_item_less{}(items[a], items[b]); });
    dev<ms_item> k(items);
    dev<unsigned char> v(tags);
    REQUIRE(cub::DeviceMergeSort::StableSortPairs(k.p, v.p, n, ms_item_less{}, cuda::stream_ref{stream}) == cudaSuccess);
    REQUIRE(cudaStreamSynchronize(stream) == cudaSuccess);
    auto gk = k.host();
    auto gv = v.host();
    bool ok = true;
    for (size_t i = 0; i < n; ++i)
    {
      ok = ok && gk[i].tag == items[p2[i]].tag && gk[i].group == items[p2[i]].group && gv[i] == tags[p2[i]];
    }
    REQUIRE(ok);
  }
}

int main()
{
  cudaStream_t stream;
  cudaStreamCreate(&stream);
  tuned_env_cases();
  segmented_cases();
  decomposer_cases();
  pstl_callsite_contract<int>(false);
  pstl_callsite_contract<unsigned long long>(true);
  // compute-sanitizer runs (tools/sanitize.sh) cap the problem size: racecheck is ~100x slower than native
  const char* cap_env  = getenv("B200RS_TEST_MAX_N");
  const size_t max_n   = cap_env != nullptr ? size_t(atoll(cap_env)) : ~size_t(0);
  env_api_goldens();
  edge_cases();
  topk_cases(stream, max_n);
  for (size_t n : {size_t(0), size_t(1), size_t(2), size_t(255), size_t(2047), size_t(2048), size_t(2049), size_t(6145),
                   size_t(100003), size_t((1u << 21) + 77)})
  {
    if (n <= max_n)
    {
      merge_sort_case(n, 1000, stream); // many ties: stability visible
      merge_sort_case(n, 0x7fffffffu, stream);
    }
  }
  for (size_t n : {size_t(3000), size_t(77777)})
  {
    partition_case(n, 1);
    partition_case(n, 4);
  }
  for (size_t n : {size_t(1000), size_t(300007)})
  {
    if (n > max_n)
    {
      continue;
    }
    half_keys<__half>(n, false);
    half_keys<__half>(n, true);
    half_keys<__nv_bfloat16>(n, false);
    half_keys<__nv_bfloat16>(n, true);
  }
  const size_t sizes[] = {1, 2, 255, 4864, 4865, 100000, (1u << 21) + 17};
  for (size_t n : sizes)
  {
    if (n > max_n)
    {
      continue;
    }
    check_keys<uint32_t>(n, false, 0, 32, false, stream);
    check_keys<int32_t>(n, true, 0, 32, true, stream);
    check_keys<float>(n, true, 8, 24, false, nullptr);
    check_keys<int64_t>(n, true, 16, 48, false, stream);
    check_keys<double>(n, false, 0, 64, true, stream);
    check_keys<uint8_t>(n, false, 0, 8, false, stream);
    check_keys<int16_t>(n, true, 3, 11, true, stream);
    check_pairs<uint64_t, uint32_t>(n, false, 0, 64, false, stream, 1);
    check_pairs<uint64_t, uint32_t>(n, false, 0, 64, true, stream, 5); // low entropy: many ties, stability visible
    check_pairs<uint32_t, uint32_t>(n, true, 0, 32, true, stream, 3);
    check_pairs<float, uint64_t>(n, false, 0, 32, false, stream, 1);
    check_pairs<uint8_t, uint16_t>(n, false, 0, 8, true, stream, 1);
    check_pairs<int32_t, uint8_t>(n, true, 5, 20, false, stream, 1);
  }
  cudaStreamDestroy(stream);
  if (g_failed == 0)
  {
    printf("test_cub_shim: all checks passed\n");
  }
  return g_failed == 0 ? 0 : 1;
}
