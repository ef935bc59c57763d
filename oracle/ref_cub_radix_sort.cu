// TEST INFRASTRUCTURE ONLY -- command-line wrapper around the UNMODIFIED reference's GPU path,
// cub::DeviceRadixSort (/root/reference/cub/cub/device/device_radix_sort.cuh), compiled for sm_100a straight from
// the reference headers where they lie (recipe: oracle/Makefile target ref_cub; output oracle/_ref/ref_cub_radix_sort).
// No reference source is copied into this repo.  Two uses on the GPU box:
//   (1) parity: sort a raw key/value file with the real CUB and write the result, so tests can compare the new
//       kernels with the reference's own GPU output bit for bit;
//   (2) context number: time cub::DeviceRadixSort on the same GPU and inputs as bench.py ("cub on same GPU").
//
// usage: ref_cub_radix_sort sort  <ktype> <vbytes> <n> <desc> <begin_bit> <end_bit> <keys.bin> <vals.bin|-> <out_keys.bin> <out_vals.bin|->
//        ref_cub_radix_sort bench <ktype> <vbytes> <log2n> <desc> <begin_bit> <end_bit> <dist> <iters>
//          dist: a number R = bitwise AND of R uniform words (1 = uniform, 3 = bit entropy 0.544, 5 = 0.201; as
//          cub/benchmarks/nvbench_helper.cu:370-400), or equal | fewK (K distinct keys, e.g. few16) | sorted
//          log2n may also be a plain item count written as n=<items>
//        ref_cub_radix_sort segsort <ktype> <vbytes> <n> <desc> <begin_bit> <end_bit> <keys.bin> <vals.bin|-> <out_keys.bin> <out_vals.bin|-> <num_segments> <begin_offsets.bin> <end_offsets.bin>
//          (cub::DeviceSegmentedRadixSort; offsets are int64; the output buffers are pre-filled with the input so that
//           positions outside every segment are defined)
//   ktype in {u8,i8,u16,i16,f16,bf16,u32,i32,f32,u64,i64,f64}; vbytes in {0,4,8}
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#define CK(x)                                                                              \
  do                                                                                       \
  {                                                                                        \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess)                                                                 \
    {                                                                                      \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

template <class K, class V>
cudaError_t do_sort(void* tmp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, size_t n, bool desc, int b,
                    int e, bool pairs)
{
  if (pairs)
  {
    return desc ? cub::DeviceRadixSort::SortPairsDescending(tmp, bytes, kin, kout, vin, vout, n, b, e)
                : cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, n, b, e);
  }
  return desc ? cub::DeviceRadixSort::SortKeysDescending(tmp, bytes, kin, kout, n, b, e)
              : cub::DeviceRadixSort::SortKeys(tmp, bytes, kin, kout, n, b, e);
}

template <class K, class V>
cudaError_t do_segsort(void* tmp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, long long n, long long segs,
                       const long long* bo, const long long* eo, bool desc, int b, int e, bool pairs)
{
  using S = cub::DeviceSegmentedRadixSort;
  if (pairs)
  {
    return desc ? S::SortPairsDescending(tmp, bytes, kin, kout, vin, vout, n, segs, bo, eo, b, e)
                : S::SortPairs(tmp, bytes, kin, kout, vin, vout, n, segs, bo, eo, b, e);
  }
  return desc ? S::SortKeysDescending(tmp, bytes, kin, kout, n, segs, bo, eo, b, e)
              : S::SortKeys(tmp, bytes, kin, kout, n, segs, bo, eo, b, e);
}

static std::vector<char> slurp(const char* path, size_t bytes)
{
  std::vector<char> buf(bytes);
  FILE* f = fopen(path, "rb");
  if (!f || fread(buf.data(), 1, bytes, f) != bytes)
  {
    fprintf(stderr, "cannot read %zu bytes from %s\n", bytes, path);
    exit(3);
  }
  fclose(f);
  return buf;
}

static void dump(const char* path, const void* p, size_t bytes)
{
  FILE* f = fopen(path, "wb");
  if (!f || fwrite(p, 1, bytes, f) != bytes)
  {
    fprintf(stderr, "cannot write %s\n", path);
    exit(3);
  }
  fclose(f);
}

// cub::DeviceSegmentedRadixSort on files (instantiated for a few key types only: it is a checker for the segmented
// oracle, SURVEY.md 8f-1, not a benchmark)
template <class K, class V>
int run_seg(int argc, char** argv, bool pairs)
{
  (void) argc;
  const bool desc = atoi(argv[5]) != 0;
  const int b = atoi(argv[6]), e = atoi(argv[7]);
  {
    const size_t n    = strtoull(argv[4], nullptr, 10);
    const size_t segs = strtoull(argv[12], nullptr, 10);
    K *kin, *kout;
    V *vin = nullptr, *vout = nullptr;
    long long *bo, *eo;
    CK(cudaMalloc(&kin, n * sizeof(K) + 16));
    CK(cudaMalloc(&kout, n * sizeof(K) + 16));
    CK(cudaMalloc(&bo, segs * 8 + 16));
    CK(cudaMalloc(&eo, segs * 8 + 16));
    auto hk = slurp(argv[8], n * sizeof(K));
    auto hb = slurp(argv[13], segs * 8);
    auto he = slurp(argv[14], segs * 8);
    CK(cudaMemcpy(kin, hk.data(), n * sizeof(K), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(kout, hk.data(), n * sizeof(K), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(bo, hb.data(), segs * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(eo, he.data(), segs * 8, cudaMemcpyHostToDevice));
    if (pairs)
    {
      CK(cudaMalloc(&vin, n * sizeof(V) + 16));
      CK(cudaMalloc(&vout, n * sizeof(V) + 16));
      auto hv = slurp(argv[9], n * sizeof(V));
      CK(cudaMemcpy(vin, hv.data(), n * sizeof(V), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(vout, hv.data(), n * sizeof(V), cudaMemcpyHostToDevice));
    }
    size_t bytes = 0;
    CK((do_segsort<K, V>(nullptr, bytes, kin, kout, vin, vout, (long long) n, (long long) segs, bo, eo, desc, b, e, pairs)));
    void* tmp;
    CK(cudaMalloc(&tmp, bytes + 16));
    CK((do_segsort<K, V>(tmp, bytes, kin, kout, vin, vout, (long long) n, (long long) segs, bo, eo, desc, b, e, pairs)));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hk.data(), kout, n * sizeof(K), cudaMemcpyDeviceToHost));
    dump(argv[10], hk.data(), n * sizeof(K));
    if (pairs)
    {
      std::vector<char> hv(n * sizeof(V));
      CK(cudaMemcpy(hv.data(), vout, n * sizeof(V), cudaMemcpyDeviceToHost));
      dump(argv[11], hv.data(), n * sizeof(V));
    }
    // context timing for tools/segmented_bench.py: REF_SEG_ITERS=<iters> repeats the same call and prints its device time
    if (const char* its = getenv("REF_SEG_ITERS"))
    {
      const int iters = atoi(its) > 0 ? atoi(its) : 1;
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0));
      for (int i = 0; i < iters; ++i)
      {
        CK((do_segsort<K, V>(tmp, bytes, kin, kout, vin, vout, (long long) n, (long long) segs, bo, eo, desc, b, e, pairs)));
      }
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("{\"impl\": \"cub-3.6.0 DeviceSegmentedRadixSort (reference, same GPU)\", \"n\": %zu, \"segments\": %zu, "
             "\"ms\": %.4f, \"temp_bytes\": %zu}\n", n, segs, ms / iters, bytes);
    }
    return 0;
  }
  return 1;
}

template <class K, class V>
int run(int argc, char** argv, bool pairs)
{
  const std::string mode = argv[1];
  const bool desc        = atoi(argv[5]) != 0;
  const int b = atoi(argv[6]), e = atoi(argv[7]);
  if (mode == "sort")
  {
    const size_t n = strtoull(argv[4], nullptr, 10);
    K *kin, *kout;
    V *vin = nullptr, *vout = nullptr;
    CK(cudaMalloc(&kin, n * sizeof(K) + 16));
    CK(cudaMalloc(&kout, n * sizeof(K) + 16));
    auto hk = slurp(argv[8], n * sizeof(K));
    CK(cudaMemcpy(kin, hk.data(), n * sizeof(K), cudaMemcpyHostToDevice));
    if (pairs)
    {
      CK(cudaMalloc(&vin, n * sizeof(V) + 16));
      CK(cudaMalloc(&vout, n * sizeof(V) + 16));
      auto hv = slurp(argv[9], n * sizeof(V));
      CK(cudaMemcpy(vin, hv.data(), n * sizeof(V), cudaMemcpyHostToDevice));
    }
    size_t bytes = 0;
    CK((do_sort<K, V>(nullptr, bytes, kin, kout, vin, vout, n, desc, b, e, pairs)));
    void* tmp;
    CK(cudaMalloc(&tmp, bytes + 16));
    CK((do_sort<K, V>(tmp, bytes, kin, kout, vin, vout, n, desc, b, e, pairs)));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hk.data(), kout, n * sizeof(K), cudaMemcpyDeviceToHost));
    dump(argv[10], hk.data(), n * sizeof(K));
    if (pairs)
    {
      std::vector<char> hv(n * sizeof(V));
      CK(cudaMemcpy(hv.data(), vout, n * sizeof(V), cudaMemcpyDeviceToHost));
      dump(argv[11], hv.data(), n * sizeof(V));
    }
    return 0;
  }
  // bench
  const size_t n = strncmp(argv[4], "n=", 2) == 0 ? strtoull(argv[4] + 2, nullptr, 10) : size_t(1) << atoi(argv[4]);
  const std::string dist = argv[8];
  const bool numeric     = !dist.empty() && dist[0] >= '0' && dist[0] <= '9';
  const int and_rounds   = numeric ? atoi(argv[8]) : 1;
  const int iters        = atoi(argv[9]);
  std::vector<K> hk(n + 8 / sizeof(K));
  {
    std::mt19937_64 rng(42);
    uint64_t* w  = reinterpret_cast<uint64_t*>(hk.data());
    size_t words = (n * sizeof(K) + 7) / 8;
    if (dist == "equal")
    {
      for (size_t i = 0; i < words; ++i)
      {
        w[i] = 0x0123456701234567ull;
      }
    }
    else if (dist.rfind("few", 0) == 0)
    {
      const int k = atoi(dist.c_str() + 3) > 0 ? atoi(dist.c_str() + 3) : 16;
      std::vector<uint64_t> pool(k);
      for (auto& x : pool)
      {
        x = rng();
      }
      for (size_t i = 0; i < n; ++i)
      {
        memcpy(reinterpret_cast<char*>(hk.data()) + i * sizeof(K), &pool[rng() % k], sizeof(K));
      }
    }
    else
    {
      for (size_t i = 0; i < words; ++i)
      {
        uint64_t x = rng();
        for (int r = 1; r < and_rounds; ++r)
        {
          x &= rng();
        }
        w[i] = x;
      }
    }
  }
  K *kin, *kout;
  V *vin = nullptr, *vout = nullptr;
  CK(cudaMalloc(&kin, n * sizeof(K)));
  CK(cudaMalloc(&kout, n * sizeof(K)));
  CK(cudaMemcpy(kin, hk.data(), n * sizeof(K), cudaMemcpyHostToDevice));
  if (pairs)
  {
    CK(cudaMalloc(&vin, n * sizeof(V)));
    CK(cudaMalloc(&vout, n * sizeof(V)));
    CK(cudaMemset(vin, 1, n * sizeof(V)));
  }
  size_t bytes = 0;
  CK((do_sort<K, V>(nullptr, bytes, kin, kout, vin, vout, n, desc, b, e, pairs)));
  void* tmp;
  CK(cudaMalloc(&tmp, bytes));
  if (dist == "sorted")
  {
    // the timed input is the already sorted array (in the requested order, all bits)
    size_t sb = 0;
    CK((do_sort<K, V>(nullptr, sb, kin, kout, vin, vout, n, desc, 0, int(sizeof(K) * 8), false)));
    void* st;
    CK(cudaMalloc(&st, sb));
    CK((do_sort<K, V>(st, sb, kin, kout, vin, vout, n, desc, 0, int(sizeof(K) * 8), false)));
    CK(cudaMemcpy(kin, kout, n * sizeof(K), cudaMemcpyDeviceToDevice));
    CK(cudaFree(st));
  }
  for (int i = 0; i < 3; ++i)
  {
    CK((do_sort<K, V>(tmp, bytes, kin, kout, vin, vout, n, desc, b, e, pairs)));
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i)
  {
    CK((do_sort<K, V>(tmp, bytes, kin, kout, vin, vout, n, desc, b, e, pairs)));
  }
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  printf("{\"impl\": \"cub-3.6.0 (reference, same GPU)\", \"ktype\": \"%s\", \"vbytes\": %d, \"n\": %zu, \"desc\": %d, "
         "\"begin_bit\": %d, \"end_bit\": %d, \"dist\": \"%s\", \"ms\": %.4f, \"gkeys_s\": %.3f, \"temp_bytes\": %zu}\n",
         argv[2], pairs ? int(sizeof(V)) : 0, n, int(desc), b, e, dist.c_str(), ms, n / ms / 1e6, bytes);
  return 0;
}

template <class K>
int by_value(int argc, char** argv)
{
  const int vb = atoi(argv[3]);
  switch (vb)
  {
    case 0:
      return run<K, uint32_t>(argc, argv, false);
    case 4:
      return run<K, uint32_t>(argc, argv, true);
    case 8:
      return run<K, uint64_t>(argc, argv, true);
    default:
      fprintf(stderr, "vbytes must be 0, 4 or 8\n");
      return 1;
  }
}

int main(int argc, char** argv)
{
  if (argc < 10)
  {
    fprintf(stderr, "see the header of oracle/ref_cub_radix_sort.cu for usage\n");
    return 1;
  }
  const std::string kt = argv[2];
  if (std::string(argv[1]) == "segsort")
  {
    if (argc < 15)
    {
      fprintf(stderr, "segsort needs 14 arguments\n");
      return 1;
    }
    const bool pairs = atoi(argv[3]) == 4;
    if (kt == "u32") return run_seg<uint32_t, uint32_t>(argc, argv, pairs);
    if (kt == "f32") return run_seg<float, uint32_t>(argc, argv, pairs);
    if (kt == "u64") return run_seg<uint64_t, uint32_t>(argc, argv, pairs);
    if (kt == "i64") return run_seg<int64_t, uint32_t>(argc, argv, pairs);
    fprintf(stderr, "segsort: ktype in {u32,f32,u64,i64}, vbytes in {0,4}\n");
    return 1;
  }
  if (kt == "u8") return by_value<uint8_t>(argc, argv);
  if (kt == "i8") return by_value<int8_t>(argc, argv);
  if (kt == "u16") return by_value<uint16_t>(argc, argv);
  if (kt == "i16") return by_value<int16_t>(argc, argv);
  if (kt == "u32") return by_value<uint32_t>(argc, argv);
  if (kt == "i32") return by_value<int32_t>(argc, argv);
  if (kt == "f16") return by_value<__half>(argc, argv);
  if (kt == "bf16") return by_value<__nv_bfloat16>(argc, argv);
  if (kt == "f32") return by_value<float>(argc, argv);
  if (kt == "u64") return by_value<uint64_t>(argc, argv);
  if (kt == "i64") return by_value<int64_t>(argc, argv);
  if (kt == "f64") return by_value<double>(argc, argv);
  fprintf(stderr, "unknown key type %s\n", kt.c_str());
  return 1;
}
