// TEST INFRASTRUCTURE, not product code: the UNMODIFIED reference cub::DeviceTopK (cub 3.6.0, compiled from
// /root/reference where it lies, sm_100a) as a command-line tool, for (1) parity of b200rs_topk against the real thing on
// the same GPU (tests/test_vs_reference_gpu.py) and (2) the "cub on the same GPU" context number of the top-k bench
// (tools/topk_bench.py; reference bench axes: cub/benchmarks/bench/topk/keys.cu:109-111).
//   ref_cub_topk run   <u32|i32|f32|u64|i64|f64> <n> <k> <largest> <keys_in.bin> <keys_out.bin>
//   ref_cub_topk bench <type> <log2 n> <log2 k> <largest> <and_rounds> <iters>
#include <cub/device/device_topk.cuh>

#include <cuda/__execution/determinism.h>
#include <cuda/__execution/output_ordering.h>
#include <cuda/__execution/require.h>
#include <cuda/stream_ref>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#define CK(x)                                                                      \
  do                                                                               \
  {                                                                                \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess)                                                         \
    {                                                                              \
      fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));  \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

template <class K>
int go(int argc, char** argv)
{
  const std::string mode = argv[1];
  const bool largest     = atoi(argv[5]) != 0;
  cudaStream_t stream;
  CK(cudaStreamCreate(&stream));
  auto env = cuda::std::execution::env{
    cuda::stream_ref{stream},
    cuda::execution::require(cuda::execution::determinism::not_guaranteed, cuda::execution::output_ordering::unsorted)};
  size_t n, k;
  std::vector<K> h;
  int iters = 1;
  if (mode == "run")
  {
    n = strtoull(argv[3], nullptr, 10);
    k = strtoull(argv[4], nullptr, 10);
    h.resize(n);
    FILE* f = fopen(argv[6], "rb");
    if (f == nullptr || fread(h.data(), sizeof(K), n, f) != n)
    {
      fprintf(stderr, "cannot read %s\n", argv[6]);
      return 2;
    }
    fclose(f);
  }
  else
  {
    n     = size_t(1) << atoi(argv[3]);
    k     = size_t(1) << atoi(argv[4]);
    iters = atoi(argv[7]);
    h.resize(n);
    std::mt19937_64 rng(12345);
    const int rounds = atoi(argv[6]);
    for (auto& x : h)
    {
      uint64_t b = rng();
      for (int r = 1; r < rounds; ++r)
      {
        b &= rng();
      }
      if (rounds == 0)
      {
        b = 4;
      }
      memcpy(&x, &b, sizeof(K));
    }
  }
  K *d_in, *d_out;
  CK(cudaMalloc(&d_in, n * sizeof(K)));
  CK(cudaMalloc(&d_out, (k ? k : 1) * sizeof(K)));
  CK(cudaMemcpy(d_in, h.data(), n * sizeof(K), cudaMemcpyHostToDevice));
  size_t bytes = 0;
  auto call    = [&](void* t) {
    return largest ? cub::DeviceTopK::MaxKeys(t, bytes, d_in, d_out, n, k, env)
                   : cub::DeviceTopK::MinKeys(t, bytes, d_in, d_out, n, k, env);
  };
  CK(call(nullptr));
  void* d_temp;
  CK(cudaMalloc(&d_temp, bytes ? bytes : 1));
  if (mode == "run")
  {
    CK(call(d_temp));
    CK(cudaStreamSynchronize(stream));
    std::vector<K> out(k);
    CK(cudaMemcpy(out.data(), d_out, k * sizeof(K), cudaMemcpyDeviceToHost));
    FILE* f = fopen(argv[7], "wb");
    fwrite(out.data(), sizeof(K), k, f);
    fclose(f);
    return 0;
  }
  for (int i = 0; i < 3; ++i)
  {
    CK(call(d_temp));
  }
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a, stream));
  for (int i = 0; i < iters; ++i)
  {
    CK(call(d_temp));
  }
  CK(cudaEventRecord(b, stream));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  ms /= iters;
  printf("{\"impl\": \"cub-3.6.0 DeviceTopK (reference, same GPU)\", \"key\": \"%s\", \"n\": %zu, \"k\": %zu, \"largest\": %d, "
         "\"ms\": %.5f, \"gkeys_per_s\": %.3f, \"temp_bytes\": %zu}\n",
         argv[2], n, k, int(largest), ms, n / ms / 1e6, bytes);
  return 0;
}

int main(int argc, char** argv)
{
  if (argc < 8)
  {
    fprintf(stderr, "usage: see the header of oracle/ref_cub_topk.cu\n");
    return 2;
  }
  const std::string kt = argv[2];
  if (kt == "u32") return go<uint32_t>(argc, argv);
  if (kt == "i32") return go<int32_t>(argc, argv);
  if (kt == "f32") return go<float>(argc, argv);
  if (kt == "u64") return go<uint64_t>(argc, argv);
  if (kt == "i64") return go<int64_t>(argc, argv);
  if (kt == "f64") return go<double>(argc, argv);
  fprintf(stderr, "unknown key type %s\n", kt.c_str());
  return 2;
}
